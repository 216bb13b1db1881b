"""CPU-only: the host-side lowering (constant folding, CSE, fused compares / ranges / multiply-adds,
32-bit forms from interval analysis, slot assignment) of RANDOM plans (tests/plan_fuzz.py), executed
by the Python model of the device VM, must reproduce the plan oracle. Join plans and string group
keys are lowered against device state and are covered by the GPU suite."""
import numpy as np
import pytest

from common import plan_tables, serialize_columns, assert_same_relation
from oracle import plan_oracle as PO
from plan_fuzz import random_plan
from resql_b200.plan import Plan
import vm_model


def _eligible(d):
    if any(p["sink_kind"] == 2 or p["source_kind"] == 3 for p in d["pipelines"]):
        return False
    return not any(PO._is_str(k[2], k[3]) for p in d["pipelines"] if p["sink_kind"] == 1 for k in p["keys"])


SEEDS = [s for s in range(160) if _eligible(random_plan(s))]


def test_enough_random_plans_are_eligible():
    assert len(SEEDS) >= 50


@pytest.mark.parametrize("level", ["host", "device"])
@pytest.mark.parametrize("agg_impl", [vm_model.IMPL_REGAGG, vm_model.IMPL_LOWAGG])
@pytest.mark.parametrize("seed", SEEDS)
def test_lowered_random_plan_matches_oracle(seed, agg_impl, level, sf001):
    d = random_plan(seed)
    tables = plan_tables(d, sf001)
    try:
        want = serialize_columns(*PO.run_plan(d, tables))
    except ZeroDivisionError:
        pytest.skip("plan divides by zero")
    plan = Plan(d)
    pool = d.get("strpool", "").encode("latin1")
    outs = []
    for pi, p in enumerate(d["pipelines"]):
        if p["source_kind"] == 1:
            t = d["tables"][p["source_id"]]
            src = [np.asarray(tables[t["name"]][c]) for c in t["columns"]]
        else:
            src = outs[p["source_id"]]
        pool_strings = {nd[4]: pool[nd[4]:].split(b"\0")[0] for nd in p["nodes"] if nd[0] == 3}
        run = vm_model.run_pipeline_vm if level == "host" else vm_model.run_pipeline_device
        outs.append(run(plan, pi, src, pool_strings, agg_impl, True))
    last = d["pipelines"][-1]
    st = [k[2] for k in last["keys"]] + [v[2] for v in last["vals"]]
    sw = [k[3] for k in last["keys"]] + [v[3] for v in last["vals"]]
    got = serialize_columns(*PO.finish(d, outs[-1], st, sw))
    assert_same_relation(got, want, d, f"random plan {seed}")
