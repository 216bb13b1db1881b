"""CPU-only check of the host-side lowering inside libresql_b200.so (constant folding, CSE,
slot assignment, accumulator-machine emission): the device program printed by rq_debug_lower is
executed by a Python model of the VM and must reproduce the plan oracle's result, which is itself
pinned against the reference engine (test_oracle_pinned.py)."""
import numpy as np
import pytest

from common import plan_names, load_plan_dict, plan_tables, serialize_columns, assert_same_relation
from oracle import plan_oracle as PO
from resql_b200.plan import Plan
import vm_model


def _has_join(plan):
    return any(p["sink_kind"] == 2 or p["source_kind"] == 3 for p in plan["pipelines"])


def _has_string_key(plan):
    return any(PO._is_str(k[2], k[3]) for p in plan["pipelines"] if p["sink_kind"] == 1 for k in p["keys"])


# level "host": the host-level units (HUnit); level "device": the encoded program (UInsn) and the device value
# references of the sinks, i.e. encode_program / to_vref as well
@pytest.mark.parametrize("level", ["host", "device"])
@pytest.mark.parametrize("with_stats", [True, False])
@pytest.mark.parametrize("agg_impl", [vm_model.IMPL_REGAGG, vm_model.IMPL_LOWAGG])
@pytest.mark.parametrize("name", plan_names())
def test_lowered_program_matches_oracle(name, agg_impl, with_stats, level, sf001):
    d = load_plan_dict(name)
    if _has_join(d) or _has_string_key(d):
        pytest.skip("hash paths are lowered on the device side (covered by the gpu tests)")
    plan = Plan(d)
    tables = plan_tables(d, sf001)
    pool = d.get("strpool", "").encode("latin1")
    outs = []
    for pi, p in enumerate(d["pipelines"]):
        if p["source_kind"] == 1:
            t = d["tables"][p["source_id"]]
            src = [np.asarray(tables[t["name"]][c]) for c in t["columns"]]
        else:
            src = outs[p["source_id"]]
        # constant strings reach the VM as device addresses pool_base + offset; base is 0 here
        pool_strings = {}
        for nd in p["nodes"]:
            if nd[0] == 3:
                pool_strings[nd[4]] = pool[nd[4]:].split(b"\0")[0]
        run = vm_model.run_pipeline_vm if level == "host" else vm_model.run_pipeline_device
        outs.append(run(plan, pi, src, pool_strings, agg_impl, with_stats))
    last = d["pipelines"][-1]
    st = [k[2] for k in last["keys"]] + [v[2] for v in last["vals"]]
    sw = [k[3] for k in last["keys"]] + [v[3] for v in last["vals"]]
    got = serialize_columns(*PO.finish(d, outs[-1], st, sw))
    want = serialize_columns(*PO.run_plan(d, tables))
    assert_same_relation(got, want, d, name)
