"""Differential test: random plans (tests/plan_fuzz.py) through the C ABI on the B200 against the
plan oracle, which is itself pinned to the reference engine (test_oracle_pinned.py). Bit-exact:
same multiset of serialized tuples, same ORDER BY key sequence."""
import pytest

from common import plan_tables, serialize_columns, assert_same_relation
from oracle.plan_oracle import run_plan
from plan_fuzz import random_plan
from resql_b200 import Plan

pytestmark = pytest.mark.gpu

SEEDS = list(range(250))


@pytest.fixture(scope="module")
def engine():
    from resql_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.shutdown()


@pytest.fixture(scope="module")
def tables_on_gpu(engine, sf001):
    from plan_fuzz import COLS
    up = {}
    for t, cols in COLS.items():
        up[t] = engine.upload(t, {c: sf001[t][c] for c, _ in cols})
    yield up
    for h in up.values():
        h.free()


@pytest.mark.parametrize("seed", SEEDS)
def test_random_plan_matches_oracle(seed, sf001, engine, tables_on_gpu):
    from resql_b200 import EngineError
    d = random_plan(seed)
    tabs = plan_tables(d, sf001)
    try:
        want = serialize_columns(*run_plan(d, tabs))
    except ZeroDivisionError:
        pytest.skip("plan divides by zero")
    try:
        res, _ = engine.execute(Plan(d), {t["name"]: tables_on_gpu[t["name"]] for t in d["tables"]})
    except EngineError as e:
        if e.code == 3:
            # a legal plan the engine refuses is a hole in the drop-in (no CPU fallback exists)
            import os
            os.makedirs("gpurun_out", exist_ok=True)
            with open("gpurun_out/fuzz_unsupported.log", "a") as f:
                f.write(f"seed {seed}: {e}\n")
            pytest.fail(f"random plan {seed} is legal but was rejected: {e}")
        raise
    got = serialize_columns(res.columns, res.sql_types, res.sql_widths)
    assert_same_relation(got, want, d, f"random plan {seed}")


def test_two_pass_forced_on_random_join_plans(sf001, engine, tables_on_gpu):
    """the same random join plans with every probe forced through the Bloom semi-join + dense pass"""
    engine.set_option("split_min_rows", 0)
    engine.set_option("split_frac", 1e18)
    try:
        _two_pass_body(sf001, engine, tables_on_gpu)
    finally:
        engine.set_option("split_min_rows", -1)
        engine.set_option("split_frac", -1)


def _two_pass_body(sf001, engine, tables_on_gpu):
    from resql_b200 import EngineError
    ran = 0
    for seed in SEEDS:
        d = random_plan(seed)
        if not any(p["sink_kind"] == 2 for p in d["pipelines"]):
            continue
        tabs = plan_tables(d, sf001)
        try:
            want = serialize_columns(*run_plan(d, tabs))
            res, _ = engine.execute(Plan(d), {t["name"]: tables_on_gpu[t["name"]] for t in d["tables"]})
        except ZeroDivisionError:
            continue
        except EngineError as e:
            if e.code == 3:
                continue
            raise
        assert_same_relation(serialize_columns(res.columns, res.sql_types, res.sql_widths), want, d, f"random plan {seed} (two-pass)")
        ran += 1
    assert ran >= 20
