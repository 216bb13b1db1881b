"""Shared helpers of the parity tests."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from resql_b200.serialize import serialize_value  # noqa: E402
from golden.queries import QUERIES  # noqa: E402


def plan_names():
    return sorted(QUERIES.keys())


def load_plan_dict(name):
    with open(os.path.join(GOLDEN, "plans", name + ".json")) as f:
        return json.load(f)


def load_golden(name):
    """-> (schema line, list of row lines) as written by the reference engine"""
    with open(os.path.join(GOLDEN, "sf001", name + ".out")) as f:
        lines = f.read().split("\n")
    assert lines[0].startswith("#schema")
    return lines[0], [l for l in lines[1:] if l != ""]


def plan_tables(plan, data):
    """restrict generated tables to the columns the plan scans, in plan order"""
    return {t["name"]: {c: data[t["name"]][c] for c in t["columns"]} for t in plan["tables"]}


def serialize_columns(cols, sql_types, sql_widths):
    n = len(cols[0]) if cols else 0
    return ["".join(serialize_value(cols[c][i], sql_types[c], sql_widths[c]) + "|" for c in range(len(cols)))
            for i in range(n)]


def assert_same_relation(got_lines, want_lines, plan, what=""):
    """Identical output as the reference defines it (test/test_common.h:125-190): same multiset
    of serialized tuples; with ORDER BY additionally the same sequence of order-key values (the
    reference's quicksort is unstable, so the order inside ties is undefined, qlib/sort.h:21)."""
    assert len(got_lines) == len(want_lines), f"{what}: {len(got_lines)} rows, reference has {len(want_lines)}"
    assert sorted(got_lines) == sorted(want_lines), f"{what}: tuple multiset differs\n got {sorted(got_lines)[:5]}\nwant {sorted(want_lines)[:5]}"
    order = plan.get("order", [])
    if order:
        def keys(lines):
            return [[l.split("|")[c] for c, _ in order] for l in lines]
        assert keys(got_lines) == keys(want_lines), f"{what}: ORDER BY key sequence differs"
