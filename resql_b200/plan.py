"""Flat plan (include/resql_b200.h rq_plan) as a JSON-serialisable Python object.

The JSON form is what the C++ host shim (resql_b200/host/gpu_executor.h) dumps with
RESQL_B200_DUMP_PLAN=1 after lowering the reference's typed operator tree; the committed
fixtures under tests/golden/plans/ were produced that way from tpch/queries/*.sql, so the GPU box
(which has no reference checkout) executes exactly the plans the reference's planner builds."""
import ctypes as C
import json

from . import native as N

OPS = ["", "COL", "CONST", "CONST_STR", "ADD", "SUB", "MUL", "DIV", "AND", "OR", "LT", "LE", "GT",
       "GE", "EQ", "NEQ", "EQ_CHAR", "EQ_VARCHAR", "NEQ_CHAR", "NEQ_VARCHAR", "LIKE", "SELECT",
       "FILTER", "PROBE", "PAYLOAD"]
OP = {n: i for i, n in enumerate(OPS) if n}
AGG = {"SUM": 1, "COUNT": 2, "MIN": 3, "MAX": 4}
SRC_TABLE, SRC_PIPELINE, SRC_CROSS, SRC_ONE_ROW = 1, 2, 3, 4
SINK_AGG, SINK_BUILD, SINK_MATERIALIZE = 1, 2, 3


class Plan:
    """
    tables:    [{"name": str, "columns": [str, ...]}]     COL nodes index `columns`
    pipelines: [{"source_kind", "source_id", "nodes": [[op, a, b, c, imm], ...], "args": [...],
                 "sink_kind", "keys": [[node, kind, sql_type, width], ...], "vals": [...],
                 "size_hint"}]
    order:     [[column, ascending], ...];  limit: int (-1 none);  strpool: str (latin1)
    """

    def __init__(self, d):
        self.d = d
        self.tables = d["tables"]
        self.pipelines = d["pipelines"]
        self.order = d.get("order", [])
        self.limit = d.get("limit", -1)
        self.strpool = d.get("strpool", "")
        self.result_names = d.get("result_names")

    @staticmethod
    def load(path):
        with open(path) as f:
            return Plan(json.load(f))

    def dumps(self):
        return json.dumps(self.d, indent=1)

    def to_c(self, tables, flags=0):
        keep = []
        handles = (C.c_void_p * len(self.tables))()
        colmap = []     # per plan table: plan column index -> column index of the uploaded table
        for i, t in enumerate(self.tables):
            tab = tables[t["name"]]
            missing = [c for c in t["columns"] if c not in tab.names]
            if missing:
                raise ValueError(f"table {t['name']}: uploaded columns {tab.names} lack {missing}")
            colmap.append([list(tab.names).index(c) for c in t["columns"]])
            handles[i] = tab.handle
        pls = (N.rq_pipeline * len(self.pipelines))()
        for i, p in enumerate(self.pipelines):
            nodes = (N.rq_node * max(1, len(p["nodes"])))()
            cm = colmap[p["source_id"]] if p["source_kind"] == SRC_TABLE else None
            for j, nd in enumerate(p["nodes"]):
                op = nd[0] if isinstance(nd[0], int) else OP[nd[0]]
                a = cm[nd[1]] if (cm is not None and op == OP["COL"]) else nd[1]
                nodes[j].op, nodes[j].a, nodes[j].b, nodes[j].c, nodes[j].imm = op, a, nd[2], nd[3], nd[4]
            args = (C.c_int32 * max(1, len(p.get("args", []))))(*p.get("args", []))
            keys = (N.rq_value * max(1, len(p["keys"])))()
            for j, k in enumerate(p["keys"]):
                keys[j].node, keys[j].kind, keys[j].sql_type, keys[j].width = k
            vals = (N.rq_value * max(1, len(p["vals"])))()
            for j, k in enumerate(p["vals"]):
                vals[j].node, vals[j].kind, vals[j].sql_type, vals[j].width = k
            keep += [nodes, args, keys, vals]
            pl = pls[i]
            pl.source_kind, pl.source_id = p["source_kind"], p["source_id"]
            pl.n_nodes, pl.nodes = len(p["nodes"]), nodes
            pl.n_args, pl.args = len(p.get("args", [])), args
            pl.sink_kind = p["sink_kind"]
            pl.n_keys, pl.keys = len(p["keys"]), keys
            pl.n_vals, pl.vals = len(p["vals"]), vals
            pl.size_hint = p.get("size_hint", 0)
            pl.source_id2 = p.get("source_id2", 0)
        order = (N.rq_order_key * max(1, len(self.order)))()
        for j, o in enumerate(self.order):
            order[j].column, order[j].ascending = o
        pool = self.strpool.encode("latin1")
        cp = N.rq_plan()
        cp.n_tables, cp.tables = len(self.tables), handles
        cp.n_pipelines, cp.pipelines = len(self.pipelines), pls
        cp.n_order, cp.order = len(self.order), order
        cp.limit = self.limit
        cp.strpool, cp.strpool_bytes = pool, len(pool)
        cp.flags = flags
        keep += [handles, pls, order, pool]
        return cp, keep
