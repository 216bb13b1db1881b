"""Random flat plans (include/resql_b200.h) over the TPC-H-shaped tables, for differential testing of
the GPU engine against the plan oracle. Shapes follow what the reference's planner can emit:
[build pipelines] -> scan -> selections -> [probes] -> (aggregate | materialize) -> projection, with
typed integer / decimal / date / CHAR(1) / string values. Test infrastructure only."""
import random

from resql_b200 import tpch

OP = {n: i for i, n in enumerate(["", "COL", "CONST", "CONST_STR", "ADD", "SUB", "MUL", "DIV", "AND", "OR", "LT", "LE", "GT",
                                   "GE", "EQ", "NEQ", "EQ_CHAR", "EQ_VARCHAR", "NEQ_CHAR", "NEQ_VARCHAR", "LIKE", "SELECT",
                                   "FILTER", "PROBE", "PAYLOAD"]) if n}
SQL_VARCHAR, SQL_CHAR, SQL_BOOL, SQL_INT, SQL_BIGINT, SQL_DECIMAL, SQL_FLOAT, SQL_DATE = range(8)
AGG_SUM, AGG_COUNT, AGG_MIN, AGG_MAX = 1, 2, 3, 4

# columns the generator uses: (name, class) with class in num (arithmetic), date, c1 (CHAR(1)), str, key
COLS = {
    "lineitem": [("l_orderkey", "key"), ("l_partkey", "key"), ("l_linenumber", "smallint"), ("l_quantity", "num"),
                 ("l_extendedprice", "num"), ("l_discount", "num"), ("l_tax", "num"), ("l_returnflag", "c1"),
                 ("l_linestatus", "c1"), ("l_shipdate", "date"), ("l_commitdate", "date"), ("l_shipmode", "str"),
                 ("l_shipinstruct", "str")],
    "orders": [("o_orderkey", "key"), ("o_custkey", "key"), ("o_totalprice", "num"), ("o_orderdate", "date"),
               ("o_orderpriority", "str"), ("o_shippriority", "smallint"), ("o_orderstatus", "c1")],
    "customer": [("c_custkey", "key"), ("c_nationkey", "smallint"), ("c_acctbal", "num"), ("c_mktsegment", "str")],
}
JOINS = [  # (build table, build key, probe table, probe key)
    ("orders", "o_orderkey", "lineitem", "l_orderkey"),
    ("customer", "c_custkey", "orders", "o_custkey"),
    ("customer", "c_nationkey", "orders", "o_shippriority"),     # duplicates on both sides
    ("customer", "c_nationkey", "lineitem", "l_linenumber"),     # duplicates on both sides
    ("orders", "o_custkey", "customer", "c_custkey"),            # duplicate build keys
]


def _sql_of(table, col):
    kind, arg = next((k, a) for (c, k, a) in tpch.SCHEMAS[table] if c == col)
    return tpch.sql_type_of(kind, arg)


class _Pipe:
    def __init__(self, rng, table, columns):
        self.rng, self.table, self.columns = rng, table, columns
        self.nodes, self.args = [], []
        self.col_node = {}

    def n(self, op, a=0, b=0, c=0, imm=0):
        self.nodes.append([OP[op], a, b, c, imm])
        return len(self.nodes) - 1

    def col(self, name):
        if name not in self.col_node:
            self.col_node[name] = self.n("COL", self.columns.index(name))
        return self.col_node[name]

    def cols_of(self, cls):
        return [c for c, k in COLS[self.table] if k == cls and c in self.columns]

    def num_expr(self, depth=0):
        """arithmetic over decimal / bigint values (wrap-around int64)"""
        rng = self.rng
        nums = self.cols_of("num") + self.cols_of("smallint")
        if depth >= 2 or rng.random() < 0.35:
            if rng.random() < 0.8 and nums:
                return self.col(rng.choice(nums))
            return self.n("CONST", imm=rng.choice([0, 1, 2, 7, 100, 1000, -3, 123456789]))
        op = rng.choice(["ADD", "SUB", "MUL", "MUL", "DIV", "CASE"])
        if op == "CASE":      # one WHEN/THEN arm with ELSE (emitCase, ExpressionsJitFlounder.h:720-754)
            return self.n("SELECT", self.predicate(2), self.num_expr(depth + 1), self.num_expr(depth + 1))
        x = self.num_expr(depth + 1)
        if op == "DIV":
            return self.n("DIV", x, self.n("CONST", imm=rng.choice([1, 2, 3, 10, 100, -7])))
        return self.n(op, x, self.num_expr(depth + 1))

    def predicate(self, depth=0):
        rng = self.rng
        if depth < 2 and rng.random() < 0.3:
            return self.n(rng.choice(["AND", "OR"]), self.predicate(depth + 1), self.predicate(depth + 1))
        kind = rng.choice(["num", "num", "date", "c1", "str", "cols"])
        cmp_ = rng.choice(["LT", "LE", "GT", "GE", "EQ", "NEQ"])
        if kind == "num" and self.cols_of("num"):
            c = rng.choice(self.cols_of("num"))
            k = {"l_quantity": rng.randint(0, 51), "l_discount": rng.randint(0, 10), "l_tax": rng.randint(0, 8)}.get(c, rng.randint(-1000, 6000000))
            x, y = self.col(c), self.n("CONST", imm=k)
            return self.n(cmp_, x, y) if rng.random() < 0.8 else self.n(cmp_, y, x)
        if kind == "date" and self.cols_of("date"):
            c = rng.choice(self.cols_of("date"))
            k = rng.choice([19920101, 19940101, 19950315, 19950617, 19980902, 19981231])
            return self.n(cmp_, self.col(c), self.n("CONST", imm=k))
        if kind == "c1" and self.cols_of("c1"):
            c = rng.choice(self.cols_of("c1"))
            return self.n(rng.choice(["EQ", "NEQ"]), self.col(c), self.n("CONST", imm=ord(rng.choice("ANRFOPX"))))
        if kind == "str" and self.cols_of("str"):
            c = rng.choice(self.cols_of("str"))
            lit = rng.choice(["AIR", "MAIL", "REG AIR", "BUILDING", "BUILDING  ", "1-URGENT", "5-LOW", "NONE", "SHIP", "zzz"])
            if rng.random() < 0.3:
                return self.n("LIKE", self.col(c), self.n("CONST_STR", imm=self.strs(rng.choice(["%AIR%", "_AIL", "%URGENT", "BUILD%", "%", "R_G%"]))))
            return self.n(rng.choice(["EQ_CHAR", "NEQ_CHAR"]), self.col(c), self.n("CONST_STR", imm=self.strs(lit)))
        nums = self.cols_of("num")
        if len(nums) >= 2:
            a, b = rng.sample(nums, 2)
            return self.n(cmp_, self.col(a), self.col(b))
        return self.n("EQ", self.n("CONST", imm=1), self.n("CONST", imm=1))

    def strs(self, lit):
        return self.pool.add(lit)


class _Pool:
    def __init__(self):
        self.s = ""

    def add(self, lit):
        off = len(self.s)
        self.s += lit + "\0"
        return off


def random_plan(seed):
    """-> plan dict; deterministic per seed"""
    rng = random.Random(seed)
    pool = _Pool()
    tables, pipelines = [], []
    join = rng.choice(JOINS) if rng.random() < 0.5 else None
    probe_table = join[2] if join else rng.choice(list(COLS))
    build_idx = None
    if join:
        bt, bk, _, _ = join
        bcols = [c for c, _ in COLS[bt]]
        tables.append({"name": bt, "columns": bcols})
        bp = _Pipe(rng, bt, bcols)
        bp.pool = pool
        for _ in range(rng.randint(0, 2)):
            bp.n("FILTER", bp.predicate())
        pay = rng.sample([c for c in bcols], rng.randint(1, 3))
        if rng.random() < 0.5 and bk not in pay:
            pay.insert(0, bk)
        pipelines.append({"source_kind": 1, "source_id": 0, "sink_kind": 2, "size_hint": rng.choice([0, 10, 100000]),
                          "nodes": bp.nodes, "args": [], "keys": [[bp.col(bk), 0, *_sql_of(bt, bk)]],
                          "vals": [[bp.col(c), 0, *_sql_of(bt, c)] for c in pay]})
        # nodes may have grown through bp.col(): keys/vals were evaluated after the filters, fine
        pipelines[-1]["nodes"] = bp.nodes
        build_idx, build_pay, build_table = 0, pay, bt
    pcols = [c for c, _ in COLS[probe_table]]
    tables.append({"name": probe_table, "columns": pcols})
    p = _Pipe(rng, probe_table, pcols)
    p.pool = pool
    if join and JOINS.index(join) in (2, 3):     # many-to-many joins: keep the result (and the oracle's work) small
        lim = {"lineitem": ("l_orderkey", 600), "orders": ("o_orderkey", 3000)}[probe_table]
        p.n("FILTER", p.n("LT", p.col(lim[0]), p.n("CONST", imm=lim[1])))
    for _ in range(rng.randint(0, 3)):
        p.n("FILTER", p.predicate())
    extra_vals = []          # (node, sql_type, width) usable as group keys / outputs
    if join:
        single = 1 if (join[1] == "o_orderkey") else 0
        p.args.append(p.col(join[3]))
        pr = p.n("PROBE", build_idx, len(p.args) - 1, 1, single)
        for j, c in enumerate(build_pay):
            if rng.random() < 0.7:
                extra_vals.append((p.n("PAYLOAD", pr, j), *_sql_of(build_table, c), dict(COLS[build_table])[c]))
        if rng.random() < 0.4:
            p.n("FILTER", p.predicate())
    values = [(p.col(c), *_sql_of(probe_table, c), k) for c, k in COLS[probe_table] if rng.random() < 0.5] + extra_vals
    if not values:
        c, k = COLS[probe_table][0]
        values = [(p.col(c), *_sql_of(probe_table, c), k)]
    src_id = len(tables) - 1
    if rng.random() < 0.65:       # aggregation
        key_pool = [v for v in values if v[3] in ("c1", "smallint", "str", "date") or (v[3] == "key" and rng.random() < 0.5)]
        keys = rng.sample(key_pool, min(len(key_pool), rng.randint(0, 3)))
        if rng.random() < 0.2:
            keys.append((p.n("MUL", p.num_expr(2), p.n("CONST", imm=rng.choice([0, 1, 2]))), SQL_BIGINT, 0, "num"))
        vals = []
        for _ in range(rng.randint(1, 8)):
            kind = rng.choice([AGG_SUM, AGG_SUM, AGG_COUNT, AGG_MIN, AGG_MAX])
            if kind == AGG_COUNT:
                vals.append([p.n("CONST", imm=1), kind, SQL_BIGINT, 0])
            elif kind == AGG_SUM or rng.random() < 0.6:
                vals.append([p.num_expr(), kind, SQL_BIGINT, 0])
            else:
                dates = p.cols_of("date")
                vals.append([p.col(rng.choice(dates)), kind, SQL_DATE, 0] if dates else [p.num_expr(), kind, SQL_BIGINT, 0])
        pipelines.append({"source_kind": 1, "source_id": src_id, "sink_kind": 1, "size_hint": rng.choice([0, 4, 100000]),
                          "nodes": p.nodes, "args": p.args, "keys": [[k[0], 0, k[1], k[2]] for k in keys], "vals": vals})
        ncols = len(keys) + len(vals)
        types = [(k[1], k[2]) for k in keys] + [(v[2], v[3]) for v in vals]
        # projection pipeline over the aggregate (plain pass-through, as ProjectionOp + MaterializeOp)
        pipelines.append({"source_kind": 2, "source_id": len(pipelines) - 1, "sink_kind": 3, "size_hint": 0,
                          "nodes": [[OP["COL"], c, 0, 0, 0] for c in range(ncols)], "args": [], "keys": [],
                          "vals": [[c, 0, types[c][0], types[c][1]] for c in range(ncols)]})
    else:                          # selection / join result rows
        outs = values[: rng.randint(1, min(6, len(values)))]
        if rng.random() < 0.4:
            outs.append((p.num_expr(), SQL_BIGINT, 0, "num"))
        pipelines.append({"source_kind": 1, "source_id": src_id, "sink_kind": 3, "size_hint": 0, "nodes": p.nodes, "args": p.args,
                          "keys": [], "vals": [[o[0], 0, o[1], o[2]] for o in outs]})
        ncols = len(outs)
        types = [(o[1], o[2]) for o in outs]
    order = [[c, rng.choice([0, 1])] for c in rng.sample(range(ncols), ncols)] if rng.random() < 0.8 else []
    return {"tables": tables, "pipelines": pipelines, "order": order, "limit": -1, "strpool": pool.s,
            "result_names": [f"c{i}" for i in range(ncols)], "result_types": [str(t) for t in types]}
