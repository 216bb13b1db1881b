"""CPU ORACLE - TEST INFRASTRUCTURE ONLY. Never imported by the product path (resql_b200/),
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.

numpy restatement of what the reference (Henning1/resql) computes for a flat plan
(include/resql_b200.h). Every rule cites the reference file:line it follows. Parity of this
restatement is PINNED: tests/test_oracle_pinned.py checks it against outputs of the reference
binary itself (oracle/_ref/resql-oracle, built by oracle/ref_build/build_ref.sh) - committed as
tests/golden/*.out together with the generating script tests/golden/make_golden.py - and,
when the binary is present, live on fresh seeded data.

Semantics:
  * all arithmetic is wrap-around int64 (ExpressionsJitFlounder.h:298-402, no overflow checks)
  * DIV truncates toward zero (cqo; idiv, :408-420)
  * compares are signed and yield 0/1 (:474-632); AND/OR bitwise (:446-468); NEQ = 1 - EQ (:1036)
  * CHAR equality ignores trailing blanks (qlib/scalar.h:27-46), VARCHAR is exact (:16-24)
  * LIKE: '%' any run, '_' one char (qlib/scalar.h:57-120)
  * selection drops tuples whose condition byte is 0 (selection.h:62-66)
  * aggregation: SUM add, COUNT inc, MIN/MAX compare-move; the first tuple initialises the
    accumulators (aggregation.h:95-152, :272-279); no input => no output row
  * hash join: inner equi-join, all matches or the first one if single (hashjoin.h:118-214)
  * ORDER BY: typed compare, strings by strcmp (types.h:264-353); LIMIT = first k (dbdata.h:407)
"""
import re
import numpy as np

OPS = ["", "COL", "CONST", "CONST_STR", "ADD", "SUB", "MUL", "DIV", "AND", "OR", "LT", "LE", "GT",
       "GE", "EQ", "NEQ", "EQ_CHAR", "EQ_VARCHAR", "NEQ_CHAR", "NEQ_VARCHAR", "LIKE", "SELECT",
       "FILTER", "PROBE", "PAYLOAD"]
SQL_VARCHAR, SQL_CHAR, SQL_BOOL, SQL_INT, SQL_BIGINT, SQL_DECIMAL, SQL_FLOAT, SQL_DATE = range(8)
AGG_SUM, AGG_COUNT, AGG_MIN, AGG_MAX = 1, 2, 3, 4


def _is_str(sql_type, width):
    return sql_type == SQL_VARCHAR or (sql_type == SQL_CHAR and width > 1)


def _cstr(x):
    return bytes(x).split(b"\0")[0]


def _to_value(col):
    """physical column -> evaluation form: int64 array, or object array of bytes for strings"""
    if col.dtype.kind == "S":
        return np.array([_cstr(v) for v in col], dtype=object)
    if col.dtype == object:
        return col
    return col.astype(np.int64)          # sign-extends int32, zero-extends uint8


def _div_trunc(a, b):
    if np.any(b == 0):
        raise ZeroDivisionError("division by zero (the reference raises SIGFPE)")
    q = np.abs(a) // np.abs(b)
    return np.where((a < 0) != (b < 0), -q, q).astype(np.int64)


def _like(s, p):
    """stringLikeCheck(string, like), restated statement by statement from src/qlib/scalar.h:57-120
    (index form of its pointer walk). The reference matches the literal prefix and the literal
    suffix independently - they may overlap ('aba' like 'ab%ba' is TRUE) - then scans the infix
    pieces left to right; '_' matches any one character (cmpLike :50-54). Quirks that follow from
    the code are kept: '%%' matches nothing, '%_' matches everything. Pinned against the reference
    engine by tests/golden/like_matrix.json."""
    s_end, p_end = len(s), len(p)
    s = s + b"\0"
    p = p + b"\0"

    def cmp(c, l):
        return c == l or l == 0x5F          # '_'

    pct = 0x25                              # '%'
    l_in_start, l_in_end, s_in_start, s_in_end = 0, p_end, 0, s_end
    l_pos = s_pos = 0
    if p[0] != pct:                         # prefix :72-81
        while l_pos < p_end and s_pos < s_end and p[l_pos] != pct:
            if not cmp(s[s_pos], p[l_pos]):
                return 0
            l_pos += 1
            s_pos += 1
        l_in_start, s_in_start = l_pos, s_pos
    if l_in_start == p_end:                 # no-'%' likes :82-85
        return 1 if s_in_start == s_end else 0
    if p[p_end - 1] != pct:                 # suffix :88-97
        s_pos, l_pos = s_end - 1, p_end - 1
        while l_pos >= 0 and s_pos >= 0 and p[l_pos] != pct:
            if not cmp(s[s_pos], p[l_pos]):
                return 0
            l_pos -= 1
            s_pos -= 1
        l_in_end, s_in_end = l_pos, s_pos + 1
    if l_in_start < l_in_end:               # infixes :100-117
        l_pos = l_in_start + 1
        s_pos = s_in_start
        while s_pos < s_in_end and l_pos < l_in_end:
            l_trace, s_trace = l_pos, s_pos
            while cmp(s[s_trace], p[l_trace]) and s_trace < s_in_end:
                l_trace += 1
                if p[l_trace] == pct:
                    l_trace += 1
                    l_pos = l_trace
                    s_pos = s_trace
                    break
                s_trace += 1
            s_pos += 1
    return 1 if l_pos >= l_in_end else 0


def _bcast(v, n):
    if isinstance(v, np.ndarray):
        return v
    if isinstance(v, bytes):
        a = np.empty(n, dtype=object)
        a[:] = [v] * n
        return a
    return np.full(n, v, dtype=np.int64)


class _Built:
    def __init__(self, keys, payload, sql):
        self.keys, self.payload, self.sql = keys, payload, sql


def run_plan(plan, tables):
    """plan: dict (the JSON form of resql_b200.plan.Plan.d); tables: name -> {column -> numpy}.
    Returns (columns, sql_types, sql_widths) with columns in the reference's physical dtypes."""
    pool = plan.get("strpool", "").encode("latin1")
    outs = []
    old = np.seterr(over="ignore")
    try:
        for p in plan["pipelines"]:
            outs.append(_run_pipeline(plan, p, tables, outs, pool))
    finally:
        np.seterr(**old)
    cols, st, sw = outs[-1][:3]
    return finish(plan, cols, st, sw)


def finish(plan, cols, st, sw):
    """ORDER BY + LIMIT + conversion to the reference's physical result types"""
    n = len(cols[0]) if cols else 0
    order = plan.get("order", [])
    idx = np.arange(n)
    if order and n > 1:
        keys = []
        for c, asc in reversed(order):
            v = cols[c]
            if v.dtype == object:
                # strcmp order == bytes order; descending via rank inversion
                ranks = np.unique(np.array([bytes(x) for x in v], dtype=object), return_inverse=True)[1]
                keys.append(ranks if asc else -ranks)
            else:
                keys.append(v if asc else -v.astype(object))
        idx = np.lexsort([np.asarray(k, dtype=object) if k.dtype == object else k for k in keys])
    lim = plan.get("limit", -1)
    if lim is not None and lim >= 0:
        idx = idx[:lim]
    res = []
    for c, v in enumerate(cols):
        v = v[idx]
        if _is_str(st[c], sw[c]):
            res.append(np.array([bytes(x) for x in v], dtype=f"S{sw[c] + 1}"))
        elif st[c] in (SQL_INT, SQL_DATE):
            res.append(v.astype(np.int64).astype(np.int32))
        elif st[c] in (SQL_BOOL, SQL_CHAR):
            res.append((v.astype(np.int64) & 0xFF).astype(np.uint8))
        else:
            res.append(v.astype(np.int64))
    return res, st, sw


def _run_pipeline(plan, p, tables, outs, pool):
    if p["source_kind"] == 1:
        t = plan["tables"][p["source_id"]]
        src = [_to_value(np.asarray(tables[t["name"]][c])) for c in t["columns"]]
    elif p["source_kind"] == 3:
        # NestedLoopsJoinOp (nestedloopsjoin.h:5-94): both children materialized, every pair produced
        left, right = list(outs[p["source_id"]][0]), list(outs[p["source_id2"]][0])
        nl = len(left[0]) if left else 0
        nr = len(right[0]) if right else 0
        src = [np.repeat(c, nr) for c in left] + [np.tile(c, nl) for c in right]
    else:
        src = list(outs[p["source_id"]][0])
    n = len(src[0]) if src else 0
    sel = np.arange(n)              # surviving tuples, as indices into the source
    vals = {}                       # node -> array aligned with `sel`
    probes = {}

    def get(i):
        return _bcast(vals[i], len(sel))

    for i, nd in enumerate(p["nodes"]):
        op = nd[0] if isinstance(nd[0], str) else OPS[nd[0]]
        a, b, c, imm = nd[1], nd[2], nd[3], nd[4]
        if op == "COL":
            vals[i] = src[a][sel]
        elif op == "CONST":
            vals[i] = int(np.int64(np.uint64(imm & 0xFFFFFFFFFFFFFFFF)))
        elif op == "CONST_STR":
            vals[i] = pool[imm:].split(b"\0")[0]
        elif op in ("ADD", "SUB", "MUL"):
            x, y = get(a).astype(np.int64), get(b).astype(np.int64)
            vals[i] = {"ADD": x + y, "SUB": x - y, "MUL": x * y}[op]
        elif op == "DIV":
            vals[i] = _div_trunc(get(a).astype(np.int64), get(b).astype(np.int64))
        elif op == "AND":
            vals[i] = get(a) & get(b)
        elif op == "OR":
            vals[i] = get(a) | get(b)
        elif op in ("LT", "LE", "GT", "GE", "EQ", "NEQ"):
            x, y = get(a), get(b)
            r = {"LT": x < y, "LE": x <= y, "GT": x > y, "GE": x >= y, "EQ": x == y, "NEQ": x != y}[op]
            vals[i] = r.astype(np.int64)
        elif op in ("EQ_CHAR", "NEQ_CHAR", "EQ_VARCHAR", "NEQ_VARCHAR", "LIKE"):
            x, y = get(a), get(b)
            if op.endswith("_CHAR"):
                r = np.array([1 if u.rstrip(b" ") == v.rstrip(b" ") else 0 for u, v in zip(x, y)], dtype=np.int64)
            elif op.endswith("VARCHAR"):
                r = np.array([1 if u == v else 0 for u, v in zip(x, y)], dtype=np.int64)
            else:
                r = np.array([_like(u, v) for u, v in zip(x, y)], dtype=np.int64)
            vals[i] = (1 - r) if op.startswith("NEQ") else r
            if len(sel) == 0:
                vals[i] = np.zeros(0, dtype=np.int64)
        elif op == "SELECT":
            cond = (get(a) & 0xFF) != 0
            x, y = get(b), get(c)
            vals[i] = np.where(cond, x, y)
        elif op == "FILTER":
            keep = (get(a) & 0xFF) != 0
            sel = sel[keep]
            for k in list(vals):
                if isinstance(vals[k], np.ndarray):
                    vals[k] = vals[k][keep]
        elif op == "PROBE":
            built = outs[a][3]
            keyvals = [get(k) for k in p["args"][b:b + c]]
            single = imm & 1
            pos, match = _probe(built, keyvals, single)
            sel = sel[pos]
            for k in list(vals):
                if isinstance(vals[k], np.ndarray):
                    vals[k] = vals[k][pos]
            probes[i] = (built, match)
        elif op == "PAYLOAD":
            built, match = probes[a]
            vals[i] = built.payload[b][match]
        else:
            raise NotImplementedError(op)

    m = len(sel)
    keys = [(get(k[0]), k[2], k[3]) for k in p["keys"]]
    if p["sink_kind"] == 3:                                   # MATERIALIZE
        cols = [get(v[0]) for v in p["vals"]]
        return cols, [v[2] for v in p["vals"]], [v[3] for v in p["vals"]], None
    if p["sink_kind"] == 2:                                   # BUILD
        built = _Built([k[0] for k in keys], [get(v[0]) for v in p["vals"]],
                       [(k[1], k[2]) for k in keys])
        return [], [], [], built
    # AGG
    st = [k[2] for k in p["keys"]] + [v[2] for v in p["vals"]]
    sw = [k[3] for k in p["keys"]] + [v[3] for v in p["vals"]]
    if m == 0:
        return [np.zeros(0, dtype=object if _is_str(t, w) else np.int64) for t, w in zip(st, sw)], st, sw, None
    if keys:
        norm = []
        for kv, t, w in keys:
            if t == SQL_CHAR and w > 1:
                norm.append(np.array([x.rstrip(b" ") for x in kv], dtype=object))   # compareChar equality
            else:
                norm.append(kv)
        if all(x.dtype != object for x in norm):
            mat = np.stack([x.astype(np.int64) for x in norm], axis=1)
            _, rep, gid = np.unique(mat, axis=0, return_index=True, return_inverse=True)
            gid = gid.reshape(-1)
            ng = len(rep)
        else:
            tup = list(zip(*[list(x) for x in norm]))
            first = {}
            gid = np.empty(m, dtype=np.int64)
            for r, k in enumerate(tup):
                g = first.get(k)
                if g is None:
                    g = len(first)
                    first[k] = g
                gid[r] = g
            ng = len(first)
            rep = np.zeros(ng, dtype=np.int64)
            for r in range(m - 1, -1, -1):
                rep[gid[r]] = r
    else:
        gid = np.zeros(m, dtype=np.int64)
        ng = 1
        rep = np.zeros(1, dtype=np.int64)
    cols = [kv[rep] for kv, _, _ in keys]
    for v in p["vals"]:
        kind = v[1]
        if kind == AGG_COUNT:
            cols.append(np.bincount(gid, minlength=ng).astype(np.int64))
            continue
        x = get(v[0]).astype(np.int64)
        if kind == AGG_SUM:
            acc = np.zeros(ng, dtype=np.uint64)
            np.add.at(acc, gid, x.view(np.uint64))
            cols.append(acc.view(np.int64))
        elif kind == AGG_MIN:
            acc = np.full(ng, np.iinfo(np.int64).max, dtype=np.int64)
            np.minimum.at(acc, gid, x)
            cols.append(acc)
        elif kind == AGG_MAX:
            acc = np.full(ng, np.iinfo(np.int64).min, dtype=np.int64)
            np.maximum.at(acc, gid, x)
            cols.append(acc)
        else:
            raise NotImplementedError(f"aggregate kind {kind}")
    return cols, st, sw, None


def _probe(built, keyvals, single):
    """-> (pos, match): probe-side positions (repeated per match) and build row of each match"""
    def norm(v, sql):
        t, w = sql
        if v.dtype == object:
            if t == SQL_CHAR and w > 1:
                return [x.rstrip(b" ") for x in v]
            return list(v)
        return list(v.astype(np.int64))
    table = {}
    bk = [norm(k, s) for k, s in zip(built.keys, built.sql)]
    for r, k in enumerate(zip(*bk)):
        table.setdefault(k, []).append(r)
    pk = [norm(k, s) for k, s in zip(keyvals, built.sql)]
    pos, match = [], []
    for r, k in enumerate(zip(*pk)):
        hits = table.get(k)
        if not hits:
            continue
        if single:
            hits = hits[:1]
        for h in hits:
            pos.append(r)
            match.append(h)
    return np.array(pos, dtype=np.int64), np.array(match, dtype=np.int64)
