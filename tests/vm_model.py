"""Python model of the device accumulator machine (resql_b200/csrc/rq_internal.h DOp) used to
check the HOST-side lowering on machines without a GPU: the program text printed by
rq_debug_lower is executed over whole numpy columns and compared with the plan oracle.
Test infrastructure only."""
import numpy as np

from oracle import plan_oracle as PO
from resql_b200 import native as N

(D_LD, D_ADD, D_SUB, D_RSUB, D_MUL, D_DIV, D_RDIV, D_AND, D_OR, D_LT, D_LE, D_GT, D_GE, D_EQ, D_NE,
 D_EQC, D_EQV, D_NEC, D_NEV, D_LIKE, D_RLIKE, D_SEL) = range(1, 23)
S_NONE, S_COL, S_SLOT, S_IMM, S_STR = range(5)
H_FCMP, H_BIN, H_MULI, H_SEL, H_PROBE, H_FRANGE = range(1, 7)
IMPL_LOWAGG, IMPL_HASHAGG, IMPL_BUILD, IMPL_EMIT, IMPL_REGAGG = 1, 2, 3, 4, 5
STR_BASE = 1 << 44      # pool base rq_debug_lower lowers string constants against


def phys_of(arr):
    if arr.dtype.kind == "S":
        return N.RQ_STR, arr.dtype.itemsize
    return {1: (N.RQ_I8, 1), 4: (N.RQ_I32, 4), 8: (N.RQ_I64, 8)}[arr.dtype.itemsize]


def parse(text):
    prog = {"cols": {}, "strcols": {}, "unit": [], "key": [], "out": [], "imm": {}, "agg": [], "aggmap": [], "aggsrc": []}
    for line in text.strip().split("\n"):
        f = line.split()
        if f[0] == "col":
            prog["cols"][int(f[1])] = int(f[3])
        elif f[0] == "strcol":
            prog["strcols"][int(f[1])] = int(f[3])
        elif f[0] == "unit":
            prog["unit"].append([int(x) for x in f[1:]])
        elif f[0] in ("key", "out"):
            prog[f[0]].append((int(f[1]), int(f[2])))
        elif f[0] == "aggsrc":
            prog[f[0]].append((int(f[1]), int(f[2]), int(f[3])))
        elif f[0] == "imm":
            prog["imm"][int(f[1])] = int(f[2])
        elif f[0] == "agg":
            prog["agg"].append(int(f[2]))
        elif f[0] == "aggmap":
            prog["aggmap"].append(int(f[2]))
    return prog


def _binop(op, a, b, valid):
    if op == D_LD: return a
    if op == D_ADD: return a + b
    if op == D_SUB: return a - b
    if op == D_RSUB: return b - a
    if op == D_MUL: return a * b
    if op == D_DIV: return np.where(valid, PO._div_trunc(a, np.where(valid, b, 1)), 0)
    if op == D_RDIV: return np.where(valid, PO._div_trunc(b, np.where(valid, a, 1)), 0)
    if op == D_AND: return a & b
    if op == D_OR: return a | b
    if op == D_LT: return (a < b).astype(np.int64)
    if op == D_LE: return (a <= b).astype(np.int64)
    if op == D_GT: return (a > b).astype(np.int64)
    if op == D_GE: return (a >= b).astype(np.int64)
    if op == D_EQ: return (a == b).astype(np.int64)
    if op == D_NE: return (a != b).astype(np.int64)
    if op in (D_EQC, D_NEC):
        r = np.array([1 if u.rstrip(b" ") == v.rstrip(b" ") else 0 for u, v in zip(a, b)], dtype=np.int64)
        return r if op == D_EQC else 1 - r
    if op in (D_EQV, D_NEV):
        r = np.array([1 if u == v else 0 for u, v in zip(a, b)], dtype=np.int64)
        return r if op == D_EQV else 1 - r
    if op == D_LIKE: return np.array([PO._like(u, v) for u, v in zip(a, b)], dtype=np.int64)
    if op == D_RLIKE: return np.array([PO._like(v, u) for u, v in zip(a, b)], dtype=np.int64)
    raise NotImplementedError(op)


def run_pipeline_vm(plan, pi, src_cols, pool_strings, agg_impl=IMPL_LOWAGG, with_stats=True):
    """src_cols: list of numpy arrays (physical dtypes; intermediates int64/object).
    Returns the list of output columns of the pipeline (evaluation form)."""
    p = plan.pipelines[pi]
    impl = agg_impl if p["sink_kind"] == 1 else IMPL_EMIT
    types, widths, mins, maxs = [], [], [], []
    for a in src_cols:
        if a.dtype == object:
            types.append(N.RQ_I64); widths.append(8)     # strings by reference inside intermediates
        else:
            t, w = phys_of(a)
            types.append(t); widths.append(w)
        if with_stats and a.dtype.kind in "iu" and len(a):
            mins.append(int(a.min())); maxs.append(int(a.max()))
        else:
            mins.append(1); maxs.append(0)               # no statistics for this column
    prog = parse(N.debug_lower(plan, pi, impl, types, widths, mins, maxs))
    vals = [PO._to_value(a) for a in src_cols]
    n = len(vals[0]) if vals else 0
    valid = np.ones(n, dtype=bool)
    slots = {}

    def operand(kind, idx, imm):
        if kind == S_COL:
            return vals[prog["cols"][idx]]
        if kind == S_STR:
            return vals[prog["strcols"][idx]]
        if kind == S_SLOT:
            return slots[idx]
        if kind == S_IMM:
            if imm >= STR_BASE and (imm - STR_BASE) in pool_strings:     # rq_debug_lower's fake pool base
                return PO._bcast(pool_strings[imm - STR_BASE], n)
            return np.full(n, imm, dtype=np.int64)
        return None

    def vref(kind, idx):
        if kind == S_IMM:
            return operand(S_IMM, 0, prog["imm"][idx])
        return operand(kind, idx, 0)

    old = np.seterr(over="ignore")
    for op, gop, xk, xi, ximm, yk, yi, yimm, zk, zi, zimm, imm, dst, filt, aux, imm2, n32 in prog["unit"]:
        x, y, z = operand(xk, xi, ximm), operand(yk, yi, yimm), operand(zk, zi, zimm)
        if op == H_FCMP:
            valid = valid & (_binop(gop, x, np.full(n, imm, dtype=np.int64), valid) != 0)
            continue
        if op == H_FRANGE:
            xs = x.astype(object)
            valid = valid & np.array([imm <= v <= imm + (imm2 & 0xFFFFFFFFFFFFFFFF) for v in xs], dtype=bool)
            continue
        def mul32(u, v):
            # the 32-bit multiply forms compute (uint32)u * (uint32)v exactly as the device does
            # (IMAD.WIDE.U32). They are only chosen when interval analysis proves both factors lie in
            # [0, 2^32) for every tuple that REACHES THE SINK (selections narrow the column bounds), so a
            # wrong choice shows up as a wrong result, not as an assertion here.
            lo = (u.astype(np.uint64) & np.uint64(0xFFFFFFFF)) * (v.astype(np.uint64) & np.uint64(0xFFFFFFFF))
            return lo.astype(np.int64)
        if op == H_BIN:
            t = mul32(x, y) if (n32 and gop == D_MUL) else _binop(gop, x, y, valid)
        elif op == H_MULI:
            inner = _binop(gop, x, np.full(n, imm, dtype=np.int64), valid)
            t = mul32(inner, y) if n32 else inner * y
        elif op == H_SEL: t = np.where((x & 0xFF) != 0, y, z)
        else: raise NotImplementedError(op)
        if dst >= 0:
            slots[dst] = t
        if filt:
            valid = valid & ((t & 0xFF) != 0)
    np.seterr(**old)
    if p["sink_kind"] != 1:
        return [vref(k, i)[valid] for k, i in prog["out"]]
    # aggregate sink: keys, then aggregates expanded through aggmap
    if not valid.any():
        return [np.zeros(0, dtype=np.int64) for _ in range(len(p["keys"]) + len(p["vals"]))]
    keys = [vref(k, i) for k, i in prog["key"]]
    groups, gid = {}, np.zeros(n, dtype=np.int64)
    for r in np.nonzero(valid)[0]:
        gid[r] = groups.setdefault(tuple(int(x[r]) for x in keys), len(groups))
    ng = max(1, len(groups))
    g = gid[valid]
    aggs = []
    old = np.seterr(over="ignore")
    for u, kind in enumerate(prog["agg"]):
        if kind == 2:
            aggs.append(np.bincount(g, minlength=ng).astype(np.int64))
            continue
        ak, ai, a32 = prog["aggsrc"][u]
        full = vref(ak, ai).astype(np.int64)
        assert not a32 or len(full) == 0 or (int(full.min()) >= 0 and int(full.max()) < 2 ** 32), "narrow sum on wide values"
        v = full[valid]
        if kind == 1:
            a = np.zeros(ng, dtype=np.uint64); np.add.at(a, g, v.view(np.uint64)); a = a.view(np.int64)
        elif kind == 3:
            a = np.full(ng, np.iinfo(np.int64).max); np.minimum.at(a, g, v)
        else:
            a = np.full(ng, np.iinfo(np.int64).min); np.maximum.at(a, g, v)
        aggs.append(a)
    np.seterr(**old)
    ks = sorted(groups.items(), key=lambda kv: kv[1])
    keycols = [np.array([k[0][j] for k in ks], dtype=np.int64) for j in range(len(keys))]
    return keycols + [aggs[u] for u in prog["aggmap"]]
