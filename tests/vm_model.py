"""Python model of the device accumulator machine (resql_b200/csrc/rq_internal.h DOp) used to
check the HOST-side lowering on machines without a GPU: the program text printed by
rq_debug_lower is executed over whole numpy columns and compared with the plan oracle.
Test infrastructure only."""
import numpy as np

from oracle import plan_oracle as PO
from resql_b200 import native as N

(D_LD, D_ADD, D_SUB, D_RSUB, D_MUL, D_DIV, D_RDIV, D_AND, D_OR, D_LT, D_LE, D_GT, D_GE, D_EQ, D_NE,
 D_EQC, D_EQV, D_NEC, D_NEV, D_LIKE, D_RLIKE, D_SEL, D_FILTER, D_GROUP, D_AGG_SUM, D_AGG_COUNT,
 D_AGG_MIN, D_AGG_MAX, D_PROBE, D_HAGG, D_BUILD, D_EMIT, D_NOP,
 D_FLT, D_FLE, D_FGT, D_FGE, D_FEQ, D_FNE) = range(1, 40)
S_NONE, S_COL, S_SLOT, S_IMM, S_STR = range(5)
IMPL_LOWAGG, IMPL_HASHAGG, IMPL_BUILD, IMPL_EMIT, IMPL_REGAGG = 1, 2, 3, 4, 5


def phys_of(arr):
    if arr.dtype.kind == "S":
        return N.RQ_STR, arr.dtype.itemsize
    return {1: (N.RQ_I8, 1), 4: (N.RQ_I32, 4), 8: (N.RQ_I64, 8)}[arr.dtype.itemsize]


def parse(text):
    prog = {"cols": {}, "strcols": {}, "insn": [], "key": [], "out": [], "imm": {}, "agg": [], "aggmap": []}
    for line in text.strip().split("\n"):
        f = line.split()
        if f[0] == "col":
            prog["cols"][int(f[1])] = int(f[3])
        elif f[0] == "strcol":
            prog["strcols"][int(f[1])] = int(f[3])
        elif f[0] == "insn":
            prog["insn"].append([int(x) for x in f[1:]])
        elif f[0] in ("key", "out"):
            prog[f[0]].append((int(f[1]), int(f[2])))
        elif f[0] == "imm":
            prog["imm"][int(f[1])] = int(f[2])
        elif f[0] == "agg":
            prog["agg"].append(int(f[2]))
        elif f[0] == "aggmap":
            prog["aggmap"].append(int(f[2]))
    return prog


def run_pipeline_vm(plan, pi, src_cols, pool_strings, agg_impl=IMPL_LOWAGG):
    """src_cols: list of numpy arrays (physical dtypes; intermediates int64/object).
    Returns the list of output columns of the pipeline (evaluation form)."""
    p = plan.pipelines[pi]
    impl = agg_impl if p["sink_kind"] == 1 else IMPL_EMIT
    types, widths = [], []
    for a in src_cols:
        if a.dtype == object:
            types.append(N.RQ_I64); widths.append(8)     # strings by reference inside intermediates
        else:
            t, w = phys_of(a)
            types.append(t); widths.append(w)
    prog = parse(N.debug_lower(plan, pi, impl, types, widths))
    vals = [PO._to_value(a) for a in src_cols]
    n = len(vals[0]) if vals else 0
    acc = np.zeros(n, dtype=np.int64)
    valid = np.ones(n, dtype=bool)
    slots = {}
    gid = np.zeros(n, dtype=np.int64)
    groups = None
    aggs = {}

    def operand(src, idx, imm):
        if src == S_COL:
            return vals[prog["cols"][idx]]       # (fused compares carry the constant in imm)
        if src == S_STR:
            return vals[prog["strcols"][idx]]
        if src == S_SLOT:
            return slots[idx]
        if src == S_IMM:
            if imm in pool_strings:
                return PO._bcast(pool_strings[imm], n)
            return np.full(n, imm, dtype=np.int64)
        return None

    def vref(kind, idx):
        if kind == S_IMM:
            return operand(S_IMM, 0, prog["imm"][idx])
        return operand(kind, idx, 0)

    old = np.seterr(over="ignore")
    for op, src, flags, dst, idx, aux, imm in prog["insn"]:
        b = operand(src, idx, imm)
        if op == D_LD: acc = b
        elif op == D_ADD: acc = acc + b
        elif op == D_SUB: acc = acc - b
        elif op == D_RSUB: acc = b - acc
        elif op == D_MUL: acc = acc * b
        elif op == D_DIV: acc = np.where(valid, PO._div_trunc(acc, np.where(valid, b, 1)), 0)
        elif op == D_RDIV: acc = np.where(valid, PO._div_trunc(b, np.where(valid, acc, 1)), 0)
        elif op == D_AND: acc = acc & b
        elif op == D_OR: acc = acc | b
        elif op == D_LT: acc = (acc < b).astype(np.int64)
        elif op == D_LE: acc = (acc <= b).astype(np.int64)
        elif op == D_GT: acc = (acc > b).astype(np.int64)
        elif op == D_GE: acc = (acc >= b).astype(np.int64)
        elif op == D_EQ: acc = (acc == b).astype(np.int64)
        elif op == D_NE: acc = (acc != b).astype(np.int64)
        elif op in (D_EQC, D_NEC):
            r = np.array([1 if u.rstrip(b" ") == v.rstrip(b" ") else 0 for u, v in zip(acc, b)], dtype=np.int64)
            acc = r if op == D_EQC else 1 - r
        elif op in (D_EQV, D_NEV):
            r = np.array([1 if u == v else 0 for u, v in zip(acc, b)], dtype=np.int64)
            acc = r if op == D_EQV else 1 - r
        elif op == D_LIKE: acc = np.array([PO._like(u, v) for u, v in zip(acc, b)], dtype=np.int64)
        elif op == D_RLIKE: acc = np.array([PO._like(v, u) for u, v in zip(acc, b)], dtype=np.int64)
        elif op == D_SEL: acc = np.where((acc & 0xFF) != 0, b, (np.full(n, prog['imm'][aux], dtype=np.int64) if flags & 2 else slots[aux]))
        elif op == D_FILTER: valid = valid & (((b if src != S_NONE else acc) & 0xFF) != 0)
        elif op == D_FLT: valid = valid & (b < imm)
        elif op == D_FLE: valid = valid & (b <= imm)
        elif op == D_FGT: valid = valid & (b > imm)
        elif op == D_FGE: valid = valid & (b >= imm)
        elif op == D_FEQ: valid = valid & (b == imm)
        elif op == D_FNE: valid = valid & (b != imm)
        elif op == D_GROUP:
            keys = [vref(k, i) for k, i in prog["key"]]
            groups = {}
            for r in np.nonzero(valid)[0]:
                k = tuple(int(x[r]) for x in keys)
                gid[r] = groups.setdefault(k, len(groups))
        elif op in (D_AGG_SUM, D_AGG_COUNT, D_AGG_MIN, D_AGG_MAX):
            v = (b if src != S_NONE else acc)
            ng = max(1, len(groups)) if groups is not None else 1
            g = gid[valid]
            if op == D_AGG_SUM:
                a = np.zeros(ng, dtype=np.uint64); np.add.at(a, g, v[valid].astype(np.int64).view(np.uint64)); a = a.view(np.int64)
            elif op == D_AGG_COUNT:
                a = np.bincount(g, minlength=ng).astype(np.int64)
            elif op == D_AGG_MIN:
                a = np.full(ng, np.iinfo(np.int64).max); np.minimum.at(a, g, v[valid].astype(np.int64))
            else:
                a = np.full(ng, np.iinfo(np.int64).min); np.maximum.at(a, g, v[valid].astype(np.int64))
            aggs[aux] = a
        elif op == D_EMIT:
            np.seterr(**old)
            return [vref(k, i)[valid] for k, i in prog["out"]]
        else:
            raise NotImplementedError(op)
        if flags & 1:
            slots[dst] = acc
    np.seterr(**old)
    # low-card aggregate output: keys, then aggregates expanded through aggmap
    if not valid.any():
        return [np.zeros(0, dtype=np.int64) for _ in range(len(p["keys"]) + len(p["vals"]))]
    if groups is None or len(prog["key"]) == 0:
        keycols = []
    else:
        ks = sorted(groups.items(), key=lambda kv: kv[1])
        keycols = [np.array([k[0][j] for k in ks], dtype=np.int64) for j in range(len(prog["key"]))]
    return keycols + [aggs[u] for u in prog["aggmap"]]
