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


# ---------------------------------------------------------------------------------------------
# Device-level model: executes the ENCODED program (the `uinsn` lines: UInsn of rq_internal.h, what the
# warps of rq_scan_kernel interpret) and resolves sink values through the device value references
# (`vkey` / `vout` / `vagg` = VRef). run_pipeline_vm above executes the host-level units; this one
# covers encode_program / to_vref: opcode selection, operand forms, immediates, offsets.
# ---------------------------------------------------------------------------------------------
K_NONE, K_M64, K_M32, K_M8, K_IMM, K_STR, K_IMM2 = range(7)
UF_XSLOT, UF_YSLOT, UF_ZSLOT, UF_FILTER, UF_STORE = 1, 2, 4, 8, 16
_BINOPS = [D_ADD, D_SUB, D_RSUB, D_MUL, D_AND, D_OR, D_LT, D_LE, D_GT, D_GE, D_EQ, D_NE]      # RQ_BINOPS order
U_MULADDI, U_MULSUBI, U_MULRSUBI, U_MULADDI32, U_MULSUBI32, U_MULRSUBI32, U_MUL32_MM, U_GEN = range(25, 33)
U_F_M64, U_F_M32, U_F_M8 = 33, 39, 45            # + index of LT LE GT GE EQ NE
U_FRANGE_M64, U_FRANGE_M32, U_FRANGE_M8, U_PROBE = 51, 52, 53, 54
KTILE = 256


def parse_device(text):
    d = {"coloff": {}, "colw": {}, "colsrc": {}, "strsrc": {}, "uinsn": [], "vkey": [], "vout": [], "vagg": [], "imm": {},
         "agg": [], "aggmap": [], "layout": (0, 0)}
    for line in text.strip().split("\n"):
        f = line.split()
        if f[0] == "col":
            d["colsrc"][int(f[1])] = int(f[3]); d["colw"][int(f[1])] = int(f[5])
        elif f[0] == "coloff":
            d["coloff"][int(f[1])] = int(f[2])
        elif f[0] == "strcol":
            d["strsrc"][int(f[1])] = int(f[3])
        elif f[0] == "layout":
            d["layout"] = (int(f[1]), int(f[2]))
        elif f[0] == "uinsn":
            d["uinsn"].append([int(x) for x in f[1:]])
        elif f[0] in ("vkey", "vout", "vagg"):
            d[f[0]].append((int(f[1]), int(f[2]), int(f[3])))
        elif f[0] == "imm":
            d["imm"][int(f[1])] = int(f[2])
        elif f[0] == "agg":
            d["agg"].append(int(f[2]))
        elif f[0] == "aggmap":
            d["aggmap"].append(int(f[2]))
    return d


def _i64(x):
    return np.asarray(x).astype(np.int64)


def run_pipeline_device(plan, pi, src_cols, pool_strings, agg_impl=IMPL_LOWAGG, with_stats=True):
    p = plan.pipelines[pi]
    impl = agg_impl if p["sink_kind"] == 1 else IMPL_EMIT
    types, widths, mins, maxs = [], [], [], []
    for a in src_cols:
        if a.dtype == object:
            types.append(N.RQ_I64); widths.append(8)
        else:
            t, w = phys_of(a)
            types.append(t); widths.append(w)
        if with_stats and a.dtype.kind in "iu" and len(a):
            mins.append(int(a.min())); maxs.append(int(a.max()))
        else:
            mins.append(1); maxs.append(0)
    D = parse_device(N.debug_lower(plan, pi, impl, types, widths, mins, maxs))
    vals = [PO._to_value(a) for a in src_cols]
    n = len(vals[0]) if vals else 0
    valid = np.ones(n, dtype=bool)
    slots = {}
    stage_bytes, slots_rel = D["layout"]
    col_at = {off: c for c, off in D["coloff"].items()}

    def const(imm):
        if imm >= STR_BASE and (imm - STR_BASE) in pool_strings:
            return PO._bcast(pool_strings[imm - STR_BASE], n)
        return np.full(n, imm, dtype=np.int64)

    def fetch(kind, in_slot, rel, imm):
        if kind in (K_M64, K_M32, K_M8):
            if in_slot:
                assert kind == K_M64 and rel >= slots_rel and (rel - slots_rel) % (KTILE * 8) == 0, "slot operand"
                return slots[(rel - slots_rel) // (KTILE * 8)]
            c = col_at[rel]
            assert {8: K_M64, 4: K_M32, 1: K_M8}[D["colw"][c]] == kind, "operand kind does not match the staged width"
            return vals[D["colsrc"][c]]
        if kind == K_STR:
            return vals[D["strsrc"][rel]]
        if kind == K_IMM2:
            return const(D["imm"][rel])
        return const(imm)

    def vref(kind, slotbits, off16):
        if kind == K_IMM:
            return const(D["imm"][off16])
        if kind == K_STR:
            return vals[D["strsrc"][off16]]
        return fetch(kind, slotbits & 1, off16 << 4, 0)

    def u32(a):
        return _i64(a).astype(np.uint64) & np.uint64(0xFFFFFFFF)

    old = np.seterr(over="ignore")
    for code, flags, dstrel, aux, gop, xk, yk, zk, xrel, yrel, zrel, imm in D["uinsn"]:
        t = None
        if 1 <= code <= 24:
            op = _BINOPS[(code - 1) // 2]
            x = fetch(K_M64, flags & UF_XSLOT, xrel, 0)
            y = fetch(K_M64, flags & UF_YSLOT, yrel, 0) if (code - 1) % 2 == 0 else np.full(n, imm, dtype=np.int64)
            t = _binop(op, _i64(x), _i64(y), valid)
        elif code in (U_MULADDI, U_MULSUBI, U_MULRSUBI):
            x, y = _i64(fetch(K_M64, flags & UF_XSLOT, xrel, 0)), _i64(fetch(K_M64, flags & UF_YSLOT, yrel, 0))
            inner = x + imm if code == U_MULADDI else (x - imm if code == U_MULSUBI else imm - x)
            t = inner * y
        elif code in (U_MULADDI32, U_MULSUBI32, U_MULRSUBI32, U_MUL32_MM):
            x, y = u32(fetch(K_M64, flags & UF_XSLOT, xrel, 0)), u32(fetch(K_M64, flags & UF_YSLOT, yrel, 0))
            k = np.uint64(imm & 0xFFFFFFFF)
            inner = {U_MULADDI32: x + k, U_MULSUBI32: x - k, U_MULRSUBI32: k - x, U_MUL32_MM: x}[code] & np.uint64(0xFFFFFFFF)
            t = (inner * y).astype(np.int64)
        elif code == U_GEN:
            x = fetch(xk, flags & UF_XSLOT, xrel, imm)
            if gop == D_LD:
                t = x
            elif gop == D_SEL:
                y = fetch(yk, flags & UF_YSLOT, yrel, imm)
                z = const(D["imm"][zrel]) if zk == K_IMM else fetch(zk, flags & UF_ZSLOT, zrel, 0)
                t = np.where((_i64(x) & 0xFF) != 0, y, z)
            else:
                y = fetch(yk, flags & UF_YSLOT, yrel, imm)
                t = _binop(gop, x, y, valid)
        elif U_F_M64 <= code < U_FRANGE_M64:
            w = (code - U_F_M64) // 6
            op = [D_LT, D_LE, D_GT, D_GE, D_EQ, D_NE][(code - U_F_M64) % 6]
            x = _i64(fetch([K_M64, K_M32, K_M8][w], 0, xrel, 0))
            k = imm if w == 0 else int(np.int64(imm).astype(np.int32))          # 4- and 1-byte columns compare in 32 bits
            valid = valid & (_binop(op, x, np.full(n, k, dtype=np.int64), valid) != 0)
            continue
        elif code in (U_FRANGE_M64, U_FRANGE_M32, U_FRANGE_M8):
            w = code - U_FRANGE_M64
            x = _i64(fetch([K_M64, K_M32, K_M8][w], 0, xrel, 0))
            if w == 0:
                span = yrel | (zrel << 32)
                ok = ((x.astype(np.uint64) - np.uint64(imm & 0xFFFFFFFFFFFFFFFF)) <= np.uint64(span))
            else:
                ok = ((u32(x) - np.uint64(imm & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)) <= np.uint64(yrel)
            valid = valid & ok
            continue
        else:
            raise NotImplementedError(f"device unit {code}")
        if flags & UF_STORE:
            assert dstrel >= slots_rel and (dstrel - slots_rel) % (KTILE * 8) == 0
            slots[(dstrel - slots_rel) // (KTILE * 8)] = t
        if flags & UF_FILTER:
            valid = valid & ((_i64(t) & 0xFF) != 0)
    np.seterr(**old)
    if p["sink_kind"] != 1:
        return [vref(*v)[valid] for v in D["vout"]]
    if not valid.any():
        return [np.zeros(0, dtype=np.int64) for _ in range(len(p["keys"]) + len(p["vals"]))]
    keys = [vref(*v) for v in D["vkey"]]
    groups, gid = {}, np.zeros(n, dtype=np.int64)
    for r in np.nonzero(valid)[0]:
        gid[r] = groups.setdefault(tuple(int(x[r]) for x in keys), len(groups))
    ng = max(1, len(groups))
    g = gid[valid]
    aggs = []
    old = np.seterr(over="ignore")
    for u, kind in enumerate(D["agg"]):
        if kind == 2:
            aggs.append(np.bincount(g, minlength=ng).astype(np.int64))
            continue
        full = _i64(vref(*D["vagg"][u]))
        if D["vagg"][u][1] & 2:           # the lowering claims the value fits unsigned 32 bits
            assert len(full) == 0 or (int(full[valid].min()) >= 0 and int(full[valid].max()) < 2 ** 32), "u32 claim on wide values"
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
    return keycols + [aggs[u] for u in D["aggmap"]]
