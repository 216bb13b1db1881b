// Included by engine.cu: program emission and plan execution.

namespace {

struct AggDedup {
    std::vector<int> uniq_of;     // sink value -> unique aggregate index
    std::vector<int> kind, node;  // per unique aggregate
};

static AggDedup dedup_aggs(const rq_pipeline& pl) {
    AggDedup d;
    d.uniq_of.assign(pl.n_vals, -1);
    for (int k = 0; k < pl.n_vals; k++) {
        const int kind = pl.vals[k].kind;
        if (kind < RQ_AGG_SUM || kind > RQ_AGG_MAX) raise(RQ_ERR_INVALID, "aggregate %d has kind %d", k, kind);
        const int node = (kind == RQ_AGG_COUNT) ? -1 : pl.vals[k].node;
        for (size_t u = 0; u < d.kind.size(); u++)
            if (d.kind[u] == kind && d.node[u] == node) { d.uniq_of[k] = (int)u; break; }
        if (d.uniq_of[k] < 0) {
            d.uniq_of[k] = (int)d.kind.size();
            d.kind.push_back(kind);
            d.node.push_back(node);
        }
    }
    if ((int)d.kind.size() > kMaxAggs) raise(RQ_ERR_UNSUPPORTED, "more than %d distinct aggregates", kMaxAggs);
    return d;
}

// ---- plan clean-up before lowering: constant folding, CSE, dead-code elimination -----------
// The reference re-materialises every constant and re-evaluates every sub-expression per tuple
// (ExpressionsJitFlounder.h emits `0.06 - 0.01` inside the scan loop); folding and sharing are
// result-identical under wrap-around int64 arithmetic.
struct SimplePipe {
    std::vector<rq_node> nodes;
    std::vector<int32_t> args;
    std::vector<rq_value> keys, vals;
    rq_pipeline pl;
};

static bool fold_binary(int op, int64_t x, int64_t y, int64_t* out) {
    const uint64_t ux = (uint64_t)x, uy = (uint64_t)y;
    switch (op) {
        case RQ_OP_ADD: *out = (int64_t)(ux + uy); return true;
        case RQ_OP_SUB: *out = (int64_t)(ux - uy); return true;
        case RQ_OP_MUL: *out = (int64_t)(ux * uy); return true;
        case RQ_OP_DIV:
            if (y == 0) return false;
            *out = (y == -1) ? (int64_t)(0ULL - ux) : x / y;
            return true;
        case RQ_OP_AND: *out = x & y; return true;
        case RQ_OP_OR:  *out = x | y; return true;
        case RQ_OP_LT:  *out = x < y; return true;
        case RQ_OP_LE:  *out = x <= y; return true;
        case RQ_OP_GT:  *out = x > y; return true;
        case RQ_OP_GE:  *out = x >= y; return true;
        case RQ_OP_EQ:  *out = x == y; return true;
        case RQ_OP_NEQ: *out = x != y; return true;
        default: return false;
    }
}

static bool is01(int op) {
    switch (op) {
        case RQ_OP_LT: case RQ_OP_LE: case RQ_OP_GT: case RQ_OP_GE: case RQ_OP_EQ: case RQ_OP_NEQ:
        case RQ_OP_EQ_CHAR: case RQ_OP_EQ_VARCHAR: case RQ_OP_NEQ_CHAR: case RQ_OP_NEQ_VARCHAR:
        case RQ_OP_LIKE:
            return true;
        default: return false;
    }
}

static void simplify_pipeline(const rq_pipeline& in, SimplePipe& sp) {
    const int n = in.n_nodes;
    std::vector<rq_node> tmp;
    std::vector<int> remap(n, -1);
    std::map<std::vector<int64_t>, int> seen;
    auto intern = [&](const rq_node& nd) -> int {
        std::vector<int64_t> key = {nd.op, nd.a, nd.b, nd.c, nd.imm};
        if (nd.op != RQ_OP_FILTER && nd.op != RQ_OP_PROBE) {
            auto it = seen.find(key);
            if (it != seen.end()) return it->second;
        }
        tmp.push_back(nd);
        const int idx = (int)tmp.size() - 1;
        seen[key] = idx;
        return idx;
    };
    auto ref = [&](int i, int r) -> int {
        if (r < 0 || r >= i) raise(RQ_ERR_INVALID, "node %d refers to node %d (must be an earlier node)", i, r);
        return remap[r];
    };
    std::vector<int32_t> args(in.args, in.args + in.n_args);
    std::function<bool(int)> is01node = [&](int t) -> bool {
        const rq_node& x = tmp[t];
        if (x.op == RQ_OP_AND || x.op == RQ_OP_OR) return is01node(x.a) && is01node(x.b);
        if (x.op == RQ_OP_CONST) return x.imm == 0 || x.imm == 1;
        return is01(x.op);
    };
    for (int i = 0; i < n; i++) {
        rq_node nd = in.nodes[i];
        if (is_binary(nd.op)) {
            nd.a = ref(i, nd.a); nd.b = ref(i, nd.b); nd.c = 0;
            const rq_node& x = tmp[nd.a];
            const rq_node& y = tmp[nd.b];
            int64_t v;
            if (x.op == RQ_OP_CONST && y.op == RQ_OP_CONST && fold_binary(nd.op, x.imm, y.imm, &v)) {
                rq_node c{RQ_OP_CONST, 0, 0, 0, v};
                remap[i] = intern(c);
                continue;
            }
            if (nd.op == RQ_OP_MUL && y.op == RQ_OP_CONST && y.imm == 1) { remap[i] = nd.a; continue; }
            if (nd.op == RQ_OP_MUL && x.op == RQ_OP_CONST && x.imm == 1) { remap[i] = nd.b; continue; }
            if (nd.op == RQ_OP_DIV && y.op == RQ_OP_CONST && y.imm == 1) { remap[i] = nd.a; continue; }
        } else if (nd.op == RQ_OP_FILTER) {
            nd.a = ref(i, nd.a); nd.b = nd.c = 0;
            // FILTER(x AND y) == FILTER x; FILTER y when both sides are 0/1 values (compare
            // results): no slot traffic, and whole warps can leave the program early
            std::vector<int> stack{nd.a}, parts;
            while (!stack.empty()) {
                const int t = stack.back(); stack.pop_back();
                if (tmp[t].op == RQ_OP_AND && is01node(tmp[t].a) && is01node(tmp[t].b)) {
                    stack.push_back(tmp[t].b); stack.push_back(tmp[t].a);
                } else parts.push_back(t);
            }
            for (size_t k = 0; k + 1 < parts.size(); k++) { rq_node f{RQ_OP_FILTER, parts[k], 0, 0, 0}; intern(f); }
            nd.a = parts.back();
        } else if (nd.op == RQ_OP_SELECT) {
            nd.a = ref(i, nd.a); nd.b = ref(i, nd.b); nd.c = ref(i, nd.c);
            // a constant condition picks its arm here (the device unit has a single immediate field,
            // which the THEN arm may need)
            if (tmp[nd.a].op == RQ_OP_CONST) { remap[i] = (tmp[nd.a].imm & 0xff) ? nd.b : nd.c; continue; }
        } else if (nd.op == RQ_OP_PROBE) {
            if (nd.b < 0 || nd.c < 0 || nd.b + nd.c > in.n_args) raise(RQ_ERR_INVALID, "node %d: probe args out of range", i);
            for (int k = 0; k < nd.c; k++) args[nd.b + k] = ref(i, in.args[nd.b + k]);
        } else if (nd.op == RQ_OP_PAYLOAD) {
            nd.a = ref(i, nd.a);
            if (tmp[nd.a].op != RQ_OP_PROBE) raise(RQ_ERR_INVALID, "node %d: PAYLOAD of a non-PROBE node", i);
        } else if (nd.op == RQ_OP_COL || nd.op == RQ_OP_CONST || nd.op == RQ_OP_CONST_STR) {
            if (nd.op != RQ_OP_COL) nd.a = 0;
            nd.b = nd.c = 0;
            if (nd.op == RQ_OP_COL) nd.imm = 0;
        } else {
            raise(RQ_ERR_INVALID, "node %d: unknown op %d", i, nd.op);
        }
        remap[i] = intern(nd);
    }
    // dead-code elimination
    const int m = (int)tmp.size();
    std::vector<char> live(m, 0);
    auto sink_ref = [&](int r) {
        if (r < 0 || r >= n) raise(RQ_ERR_INVALID, "sink refers to node %d", r);
        return remap[r];
    };
    sp.keys.assign(in.keys, in.keys + in.n_keys);
    sp.vals.assign(in.vals, in.vals + in.n_vals);
    for (auto& k : sp.keys) { k.node = sink_ref(k.node); live[k.node] = 1; }
    for (auto& v : sp.vals) {
        if (in.sink_kind == RQ_SINK_AGG && v.kind == RQ_AGG_COUNT) { v.node = 0; continue; }
        v.node = sink_ref(v.node); live[v.node] = 1;
    }
    for (int i = m - 1; i >= 0; i--) {
        const rq_node& nd = tmp[i];
        if (nd.op == RQ_OP_FILTER || nd.op == RQ_OP_PROBE) live[i] = 1;
        if (!live[i]) continue;
        if (is_binary(nd.op)) { live[nd.a] = live[nd.b] = 1; }
        else if (nd.op == RQ_OP_FILTER) live[nd.a] = 1;
        else if (nd.op == RQ_OP_SELECT) { live[nd.a] = live[nd.b] = live[nd.c] = 1; }
        else if (nd.op == RQ_OP_PAYLOAD) live[nd.a] = 1;
        else if (nd.op == RQ_OP_PROBE) for (int k = 0; k < nd.c; k++) live[args[nd.b + k]] = 1;
    }
    std::vector<int> remap2(m, -1);
    for (int i = 0; i < m; i++) {
        if (!live[i]) continue;
        rq_node nd = tmp[i];
        if (is_binary(nd.op)) { nd.a = remap2[nd.a]; nd.b = remap2[nd.b]; }
        else if (nd.op == RQ_OP_FILTER) nd.a = remap2[nd.a];
        else if (nd.op == RQ_OP_SELECT) { nd.a = remap2[nd.a]; nd.b = remap2[nd.b]; nd.c = remap2[nd.c]; }
        else if (nd.op == RQ_OP_PAYLOAD) nd.a = remap2[nd.a];
        else if (nd.op == RQ_OP_PROBE) for (int k = 0; k < nd.c; k++) args[nd.b + k] = remap2[args[nd.b + k]];
        sp.nodes.push_back(nd);
        remap2[i] = (int)sp.nodes.size() - 1;
    }
    for (auto& k : sp.keys) k.node = remap2[k.node];
    for (auto& v : sp.vals)
        if (!(in.sink_kind == RQ_SINK_AGG && v.kind == RQ_AGG_COUNT)) v.node = remap2[v.node];
    // hoist every FILTER to right behind the value it tests: the value is still in the
    // accumulator and later work is skipped for dropped tuples (expressions are side-effect free)
    {
        // Source columns come first (they depend on nothing): the pipeline splitters cut in front of
        // a probe and rely on every column read sitting before the cut.
        const int k = (int)sp.nodes.size();
        std::vector<int> order_, pos(k, -1);
        for (int pass = 0; pass < 2; pass++)
            for (int i = 0; i < k; i++) {
                if (sp.nodes[i].op == RQ_OP_FILTER || (sp.nodes[i].op == RQ_OP_COL) != (pass == 0)) continue;
                order_.push_back(i);
                for (int f = 0; f < k; f++)
                    if (sp.nodes[f].op == RQ_OP_FILTER && sp.nodes[f].a == i) order_.push_back(f);
            }
        for (int i = 0; i < k; i++) pos[order_[i]] = i;
        std::vector<rq_node> re(k);
        for (int i = 0; i < k; i++) {
            rq_node nd = sp.nodes[order_[i]];
            if (is_binary(nd.op)) { nd.a = pos[nd.a]; nd.b = pos[nd.b]; }
            else if (nd.op == RQ_OP_FILTER) nd.a = pos[nd.a];
            else if (nd.op == RQ_OP_SELECT) { nd.a = pos[nd.a]; nd.b = pos[nd.b]; nd.c = pos[nd.c]; }
            else if (nd.op == RQ_OP_PAYLOAD) nd.a = pos[nd.a];
            else if (nd.op == RQ_OP_PROBE) for (int q = 0; q < nd.c; q++) args[nd.b + q] = pos[args[nd.b + q]];
            re[i] = nd;
        }
        sp.nodes = re;
        for (auto& kk : sp.keys) kk.node = pos[kk.node];
        for (auto& v : sp.vals)
            if (!(in.sink_kind == RQ_SINK_AGG && v.kind == RQ_AGG_COUNT)) v.node = pos[v.node];
    }
    sp.args = args;
    sp.pl = in;
    sp.pl.n_nodes = (int)sp.nodes.size(); sp.pl.nodes = sp.nodes.data();
    sp.pl.n_args = (int)sp.args.size();   sp.pl.args = sp.args.data();
    sp.pl.keys = sp.keys.data();          sp.pl.vals = sp.vals.data();
}

// Selection compares on the same column intersect to one interval test (BETWEEN, date ranges):
// lo <= x <= hi is a single unsigned compare of (x - lo) against (hi - lo). Selection compares
// only narrow the mask and read columns, so the later one may move up to the earlier one.
static void fuse_ranges(Lowerer& L) {
    auto interval = [&](const HUnit& u, __int128* lo, __int128* hi) -> bool {
        const __int128 NEG = -((__int128)1 << 100), POS = (__int128)1 << 100;
        *lo = NEG; *hi = POS;
        if (u.op == H_FRANGE) { *lo = u.imm; *hi = (__int128)u.imm + (__int128)(uint64_t)u.imm2; return true; }
        if (u.op != H_FCMP) return false;
        switch (u.gop) {
            case D_GE: *lo = u.imm; return true;
            case D_GT: *lo = (__int128)u.imm + 1; return true;
            case D_LE: *hi = u.imm; return true;
            case D_LT: *hi = (__int128)u.imm - 1; return true;
            case D_EQ: *lo = *hi = u.imm; return true;
            default: return false;
        }
    };
    const __int128 NEG = -((__int128)1 << 100), POS = (__int128)1 << 100;
    for (size_t i = 0; i < L.prog.size(); i++) {
        __int128 lo, hi;
        if (!interval(L.prog[i], &lo, &hi)) continue;
        for (size_t j = i + 1; j < L.prog.size();) {
            __int128 lo2, hi2;
            const HUnit& v = L.prog[j];
            if ((v.op == H_FCMP || v.op == H_FRANGE) && v.x.kind == L.prog[i].x.kind && v.x.idx == L.prog[i].x.idx &&
                interval(v, &lo2, &hi2)) {
                const __int128 nlo = lo > lo2 ? lo : lo2, nhi = hi < hi2 ? hi : hi2;
                if (nlo <= nhi) {           // (an empty intersection is left to the two compares)
                    lo = nlo; hi = nhi;
                    L.prog.erase(L.prog.begin() + j);
                    continue;
                }
            }
            j++;
        }
        HUnit& u = L.prog[i];
        // value range of the compared column as the device sees it: 8-byte columns are int64, 4-byte
        // columns compare as int32, 1-byte columns are zero-extended bytes. A bound outside it (from
        // `col > INT32_MAX`, which became lo = 2^31) must not be truncated into the compare constant.
        const int w = L.P.col_w[u.x.idx];
        const __int128 tmin = w == 8 ? (__int128)INT64_MIN : w == 4 ? (__int128)INT32_MIN : 0;
        const __int128 tmax = w == 8 ? (__int128)INT64_MAX : w == 4 ? (__int128)INT32_MAX : 255;
        if ((lo != NEG && lo > tmax) || (hi != POS && hi < tmin)) {
            // no value of the column can pass: x < (smallest compare constant) is always false
            u.op = H_FCMP; u.gop = D_LT; u.imm = w == 8 ? INT64_MIN : INT32_MIN; u.imm2 = 0;
            continue;
        }
        if (lo != NEG && lo <= tmin) lo = NEG;       // every value passes this side
        if (hi != POS && hi >= tmax) hi = POS;
        if (lo != NEG && hi != POS) {
            u.op = H_FRANGE; u.gop = 0; u.imm = (int64_t)lo; u.imm2 = (int64_t)(uint64_t)(hi - lo);
        } else if (lo != NEG) { u.op = H_FCMP; u.gop = D_GE; u.imm = (int64_t)lo; }
        else if (hi != POS) { u.op = H_FCMP; u.gop = D_LE; u.imm = (int64_t)hi; }
        else { u.op = H_FCMP; u.gop = D_GE; u.imm = w == 8 ? INT64_MIN : INT32_MIN; u.imm2 = 0; }   // always true
    }
}

// Emits the host-level program (units) for one pipeline.
static void emit_program(Lowerer& L, int impl, const AggDedup& ad) {
    KParams& P = L.P;
    const rq_pipeline& pl = L.pl;
    const int n = L.n;

    // a value needs a slot when somebody other than a selection directly behind it reads it
    auto needs_slot = [&](int i) -> bool {
        if (L.sink_ref[i]) return true;
        for (int j = i + 1; j < n; j++) {
            const rq_node& nd = pl.nodes[j];
            if (L.fused[j] && nd.op != RQ_OP_FILTER) {
                // a folded inner node reads only leaves
                continue;
            }
            if (is_binary(nd.op) && (nd.a == i || nd.b == i)) return true;
            if (nd.op == RQ_OP_SELECT && (nd.a == i || nd.b == i || nd.c == i)) return true;
        }
        return false;
    };
    auto filter_behind = [&](int i) -> int {     // FILTER node testing value i right behind it
        for (int j = i + 1; j < n && pl.nodes[j].op == RQ_OP_FILTER; j++)
            if (pl.nodes[j].a == i) return j;
        return -1;
    };
    auto finish_value = [&](int i) {
        HUnit& u = L.prog.back();
        if (needs_slot(i)) {
            const int s = L.alloc_slot();
            L.slot[i] = s;
            u.dst = s;
        }
        if (filter_behind(i) >= 0) u.filt = true;
    };

    for (int i = 0; i < n; i++) {
        const rq_node& nd = pl.nodes[i];
        if (is_leaf(nd.op) || nd.op == RQ_OP_PAYLOAD) {
            // no unit
        } else if (nd.op == RQ_OP_FILTER) {
            HOpnd col; int64_t k;
            const int fc = L.fcmp_of(i, &col, &k);
            if (fc) {
                HUnit& u = L.emit(H_FCMP, (uint8_t)fc);
                u.x = col; u.imm = k;
            } else if (is_leaf(pl.nodes[nd.a].op)) {
                HUnit& u = L.emit(H_BIN, D_LD);      // selection on a plain (BOOL) operand
                u.x = L.operand_of(nd.a);
                u.filt = true;
            }
            // otherwise folded into the unit that computed the value (finish_value)
        } else if (L.fused[i]) {
            // folded into its consumer
        } else if (is_binary(nd.op)) {
            int inner, other; uint8_t gop; int64_t k; HOpnd x;
            if (L.muli_of(i, &inner, &other, &gop, &k, &x)) {
                HUnit& u = L.emit(H_MULI, gop);
                u.x = x; u.imm = k; u.y = L.operand_of(other);
                u.n32 = L.is_u32(inner) && L.is_u32(other) && k >= 0 && k <= 0xffffffffLL;
            } else {
                HUnit& u = L.emit(H_BIN, dop_left(nd.op));
                u.x = L.operand_of(nd.a); u.y = L.operand_of(nd.b);
                u.n32 = nd.op == RQ_OP_MUL && L.is_u32(nd.a) && L.is_u32(nd.b);
            }
            finish_value(i);
        } else if (nd.op == RQ_OP_SELECT) {
            HUnit& u = L.emit(H_SEL, D_SEL);
            u.x = L.operand_of(nd.a); u.y = L.operand_of(nd.b); u.z = L.operand_of(nd.c);
            finish_value(i);
        } else if (nd.op == RQ_OP_PROBE) {
            if (P.n_probes >= kMaxProbes) raise(RQ_ERR_UNSUPPORTED, "more than %d joins in one pipeline", kMaxProbes);
            if (nd.a < 0 || nd.a >= (int)L.outs.size() || !L.outs[nd.a].ht)
                raise(RQ_ERR_INVALID, "PROBE refers to pipeline %d which built no hash table", nd.a);
            DProbe& pr = P.probe[P.n_probes];
            pr.ht = L.outs[nd.a].ht->d;
            // internal modes (set by split_at_probe, never by the ABI caller): bit1 = Bloom-only
            // semi-join pass, bit2 = expansion pass (Bloom test here, one output row per match in
            // the materialize sink), bit3 = fetch the payload of the entry a previous pass found
            pr.fetch = (int32_t)((nd.imm >> 3) & 1);
            if (!pr.fetch && nd.c != pr.ht.nk) raise(RQ_ERR_INVALID, "PROBE has %d keys, the build side %d", nd.c, pr.ht.nk);
            if (pr.fetch && nd.c != 1) raise(RQ_ERR_INVALID, "internal: fetch probe needs exactly the entry index");
            for (int k = 0; k < nd.c; k++) L.hprobe_key[P.n_probes][k] = L.href_of(pl.args[nd.b + k]);
            pr.single = (int32_t)(nd.imm & 1);
            pr.bloom_only = (int32_t)(((nd.imm >> 1) | (nd.imm >> 2)) & 1);
            if (pr.bloom_only && pr.ht.bloom == nullptr && !pr.ht.direct) raise(RQ_ERR_INVALID, "internal: semi-join pass without a filter");
            if ((nd.imm >> 2) & 1) {
                if (impl != IMPL_EMIT || i != n - 1) raise(RQ_ERR_INVALID, "internal: expansion probe must end a materialize pipeline");
                P.expand_probe = P.n_probes;
            }
            const std::vector<uint8_t>& pw = L.outs[nd.a].pay_word;
            pr.n_out = (int32_t)pw.size();
            if (pr.n_out > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "more than %d payload columns", kMaxOut);
            memset(pr.out_slot, 0xff, sizeof(pr.out_slot));
            memset(pr.pay_word, 0, sizeof(pr.pay_word));
            pr.dup_counter = (unsigned long long*)(E.flags + 4);
            for (int j = i + 1; j < n; j++) {
                if (pl.nodes[j].op != RQ_OP_PAYLOAD || pl.nodes[j].a != i) continue;
                const int b = pl.nodes[j].b;
                if (b < 0 || b >= pr.n_out) raise(RQ_ERR_INVALID, "PAYLOAD %d out of range", b);
                if (pw[b] == 0xff) raise(RQ_ERR_INVALID, "internal: payload %d was not stored by the build", b);
                pr.pay_word[b] = pw[b];
                const int sl = L.alloc_slot();
                L.slot[j] = sl;
                pr.out_slot[b] = (uint8_t)sl;
            }
            HUnit& u = L.emit(H_PROBE, 0);
            u.aux = P.n_probes;
            P.n_probes++;
        }
        L.release_dead(i);
    }

    fuse_ranges(L);

    // sinks
    if (pl.n_keys > kMaxKeys) raise(RQ_ERR_UNSUPPORTED, "more than %d key columns", kMaxKeys);
    for (int k = 0; k < pl.n_keys; k++) L.hkey.push_back(L.href_of(pl.keys[k].node));
    if (impl == IMPL_BUILD || impl == IMPL_EMIT) {
        if (pl.n_vals > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "more than %d output columns", kMaxOut);
        for (int k = 0; k < pl.n_vals; k++) L.hout.push_back(L.href_of(pl.vals[k].node));
    } else {
        P.na = (int)ad.kind.size();
        for (int u = 0; u < P.na; u++) {
            P.agg_kind[u] = (uint8_t)ad.kind[u];
            if (ad.kind[u] != RQ_AGG_COUNT) L.hagg_src[u] = L.href_of(ad.node[u]);
        }
    }
    P.sink = impl;
}

// ---- shared-memory layout and launch geometry --------------------------------------------
// [mbarriers: kMaxWarps x kMaxStages][program][warp region 0][warp region 1]...; a warp region is
// [stages][slots][lane-private accumulators].
static bool layout_smem(KParams& P, int n_slots, int acc_bytes, int max_warps) {
    const uint32_t bars = kMaxWarps * kMaxStages * 8 + kMaxInsn * (uint32_t)sizeof(UInsn);   // + program copy
    auto warp_bytes = [&](int S) {
        uint32_t wb = (uint32_t)S * P.stage_bytes + (uint32_t)n_slots * kTile * 8 + (uint32_t)acc_bytes;
        wb = (wb + 127) & ~127u;
        return wb == 0 ? 128u : wb;
    };
    auto fit = [&](int S) { return (int)std::min<int64_t>(max_warps, ((int64_t)kSmemMax - bars) / warp_bytes(S)); };
    // Warps hide latency better than a second stage does (measured: throughput grows almost
    // linearly with resident warps, single- and double-buffered runs tie at equal warp counts), so
    // take the most warps; double buffering only when it costs no warp.
    int bestW = fit(1), bestS = 1;
    if (bestW < 1) return false;
    if (fit(2) >= bestW) bestS = 2;
    // rq_set_option("stages" / "warps") overrides the choice (profiles/r01_q1_sweep.txt was made that way)
    if (E.opt.stages >= 1 && E.opt.stages <= kMaxStages && fit(E.opt.stages) >= 1) { bestS = E.opt.stages; bestW = fit(bestS); }
    if (E.opt.warps >= 1 && E.opt.warps <= bestW) bestW = E.opt.warps;
    P.stages = bestS;
    P.warps = bestW;
    P.n_slots = n_slots;
    const uint32_t wb = warp_bytes(bestS);
    P.prog_off = kMaxWarps * kMaxStages * 8;
    P.warp_off = bars;
    P.warp_bytes = wb;
    P.slots_rel = (uint32_t)bestS * P.stage_bytes;
    P.acc_rel = P.slots_rel + (uint32_t)n_slots * kTile * 8;
    P.smem_bytes = bars + (uint32_t)bestW * wb;
    return true;
}

// ---- encode: host-level units -> fused device instructions (needs the layout) -----------------
struct UOperand { uint8_t kind; uint8_t slot; uint32_t off; };

static UOperand resolve(const KParams& P, const HOpnd& h) {
    UOperand u{K_NONE, 0, 0};
    switch (h.kind) {
        case S_COL:
            u.kind = P.col_w[h.idx] == 8 ? K_M64 : (P.col_w[h.idx] == 4 ? K_M32 : K_M8);
            u.off = P.col_off[h.idx];
            break;
        case S_SLOT: u.kind = K_M64; u.slot = 1; u.off = P.slots_rel + (uint32_t)h.idx * kTile * 8; break;
        case S_IMM: u.kind = K_IMM; break;
        case S_STR: u.kind = K_STR; u.off = (uint32_t)h.idx << 4; break;
        default: break;
    }
    return u;
}
static VRef to_vref(const KParams& P, const HRef& h) {
    VRef v; v.kind = K_NONE; v.slot = 0; v.off16 = 0;
    if (h.kind == S_IMM) { v.kind = K_IMM; v.off16 = h.idx; return v; }
    const UOperand u = resolve(P, h);
    v.kind = u.kind; v.slot = (uint8_t)(u.slot | (h.u32 ? 2 : 0)); v.off16 = (uint16_t)(u.off >> 4);
    return v;
}

static int bin_index(uint8_t op) {   // position in RQ_BINOPS, -1 if not a fusable binary op
    switch (op) {
        case D_ADD: return 0; case D_SUB: return 1; case D_RSUB: return 2; case D_MUL: return 3;
        case D_AND: return 4; case D_OR: return 5; case D_LT: return 6; case D_LE: return 7;
        case D_GT: return 8; case D_GE: return 9; case D_EQ: return 10; case D_NE: return 11;
    }
    return -1;
}

static void encode_program(Lowerer& L, KParams& P) {
    P.n_insn = 0;
    for (size_t i = 0; i < L.prog.size(); i++) {
        const HUnit& h = L.prog[i];
        UOperand x = resolve(P, h.x), y = resolve(P, h.y), z = resolve(P, h.z);
        UInsn u;
        memset(&u, 0, sizeof(u));
        if (h.dst >= 0) { u.flags |= UF_STORE; u.dstrel = P.slots_rel + (uint32_t)h.dst * kTile * 8; }
        if (h.filt) u.flags |= UF_FILTER;
        u.aux = (uint8_t)h.aux;
        u.gop = h.gop;
        u.imm = h.imm;
        if (h.op == H_FCMP) {
            const int ci = h.gop - D_LT;
            if (ci < 0 || ci > 5) raise(RQ_ERR_INVALID, "internal: bad fused compare");
            if (x.kind == K_M64) u.code = (uint8_t)(U_FLT_M64 + ci);
            else if (x.kind == K_M32) u.code = (uint8_t)(U_FLT_M32 + ci);
            else if (x.kind == K_M8) u.code = (uint8_t)(U_FLT_M8 + ci);
            else raise(RQ_ERR_INVALID, "internal: fused compare on a non-column operand");
        } else if (h.op == H_FRANGE) {
            u.code = x.kind == K_M64 ? U_FRANGE_M64 : (x.kind == K_M32 ? U_FRANGE_M32 : U_FRANGE_M8);
            if (x.kind != K_M64 && x.kind != K_M32 && x.kind != K_M8) raise(RQ_ERR_INVALID, "internal: range compare on a non-column operand");
            y = UOperand{K_NONE, 0, (uint32_t)((uint64_t)h.imm2 & 0xffffffffu)};
            z = UOperand{K_NONE, 0, (uint32_t)((uint64_t)h.imm2 >> 32)};
        } else if (h.op == H_MULI) {
            if (x.kind != K_M64 || y.kind != K_M64) raise(RQ_ERR_INVALID, "internal: MULI operands");
            u.code = h.gop == D_ADD ? U_MULADDI : (h.gop == D_SUB ? U_MULSUBI : U_MULRSUBI);
            if (h.n32) u.code = h.gop == D_ADD ? U_MULADDI32 : (h.gop == D_SUB ? U_MULSUBI32 : U_MULRSUBI32);
        } else if (h.op == H_PROBE) {
            u.code = U_PROBE;
        } else if (h.op == H_SEL) {
            u.code = U_GEN;
            if (y.kind == K_IMM) u.imm = h.y.imm;                 // (a constant condition was folded: x is never K_IMM)
            if (x.kind == K_IMM) raise(RQ_ERR_INVALID, "internal: SELECT with a constant condition reached the device program");
            if (z.kind == K_IMM) z.off = (uint32_t)L.imm_index(h.z.imm);
        } else {   // H_BIN
            const int bi = bin_index(h.gop);
            static const int swapped[12] = {0, 2, 1, 3, 4, 5, 8, 9, 6, 7, 10, 11};
            if (bi >= 0 && x.kind == K_M64 && y.kind == K_M64) {
                u.code = (uint8_t)(U_ADD_MM + 2 * bi);
                if (h.n32 && h.gop == D_MUL) u.code = U_MUL32_MM;
            } else if (bi >= 0 && x.kind == K_M64 && y.kind == K_IMM) {
                u.code = (uint8_t)(U_ADD_MM + 2 * bi + 1);
                u.imm = h.y.imm;
            } else if (bi >= 0 && x.kind == K_IMM && y.kind == K_M64) {
                // imm OP m64  ==  m64 OP' imm
                u.code = (uint8_t)(U_ADD_MM + 2 * swapped[bi] + 1);
                u.imm = h.x.imm;
                x = y;
                y = UOperand{K_NONE, 0, 0};
            } else {
                u.code = U_GEN;
                if (y.kind == K_IMM) u.imm = h.y.imm;
                if (x.kind == K_IMM) {
                    // both operands constant (string predicates on two literals are not folded): the unit has
                    // one immediate field, the second constant goes through the immediate table
                    if (y.kind == K_IMM && h.gop != D_LD) { y.kind = K_IMM2; y.off = (uint32_t)L.imm_index(h.y.imm); }
                    u.imm = h.x.imm;
                }
            }
        }
        if (x.slot) u.flags |= UF_XSLOT;
        if (y.slot) u.flags |= UF_YSLOT;
        if (z.slot) u.flags |= UF_ZSLOT;
        u.xkind = x.kind; u.ykind = y.kind; u.zkind = z.kind;
        u.xrel = x.kind == K_STR ? (x.off >> 4) : x.off;
        u.yrel = y.kind == K_STR ? (y.off >> 4) : y.off;
        u.zrel = z.kind == K_STR ? (z.off >> 4) : z.off;
        if (P.n_insn >= kMaxInsn) raise(RQ_ERR_UNSUPPORTED, "program longer than %d units", kMaxInsn);
        P.insn[P.n_insn++] = u;
    }
    // sinks
    P.nk = (int)L.hkey.size();
    for (int k = 0; k < P.nk; k++) P.key[k] = to_vref(P, L.hkey[k]);
    P.n_out = (int)L.hout.size();
    for (int k = 0; k < P.n_out; k++) P.out[k] = to_vref(P, L.hout[k]);
    for (int u = 0; u < kMaxAggs; u++) P.agg_src[u] = to_vref(P, L.hagg_src[u]);
    for (int p = 0; p < P.n_probes; p++)
        for (int k = 0; k < (P.probe[p].fetch ? 1 : P.probe[p].ht.nk); k++) P.probe[p].key[k] = to_vref(P, L.hprobe_key[p][k]);
}

// packed group key of the low-cardinality paths: every key is a bit field of one 64-bit word
static bool pack_group_key(const Lowerer& L, KParams& P, KeyUnpack& ku) {
    memset(&ku, 0, sizeof(ku));
    int total = 0;
    ku.nk = (int)L.hkey.size();
    if (ku.nk > kMaxKeys) return false;
    for (int k = 0; k < ku.nk; k++) {
        const HRef& h = L.hkey[k];
        int bits = 64, sign = 1;
        if (h.kind == S_STR) return false;
        if (h.kind == S_COL) {
            const int w = P.col_w[h.idx];
            bits = 8 * w;
            sign = (w == 1) ? 0 : 1;
        }
        if (h.lo >= 0) {        // non-negative values need no sign and only as many bits as the bound
            int need = 1;
            while (need < 63 && ((int64_t)1 << need) <= h.hi) need++;
            if (need < bits) { bits = need; sign = 0; }
        }
        if (total + bits > 64) return false;
        P.key_shift[k] = (uint8_t)total; P.key_bits[k] = (uint8_t)bits;
        ku.shift[k] = (uint8_t)total; ku.bits[k] = (uint8_t)bits; ku.sign[k] = (uint8_t)sign;
        total += bits;
    }
    P.key32 = total <= 31 ? 1 : 0;    // leaves room for the two sentinels of the register path
    return true;
}

// Register-aggregation path: accumulation form of every SUM (rq_internal.h AggMode) from the value
// bounds, and how often the 32-bit piece sums must be flushed so that none can wrap. A lane adds
// at most kR tuples per tile to an accumulator.
static void choose_agg_modes(const Lowerer& L, KParams& P, int64_t src_rows) {
    uint64_t flush = UINT64_MAX;          // tiles between two flushes
    auto bound_tiles = [](uint64_t piece_max) -> uint64_t {
        return 0xffffffffULL / ((uint64_t)kR * std::max<uint64_t>(piece_max, 1));
    };
    for (int u = 0; u < P.na; u++) {
        P.agg_mode[u] = AM_FULL;
        P.agg_shift[u] = 0;
        if (P.agg_kind[u] != RQ_AGG_SUM) continue;
        const HRef& h = L.hagg_src[u];
        if (h.lo < 0) continue;
        const uint64_t hi = (uint64_t)h.hi;
        if (hi < (1ULL << 25)) {
            P.agg_mode[u] = AM_P1;
            flush = std::min(flush, bound_tiles(hi));
        } else if (hi <= 0xffffffffULL) {
            P.agg_mode[u] = AM_W64;
        } else if (hi < (1ULL << 48)) {
            int bits = 0;
            while (bits < 63 && (1ULL << bits) <= hi) bits++;
            const int sh = (bits + 1) / 2;
            P.agg_mode[u] = AM_P2;
            P.agg_shift[u] = (uint8_t)sh;
            flush = std::min(flush, bound_tiles((1ULL << sh) - 1));
            flush = std::min(flush, bound_tiles(hi >> sh));
        }
    }
    // tuple counters are 32-bit as well: kR per tile
    flush = std::min(flush, bound_tiles(1));
    const int64_t tiles = std::max<int64_t>(1, (src_rows + kTile - 1) / kTile);
    const int64_t grid = std::min<int64_t>((tiles + P.warps - 1) / P.warps, (int64_t)E.sm_count > 0 ? E.sm_count : 148);
    const uint64_t per_warp = (uint64_t)((tiles + grid * P.warps - 1) / (grid * P.warps));
    P.flush_tiles = flush >= per_warp ? 0 : (int32_t)std::min<uint64_t>(flush, 1u << 30);
}

// ---- host-side trace (RQ_TRACE): wall-clock offsets since the plan started, after a stream sync --
static std::chrono::steady_clock::time_point g_trace_t0;
static bool g_trace = false;
static void trace_point(const char* what, int pi = -1) {
    if (!g_trace) return;
    const double host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_trace_t0).count();
    cudaStreamSynchronize(E.stream);
    const double dev = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_trace_t0).count();
    fprintf(stderr, "[rq] t=%8.3f ms (stream idle at %8.3f)  %s %d\n", host, dev, what, pi);
}

// ---- event pool -------------------------------------------------------------------------
struct EventPair { cudaEvent_t a, b; };
static std::vector<EventPair> g_event_pool;
static EventPair& event_pair(size_t i) {
    while (g_event_pool.size() <= i) {
        EventPair p;
        CK(cudaEventCreate(&p.a));
        CK(cudaEventCreate(&p.b));
        g_event_pool.push_back(p);
    }
    return g_event_pool[i];
}

static std::unique_ptr<rq_table> new_intermediate(int n_cols, int64_t cap_rows) {
    std::unique_ptr<rq_table> t(new rq_table());
    t->n_rows = -1;
    t->cap_rows = round_up(std::max<int64_t>(cap_rows, 1), kPadRows);
    for (int c = 0; c < n_cols; c++) {
        DevColumn dc;
        dc.type = RQ_I64;
        dc.width = 8;
        dc.tile_stride = (int64_t)kTile * 8;
        CK(dmalloc(&dc.d, (size_t)t->cap_rows * 8));
        t->cols.push_back(dc);
    }
    CK(dmalloc(&t->d_n_rows, 8));
    CK(cudaMemsetAsync(t->d_n_rows, 0, 8, E.stream));
    return t;
}

static void set_types(rq_table& t, const rq_pipeline& pl) {
    t.sql_type.clear();
    t.sql_width.clear();
    for (int k = 0; k < pl.n_keys; k++) { t.sql_type.push_back(pl.keys[k].sql_type); t.sql_width.push_back(pl.keys[k].width); }
    for (int k = 0; k < pl.n_vals; k++) { t.sql_type.push_back(pl.vals[k].sql_type); t.sql_width.push_back(pl.vals[k].width); }
}

static void record_event(cudaEvent_t ev);
static int max_warps_of(int gr) { return (gr == 4 ? ScanCfg<4>::kThreads : gr == 1 ? ScanCfg<1>::kThreads : ScanCfg<0>::kThreads) / 32; }

static void launch_pipeline(const KParams& P, int gr, int64_t rows, rq_timings* tm, bool is_scan,
                            size_t& ev_idx, std::vector<std::pair<size_t, int>>& ev_used) {
    const int64_t tiles = std::max<int64_t>(1, (rows + kTile - 1) / kTile);
    const int W = P.warps;
    const int grid = (int)std::min<int64_t>((tiles + W - 1) / W, (int64_t)E.sm_count);
    if (tiles + (int64_t)grid * W >= ((int64_t)1 << 32))
        raise(RQ_ERR_UNSUPPORTED, "a pipeline over more than 2^40 rows (the kernel counts tiles in 32 bits)");
    EventPair& ep = event_pair(ev_idx);
    record_event(ep.a);
    if (gr == 4) rq_scan_kernel<4><<<grid, 32 * W, P.smem_bytes, E.stream>>>(P);
    else if (gr == 1) rq_scan_kernel<1><<<grid, 32 * W, P.smem_bytes, E.stream>>>(P);
    else rq_scan_kernel<0><<<grid, 32 * W, P.smem_bytes, E.stream>>>(P);
    CK(cudaGetLastError());
    record_event(ep.b);
    ev_used.push_back({ev_idx, is_scan ? 1 : 0});
    ev_idx++;
    if (tm) tm->kernel_launches++;
}

// hash-table capacities that worked, per pipeline shape and input size
static std::map<uint64_t, uint64_t> g_ht_capacity;
static uint64_t pipeline_signature(const rq_pipeline& pl, int64_t rows) {
    uint64_t h = 0xcbf29ce484222325ULL;
    auto mixin = [&](const void* p, size_t n) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ULL; }
    };
    mixin(&rows, sizeof rows);
    mixin(&pl.sink_kind, sizeof pl.sink_kind);
    mixin(&pl.source_kind, sizeof pl.source_kind);
    mixin(&pl.source_id, sizeof pl.source_id);
    for (int i = 0; i < pl.n_nodes; i++) mixin(&pl.nodes[i], sizeof(rq_node));
    for (int i = 0; i < pl.n_keys; i++) mixin(&pl.keys[i], sizeof(rq_value));
    for (int i = 0; i < pl.n_vals; i++) mixin(&pl.vals[i], sizeof(rq_value));
    return h;
}

static bool has_str_key(const rq_pipeline& pl) {
    for (int k = 0; k < pl.n_keys; k++)
        if (pl.keys[k].sql_type == RQ_SQL_VARCHAR || (pl.keys[k].sql_type == RQ_SQL_CHAR && pl.keys[k].width > 1)) return true;
    return false;
}

// thrown when a multi-match probe met a tuple with several matches: the pipeline is rerun in its
// expanded two-pass form (split_at_probe, mode SPLIT_EXPAND)
struct NeedExpand {};

// ---- host reads of device values: recorded once, predicted afterwards ---------------------------
// Executing a plan makes the host look at device values now and then: overflow / table-full /
// error flags, the number of groups or materialized rows, the row counts of the other ranks. Every
// such read goes through host_read(). The first clean execution of a plan on given tables RECORDS
// the sequence of values it saw. Later executions REPLAY: host_read() returns the recorded value at
// once (no stream synchronisation) and only enqueues a copy of the actual device value into a log;
// the whole plan - all pipelines, NCCL exchanges, ORDER BY, result read-back - is enqueued without
// a single host wait. At the end one small kernel compares the log with the recorded values and the
// verdict comes back with the result (sharded plans: all-reduced, so every rank takes the same
// decision). Any difference - other data under the same table name, a table that filled up, a
// division by zero - discards the result and runs the plan again the careful way, which reports
// errors exactly as before. Device code never trusts a predicted value: outputs and hash tables are
// bounds-checked against their allocation, so a wrong prediction can only produce a discarded result.
// A replayed plan is a fixed sequence of launches with fixed arguments (table identities are part
// of the memo key), so the second replay is recorded into a CUDA graph: later executions are ONE
// graph launch and one wait - no per-launch host cost, which is what bounds short plans and the
// 8-GPU runs, where the kernels of a shard take fractions of a millisecond.
// a result and the one host block its columns live in (rq_result_col::data point into it)
struct ResultBox {
    rq_result r;
    unsigned char* block;
};

struct PlanMemo {
    std::vector<std::vector<unsigned char>> reads;
    bool valid = false;
    int replays = 0;                   // plain replays that succeeded
    bool no_graph = false;             // capture failed once: stay with plain replays
    // buffers that live as long as the memo (so that a captured graph may refer to them)
    char* d_strpool = nullptr;
    unsigned char* d_expect = nullptr;
    unsigned char* d_log = nullptr;
    int32_t* d_ok = nullptr;
    size_t log_cap = 0;
    // the graph and what is needed to hand out its result
    cudaGraphExec_t gexec = nullptr;
    struct Col { int type = 0, width = 0, sql_type = 0, sql_width = 0; size_t off = 0; };
    std::vector<Col> cols;             // result columns inside the packed result
    unsigned char* h_pack = nullptr;   // pinned landing buffer of the packed result
    size_t pack_bytes = 0;
    int64_t res_rows = 0;
    std::vector<std::pair<size_t, int>> ev_used;
    rq_timings tm_static{};            // launches / lowering time of the captured run
    void release() {
        if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        if (h_pack) cudaFreeHost(h_pack);
        h_pack = nullptr; pack_bytes = 0;
        cols.clear();
        if (d_strpool) cudaFree(d_strpool);
        if (d_expect) cudaFree(d_expect);
        if (d_log) cudaFree(d_log);
        if (d_ok) cudaFree(d_ok);
        d_strpool = nullptr; d_expect = nullptr; d_log = nullptr; d_ok = nullptr;
        log_cap = 0; replays = 0; valid = false; reads.clear();
    }
};
static std::map<uint64_t, PlanMemo> g_plan_memo;
struct ReplayState {
    int mode = 0;                      // 0 careful, 1 careful + recording, 2 replaying
    PlanMemo* memo = nullptr;
    size_t idx = 0;
    unsigned char* d_log = nullptr;    // replay: what the host would have read, in order
    size_t log_bytes = 0, log_cap = 0;
    int retries = 0;                   // recording: a retried pipeline makes the run unfit as a script
    int syncs = 0;                     // host waits of this execution (reported through rq_timings)
    const char* why = "";             // last retry reason (trace)
    bool capturing = false;            // the stream is being captured into a CUDA graph
};
static ReplayState RP;
struct ReplayDiverged {};
// (event records inside a stream capture must be external nodes to stay usable for timing)
static void record_event(cudaEvent_t ev);
#define g_pinned (E.pinned)
static constexpr size_t kPinnedRead = 36864;

static void record_event(cudaEvent_t ev) {
    CK(cudaEventRecordWithFlags(ev, E.stream, RP.capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
}

static void stream_sync() {
    CK(cudaStreamSynchronize(E.stream));
    RP.syncs++;
}

static void host_read(void* dst, const void* d_src, size_t bytes) {
    if (bytes > kPinnedRead) raise(RQ_ERR_INVALID, "internal: host_read of %zu bytes", bytes);
    if (RP.mode == 2) {
        if (RP.idx >= RP.memo->reads.size() || RP.memo->reads[RP.idx].size() != bytes) throw ReplayDiverged{};
        memcpy(dst, RP.memo->reads[RP.idx].data(), bytes);
        const size_t padded = (bytes + 7) & ~(size_t)7;
        if (RP.log_bytes + padded > RP.log_cap) throw ReplayDiverged{};
        CK(cudaMemcpyAsync(RP.d_log + RP.log_bytes, d_src, bytes, cudaMemcpyDeviceToDevice, E.stream));
        RP.log_bytes += padded;
        RP.idx++;
        return;
    }
    CK(cudaMemcpyAsync(g_pinned, d_src, bytes, cudaMemcpyDeviceToHost, E.stream));
    stream_sync();
    memcpy(dst, g_pinned, bytes);
    if (RP.mode == 1) RP.memo->reads.emplace_back((const unsigned char*)dst, (const unsigned char*)dst + bytes);
}

// small host value -> device without a host wait and without a host buffer: the bytes travel as
// kernel arguments
static void upload_small(void* d_dst, const void* src, size_t bytes) {
    if (bytes > sizeof(SmallBytes)) raise(RQ_ERR_INVALID, "internal: upload_small of %zu bytes", bytes);
    SmallBytes v;
    memset(&v, 0, sizeof(v));
    memcpy(v.b, src, bytes);
    rq_store_bytes<<<1, 64, 0, E.stream>>>((unsigned char*)d_dst, v, (int)bytes);
    CK(cudaGetLastError());
}

static void check_flags(const char* what, const void* d_extra = nullptr) {
    if (d_extra) {        // one round trip for both: the extra word is parked behind the flags
        CK(cudaMemcpyAsync(E.flags + 8, d_extra, 8, cudaMemcpyDeviceToDevice, E.stream));
        host_read(E.h_flags, E.flags, 40);
    } else {
        host_read(E.h_flags, E.flags, 32);
    }
    if (E.h_flags[2] == 2) raise(RQ_ERR_INVALID, "internal: %s was lowered for a kernel variant without probe support", what);
    if (E.h_flags[2]) raise(RQ_ERR_RUNTIME, "division by zero in %s (the reference raises SIGFPE here)", what);
    if (*(unsigned long long*)(E.h_flags + 4) != 0) throw NeedExpand{};
}

// ---- tile ranges (KParams::tile_range) ----------------------------------------------------------
// (1) Zone skipping. A selection `col CMP const` / `lo <= col <= hi` on a source column whose values
// are non-decreasing over the rows (DevColumn::sorted, taken with the upload statistics: o_orderkey,
// l_orderkey, any surrogate key of a table loaded in key order) can only pass inside one row range.
// Two binary searches on the device (rq_sorted_tile_range) turn it into the tile range the scan
// visits; the selection itself stays in the program, so the result is exact. This is what makes a
// pruned build (prune_build_by_probe_stats) cheap on a sharded plan: a rank whose lineitem shard
// covers 1/8 of the order keys reads 1/8 of orders instead of filtering all of it.
// (2) Shared builds. A join build over a table that is complete on every rank and whose program is
// the same on every rank is split: rank r scans tiles [T*r/W, T*(r+1)/W) and the direct-address
// tables (bitmap; disjoint bits because the keys are unique) are summed with one ncclAllReduce.
struct SharedBuild { bool active = false; };
static SharedBuild g_shared_build;       // set by run_pipeline for the pipeline it is about to run
struct TileRangeBuf {
    uint32_t* d = nullptr;
    ~TileRangeBuf() { if (d) dfree(d); }     // (stream-ordered: after the launches that read it)
};
static constexpr int64_t kZoneSkipMinRows = 1 << 18;

static void restrict_tiles(const Lowerer& L, KParams& P, const rq_table& src, TileRangeBuf& buf, bool share, rq_timings* tm) {
    if (src.n_rows < 0 || src.n_rows < kZoneSkipMinRows) { if (!share) return; }
    if (src.n_rows < 0) return;
    const uint32_t n_tiles = (uint32_t)((src.n_rows + kTile - 1) / kTile);
    uint32_t init[2] = {0u, n_tiles};
    if (share) {
        const uint64_t W = (uint64_t)E.dist.world, r = (uint64_t)E.dist.rank;
        init[0] = (uint32_t)((uint64_t)n_tiles * r / W);
        init[1] = (uint32_t)((uint64_t)n_tiles * (r + 1) / W);
    }
    struct Zone { int col; int64_t lo, hi; };
    std::vector<Zone> zones;
    std::vector<int> src_of(P.n_cols, -1);
    for (size_t k = 0; k < L.staged_of_col.size(); k++)
        if (src.cols[k].type != RQ_STR && L.staged_of_col[k] >= 0) src_of[L.staged_of_col[k]] = (int)k;
    for (const HUnit& u : L.prog) {
        if ((u.op != H_FCMP && u.op != H_FRANGE) || u.x.kind != S_COL || u.x.idx >= (uint16_t)P.n_cols) continue;
        const int sc = src_of[u.x.idx];
        if (sc < 0 || !src.cols[sc].sorted) continue;
        int64_t lo = INT64_MIN, hi = INT64_MAX;
        if (u.op == H_FRANGE) {
            lo = u.imm;
            const __int128 h = (__int128)u.imm + (__int128)(uint64_t)u.imm2;
            hi = h > (__int128)INT64_MAX ? INT64_MAX : (int64_t)h;
        } else if (u.gop == D_GE) lo = u.imm;
        else if (u.gop == D_GT) { if (u.imm == INT64_MAX) continue; lo = u.imm + 1; }
        else if (u.gop == D_LE) hi = u.imm;
        else if (u.gop == D_LT) { if (u.imm == INT64_MIN) continue; hi = u.imm - 1; }
        else if (u.gop == D_EQ) { lo = u.imm; hi = u.imm; }
        else continue;
        zones.push_back({sc, lo, hi});
    }
    if (zones.empty() && !share) return;
    CK(dmalloc(&buf.d, 8));
    upload_small(buf.d, init, 8);
    for (const Zone& z : zones) {
        const DevColumn& dc = src.cols[z.col];
        rq_sorted_tile_range<<<1, 32, 0, E.stream>>>(dc.d, dc.width, dc.tile_stride ? dc.tile_stride : (int64_t)kTile * dc.width,
                                                      src.n_rows, z.lo, z.hi, buf.d);
        if (tm) tm->kernel_launches++;
    }
    CK(cudaGetLastError());
    P.tile_range = buf.d;
}

// ---- one pipeline -------------------------------------------------------------------------
struct SplitPipes {
    std::vector<rq_node> a_nodes, b_nodes;
    std::vector<int32_t> a_args, b_args;
    std::vector<rq_value> a_vals, b_keys, b_vals;
    std::vector<int> live_src_col;     // per materialized value: source column it copies, or -1
    rq_pipeline a, b;
};
enum SplitMode { SPLIT_SEMI = 0, SPLIT_EXPAND = 1, SPLIT_EXCHANGE = 2 };
static bool split_at_probe(const rq_plan& plan, const rq_pipeline& pl, const rq_table& src,
                           const std::vector<PipeOut>& outs, SplitPipes& sp, SplitMode mode, int ordinal = 0);
static std::map<uint64_t, int64_t> g_emit_rows;    // rows a materialize pipeline produced last time

static void run_pipeline_one(const rq_plan& plan, const rq_pipeline& pl_in, int pi, std::vector<PipeOut>& outs,
                             const char* d_strpool, rq_timings* tm, size_t& ev_idx,
                             std::vector<std::pair<size_t, int>>& ev_used, double& lower_ms,
                             bool is_fact_scan, const rq_table* src_override, PipeOut& result, bool allow_split);

// ---- build-side pruning from probe-side statistics --------------------------------------------
// An inner equi-join can only match build tuples whose key lies inside the value range of the
// probe key. When every pipeline that probes build pipeline `pi` does so with a plain column whose
// min/max were taken at upload, `lo <= key <= hi` is added to the build pipeline right behind the
// key value (two selections the lowering fuses into one range compare). On one GPU the range is
// usually the whole key domain; with a row-range sharded fact table that is clustered on the join
// key (lineitem on l_orderkey) each rank builds only the slice of the build side it can ever probe,
// so the build work shards with the fact table although the build table itself is replicated.
struct PrunedBuild {
    std::vector<rq_node> nodes;
    std::vector<int32_t> args;
    std::vector<rq_value> keys, vals;
    rq_pipeline pl;
};

// Which payload columns of build pipeline `pi` does the rest of the plan read, and where do they
// live in a hash-table entry? A payload that repeats a join key (the reference stores the join
// keys AND all requested build attributes, hashjoin.h:233-236) is read from the key word; a payload
// no later PAYLOAD node reads is not stored at all. Returns the build pipeline with the stored
// payloads only; pay_word[q] = entry word of the plan's payload q (0xff = not stored).
struct BuildLayout {
    std::vector<rq_value> vals;
    std::vector<uint8_t> pay_word;
    std::vector<int> sql_type, sql_width;
    rq_pipeline pl;
};
static void layout_build_payload(const rq_plan& plan, int pi, const rq_pipeline& b, BuildLayout& out) {
    const int nv = b.n_vals, nk = b.n_keys;
    std::vector<char> needed(nv, 0);
    for (int q = pi + 1; q < plan.n_pipelines; q++) {
        const rq_pipeline& pq = plan.pipelines[q];
        std::vector<char> used(pq.n_nodes, 0);
        auto use = [&](int r) { if (r >= 0 && r < pq.n_nodes) used[r] = 1; };
        for (int i = 0; i < pq.n_nodes; i++) {
            const rq_node& nd = pq.nodes[i];
            if (is_binary(nd.op)) { use(nd.a); use(nd.b); }
            else if (nd.op == RQ_OP_FILTER) use(nd.a);
            else if (nd.op == RQ_OP_SELECT) { use(nd.a); use(nd.b); use(nd.c); }
            else if (nd.op == RQ_OP_PROBE) for (int k = 0; k < nd.c && nd.b + k < pq.n_args; k++) use(pq.args[nd.b + k]);
        }
        for (int k = 0; k < pq.n_keys; k++) use(pq.keys[k].node);
        for (int k = 0; k < pq.n_vals; k++) use(pq.vals[k].node);
        for (int i = 0; i < pq.n_nodes; i++) {
            const rq_node& nd = pq.nodes[i];
            if (nd.op != RQ_OP_PAYLOAD || !used[i]) continue;
            if (nd.a < 0 || nd.a >= pq.n_nodes || pq.nodes[nd.a].op != RQ_OP_PROBE || pq.nodes[nd.a].a != pi) continue;
            if (nd.b >= 0 && nd.b < nv) needed[nd.b] = 1;
        }
    }
    out.pay_word.assign(nv, 0xff);
    for (int q = 0; q < nv; q++) {
        out.sql_type.push_back(b.vals[q].sql_type);
        out.sql_width.push_back(b.vals[q].width);
        if (!needed[q]) continue;
        int alias = -1;
        for (int k = 0; k < nk; k++) if (b.keys[k].node == b.vals[q].node) alias = k;
        if (alias >= 0) { out.pay_word[q] = (uint8_t)(1 + alias); continue; }
        out.pay_word[q] = (uint8_t)(1 + nk + out.vals.size());
        out.vals.push_back(b.vals[q]);
    }
    out.pl = b;
    out.pl.n_vals = (int)out.vals.size();
    out.pl.vals = out.vals.data();
}
static bool prune_build_by_probe_stats(const rq_plan& plan, int pi, PrunedBuild& out) {
    if (!E.opt.prune_builds) return false;
    const rq_pipeline& b = plan.pipelines[pi];
    const int nk = b.n_keys;
    if (nk < 1 || nk > kMaxKeys) return false;
    std::vector<int64_t> lo(nk, INT64_MAX), hi(nk, INT64_MIN);
    std::vector<char> ok(nk, 1);
    int probers = 0;
    for (int q = pi + 1; q < plan.n_pipelines; q++) {
        const rq_pipeline& pq = plan.pipelines[q];
        for (int i = 0; i < pq.n_nodes; i++) {
            const rq_node& nd = pq.nodes[i];
            if (nd.op != RQ_OP_PROBE || nd.a != pi) continue;
            probers++;
            if (nd.c != nk || nd.b < 0 || nd.b + nd.c > pq.n_args) return false;
            const rq_table* t = (pq.source_kind == RQ_SRC_TABLE && pq.source_id >= 0 && pq.source_id < plan.n_tables)
                                    ? plan.tables[pq.source_id] : nullptr;
            for (int k = 0; k < nk; k++) {
                const int a = pq.args[nd.b + k];
                const rq_node* kn = (a >= 0 && a < pq.n_nodes) ? &pq.nodes[a] : nullptr;
                if (!t || !kn || kn->op != RQ_OP_COL || kn->a < 0 || kn->a >= (int)t->cols.size() ||
                    !t->cols[kn->a].has_stats || t->cols[kn->a].type == RQ_STR) { ok[k] = 0; continue; }
                lo[k] = std::min(lo[k], t->cols[kn->a].vmin);
                hi[k] = std::max(hi[k], t->cols[kn->a].vmax);
            }
        }
    }
    if (probers == 0) return false;
    bool any = false;
    for (int k = 0; k < nk; k++) {
        const int st = b.keys[k].sql_type;
        if (st == RQ_SQL_VARCHAR || st == RQ_SQL_CHAR) ok[k] = 0;     // CHAR(1) included: keep it simple
        if (ok[k] && lo[k] <= hi[k]) any = true; else ok[k] = 0;
    }
    if (!any) return false;
    out.nodes.assign(b.nodes, b.nodes + b.n_nodes);
    out.args.assign(b.args, b.args + b.n_args);
    out.keys.assign(b.keys, b.keys + b.n_keys);
    out.vals.assign(b.vals, b.vals + b.n_vals);
    for (int k = 0; k < nk; k++) {
        if (!ok[k]) continue;
        const int pos = out.keys[k].node;                 // insert the six nodes right behind it
        if (pos < 0 || pos >= (int)out.nodes.size()) return false;
        const int K = 6;
        auto mv = [&](int r) { return r > pos ? r + K : r; };
        std::vector<rq_node> nn;
        nn.reserve(out.nodes.size() + K);
        for (int i = 0; i <= pos; i++) nn.push_back(out.nodes[i]);
        nn.push_back(rq_node{RQ_OP_CONST, 0, 0, 0, lo[k]});
        nn.push_back(rq_node{RQ_OP_GE, pos, pos + 1, 0, 0});
        nn.push_back(rq_node{RQ_OP_FILTER, pos + 2, 0, 0, 0});
        nn.push_back(rq_node{RQ_OP_CONST, 0, 0, 0, hi[k]});
        nn.push_back(rq_node{RQ_OP_LE, pos, pos + 4, 0, 0});
        nn.push_back(rq_node{RQ_OP_FILTER, pos + 5, 0, 0, 0});
        for (int i = pos + 1; i < (int)out.nodes.size(); i++) {
            rq_node nd = out.nodes[i];
            if (is_binary(nd.op)) { nd.a = mv(nd.a); nd.b = mv(nd.b); }
            else if (nd.op == RQ_OP_FILTER || nd.op == RQ_OP_PAYLOAD) nd.a = mv(nd.a);
            else if (nd.op == RQ_OP_SELECT) { nd.a = mv(nd.a); nd.b = mv(nd.b); nd.c = mv(nd.c); }
            nn.push_back(nd);
        }
        for (auto& a : out.args) a = mv(a);
        for (auto& kk : out.keys) kk.node = mv(kk.node);
        for (auto& v : out.vals) v.node = mv(v.node);
        out.nodes.swap(nn);
    }
    out.pl = b;
    out.pl.n_nodes = (int)out.nodes.size(); out.pl.nodes = out.nodes.data();
    out.pl.n_args = (int)out.args.size(); out.pl.args = out.args.data();
    out.pl.keys = out.keys.data(); out.pl.vals = out.vals.data();
    return true;
}

static bool is_partitioned_plan(const rq_plan& plan);
static void run_pipeline_partitioned(const rq_plan& plan, const rq_pipeline& pl_in, int pi, std::vector<PipeOut>& outs,
                                     const char* d_strpool, rq_timings* tm, size_t& ev_idx,
                                     std::vector<std::pair<size_t, int>>& ev_used, double& lower_ms,
                                     const rq_table* src_override, bool first_probe_aligned, PipeOut& result);

// May build pipeline `pi` of a sharded plan be split over the ranks ("shared builds" above)? Every rank
// must take the same decisions and meet the same errors up to and including this pipeline, so all of
// pipelines 0..pi must be rank-invariant: scans of tables that are complete on every rank (not the
// fact table, which is the one the merged pipeline scans), probing only rank-invariant builds, and
// pruned (prune_build_by_probe_stats) only by statistics of complete tables. The pipeline itself must
// be a plain scan of a table large enough to be worth a collective.
static bool build_is_shareable(const rq_plan& plan, int pi) {
    if (!E.opt.share_builds || !(plan.flags & RQ_PLAN_SHARDED) || (plan.flags & RQ_PLAN_PARTITIONED)) return false;
    if (!E.dist.comm || E.dist.world <= 1) return false;
    int merge = plan.n_pipelines - 1;
    for (int q = 0; q < plan.n_pipelines; q++) if (plan.pipelines[q].sink_kind == RQ_SINK_AGG) merge = q;
    const rq_pipeline& mp = plan.pipelines[merge];
    const int fact = mp.source_kind == RQ_SRC_TABLE ? mp.source_id : -1;
    if (fact < 0) return false;         // the sharded table is not scanned directly: nothing is known
    for (int q = 0; q <= pi; q++) {
        const rq_pipeline& p = plan.pipelines[q];
        if (p.source_kind != RQ_SRC_TABLE || p.source_id == fact || p.source_id < 0 || p.source_id >= plan.n_tables) return false;
        if (p.sink_kind != RQ_SINK_BUILD) return false;
        for (int i = 0; i < p.n_nodes; i++)
            if (p.nodes[i].op == RQ_OP_PROBE && (q == pi || p.nodes[i].a < 0 || p.nodes[i].a >= q)) return false;
        for (int r = q + 1; r < plan.n_pipelines; r++) {
            const rq_pipeline& pr = plan.pipelines[r];
            bool probes_q = false;
            for (int i = 0; i < pr.n_nodes; i++) probes_q |= pr.nodes[i].op == RQ_OP_PROBE && pr.nodes[i].a == q;
            if (probes_q && pr.source_kind == RQ_SRC_TABLE && pr.source_id == fact) return false;
        }
    }
    const rq_table* t = plan.tables[plan.pipelines[pi].source_id];
    return t && t->n_rows >= E.opt.share_min_rows;
}

static void run_pipeline(const rq_plan& plan, int pi, std::vector<PipeOut>& outs,
                         const char* d_strpool, rq_timings* tm, size_t& ev_idx,
                         std::vector<std::pair<size_t, int>>& ev_used, double& lower_ms,
                         bool is_fact_scan = true) {
    PrunedBuild pb;
    BuildLayout bl;
    const rq_pipeline* pl = &plan.pipelines[pi];
    const bool part = is_partitioned_plan(plan);
    if (pl->sink_kind == RQ_SINK_BUILD) {
        // (the probe-side statistics of ONE shard say nothing about the keys other ranks probe with)
        if (!part && prune_build_by_probe_stats(plan, pi, pb)) pl = &pb.pl;
        layout_build_payload(plan, pi, *pl, bl);
        pl = &bl.pl;
    }
    g_shared_build.active = !part && pl->sink_kind == RQ_SINK_BUILD && build_is_shareable(plan, pi);
    struct Reset { ~Reset() { g_shared_build.active = false; } } reset_share;
    if (part)
        run_pipeline_partitioned(plan, *pl, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, nullptr, false, outs[pi]);
    else
        run_pipeline_one(plan, *pl, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, is_fact_scan,
                         nullptr, outs[pi], true);
    if (pl->sink_kind == RQ_SINK_BUILD) {
        outs[pi].pay_word = bl.pay_word;
        outs[pi].payload_sql_type = bl.sql_type;
        outs[pi].payload_sql_width = bl.sql_width;
    }
}

static void run_pipeline_impl(const rq_plan& plan, const rq_pipeline& pl_in, int pi, std::vector<PipeOut>& outs,
                              const char* d_strpool, rq_timings* tm, size_t& ev_idx,
                              std::vector<std::pair<size_t, int>>& ev_used, double& lower_ms,
                              bool is_fact_scan, const rq_table* src_override, PipeOut& result, bool allow_split,
                              bool expand);

// pipelines that met duplicate matches once are run in expanded form straight away afterwards
static std::set<uint64_t> g_needs_expand;
// aggregation pipelines: the implementation (register / shared-memory / hash) that did not overflow
static std::map<uint64_t, int> g_agg_impl;
// join builds that met duplicate keys: hash form, not direct-address
static std::set<uint64_t> g_no_direct;

static void run_pipeline_one(const rq_plan& plan, const rq_pipeline& pl_in, int pi, std::vector<PipeOut>& outs,
                             const char* d_strpool, rq_timings* tm, size_t& ev_idx,
                             std::vector<std::pair<size_t, int>>& ev_used, double& lower_ms,
                             bool is_fact_scan, const rq_table* src_override, PipeOut& result, bool allow_split) {
    const uint64_t sig = pipeline_signature(pl_in, src_override ? -2 : -1) ^ 0x9e3779b97f4a7c15ULL;
    if (!g_needs_expand.count(sig)) {
        try {
            run_pipeline_impl(plan, pl_in, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, is_fact_scan,
                              src_override, result, allow_split, false);
            return;
        } catch (NeedExpand&) {
            cudaStreamSynchronize(E.stream);
            RP.retries++; RP.why = "multi-match expansion";
            result = PipeOut();
            g_needs_expand.insert(sig);
        }
    }
    run_pipeline_impl(plan, pl_in, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, is_fact_scan,
                      src_override, result, allow_split, true);
}

static void run_pipeline_impl(const rq_plan& plan, const rq_pipeline& pl_in, int pi, std::vector<PipeOut>& outs,
                              const char* d_strpool, rq_timings* tm, size_t& ev_idx,
                              std::vector<std::pair<size_t, int>>& ev_used, double& lower_ms,
                              bool is_fact_scan, const rq_table* src_override, PipeOut& result, bool allow_split,
                              bool expand) {
    SimplePipe sp;
    simplify_pipeline(pl_in, sp);
    const rq_pipeline& pl = sp.pl;
    const rq_table* src = nullptr;
    std::unique_ptr<rq_table> cross_holder;
    if (src_override) {
        src = src_override;
    } else if (pl.source_kind == RQ_SRC_TABLE) {
        if (pl.source_id < 0 || pl.source_id >= plan.n_tables || !plan.tables[pl.source_id])
            raise(RQ_ERR_INVALID, "pipeline %d: table %d out of range", pi, pl.source_id);
        src = plan.tables[pl.source_id];
    } else if (pl.source_kind == RQ_SRC_CROSS) {
        // NestedLoopsJoinOp: materialize the cross product of two earlier relations, then run the
        // pipeline (condition, projection, sink) over it like over any other relation
        const int a = pl.source_id, b = pl_in.source_id2;
        if (a < 0 || a >= pi || b < 0 || b >= pi || !outs[a].table || !outs[b].table)
            raise(RQ_ERR_INVALID, "pipeline %d: cross product needs two earlier relation outputs", pi);
        const rq_table& L = *outs[a].table;
        const rq_table& R = *outs[b].table;
        int64_t nl = L.n_rows, nr = R.n_rows;
        if (nl < 0) host_read(&nl, L.d_n_rows, 8);
        if (nr < 0) host_read(&nr, R.d_n_rows, 8);
        const int ncols = (int)(L.cols.size() + R.cols.size());
        if (ncols > kMaxStagedCols) raise(RQ_ERR_UNSUPPORTED, "nested-loops join over more than %d columns", kMaxStagedCols);
        if (nl > 0 && nr > ((int64_t)1 << 28) / nl)
            raise(RQ_ERR_UNSUPPORTED, "nested-loops join of %lld x %lld tuples exceeds the materialization limit (2^28 pairs)",
                  (long long)nl, (long long)nr);
        cross_holder = new_intermediate(ncols, nl * nr);
        CrossCols cc;
        memset(&cc, 0, sizeof(cc));
        cc.n_left = (int32_t)L.cols.size(); cc.n_right = (int32_t)R.cols.size();
        for (int c = 0; c < cc.n_left; c++) cc.in[c] = (const int64_t*)L.cols[c].d;
        for (int c = 0; c < cc.n_right; c++) cc.in[cc.n_left + c] = (const int64_t*)R.cols[c].d;
        for (int c = 0; c < ncols; c++) cc.out[c] = (int64_t*)cross_holder->cols[c].d;
        const int64_t total = nl * nr;
        if (total > 0) {
            rq_cross_product<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, E.stream>>>(cc, nl, nr);
            if (tm) tm->kernel_launches++;
            CK(cudaGetLastError());
        }
        upload_small(cross_holder->d_n_rows, &total, 8);
        cross_holder->n_rows = total;
        src = cross_holder.get();
    } else if (pl.source_kind == RQ_SRC_PIPELINE) {
        if (pl.source_id < 0 || pl.source_id >= pi || !outs[pl.source_id].table)
            raise(RQ_ERR_INVALID, "pipeline %d: source pipeline %d has no relation output", pi, pl.source_id);
        src = outs[pl.source_id].table.get();
    } else if (pl.source_kind == RQ_SRC_ONE_ROW) {
        // leaf projection (projection.h:49-58): one tuple, no attributes
        cross_holder.reset(new rq_table());
        cross_holder->n_rows = 1;
        cross_holder->cap_rows = kPadRows;
        src = cross_holder.get();
    } else {
        raise(RQ_ERR_INVALID, "pipeline %d: bad source kind %d", pi, pl.source_kind);
    }
    const bool is_scan = pl.source_kind == RQ_SRC_TABLE && is_fact_scan && !src_override;

    // A selective hash-join probe inside a big scan is run as two passes: scan -> filter -> Bloom
    // test -> materialize the few survivors, then probe/aggregate/build over dense tiles of them.
    if (expand) {
        SplitPipes sx;
        if (!split_at_probe(plan, pl, *src, outs, sx, SPLIT_EXPAND))
            raise(RQ_ERR_INVALID, "pipeline %d: duplicate matches reported but no multi-match probe found", pi);
        PipeOut mid;
        run_pipeline_one(plan, sx.a, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, is_fact_scan, src_override, mid, false);
        run_pipeline_one(plan, sx.b, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, false, mid.table.get(), result, false);
        return;
    }
    if (allow_split && !src_override && pl.source_kind == RQ_SRC_TABLE) {
        SplitPipes sx;
        if (split_at_probe(plan, pl, *src, outs, sx, SPLIT_SEMI)) {
            PipeOut mid;
            run_pipeline_one(plan, sx.a, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, is_fact_scan, nullptr, mid, false);
            for (size_t c = 0; c < sx.live_src_col.size(); c++) {      // value bounds survive the copy
                const int sc = sx.live_src_col[c];
                if (sc >= 0 && src->cols[sc].has_stats) {
                    mid.table->cols[c].has_stats = true;
                    mid.table->cols[c].vmin = src->cols[sc].vmin;
                    mid.table->cols[c].vmax = src->cols[sc].vmax;
                }
            }
            run_pipeline_one(plan, sx.b, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, false, mid.table.get(), result, false);
            return;
        }
    }

    auto t0 = std::chrono::steady_clock::now();
    std::vector<int> impls;
    AggDedup ad;
    if (pl.sink_kind == RQ_SINK_AGG) {
        ad = dedup_aggs(pl);
        if (!has_str_key(pl)) {
            // the register-aggregation kernels (GR > 0) carry no probe code: pipelines that join
            // take the shared-memory accumulator path of the generic kernel
            bool has_probe = false;
            for (int i = 0; i < pl.n_nodes; i++) has_probe |= pl.nodes[i].op == RQ_OP_PROBE;
            if ((int)ad.kind.size() <= kNAR && !has_probe) impls.push_back(IMPL_REGAGG);
            impls.push_back(IMPL_LOWAGG);
        }
        impls.push_back(IMPL_HASHAGG);
    } else if (pl.sink_kind == RQ_SINK_MATERIALIZE) {
        impls.push_back(IMPL_EMIT);
    } else if (pl.sink_kind == RQ_SINK_BUILD) {
        impls.push_back(IMPL_BUILD);
    } else {
        raise(RQ_ERR_INVALID, "pipeline %d: bad sink kind %d", pi, pl.sink_kind);
    }
    const int64_t src_rows = src->n_rows >= 0 ? src->n_rows : src->cap_rows;
    // the aggregation path that worked for this pipeline last time is tried first
    const uint64_t impl_sig = pipeline_signature(pl_in, src_override ? -2 : -1) ^ 0x696d706cULL;
    size_t first_attempt = 0;
    {
        auto known = g_agg_impl.find(impl_sig);
        if (known != g_agg_impl.end())
            for (size_t k = 0; k < impls.size(); k++) if (impls[k] == known->second) first_attempt = k;
    }

    for (size_t attempt = first_attempt; attempt < impls.size(); attempt++) {
        const int impl = impls[attempt];
        g_agg_impl[impl_sig] = impl;
        KParams P;
        memset(&P, 0, sizeof(P));
        P.expand_probe = -1;
        P.n_rows = src->n_rows;
        P.n_rows_ptr = src->n_rows < 0 ? src->d_n_rows : nullptr;
        P.n_rows_cap = src->cap_rows;
        P.borrowed = src->borrowed ? 1 : 0;
        P.stream_hint = 1;
        P.overflow = E.flags + 0;
        P.ht_full = E.flags + 1;
        P.err = E.flags + 2;
        P.ht_entries = (unsigned long long*)(E.flags + 6);
        Lowerer L(plan, pl, *src, outs, d_strpool, P);
        L.prepare();
        emit_program(L, impl, ad);

        std::unique_ptr<rq_table> out;
        int gr = 0;
        KeyUnpack ku;
        memset(&ku, 0, sizeof(ku));
        if (impl == IMPL_REGAGG || impl == IMPL_LOWAGG) {
            P.na = (int)ad.kind.size();
            for (int u = 0; u < P.na; u++) P.agg_kind[u] = (uint8_t)ad.kind[u];
            P.g_state = E.g_state; P.g_keys = E.g_keys; P.g_acc = E.g_acc;
            if (!pack_group_key(L, P, ku)) continue;     // keys wider than 64 bits: hash aggregation
            if (impl == IMPL_REGAGG) {
                gr = pl.n_keys == 0 ? 1 : kRegGroups;
                P.G = 0;
                if (!layout_smem(P, L.n_slots, 0, max_warps_of(gr))) continue;
                choose_agg_modes(L, P, src_rows);
            } else {
                bool fits = false;
                for (int G = (pl.n_keys == 0 ? 1 : kLowCardMaxGroups); G >= 1 && !fits; G >>= 1) {
                    P.G = G;
                    fits = layout_smem(P, L.n_slots, G * P.na * 32 * 8, max_warps_of(0)) && P.warps >= 8;
                    if (pl.n_keys == 0) break;
                }
                if (!fits) continue;
            }
        } else {
            P.G = 0;
            if (!layout_smem(P, L.n_slots, 0, max_warps_of(0)))
                raise(RQ_ERR_UNSUPPORTED, "pipeline %d does not fit in shared memory", pi);
        }
        encode_program(L, P);
        if (impl == IMPL_REGAGG) {
            for (int u = 0; u < kMaxAggs; u++) {
                uint32_t form = AF_NONE;
                if (u < P.na && P.agg_kind[u] != RQ_AGG_COUNT) {
                    if (P.agg_kind[u] == RQ_AGG_MIN) form = AF_MIN;
                    else if (P.agg_kind[u] == RQ_AGG_MAX) form = AF_MAX;
                    else form = P.agg_mode[u] == AM_P1 ? AF_P1 : P.agg_mode[u] == AM_P2 ? AF_P2 : P.agg_mode[u] == AM_W64 ? AF_W64 : AF_FULL;
                }
                const VRef vr = P.agg_src[u];
                uint32_t d = form;
                if (form != AF_NONE && vr.kind == K_M64) d |= 8u | ((vr.slot & 1) ? 16u : 0u) | (((uint32_t)vr.off16 << 4) << 8);
                P.agg_desc[u] = d;
            }
        }
        P.l2_prefetch = (P.n_cols > 0 && P.n_probes == 0 && impl != IMPL_BUILD && impl != IMPL_HASHAGG) ? 1 : 0;
        lower_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        // zone skipping on sorted source columns (table scans only; a shared build adds its rank share below)
        TileRangeBuf tile_buf;
        if (pl.source_kind == RQ_SRC_TABLE && !src_override && E.opt.zone_skip && gr != 4) restrict_tiles(L, P, *src, tile_buf, false, tm);

        CK(cudaMemsetAsync(E.flags, 0, 32, E.stream));
        if (impl == IMPL_REGAGG || impl == IMPL_LOWAGG) {
            AggKinds gk;
            memcpy(gk.kind, P.agg_kind, kMaxAggs);
            rq_group_table_init<<<(kGroupTableCap + 255) / 256, 256, 0, E.stream>>>(E.g_state, E.g_acc, gk, P.na);
            if (tm) tm->kernel_launches++;
            launch_pipeline(P, gr, src_rows, tm, is_scan, ev_idx, ev_used);
            // dense output: keys, then every requested aggregate (duplicates expanded)
            const int ncols = pl.n_keys + pl.n_vals;
            out = new_intermediate(ncols, kGroupTableCap);
            const int nuniq = P.na;
            std::unique_ptr<rq_table> dense = new_intermediate(pl.n_keys + nuniq, kGroupTableCap);
            std::vector<int64_t*> h_dense(pl.n_keys + nuniq);
            for (int c = 0; c < pl.n_keys + nuniq; c++) h_dense[c] = (int64_t*)dense->cols[c].d;
            if ((int)h_dense.size() > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "more than %d output columns", kMaxOut);
            ColPtrs cp;
            memset(&cp, 0, sizeof(cp));
            for (size_t c = 0; c < h_dense.size(); c++) cp.p[c] = h_dense[c];
            rq_group_table_compact<<<(kGroupTableCap + 255) / 256, 256, 0, E.stream>>>(
                E.g_state, E.g_keys, E.g_acc, ku, nuniq, cp, dense->d_n_rows);
            if (tm) tm->kernel_launches++;
            CK(cudaGetLastError());
            check_flags("aggregation pipeline", dense->d_n_rows);
            const int64_t n_groups_host = *(const int64_t*)(E.h_flags + 8);
            if (E.h_flags[0]) { RP.retries++; RP.why = "group overflow"; continue; }   // more groups than this path tracks: next implementation
            // expand duplicates by aliasing: copy the columns (tiny)
            {
                if (ncols > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "more than %d output columns", kMaxOut);
                CopyCols cc;
                memset(&cc, 0, sizeof(cc));
                for (int k = 0; k < pl.n_keys; k++) { cc.in[k] = (const int64_t*)dense->cols[k].d; cc.out[k] = (int64_t*)out->cols[k].d; cc.len[k] = kGroupTableCap; }
                for (int k = 0; k < pl.n_vals; k++) {
                    cc.in[pl.n_keys + k] = (const int64_t*)dense->cols[pl.n_keys + ad.uniq_of[k]].d;
                    cc.out[pl.n_keys + k] = (int64_t*)out->cols[pl.n_keys + k].d;
                    cc.len[pl.n_keys + k] = kGroupTableCap;
                }
                cc.in[ncols] = dense->d_n_rows; cc.out[ncols] = out->d_n_rows; cc.len[ncols] = 1;
                cc.ncols = ncols + 1;
                rq_copy_cols<<<dim3(2, (unsigned)cc.ncols), 256, 0, E.stream>>>(cc);
                if (tm) tm->kernel_launches++;
                CK(cudaGetLastError());
            }
            out->n_rows = n_groups_host;      // (the copies above are stream-ordered; no host wait needed)
            set_types(*out, pl);
            result.table = std::move(out);
            return;
        }
        if (impl == IMPL_BUILD || impl == IMPL_HASHAGG) {
            const int64_t rows_bound = std::max<int64_t>(1, src->n_rows >= 0 ? src->n_rows : src->cap_rows);
            int64_t want = rows_bound;
            // the planner's cardinality guess (RelOperator::getSize) sizes the first attempt of a table
            // scan; the dense second pass of a split pipeline inserts (nearly) every row it reads
            if (pl.size_hint > 0 && !src_override) want = std::min<int64_t>(rows_bound, std::max<int64_t>(pl.size_hint, 2048));
            uint64_t max_load_den = impl == IMPL_BUILD ? 3 : 2;   // joins: load factor <= 1/3
            uint64_t cap = 4096;
            while (cap < (uint64_t)(max_load_den * want)) cap <<= 1;
            uint64_t cap_max = 4096;
            while (cap_max < (uint64_t)(max_load_den * rows_bound)) cap_max <<= 1;
            // capacities that worked for this pipeline on this many rows are remembered, so a
            // repeated query does not pay for regrowth again
            const uint64_t sig = pipeline_signature(pl_in, rows_bound);
            auto known = g_ht_capacity.find(sig);
            if (known != g_ht_capacity.end()) cap = std::min<uint64_t>(known->second, cap_max);
            const int nk = pl.n_keys;
            if (nk > kMaxKeys) raise(RQ_ERR_UNSUPPORTED, "more than %d key columns", kMaxKeys);
            const int nv = impl == IMPL_BUILD ? pl.n_vals : (int)ad.kind.size();
            // hash aggregation on keys that pack into < 64 bits: the entry's first word is the key
            bool packed = false;
            if (impl == IMPL_HASHAGG && nk >= 1 && pl.n_keys + pl.n_vals <= kMaxOut && !has_str_key(pl) && pack_group_key(L, P, ku)) {
                int total = 0;
                for (int k = 0; k < ku.nk; k++) total += ku.bits[k];
                packed = total <= 63;
            }
            std::unique_ptr<HashTableDev> ht;
            unsigned long long n_used = 0;
            // Direct-address form (rq_internal.h DHashTable): one integer key with a proven, dense
            // enough value domain. The key bounds come from the upload statistics narrowed by the
            // pipeline's own range selections (so a build pruned to the probe side's key range - or
            // to one rank's share of it - gets a table of exactly that range).
            if (impl == IMPL_BUILD && nk == 1 && E.opt.direct_joins && !g_no_direct.count(sig)) {
                const HRef& hk = L.hkey[0];
                const int st = pl.keys[0].sql_type;
                const bool int_key = !(st == RQ_SQL_VARCHAR || (st == RQ_SQL_CHAR && pl.keys[0].width > 1));
                if (int_key && hk.lo > INT64_MIN / 2 && hk.hi < INT64_MAX / 2 && hk.hi >= hk.lo) {
                    const uint64_t dsize = (uint64_t)(hk.hi - hk.lo) + 1;
                    const uint64_t bytes = dsize / 8 + dsize * (uint64_t)nv * 8;
                    if (dsize <= (1ULL << 33) && bytes <= (24ULL << 30) &&
                        dsize <= 64ULL * (uint64_t)std::max<int64_t>(rows_bound, 4096)) {
                        ht.reset(new HashTableDev());
                        ht->capacity = dsize;
                        ht->d.direct = 1;
                        ht->d.dlo = hk.lo;
                        ht->d.dsize = dsize;
                        ht->d.dnv = (uint32_t)nv;
                        ht->d.nk = nk; ht->d.nv = nv;
                        ht->d.cap_mask = 0; ht->d.shift = 63;
                        ht->d.limit = ~0ULL;
                        const size_t words = (size_t)(dsize + 31) / 32;
                        CK(dmalloc(&ht->d.dbits, words * 4));
                        CK(cudaMemsetAsync(ht->d.dbits, 0, words * 4, E.stream));
                        if (nv > 0) CK(dmalloc(&ht->d.darr, (size_t)dsize * nv * 8));    // (not cleared: the bitmap says what is valid)
                        P.ht = ht->d;
                        CK(cudaMemsetAsync(E.flags, 0, 32, E.stream));
                        const bool shared = g_shared_build.active && nv == 0 && pl.source_kind == RQ_SRC_TABLE && !src_override;
                        TileRangeBuf share_buf;
                        const uint32_t* zone_range = P.tile_range;
                        if (shared) {
                            // this rank's share of the tiles (intersected with the zone range, if any)
                            P.tile_range = nullptr;
                            restrict_tiles(L, P, *src, share_buf, true, tm);
                        }
                        launch_pipeline(P, 0, rows_bound, tm, is_scan, ev_idx, ev_used);
                        if (shared) {
                            // bitmaps are summed (unique keys: disjoint bits), and so are the flag words, so that
                            // every rank sees the same entry count, error and duplicate flags and decides alike;
                            // a key present on two ranks shows as population count != entry count
                            Dist& D = E.dist;
                            const EventPair ep = event_pair(ev_idx);
                            ev_used.push_back({ev_idx, 2});
                            ev_idx++;
                            record_event(ep.a);
                            int rc = D.all_reduce(ht->d.dbits, ht->d.dbits, words, 3 /* ncclUint32 */, 0 /* ncclSum */, D.comm, E.stream);
                            if (rc == 0) rc = D.all_reduce(E.flags, E.flags, 4, 5 /* ncclUint64 */, 0, D.comm, E.stream);
                            if (rc != 0) raise(RQ_ERR_NCCL, "ncclAllReduce(shared build) failed: %s", D.get_error_string ? D.get_error_string(rc) : "?");
                            CK(cudaMemsetAsync(E.flags + 12, 0, 8, E.stream));
                            rq_popcount_words<<<(unsigned)std::min<size_t>((words + 255) / 256, 1184), 256, 0, E.stream>>>(ht->d.dbits, words, (unsigned long long*)(E.flags + 12));
                            rq_shared_build_check<<<1, 32, 0, E.stream>>>(E.flags);
                            if (tm) tm->kernel_launches += 2;
                            CK(cudaGetLastError());
                            record_event(ep.b);
                        }
                        P.tile_range = zone_range;      // (a fallback to the hash form below builds the whole table locally)
                        check_flags("join build pipeline");
                        if (E.h_flags[3] == 0) {
                            ht->entries = *(unsigned long long*)(E.h_flags + 6);
                            result.ht = std::move(ht);
                            return;
                        }
                        // duplicate keys: this build needs the hash form (and will get it right away next time)
                        g_no_direct.insert(sig);
                        RP.retries++; RP.why = "direct-address build met duplicate keys";
                        ht.reset();
                    }
                }
            }
            for (;;) {
                ht.reset(new HashTableDev());
                ht->capacity = cap;
                ht->d.cap_mask = cap - 1;
                { uint32_t lg = 0; while ((1ULL << lg) < cap) lg++; ht->d.shift = 64 - lg; }
                if (impl == IMPL_BUILD) {
                    // 8 filter bits per slot (16+ per key at load factor <= 1/2), L2 resident
                    const uint64_t words = std::max<uint64_t>(cap / 4, 1024);
                    ht->d.bloom_mask = (uint32_t)(words - 1);
                    CK(dmalloc(&ht->d.bloom, words * 4));
                    CK(cudaMemsetAsync(ht->d.bloom, 0, words * 4, E.stream));
                }
                ht->d.nk = nk; ht->d.nv = nv;
                ht->d.hmul = 0x9E3779B97F4A7C15ULL; ht->d.hsub = 0; ht->d.bloom_shift = 0;
                for (int k = 0; k < nk && k < kMaxKeys; k++) {
                    const int st = pl.keys[k].sql_type;
                    ht->d.key_kind[k] = st == RQ_SQL_VARCHAR ? 2 : (st == RQ_SQL_CHAR && pl.keys[k].width > 1) ? 1 : 0;
                }
                ht->d.packed = packed ? 1 : 0;
                ht->d.stride = (uint32_t)((1 + (packed ? 0 : nk) + nv + 3) / 4 * 4);
                CK(dmalloc(&ht->d.ent, cap * ht->d.stride * 8));
                if (impl == IMPL_HASHAGG) {
                    { HtKinds hk; memcpy(hk.kind, P.agg_kind, kMaxAggs); rq_ht_init<<<(unsigned)((cap + 255) / 256), 256, 0, E.stream>>>(ht->d, hk, 1); }
                    if (tm) tm->kernel_launches++;
                } else {
                    // whole sectors are cleared (a tag-only clear is a partial write per sector)
                    CK(cudaMemsetAsync(ht->d.ent, 0, cap * ht->d.stride * 8, E.stream));
                }
                ht->d.limit = cap / max_load_den + 1;      // beyond this load the table is regrown anyway
                P.ht = ht->d;
                CK(cudaMemsetAsync(E.flags, 0, 32, E.stream));
                trace_point("hash table allocated + cleared", pi);
                launch_pipeline(P, 0, rows_bound, tm, is_scan, ev_idx, ev_used);
                check_flags(impl == IMPL_BUILD ? "join build pipeline" : "hash aggregation pipeline");
                trace_point("hash sink kernel done", pi);
                bool regrow = E.h_flags[1] != 0;
                n_used = *(unsigned long long*)(E.h_flags + 6);     // counted by the kernel itself
                if (!regrow) regrow = n_used * max_load_den > cap && cap < cap_max;
                if (!regrow) break;
                RP.retries++; RP.why = "hash table regrown";
                if (cap >= cap_max) { raise(RQ_ERR_RUNTIME, "pipeline %d: hash table overflow at capacity %llu", pi, (unsigned long long)cap); }
                uint64_t next = cap * 8;
                if (!E.h_flags[1]) { next = cap; while (next < max_load_den * n_used) next <<= 1; }
                cap = std::min<uint64_t>(next, cap_max);
            }
            {   // remember the capacity this pipeline needs, not the one the search happened to end at
                uint64_t ideal = 4096;
                while (ideal < max_load_den * n_used) ideal <<= 1;
                g_ht_capacity[sig] = std::min<uint64_t>(ideal, cap_max);
            }
            if (impl == IMPL_BUILD) {
                ht->entries = n_used;
                result.ht = std::move(ht);
                return;
            }
            // dense relation out of the aggregation table
            const unsigned long long n_groups = n_used;
            const int ncols = pl.n_keys + pl.n_vals;
            out = new_intermediate(ncols, (int64_t)n_groups);
            std::vector<int> colmap(ncols);
            std::vector<int64_t*> h_cols(ncols);
            for (int k = 0; k < pl.n_keys; k++) colmap[k] = k;
            for (int k = 0; k < pl.n_vals; k++) colmap[pl.n_keys + k] = pl.n_keys + ad.uniq_of[k];
            for (int c = 0; c < ncols; c++) h_cols[c] = (int64_t*)out->cols[c].d;
            if (packed) {
                PackedCompact pc;
                memset(&pc, 0, sizeof(pc));
                pc.nk = nk; pc.n_out = ncols;
                for (int k = 0; k < nk; k++) { pc.shift[k] = ku.shift[k]; pc.bits[k] = ku.bits[k]; pc.sign[k] = ku.sign[k]; }
                for (int c = 0; c < ncols; c++) { pc.colmap[c] = colmap[c]; pc.out[c] = h_cols[c]; }
                rq_ht_compact_packed<<<(unsigned)((cap + 255) / 256), 256, 0, E.stream>>>(ht->d, pc, (unsigned long long*)out->d_n_rows, (unsigned long long)out->cap_rows);
                if (tm) tm->kernel_launches++;
                CK(cudaGetLastError());
                out->n_rows = (int64_t)n_groups;
                set_types(*out, pl);
                result.table = std::move(out);
                return;
            }
            if (ncols > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "more than %d output columns", kMaxOut);
            HtCompact hc;
            memset(&hc, 0, sizeof(hc));
            for (int c = 0; c < ncols; c++) { hc.colmap[c] = colmap[c]; hc.out[c] = h_cols[c]; }
            rq_ht_compact<<<(unsigned)((cap + 255) / 256), 256, 0, E.stream>>>(ht->d, hc, ncols, (unsigned long long*)out->d_n_rows, (unsigned long long)out->cap_rows);
            if (tm) tm->kernel_launches++;
            CK(cudaGetLastError());
            out->n_rows = (int64_t)n_groups;
            set_types(*out, pl);
            result.table = std::move(out);
            return;
        }
        if (impl == IMPL_EMIT) {
            int64_t cap = src->n_rows >= 0 ? src->n_rows : src->cap_rows;
            cap = std::min<int64_t>(std::max<int64_t>(cap, 1), (int64_t)1 << 24);
            const uint64_t esig = pipeline_signature(pl_in, src_rows) ^ 0x5bd1e995ULL;
            {   // a repeated query sizes the output from what it produced last time
                auto known = g_emit_rows.find(esig);
                if (known != g_emit_rows.end())
                    cap = std::min<int64_t>(std::max<int64_t>(src_rows, 1), known->second + known->second / 8 + 1024);
            }
            for (int round = 0; round < 2; round++) {
                const int n_emit = pl.n_vals + (P.expand_probe >= 0 ? 1 : 0);     // + entry index of the match
                if (n_emit > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "more than %d output columns", kMaxOut);
                out = new_intermediate(n_emit, cap);
                P.out_cap = out->cap_rows;
                P.out_count = (unsigned long long*)out->d_n_rows;
                for (int k = 0; k < n_emit; k++) P.out_col[k] = (int64_t*)out->cols[k].d;
                trace_point("materialize output allocated", pi);
                launch_pipeline(P, 0, src_rows, tm, is_scan, ev_idx, ev_used);
                check_flags("materialize pipeline", out->d_n_rows);
                trace_point("materialize kernel done", pi);
                const int64_t produced = *(const int64_t*)(E.h_flags + 8);
                g_emit_rows[esig] = produced;
                if (produced <= out->cap_rows) { out->n_rows = produced; break; }
                if (round == 1) raise(RQ_ERR_RUNTIME, "materialize overflow");
                RP.retries++; RP.why = "materialize output regrown";
                cap = produced;
            }
            set_types(*out, pl);
            if (P.expand_probe >= 0) { out->sql_type.push_back(RQ_SQL_BIGINT); out->sql_width.push_back(0); }
            result.table = std::move(out);
            return;
        }
    }
    raise(RQ_ERR_UNSUPPORTED, "pipeline %d: no implementation fits", pi);
}

// ---- two-pass execution of a selective probe ---------------------------------------------------
// pl is a simplified pipeline over a base table. If its first PROBE is selective (the build side
// holds far fewer keys than the probe key's value domain, taken from the upload statistics), the
// pipeline is cut in front of that PROBE:
//   pass A  nodes[0 .. p0) + Bloom-only PROBE  -> MATERIALIZE(values that are still needed)
//   pass B  source = pass A's relation: COL per value, then nodes[p0 .. n) and the original sink
// The Bloom filter has no false negatives, so pass B sees every tuple the one-pass form would have
// joined; results are identical. Pass A is a pure streaming kernel (no dependent table walks), pass
// B runs the walks / atomics over dense tiles where every lane has work.
static bool split_at_probe(const rq_plan& plan, const rq_pipeline& pl, const rq_table& src,
                           const std::vector<PipeOut>& outs, SplitPipes& sp, SplitMode mode, int ordinal) {
    (void)plan;
    const int n = pl.n_nodes;
    int p0 = -1;
    if (mode == SPLIT_EXCHANGE) {
        // partitioned plans: cut in front of the probe number `ordinal` (no probe node in pass A: the
        // rows are shipped to the rank that owns their key before they probe)
        int seen = 0;
        for (int i = 0; i < n && p0 < 0; i++)
            if (pl.nodes[i].op == RQ_OP_PROBE && seen++ == ordinal) p0 = i;
        if (p0 < 0) return false;
    } else if (mode == SPLIT_EXPAND) {
        // multi-match expansion (hashjoin.h:118-165): cut at the first probe that may match more than
        // one build tuple. Pass A emits one row per match together with the entry index, pass B
        // fetches the payload through that index.
        for (int i = 0; i < n && p0 < 0; i++)
            if (pl.nodes[i].op == RQ_OP_PROBE && (pl.nodes[i].imm & (1 | 8)) == 0) p0 = i;
        if (p0 < 0) return false;
        const rq_node& pr = pl.nodes[p0];
        if (pr.a < 0 || pr.a >= (int)outs.size() || !outs[pr.a].ht) return false;
    } else {
        int64_t min_rows = 4 << 20;
        double max_frac = 0.3;
        if (E.opt.split_min_rows >= 0) min_rows = E.opt.split_min_rows;
        if (E.opt.split_frac >= 0) max_frac = E.opt.split_frac;
        if (src.n_rows < min_rows) return false;
        for (int i = 0; i < n && p0 < 0; i++) if (pl.nodes[i].op == RQ_OP_PROBE) p0 = i;
        if (p0 < 0 || (pl.nodes[p0].imm & ~1LL) != 0) return false;
        const rq_node& pr = pl.nodes[p0];
        if (pr.a < 0 || pr.a >= (int)outs.size() || !outs[pr.a].ht || !(outs[pr.a].ht->d.bloom || outs[pr.a].ht->d.direct)) return false;
        // selectivity estimate: build entries / size of the probe key's value domain
        double frac = 1.0;
        if (pr.c == 1) {
            const rq_node& kn = pl.nodes[pl.args[pr.b]];
            if (kn.op == RQ_OP_COL && kn.a >= 0 && kn.a < (int)src.cols.size() && src.cols[kn.a].has_stats) {
                const double dom = (double)src.cols[kn.a].vmax - (double)src.cols[kn.a].vmin + 1.0;
                if (dom > 0) frac = (double)outs[pr.a].ht->entries / dom;
            }
        }
        if (frac > max_frac) return false;
    }
    const rq_node& pr = pl.nodes[p0];

    // values computed before the cut and read behind it
    std::vector<char> need(n, 0);
    auto mark = [&](int r) { if (r >= 0 && r < p0) need[r] = 1; };
    for (int i = p0; i < n; i++) {
        const rq_node& nd = pl.nodes[i];
        if (is_binary(nd.op)) { mark(nd.a); mark(nd.b); }
        else if (nd.op == RQ_OP_FILTER) mark(nd.a);
        else if (nd.op == RQ_OP_SELECT) { mark(nd.a); mark(nd.b); mark(nd.c); }
        else if (nd.op == RQ_OP_PROBE && !(i == p0 && mode == SPLIT_EXPAND))
            for (int k = 0; k < nd.c; k++) mark(pl.args[nd.b + k]);
    }
    for (int k = 0; k < pl.n_keys; k++) mark(pl.keys[k].node);
    for (int k = 0; k < pl.n_vals; k++)
        if (!(pl.sink_kind == RQ_SINK_AGG && pl.vals[k].kind == RQ_AGG_COUNT)) mark(pl.vals[k].node);
    for (int i = 0; i < p0; i++)
        if (need[i] && pl.nodes[i].op == RQ_OP_FILTER) return false;
    // A PAYLOAD of a probe in front of the cut is an ordinary live value (it sits in a value slot of
    // pass A and is materialized like any other; string payloads are addresses into resident tables).
    // PAYLOAD nodes that stand behind the cut but belong to a probe in front of it are evaluated by
    // pass A as well: they are appended to its node list.
    std::vector<int> late_payload;       // nodes i >= p0: PAYLOAD of a probe < p0
    for (int i = p0; i < n; i++)
        if (pl.nodes[i].op == RQ_OP_PAYLOAD && pl.nodes[i].a < p0) late_payload.push_back(i);

    // pass A
    sp.a_nodes.assign(pl.nodes, pl.nodes + p0);
    std::vector<int> late_pos(n, -1);
    for (int i : late_payload) { late_pos[i] = (int)sp.a_nodes.size(); sp.a_nodes.push_back(pl.nodes[i]); }
    if (mode != SPLIT_EXCHANGE) {
        rq_node semi = pr;
        semi.imm |= (mode == SPLIT_EXPAND ? 4 : 2);
        sp.a_nodes.push_back(semi);
    }
    sp.a_args.assign(pl.args, pl.args + pl.n_args);
    std::vector<int> newidx(n, -1);
    for (int i = 0; i < p0; i++) {
        const int op = pl.nodes[i].op;
        if (!need[i] || op == RQ_OP_CONST || op == RQ_OP_CONST_STR) continue;
        rq_value v;
        v.node = i; v.kind = 0; v.sql_type = RQ_SQL_BIGINT; v.width = 0;
        newidx[i] = (int)sp.a_vals.size();
        sp.a_vals.push_back(v);
        sp.live_src_col.push_back(op == RQ_OP_COL ? pl.nodes[i].a : -1);
    }
    for (int i : late_payload) {
        rq_value v;
        v.node = late_pos[i]; v.kind = 0; v.sql_type = RQ_SQL_BIGINT; v.width = 0;
        newidx[i] = (int)sp.a_vals.size();
        sp.a_vals.push_back(v);
        sp.live_src_col.push_back(-1);
    }
    if (mode != SPLIT_EXPAND && sp.a_vals.empty()) return false;
    if ((int)sp.a_vals.size() + 1 > kMaxOut || (int)sp.a_vals.size() + 1 > kMaxStagedCols) {
        if (mode == SPLIT_EXPAND) raise(RQ_ERR_UNSUPPORTED, "multi-match join carries more than %d live values", kMaxStagedCols - 1);
        return false;
    }
    memset(&sp.a, 0, sizeof(sp.a));
    sp.a.source_kind = pl.source_kind; sp.a.source_id = pl.source_id; sp.a.source_id2 = pl.source_id2;
    sp.a.n_nodes = (int)sp.a_nodes.size(); sp.a.nodes = sp.a_nodes.data();
    sp.a.n_args = (int)sp.a_args.size(); sp.a.args = sp.a_args.data();
    sp.a.sink_kind = RQ_SINK_MATERIALIZE;
    sp.a.n_vals = (int)sp.a_vals.size(); sp.a.vals = sp.a_vals.data();

    // pass B
    for (size_t c = 0; c < sp.a_vals.size(); c++) sp.b_nodes.push_back(rq_node{RQ_OP_COL, (int)c, 0, 0, 0});
    int slot_col_node = -1;
    if (mode == SPLIT_EXPAND) {          // the entry index pass A appended as its last column
        slot_col_node = (int)sp.b_nodes.size();
        sp.b_nodes.push_back(rq_node{RQ_OP_COL, (int)sp.a_vals.size(), 0, 0, 0});
    }
    for (int i = 0; i < p0; i++) {
        const int op = pl.nodes[i].op;
        if (need[i] && (op == RQ_OP_CONST || op == RQ_OP_CONST_STR)) {
            newidx[i] = (int)sp.b_nodes.size();
            sp.b_nodes.push_back(pl.nodes[i]);
        }
    }
    for (int i = p0; i < n; i++) {
        rq_node nd = pl.nodes[i];
        if (late_pos[i] >= 0) continue;          // evaluated by pass A, a COL of pass B (newidx set above)
        if (is_binary(nd.op)) { nd.a = newidx[nd.a]; nd.b = newidx[nd.b]; }
        else if (nd.op == RQ_OP_FILTER) nd.a = newidx[nd.a];
        else if (nd.op == RQ_OP_SELECT) { nd.a = newidx[nd.a]; nd.b = newidx[nd.b]; nd.c = newidx[nd.c]; }
        else if (nd.op == RQ_OP_PAYLOAD) nd.a = newidx[nd.a];
        else if (nd.op == RQ_OP_PROBE) {
            const int b0 = (int)sp.b_args.size();
            if (i == p0 && mode == SPLIT_EXPAND) {
                sp.b_args.push_back(slot_col_node);
                nd.c = 1;
                nd.imm |= 8;
            } else {
                for (int k = 0; k < nd.c; k++) sp.b_args.push_back(newidx[pl.args[nd.b + k]]);
            }
            nd.b = b0;
        }
        newidx[i] = (int)sp.b_nodes.size();
        sp.b_nodes.push_back(nd);
    }
    sp.b_keys.assign(pl.keys, pl.keys + pl.n_keys);
    sp.b_vals.assign(pl.vals, pl.vals + pl.n_vals);
    for (auto& k : sp.b_keys) k.node = newidx[k.node];
    for (auto& v : sp.b_vals)
        if (!(pl.sink_kind == RQ_SINK_AGG && v.kind == RQ_AGG_COUNT)) v.node = newidx[v.node];
    sp.b = pl;
    sp.b.source_kind = RQ_SRC_PIPELINE; sp.b.source_id = 0;
    sp.b.n_nodes = (int)sp.b_nodes.size(); sp.b.nodes = sp.b_nodes.data();
    sp.b.n_args = (int)sp.b_args.size(); sp.b.args = sp.b_args.data();
    sp.b.keys = sp.b_keys.data(); sp.b.vals = sp.b_vals.data();
    return true;
}

// ---- multi-GPU: merge of the per-rank partial results (RQ_PLAN_SHARDED) -----------------------
// Fact tables are row-range sharded, one process per GPU. After the plan's last aggregation ran on
// the local shard, every rank holds a dense group table (keys, partial SUM/COUNT/MIN/MAX). The
// tables are exchanged with ONE ncclAllGather (padded to the largest rank) and re-aggregated on
// every rank by the ordinary aggregation pipeline (SUM and COUNT partials add, MIN/MAX take
// min/max). mod-2^64 addition is associative, so the merged sums are bit-identical to a single-GPU
// run; AVG is finalised by the following pipeline, i.e. after the merge (aggregation.h:182-204).
// A plan without aggregation concatenates the per-rank relations instead.
static bool is_str_type(int sql_type, int sql_width) {
    return sql_type == RQ_SQL_VARCHAR || (sql_type == RQ_SQL_CHAR && sql_width > 1);
}

static std::unique_ptr<rq_table> gather_relation(const rq_table& local, rq_timings* tm, std::vector<void*>& owned,
                                                 size_t& ev_idx, std::vector<std::pair<size_t, int>>& ev_used) {
    Dist& D = E.dist;
    const int W = D.world;
    const int ncols = (int)local.cols.size();
    if (W > 64) raise(RQ_ERR_UNSUPPORTED, "sharded plans support up to 64 ranks");
    if (ncols > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "sharded merge of more than %d columns", kMaxOut);
    const EventPair ep = event_pair(ev_idx);
    ev_used.push_back({ev_idx, 2});
    ev_idx++;
    record_event(ep.a);
    // 1. row counts
    int64_t* d_counts = nullptr;
    CK(dmalloc(&d_counts, sizeof(int64_t) * W));
    const int64_t* d_n = local.d_n_rows;
    int64_t* d_tmp = nullptr;
    if (local.n_rows >= 0) {
        CK(dmalloc(&d_tmp, 8));
        upload_small(d_tmp, &local.n_rows, 8);
        d_n = d_tmp;
    }
    auto nccl_ck = [&](int rc, const char* what) {
        if (rc != 0) raise(RQ_ERR_NCCL, "%s failed: %s", what, D.get_error_string ? D.get_error_string(rc) : "?");
    };
    nccl_ck(D.all_gather(d_n, d_counts, 1, 4 /* ncclInt64 */, D.comm, E.stream), "ncclAllGather(counts)");
    std::vector<int64_t> counts(W);
    host_read(counts.data(), d_counts, sizeof(int64_t) * W);
    int64_t maxn = 0, total = 0;
    for (int r = 0; r < W; r++) { maxn = std::max(maxn, counts[r]); total += counts[r]; }
    std::unique_ptr<rq_table> all = new_intermediate(ncols, total);
    all->sql_type = local.sql_type;
    all->sql_width = local.sql_width;
    if (maxn > 0 && ncols > 0) {
        // 2. [ncols][maxn] per rank -> [world][ncols][maxn], one pack and one unpack kernel
        int64_t *send = nullptr, *recv = nullptr;
        const size_t per_rank = (size_t)ncols * (size_t)maxn;
        CK(dmalloc(&send, per_rank * 8));
        CK(dmalloc(&recv, per_rank * 8 * W));
        GatherCols gc;
        memset(&gc, 0, sizeof(gc));
        gc.ncols = ncols; gc.world = W;
        for (int c = 0; c < ncols; c++) { gc.in[c] = (const int64_t*)local.cols[c].d; gc.out[c] = (int64_t*)all->cols[c].d; }
        GatherCounts gn;
        memset(&gn, 0, sizeof(gn));
        int64_t off = 0;
        for (int r = 0; r < W; r++) { gn.count[r] = counts[r]; gn.off[r] = off; off += counts[r]; }
        const unsigned pb = (unsigned)std::min<size_t>((per_rank + 255) / 256, 148 * 8);
        rq_gather_pack<<<pb, 256, 0, E.stream>>>(gc, counts[D.rank], maxn, send);
        nccl_ck(D.all_gather(send, recv, per_rank, 4, D.comm, E.stream), "ncclAllGather(partials)");
        const unsigned ub = (unsigned)std::min<size_t>((per_rank * W + 255) / 256, 148 * 8);
        rq_gather_unpack<<<ub, 256, 0, E.stream>>>(gc, gn, maxn, recv);
        if (tm) tm->kernel_launches += 2;
        dfree(send);
        dfree(recv);
    }
    // string columns hold rank-local addresses: ship the bytes (sql_width + 1 per row) and point the
    // gathered column at the received copies
    for (int c = 0; c < ncols && maxn > 0; c++) {
        if (!(c < (int)local.sql_type.size() && is_str_type(local.sql_type[c], local.sql_width[c]))) continue;
        const int w = local.sql_width[c] + 1;
        unsigned char *sb = nullptr, *rb = nullptr, *cat = nullptr;
        CK(dmalloc(&sb, (size_t)maxn * w));
        CK(dmalloc(&rb, (size_t)maxn * w * W));
        CK(dmalloc(&cat, (size_t)std::max<int64_t>(total, 1) * w));
        owned.push_back(cat);
        const int64_t mine = counts[D.rank];
        if (mine > 0) rq_gather_str<<<(unsigned)((mine + 255) / 256), 256, 0, E.stream>>>((const int64_t*)local.cols[c].d, sb, w, mine);
        nccl_ck(D.all_gather(sb, rb, (size_t)maxn * w, 0 /* ncclInt8 */, D.comm, E.stream), "ncclAllGather(strings)");
        int64_t off = 0;
        for (int r = 0; r < W; r++) {
            if (counts[r] > 0)
                CK(cudaMemcpyAsync(cat + (size_t)off * w, rb + (size_t)r * maxn * w, (size_t)counts[r] * w, cudaMemcpyDeviceToDevice, E.stream));
            off += counts[r];
        }
        if (total > 0) rq_str_addrs<<<(unsigned)((total + 255) / 256), 256, 0, E.stream>>>(cat, w, total, (int64_t*)all->cols[c].d);
        if (tm) tm->kernel_launches += 2;
        dfree(sb);
        dfree(rb);
    }
    CK(cudaGetLastError());
    upload_small(all->d_n_rows, &total, 8);
    record_event(ep.b);
    all->n_rows = total;
    dfree(d_counts);
    if (d_tmp) dfree(d_tmp);
    return all;
}

// Sharded plans: before the first exchange every rank learns whether some rank failed while running
// its pipelines (division by zero on one shard, a table that cannot grow, out of memory ...), so that
// all ranks leave the plan with an error instead of waiting in a collective the failed rank never joins.
// The same exchange carries one more bit per rank (`big`: my partial result is large), so that all
// ranks pick the same merge strategy. Returns whether any rank said so.
static bool agree_on_status(int local_code, const std::string& local_msg, bool big) {
    Dist& D = E.dist;
    int32_t* d_st = nullptr;
    CK(dmalloc(&d_st, sizeof(int32_t) * (D.world + 1)));
    const int32_t mine = local_code != 0 ? local_code : (big ? -1 : 0);
    upload_small(d_st + D.world, &mine, 4);
    const int rc = D.all_gather(d_st + D.world, d_st, 1, 2 /* ncclInt32 */, D.comm, E.stream);
    if (rc != 0) raise(RQ_ERR_NCCL, "ncclAllGather(status) failed: %s", D.get_error_string ? D.get_error_string(rc) : "?");
    std::vector<int32_t> st(D.world);
    host_read(st.data(), d_st, sizeof(int32_t) * D.world);
    dfree(d_st);
    if (local_code != 0) raise(local_code, "%s", local_msg.c_str());
    bool any_big = false;
    for (int r = 0; r < D.world; r++) {
        if (st[r] > 0) raise(st[r], "rank %d failed while executing its shard of the plan (status %d); the plan was abandoned on all ranks", r, st[r]);
        any_big |= st[r] < 0;
    }
    return any_big;
}

// ---- hash-partitioned exchange (RQ_PLAN_PARTITIONED; kernels in exchange_kernels.cuh) -----------
// Every row of `local` (may be null: this rank failed before it had anything to send) goes to rank
// hash(key columns) mod world. The counts exchange doubles as the status agreement: a rank that failed
// in the pipeline in front of the exchange reports its error code there and every rank raises.
static std::unique_ptr<rq_table> exchange_relation(const rq_table* local, const std::vector<int>& key_cols,
                                                   const std::vector<uint8_t>& key_kinds, int local_code,
                                                   const std::string& local_msg, rq_timings* tm, size_t& ev_idx,
                                                   std::vector<std::pair<size_t, int>>& ev_used) {
    Dist& D = E.dist;
    const int W = D.world;
    if (W > kMaxRanks) raise(RQ_ERR_UNSUPPORTED, "partitioned plans support up to %d ranks", kMaxRanks);
    const int ncols = local ? (int)local->cols.size() : 0;
    if (ncols > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "exchange of more than %d columns", kMaxOut);
    if (local)
        for (int c = 0; c < ncols; c++)
            if (c < (int)local->sql_type.size() && is_str_type(local->sql_type[c], local->sql_width[c]))
                raise(RQ_ERR_UNSUPPORTED, "partitioned plans cannot ship string values between ranks yet (column %d)", c);
    auto nccl_ck = [&](int rc, const char* what) {
        if (rc != 0) raise(RQ_ERR_NCCL, "%s failed: %s", what, D.get_error_string ? D.get_error_string(rc) : "?");
    };
    const EventPair ep = event_pair(ev_idx);
    ev_used.push_back({ev_idx, 2});
    ev_idx++;
    record_event(ep.a);

    const int64_t rows_bound = local ? std::max<int64_t>(local->n_rows >= 0 ? local->n_rows : local->cap_rows, 0) : 0;
    // [cnt W][off W][cursor W][status 1] and the gathered matrix [W][W+1]
    unsigned long long* d_meta = nullptr;
    unsigned long long* d_matrix = nullptr;
    CK(dmalloc(&d_meta, sizeof(unsigned long long) * (3 * W + 1)));
    CK(dmalloc(&d_matrix, sizeof(unsigned long long) * W * (W + 1)));
    CK(cudaMemsetAsync(d_meta, 0, sizeof(unsigned long long) * (3 * W + 1), E.stream));
    unsigned long long *d_cnt = d_meta, *d_off = d_meta + W, *d_cur = d_meta + 2 * W;
    uint8_t* d_dest = nullptr;
    int64_t* d_send = nullptr;
    ExCols X;
    memset(&X, 0, sizeof(X));
    X.ncols = ncols; X.world = W; X.nkeys = (int)key_cols.size();
    if (X.nkeys > kMaxKeys) raise(RQ_ERR_UNSUPPORTED, "more than %d key columns", kMaxKeys);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, (rows_bound + kExBlockRows - 1) / kExBlockRows);
    if (local && local_code == 0 && ncols > 0) {
        for (int c = 0; c < ncols; c++) X.in[c] = (const int64_t*)local->cols[c].d;
        for (int k = 0; k < X.nkeys; k++) { X.key_col[k] = key_cols[k]; X.key_kind[k] = key_kinds[k]; }
        CK(dmalloc(&d_dest, (size_t)std::max<int64_t>(rows_bound, 1)));
        CK(dmalloc(&d_send, (size_t)std::max<int64_t>(rows_bound, 1) * ncols * 8));
        const int64_t* n_ptr = local->n_rows < 0 ? local->d_n_rows : nullptr;
        rq_ex_count<<<blocks, kExThreads, 0, E.stream>>>(X, n_ptr, local->n_rows, rows_bound, d_dest, d_cnt);
        rq_ex_offsets<<<1, 32, 0, E.stream>>>(d_cnt, d_off, d_cur, W);
        rq_ex_scatter<<<blocks, kExThreads, 0, E.stream>>>(X, n_ptr, local->n_rows, rows_bound, d_dest, d_cnt, d_off, d_cur, d_send);
        if (tm) tm->kernel_launches += 3;
        CK(cudaGetLastError());
    }
    // counts (+ status) of every rank
    {
        const unsigned long long st = (unsigned long long)(unsigned)local_code;
        // row r of the matrix = [cnt of rank r for every destination][status of rank r]; the local row is
        // assembled in place: cnt is followed by the status word
        unsigned long long* d_row = nullptr;
        CK(dmalloc(&d_row, sizeof(unsigned long long) * (W + 1)));
        CK(cudaMemcpyAsync(d_row, d_cnt, sizeof(unsigned long long) * W, cudaMemcpyDeviceToDevice, E.stream));
        upload_small(d_row + W, &st, 8);
        nccl_ck(D.all_gather(d_row, d_matrix, (size_t)(W + 1), 5 /* ncclUint64 */, D.comm, E.stream), "ncclAllGather(exchange counts)");
        dfree(d_row);
    }
    std::vector<unsigned long long> M((size_t)W * (W + 1));
    host_read(M.data(), d_matrix, sizeof(unsigned long long) * M.size());
    if (local_code != 0) raise(local_code, "%s", local_msg.c_str());
    for (int r = 0; r < W; r++)
        if (M[(size_t)r * (W + 1) + W] != 0)
            raise((int)M[(size_t)r * (W + 1) + W], "rank %d failed in front of an exchange (status %d); the plan was abandoned on all ranks",
                  r, (int)M[(size_t)r * (W + 1) + W]);
    // every rank must ship the same columns (same plan): ncols is symmetric by construction
    ExUnpack U;
    memset(&U, 0, sizeof(U));
    U.ncols = ncols; U.world = W;
    int64_t total = 0;
    for (int r = 0; r < W; r++) { U.count[r] = (int64_t)M[(size_t)r * (W + 1) + D.rank]; U.off[r] = total; total += U.count[r]; }
    std::unique_ptr<rq_table> out = new_intermediate(ncols, total);
    if (local) { out->sql_type = local->sql_type; out->sql_width = local->sql_width; }
    int64_t* d_recv = nullptr;
    CK(dmalloc(&d_recv, (size_t)std::max<int64_t>(total, 1) * std::max(ncols, 1) * 8));
    if (ncols > 0) {
        nccl_ck(D.group_start(), "ncclGroupStart");
        int64_t soff = 0;
        for (int r = 0; r < W; r++) {
            const int64_t sc = (int64_t)M[(size_t)D.rank * (W + 1) + r];
            if (sc > 0) nccl_ck(D.send(d_send + (size_t)soff * ncols, (size_t)sc * ncols, 4 /* ncclInt64 */, r, D.comm, E.stream), "ncclSend");
            soff += sc;
            if (U.count[r] > 0) nccl_ck(D.recv(d_recv + (size_t)U.off[r] * ncols, (size_t)U.count[r] * ncols, 4, r, D.comm, E.stream), "ncclRecv");
        }
        nccl_ck(D.group_end(), "ncclGroupEnd");
        if (total > 0) {
            for (int c = 0; c < ncols; c++) U.out[c] = (int64_t*)out->cols[c].d;
            const unsigned ub = (unsigned)std::min<int64_t>((total * ncols + 255) / 256, 148 * 16);
            rq_ex_unpack<<<ub, 256, 0, E.stream>>>(U, d_recv, total);
            if (tm) tm->kernel_launches++;
            CK(cudaGetLastError());
        }
    }
    upload_small(out->d_n_rows, &total, 8);
    out->n_rows = total;
    record_event(ep.b);
    dfree(d_meta); dfree(d_matrix); dfree(d_dest); dfree(d_send); dfree(d_recv);
    return out;
}

// re-aggregation of exchanged partial groups: column i of `all` is key / partial aggregate i of
// pipeline pl (SUM and COUNT partials add, MIN / MAX take min / max)
static std::unique_ptr<rq_table> reaggregate(const rq_plan& plan, const rq_pipeline& pl, rq_table* all, const char* d_strpool,
                                             rq_timings* tm, size_t& ev_idx, std::vector<std::pair<size_t, int>>& ev_used,
                                             double& lower_ms) {
    const int nk = pl.n_keys, nv = pl.n_vals;
    std::vector<rq_node> nodes(nk + nv);
    std::vector<rq_value> keys(pl.keys, pl.keys + nk), vals(pl.vals, pl.vals + nv);
    for (int i = 0; i < nk + nv; i++) { nodes[i] = rq_node{RQ_OP_COL, i, 0, 0, 0}; }
    for (int k = 0; k < nk; k++) keys[k].node = k;
    for (int v = 0; v < nv; v++) {
        vals[v].node = nk + v;
        if (vals[v].kind == RQ_AGG_COUNT) vals[v].kind = RQ_AGG_SUM;     // partial counts add up
    }
    rq_pipeline mp;
    memset(&mp, 0, sizeof(mp));
    mp.source_kind = RQ_SRC_TABLE; mp.source_id = 0;
    mp.n_nodes = nk + nv; mp.nodes = nodes.data();
    mp.sink_kind = RQ_SINK_AGG;
    mp.n_keys = nk; mp.keys = keys.data();
    mp.n_vals = nv; mp.vals = vals.data();
    mp.size_hint = all->n_rows;
    rq_table* tabs[1] = {all};
    rq_plan mplan;
    memset(&mplan, 0, sizeof(mplan));
    mplan.n_tables = 1; mplan.tables = tabs;
    mplan.n_pipelines = 1; mplan.pipelines = &mp;
    mplan.limit = -1;
    mplan.strpool = plan.strpool; mplan.strpool_bytes = plan.strpool_bytes;
    std::vector<PipeOut> mouts(1);
    run_pipeline(mplan, 0, mouts, d_strpool, tm, ev_idx, ev_used, lower_ms, /*is_fact_scan=*/false);
    return std::move(mouts[0].table);
}

static void merge_sharded(const rq_plan& plan, int pi, std::vector<PipeOut>& outs, const char* d_strpool,
                          rq_timings* tm, size_t& ev_idx, std::vector<std::pair<size_t, int>>& ev_used,
                          double& lower_ms) {
    if (!outs[pi].table) raise(RQ_ERR_INVALID, "sharded merge: pipeline %d has no relation output", pi);
    const rq_pipeline& pl = plan.pipelines[pi];
    std::unique_ptr<rq_table> all = gather_relation(*outs[pi].table, tm, outs[pi].owned, ev_idx, ev_used);
    trace_point("partials gathered", pi);
    if (pl.sink_kind != RQ_SINK_AGG) {       // no aggregation: the concatenation is the result
        outs[pi].table = std::move(all);
        return;
    }
    outs[pi].table = reaggregate(plan, pl, all.get(), d_strpool, tm, ev_idx, ev_used, lower_ms);
}

// Large partial results: every group goes to the rank that owns hash(group key) and is merged there
// (each rank re-aggregates 1/world of the groups instead of all of them). The merged relation stays
// partitioned; the final relation of the plan is concatenated before ORDER BY / LIMIT.
static void merge_partitioned(const rq_plan& plan, int pi, std::vector<PipeOut>& outs, const char* d_strpool,
                              rq_timings* tm, size_t& ev_idx, std::vector<std::pair<size_t, int>>& ev_used,
                              double& lower_ms) {
    if (!outs[pi].table) raise(RQ_ERR_INVALID, "partitioned merge: pipeline %d has no relation output", pi);
    const rq_pipeline& pl = plan.pipelines[pi];
    if (pl.sink_kind != RQ_SINK_AGG) return;      // plain relation: concatenated at the end of the plan
    std::vector<int> key_cols;
    std::vector<uint8_t> kinds;
    for (int k = 0; k < pl.n_keys; k++) { key_cols.push_back(k); kinds.push_back(0); }
    std::unique_ptr<rq_table> ex = exchange_relation(outs[pi].table.get(), key_cols, kinds, 0, "", tm, ev_idx, ev_used);
    trace_point("partials exchanged", pi);
    outs[pi].table = reaggregate(plan, pl, ex.get(), d_strpool, tm, ev_idx, ev_used, lower_ms);
}

// ---- RQ_PLAN_PARTITIONED: every table is a row range; joins meet by key hash -------------------
static bool is_partitioned_plan(const rq_plan& plan) {
    return (plan.flags & RQ_PLAN_PARTITIONED) && E.dist.comm && E.dist.world > 1;
}

static void run_pipeline_partitioned(const rq_plan& plan, const rq_pipeline& pl_in, int pi, std::vector<PipeOut>& outs,
                                     const char* d_strpool, rq_timings* tm, size_t& ev_idx,
                                     std::vector<std::pair<size_t, int>>& ev_used, double& lower_ms,
                                     const rq_table* src_override, bool first_probe_aligned, PipeOut& result) {
    SimplePipe sp;
    simplify_pipeline(pl_in, sp);
    const rq_pipeline& pl = sp.pl;
    if (pl.source_kind == RQ_SRC_CROSS)
        raise(RQ_ERR_UNSUPPORTED, "pipeline %d: nested-loops joins are not available in partitioned plans", pi);
    const rq_table* src = src_override;
    if (!src) {
        if (pl.source_kind == RQ_SRC_TABLE) {
            if (pl.source_id < 0 || pl.source_id >= plan.n_tables || !plan.tables[pl.source_id])
                raise(RQ_ERR_INVALID, "pipeline %d: table %d out of range", pi, pl.source_id);
            src = plan.tables[pl.source_id];
        } else if (pl.source_kind == RQ_SRC_PIPELINE) {
            if (pl.source_id < 0 || pl.source_id >= pi || !outs[pl.source_id].table)
                raise(RQ_ERR_INVALID, "pipeline %d: source pipeline %d has no relation output", pi, pl.source_id);
            src = outs[pl.source_id].table.get();
        } else {
            raise(RQ_ERR_INVALID, "pipeline %d: bad source kind %d", pi, pl.source_kind);
        }
    }
    auto run_local = [&](const rq_pipeline& p, const rq_table* s, PipeOut& r) {
        run_pipeline_one(plan, p, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, s == nullptr, s, r, false);
    };
    const rq_table* local_src = src_override;      // nullptr = the pipeline's own table

    // 1. a probe whose input is not yet partitioned by its key: ship the rows first
    SplitPipes sx;
    if (split_at_probe(plan, pl, *src, outs, sx, SPLIT_EXCHANGE, first_probe_aligned ? 1 : 0)) {
        int code = 0;
        std::string msg;
        PipeOut mid;
        try {
            run_local(sx.a, local_src, mid);
        } catch (RqError& e) {
            cudaStreamSynchronize(E.stream);
            code = e.code; msg = e.msg;
        }
        // key columns of the probe in pass A's output: the probe is pass B's first PROBE node
        std::vector<int> key_cols;
        std::vector<uint8_t> kinds;
        for (int i = 0; i < sx.b.n_nodes && key_cols.empty(); i++) {
            const rq_node& nd = sx.b.nodes[i];
            if (nd.op != RQ_OP_PROBE) continue;
            if (nd.a < 0 || nd.a >= (int)outs.size() || !outs[nd.a].ht) raise(RQ_ERR_INVALID, "PROBE refers to pipeline %d which built no hash table", nd.a);
            for (int k = 0; k < nd.c; k++) {
                const rq_node& kn = sx.b.nodes[sx.b.args[nd.b + k]];
                if (kn.op != RQ_OP_COL) raise(RQ_ERR_UNSUPPORTED, "pipeline %d: constant join key in a partitioned plan", pi);
                key_cols.push_back(kn.a);
                kinds.push_back(outs[nd.a].ht->d.key_kind[k]);
            }
        }
        std::unique_ptr<rq_table> ex = exchange_relation(code == 0 ? mid.table.get() : nullptr, key_cols, kinds, code, msg, tm, ev_idx, ev_used);
        run_pipeline_partitioned(plan, sx.b, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms, ex.get(), true, result);
        return;
    }
    // 2. a build: its rows go to the rank that owns their key, the table is built there
    if (pl.sink_kind == RQ_SINK_BUILD) {
        std::vector<rq_value> mvals;
        for (int k = 0; k < pl.n_keys; k++) { rq_value v = pl.keys[k]; v.kind = 0; mvals.push_back(v); }
        for (int k = 0; k < pl.n_vals; k++) { rq_value v = pl.vals[k]; v.kind = 0; mvals.push_back(v); }
        rq_pipeline a = pl;
        a.sink_kind = RQ_SINK_MATERIALIZE;
        a.n_keys = 0; a.keys = nullptr;
        a.n_vals = (int)mvals.size(); a.vals = mvals.data();
        int code = 0;
        std::string msg;
        PipeOut mid;
        try {
            run_local(a, local_src, mid);
        } catch (RqError& e) {
            cudaStreamSynchronize(E.stream);
            code = e.code; msg = e.msg;
        }
        std::vector<int> key_cols;
        std::vector<uint8_t> kinds;
        for (int k = 0; k < pl.n_keys; k++) {
            key_cols.push_back(k);
            const int st = pl.keys[k].sql_type;
            kinds.push_back(st == RQ_SQL_VARCHAR ? 2 : (st == RQ_SQL_CHAR && pl.keys[k].width > 1) ? 1 : 0);
        }
        std::unique_ptr<rq_table> ex = exchange_relation(code == 0 ? mid.table.get() : nullptr, key_cols, kinds, code, msg, tm, ev_idx, ev_used);
        const int nc = pl.n_keys + pl.n_vals;
        std::vector<rq_node> bn(nc);
        std::vector<rq_value> bk(pl.keys, pl.keys + pl.n_keys), bv(pl.vals, pl.vals + pl.n_vals);
        for (int i = 0; i < nc; i++) bn[i] = rq_node{RQ_OP_COL, i, 0, 0, 0};
        for (int k = 0; k < pl.n_keys; k++) bk[k].node = k;
        for (int k = 0; k < pl.n_vals; k++) bv[k].node = pl.n_keys + k;
        rq_pipeline b = pl;
        b.source_kind = RQ_SRC_PIPELINE; b.source_id = 0;
        b.n_nodes = nc; b.nodes = bn.data();
        b.n_args = 0; b.args = nullptr;
        b.keys = bk.data(); b.vals = bv.data();
        b.size_hint = 0;
        run_local(b, ex.get(), result);
        return;
    }
    // 3. everything this rank needs is local now
    run_local(pl, local_src, result);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// result assembly
// ------------------------------------------------------------------------------------------
static int phys_type(int sql_type, int sql_width, int* width) {
    switch (sql_type) {
        case RQ_SQL_BOOL: *width = 1; return RQ_I8;
        case RQ_SQL_CHAR:
            if (sql_width <= 1) { *width = 1; return RQ_I8; }
            *width = sql_width + 1; return RQ_STR;
        case RQ_SQL_VARCHAR: *width = sql_width + 1; return RQ_STR;
        case RQ_SQL_INT: case RQ_SQL_DATE: *width = 4; return RQ_I32;
        default: *width = 8; return RQ_I64;
    }
}

extern "C" int rq_result_free(rq_result* r) {
    if (!r) return RQ_OK;
    ResultBox* box = reinterpret_cast<ResultBox*>(r);      // every result is allocated as a ResultBox
    free(box->block);
    free(r->cols);
    free(box);
    return RQ_OK;
}

// compares the logged device values of a replayed plan with the recorded ones
__global__ void rq_validate_log(const uint64_t* log, const uint64_t* expect, int64_t n_words, int32_t* ok) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words && log[i] != expect[i]) *ok = 0;
}

static void read_timings(const std::vector<std::pair<size_t, int>>& ev_used, rq_timings* tm) {
    float ms = 0;
    for (auto& u : ev_used) {
        if (cudaEventElapsedTime(&ms, g_event_pool[u.first].a, g_event_pool[u.first].b) != cudaSuccess) { cudaGetLastError(); continue; }
        if (u.second == 2) { tm->nccl_ms += ms; continue; }
        tm->kernel_ms += ms;
        if (u.second) { tm->scan_kernel_ms += ms; tm->fact_scan_ms = ms; }
        if (g_trace) fprintf(stderr, "[rq] scan kernel launch %zu: %.3f ms%s\n", u.first, ms, u.second ? " (table scan)" : "");
    }
    if (cudaEventElapsedTime(&ms, E.ev[1], E.ev[2]) == cudaSuccess) tm->d2h_ms = ms; else cudaGetLastError();
    if (g_trace && cudaEventElapsedTime(&ms, E.ev[0], E.ev[2]) == cudaSuccess)
        fprintf(stderr, "[rq] plan: %.3f ms on the stream from first launch to result read-back, %.3f ms in pipeline kernels\n", ms, tm->kernel_ms);
}

// One execution of the plan in the mode RP describes. `verdict` (may be null): careful runs report
// whether the run was clean on every rank (fit to be replayed), replays whether every predicted
// value was right on every rank.
static int execute_once(const rq_plan* plan, rq_result** out, rq_timings* tm, bool* verdict) {
    if (tm) memset(tm, 0, sizeof(*tm));
    char* d_strpool = nullptr;
    rq_result* res = nullptr;
    std::vector<void*> scratch;
    const bool sharded = (plan->flags & (RQ_PLAN_SHARDED | RQ_PLAN_PARTITIONED)) && E.dist.comm && E.dist.world > 1;
    const bool partitioned = is_partitioned_plan(*plan);
    unsigned char* d_expect = nullptr;
    int32_t* d_ok = nullptr;
    bool capture_open = RP.capturing;
    try {
        if (RP.mode == 2) {
            // recorded values and log live in the memo (memo_prepare); string constants too
            RP.log_cap = RP.memo->log_cap;
            RP.d_log = RP.memo->d_log;
            d_expect = RP.memo->d_expect;
            d_ok = RP.memo->d_ok;
            d_strpool = RP.memo->d_strpool;
            if (RP.capturing) CK(cudaStreamBeginCapture(E.stream, cudaStreamCaptureModeRelaxed));
            if (RP.log_cap) CK(cudaMemsetAsync(RP.d_log, 0, RP.log_cap, E.stream));
        } else if (plan->strpool_bytes > 0) {
            CK(dmalloc(&d_strpool, plan->strpool_bytes));
            CK(cudaMemcpyAsync(d_strpool, plan->strpool, plan->strpool_bytes, cudaMemcpyHostToDevice, E.stream));
        }
        std::vector<PipeOut> outs(plan->n_pipelines);
        size_t ev_idx = 0;
        std::vector<std::pair<size_t, int>> ev_used;
        double lower_ms = 0;
        g_trace = E.opt.trace && E.dist.rank == 0;
        g_trace_t0 = std::chrono::steady_clock::now();
        record_event(E.ev[0]);
        // sharded plans: merge after the last aggregation (or concatenate the final relation)
        bool final_gather = false;              // the last relation is partitioned over the ranks
        std::vector<void*> gather_owned;        // string bytes a final gather received
        int merge_after = -1;
        if (sharded) {
            for (int pi = 0; pi < plan->n_pipelines; pi++)
                if (plan->pipelines[pi].sink_kind == RQ_SINK_AGG) merge_after = pi;
            if (merge_after < 0) merge_after = plan->n_pipelines - 1;
        }
        int local_code = 0;
        std::string local_msg;
        for (int pi = 0; pi < plan->n_pipelines; pi++) {
            trace_point("pipeline start", pi);
            if (local_code == 0) {
                if (partitioned) {
                    // (exchanges inside the pipeline carry the status of every rank themselves)
                    run_pipeline(*plan, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms);
                } else if (pi <= merge_after) {
                    // a failure on this rank's shard must not leave the other ranks waiting in the merge
                    try {
                        run_pipeline(*plan, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms);
                    } catch (RqError& e) {
                        cudaStreamSynchronize(E.stream);
                        local_code = e.code; local_msg = e.msg;
                    }
                } else {
                    run_pipeline(*plan, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms);
                }
            }
            trace_point("pipeline done", pi);
            if (pi == merge_after) {
                // small partial results (a few groups) are all-gathered and merged on every rank;
                // large ones are hash-partitioned so that each rank merges its share
                const rq_table* part_tab = outs[pi].table.get();
                const bool big = partitioned || (local_code == 0 && part_tab && plan->pipelines[pi].sink_kind == RQ_SINK_AGG &&
                                                 plan->pipelines[pi].n_keys > 0 && !has_str_key(plan->pipelines[pi]) &&
                                                 (part_tab->n_rows < 0 || part_tab->n_rows > kGroupTableCap));
                const bool any_big = agree_on_status(local_code, local_msg, big);
                if (any_big && plan->pipelines[pi].sink_kind == RQ_SINK_AGG && !has_str_key(plan->pipelines[pi])) {
                    merge_partitioned(*plan, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms);
                    final_gather = true;
                } else if (partitioned && plan->pipelines[pi].sink_kind != RQ_SINK_AGG) {
                    final_gather = true;      // no aggregation: the per-rank relations are concatenated at the end
                } else {
                    merge_sharded(*plan, pi, outs, d_strpool, tm, ev_idx, ev_used, lower_ms);
                }
                trace_point("sharded merge done", pi);
            }
        }

        rq_table* fin = outs[plan->n_pipelines - 1].table.get();
        if (!fin) raise(RQ_ERR_INVALID, "last pipeline must produce a relation");
        int64_t n = fin->n_rows;
        if (n < 0) host_read(&n, fin->d_n_rows, 8);
        const int ncols = (int)fin->cols.size();
        trace_point("row count read");
        // ORDER BY + LIMIT over relation t (n rows): cols = the ordered columns, n_out = rows kept
        // (defer_perm: the columns stay as they are and the permutation is handed to the result kernel)
        bool defer_perm = false;
        const uint32_t* final_perm = nullptr;
        auto order_limit = [&](rq_table* t, int64_t n, std::vector<int64_t*>& cols, int64_t& n_out) {
            const int ncols = (int)t->cols.size();
            cols.assign(ncols, nullptr);
            for (int c = 0; c < ncols; c++) cols[c] = (int64_t*)t->cols[c].d;
            n_out = n;
            if (plan->limit >= 0 && plan->limit < n_out) n_out = plan->limit;
        if (plan->n_order > 0 && n > 1) {
            if (plan->n_order > kMaxSortKeys) raise(RQ_ERR_UNSUPPORTED, "more than %d ORDER BY keys", kMaxSortKeys);
            SortKeys K;
            memset(&K, 0, sizeof(K));
            K.n_keys = plan->n_order;
            for (int k = 0; k < plan->n_order; k++) {
                const int c = plan->order[k].column;
                if (c < 0 || c >= ncols) raise(RQ_ERR_INVALID, "ORDER BY column %d out of range", c);
                int w;
                K.col[k] = cols[c];
                K.is_str[k] = phys_type(t->sql_type[c], t->sql_width[c], &w) == RQ_STR;
                K.desc[k] = plan->order[k].ascending ? 0 : 1;
            }
            uint32_t* perm = nullptr;
            CK(dmalloc(&perm, sizeof(uint32_t) * std::max<int64_t>(n, kBitonicMax)));
            scratch.push_back(perm);
            bool sorted_by_topk = false;
            if (n > kBitonicMax && plan->limit >= 0 && plan->limit <= kBitonicMax / 2 && n <= 0xffffffffLL && E.opt.topk) {
                // ORDER BY ... LIMIT k: radix select of the k-th smallest first-key value (6 passes of
                // 11 bits, all on the device), keep the rows up to it (ties included), sort those.
                uint64_t* k1 = nullptr; uint32_t* hist = nullptr; unsigned long long* st = nullptr; uint32_t* cand = nullptr;
                CK(dmalloc(&k1, sizeof(uint64_t) * n)); scratch.push_back(k1);
                CK(dmalloc(&hist, sizeof(uint32_t) * kSelBins)); scratch.push_back(hist);
                CK(dmalloc(&st, 32)); scratch.push_back(st);
                CK(dmalloc(&cand, sizeof(uint32_t) * kBitonicMax)); scratch.push_back(cand);
                const unsigned eb = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
                rq_sort_iota<<<(unsigned)((n + 255) / 256), 256, 0, E.stream>>>(perm, n);
                rq_sort_make_keys<<<(unsigned)((n + 255) / 256), 256, 0, E.stream>>>(K.col[0], perm, k1, n, K.is_str[0], 0, K.desc[0]);
                const unsigned long long init[4] = {0ULL, (unsigned long long)std::max<int64_t>(plan->limit, 1), 0ULL, 0ULL};
                upload_small(st, init, 32);
                CK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * kSelBins, E.stream));
                int shift = 64;
                while (shift > 0) {
                    const int bits = std::min(kSelBits, shift);
                    shift -= bits;
                    rq_topk_hist<<<eb, 256, 0, E.stream>>>(k1, n, st, shift, bits, hist);
                    rq_topk_pick<<<1, 1024, 0, E.stream>>>(hist, st, bits);
                    if (tm) tm->kernel_launches += 2;
                }
                rq_topk_compact<<<eb, 256, 0, E.stream>>>(k1, n, st, cand, kBitonicMax, st + 2);
                if (tm) tm->kernel_launches += 3;
                unsigned long long n_cand = 0;
                host_read(&n_cand, st + 2, 8);
                if (n_cand <= (unsigned long long)kBitonicMax) {
                    rq_sort_small<<<1, 1024, 0, E.stream>>>(K, (const int64_t*)(st + 2), (int64_t)kBitonicMax, perm, cand);
                    if (tm) tm->kernel_launches++;
                    sorted_by_topk = true;
                }
            }
            if (sorted_by_topk) {
                // perm holds the first rows of the order; LIMIT cuts it below
            } else if (n <= kBitonicMax) {
                rq_sort_small<<<1, 1024, 0, E.stream>>>(K, t->d_n_rows, n, perm, nullptr);
                if (tm) tm->kernel_launches++;
            } else {
                // LSD radix sort, least significant ORDER BY key first; every pass is stable
                if (n > 0xffffffffLL) raise(RQ_ERR_UNSUPPORTED, "ORDER BY over more than 2^32 rows");
                uint32_t* perm2 = nullptr; uint64_t *k1 = nullptr, *k2 = nullptr; uint32_t* hist = nullptr;
                unsigned long long* bits = nullptr;
                const unsigned nb = (unsigned)((n + kRadixChunk - 1) / kRadixChunk);
                const unsigned eb = (unsigned)((n + 255) / 256);
                CK(dmalloc(&perm2, sizeof(uint32_t) * n)); scratch.push_back(perm2);
                CK(dmalloc(&k1, sizeof(uint64_t) * n)); scratch.push_back(k1);
                CK(dmalloc(&k2, sizeof(uint64_t) * n)); scratch.push_back(k2);
                CK(dmalloc(&hist, sizeof(uint32_t) * 16 * nb)); scratch.push_back(hist);
                CK(dmalloc(&bits, 16)); scratch.push_back(bits);
                rq_sort_iota<<<eb, 256, 0, E.stream>>>(perm, n);
                if (tm) tm->kernel_launches++;
                for (int k = plan->n_order - 1; k >= 0; k--) {
                    const int c = plan->order[k].column;
                    const int words = K.is_str[k] ? (t->sql_width[c] + 7) / 8 : 1;
                    for (int w = words - 1; w >= 0; w--) {
                        rq_sort_make_keys<<<eb, 256, 0, E.stream>>>(K.col[k], perm, k1, n, K.is_str[k], w, K.desc[k]);
                        const unsigned long long init[2] = {0ULL, ~0ULL};
                        upload_small(bits, init, 16);
                        rq_sort_key_bits<<<std::min(eb, 1024u), 256, 0, E.stream>>>(k1, n, bits);
                        unsigned long long h_bits[2];
                        host_read(h_bits, bits, 16);
                        if (tm) tm->kernel_launches += 2;
                        const uint64_t varying = h_bits[0] ^ h_bits[1];
                        for (int shift = 0; shift < 64; shift += 4) {
                            if (((varying >> shift) & 15) == 0) continue;
                            rq_radix_hist<<<nb, kRadixThreads, 0, E.stream>>>(k1, n, shift, hist);
                            rq_radix_scan<<<1, 1024, 0, E.stream>>>(hist, (int64_t)16 * nb);
                            rq_radix_scatter<<<nb, kRadixThreads, 0, E.stream>>>(k1, perm, k2, perm2, n, shift, hist);
                            if (tm) tm->kernel_launches += 3;
                            std::swap(k1, k2);
                            std::swap(perm, perm2);
                        }
                    }
                }
                CK(cudaGetLastError());
            }
            if (defer_perm) {
                final_perm = perm;           // applied by rq_finish_result together with narrowing
            } else {
                for (int c = 0; c < ncols; c++) {
                    int64_t* sorted = nullptr;
                    CK(dmalloc(&sorted, sizeof(int64_t) * std::max<int64_t>(n_out, 1)));
                    scratch.push_back(sorted);
                    rq_apply_perm<<<(unsigned)((n_out + 255) / 256), 256, 0, E.stream>>>(cols[c], sorted, perm, t->d_n_rows, n, plan->limit);
                    if (tm) tm->kernel_launches++;
                    cols[c] = sorted;
                }
                CK(cudaGetLastError());
            }
        }

        };
        std::vector<int64_t*> cols;
        int64_t n_out = 0;
        std::unique_ptr<rq_table> gathered;
        if (final_gather) {
            // The relation is hash-partitioned over the ranks (merge_partitioned): with ORDER BY ...
            // LIMIT k every rank contributes only its first k rows, then the concatenation is ordered.
            rq_table part;
            std::vector<int64_t*> pc;
            int64_t pn = n;
            if (plan->n_order > 0 && plan->limit >= 0) order_limit(fin, n, pc, pn);
            else { pc.resize(ncols); for (int c = 0; c < ncols; c++) pc[c] = (int64_t*)fin->cols[c].d; }
            part.n_rows = pn;
            part.cap_rows = pn;
            for (int c = 0; c < ncols; c++) {
                DevColumn dc;
                dc.type = RQ_I64; dc.width = 8; dc.d = (unsigned char*)pc[c]; dc.owned = false; dc.tile_stride = (int64_t)kTile * 8;
                part.cols.push_back(dc);
            }
            part.sql_type = fin->sql_type;
            part.sql_width = fin->sql_width;
            gathered = gather_relation(part, tm, gather_owned, ev_idx, ev_used);
            fin = gathered.get();
            n = fin->n_rows;
        }
        defer_perm = true;
        order_limit(fin, n, cols, n_out);

        trace_point("sorted");
        // one kernel permutes, cuts, narrows to the reference's physical widths and copies strings by
        // value for all columns into one packed buffer; one copy brings it to the host
        ResultBox* box = (ResultBox*)calloc(1, sizeof(ResultBox));
        res = &box->r;
        res->n_rows = n_out;
        res->n_cols = ncols;
        res->cols = (rq_result_col*)calloc(ncols, sizeof(rq_result_col));
        if (ncols > kMaxOut) raise(RQ_ERR_UNSUPPORTED, "more than %d result columns", kMaxOut);
        FinishCols F;
        memset(&F, 0, sizeof(F));
        F.ncols = ncols;
        size_t total = 0;
        if (RP.capturing) { RP.memo->cols.assign(ncols, PlanMemo::Col()); RP.memo->res_rows = n_out; }
        for (int c = 0; c < ncols; c++) {
            int w = 8;
            const int pt = phys_type(fin->sql_type[c], fin->sql_width[c], &w);
            rq_result_col& rc = res->cols[c];
            rc.type = pt; rc.width = w; rc.sql_type = fin->sql_type[c]; rc.sql_width = fin->sql_width[c];
            if (w > 0xffff) raise(RQ_ERR_UNSUPPORTED, "result column %d is %d bytes wide", c, w);
            F.in[c] = cols[c];
            F.off[c] = (uint64_t)total;
            F.width[c] = (uint16_t)w;
            F.kind[c] = pt == RQ_I64 ? 0 : pt == RQ_I32 ? 1 : pt == RQ_I8 ? 2 : 3;
            if (RP.capturing) {
                PlanMemo::Col& mc = RP.memo->cols[c];
                mc.type = pt; mc.width = w; mc.sql_type = rc.sql_type; mc.sql_width = rc.sql_width; mc.off = total;
            }
            total = (total + (size_t)std::max<int64_t>(n_out, 1) * w + 15) & ~(size_t)15;
        }
        box->block = (unsigned char*)malloc(std::max<size_t>(total, 16));
        for (int c = 0; c < ncols; c++) res->cols[c].data = box->block + F.off[c];
        unsigned char* d_pack = nullptr;
        if (n_out > 0) {
            CK(dmalloc(&d_pack, total));
            scratch.push_back(d_pack);
            const unsigned blocks = (unsigned)std::min<int64_t>((n_out + 255) / 256, 148 * 16);
            rq_finish_result<<<blocks, 256, 0, E.stream>>>(F, final_perm, fin->d_n_rows, n, plan->limit, d_pack);
            if (tm) tm->kernel_launches++;
        }
        CK(cudaGetLastError());
        record_event(E.ev[1]);
        if (RP.capturing) {
            RP.memo->pack_bytes = total;
            CK(cudaMallocHost(&RP.memo->h_pack, std::max<size_t>(total, 16)));
        }
        if (n_out > 0)
            CK(cudaMemcpyAsync(RP.capturing ? (void*)RP.memo->h_pack : (void*)box->block, d_pack, total, cudaMemcpyDeviceToHost, E.stream));
        // the verdict travels with the result: replay = all predictions right, careful = run was clean
        // (sharded plans: the minimum over the ranks, so every rank draws the same conclusion)
        int32_t h_ok_local = RP.retries == 0 ? 1 : 0;
        if (RP.mode == 2 || sharded) {
            if (!d_ok) {
                CK(dmalloc(&d_ok, 8));
                scratch.push_back(d_ok);
            }
            upload_small(d_ok, &h_ok_local, 4);
            if (RP.mode == 2) {
                if (RP.idx != RP.memo->reads.size()) throw ReplayDiverged{};
                const int64_t words = (int64_t)(RP.log_cap / 8);
                if (words > 0)
                    rq_validate_log<<<(unsigned)((words + 255) / 256), 256, 0, E.stream>>>((const uint64_t*)RP.d_log, (const uint64_t*)d_expect, words, d_ok);
            }
            if (sharded) {
                const int rc = E.dist.all_reduce(d_ok, d_ok, 1, 2 /* ncclInt32 */, 3 /* ncclMin */, E.dist.comm, E.stream);
                if (rc != 0) raise(RQ_ERR_NCCL, "ncclAllReduce(verdict) failed");
            }
            CK(cudaMemcpyAsync(g_pinned, d_ok, 4, cudaMemcpyDeviceToHost, E.stream));
        }
        record_event(E.ev[2]);
        // everything that lives on the device is released here, in stream order (inside a capture the
        // frees must be part of the graph)
        outs.clear();
        gathered.reset();
        for (void* p : scratch) dfree(p);
        scratch.clear();
        for (void* p : gather_owned) dfree(p);
        gather_owned.clear();
        if (d_strpool && RP.mode != 2) { dfree(d_strpool); d_strpool = nullptr; }
        if (RP.capturing) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamEndCapture(E.stream, &graph));
            capture_open = false;
            cudaError_t ge = cudaGraphInstantiate(&RP.memo->gexec, graph, 0);
            cudaGraphDestroy(graph);
            if (ge != cudaSuccess) { RP.memo->gexec = nullptr; raise(RQ_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ge)); }
            CK(cudaGraphLaunch(RP.memo->gexec, E.stream));
        }
        stream_sync();
        if (RP.capturing && n_out > 0) memcpy(box->block, RP.memo->h_pack, RP.memo->pack_bytes);
        if (verdict) *verdict = (RP.mode == 2 || sharded) ? (*(const int32_t*)g_pinned != 0) : (h_ok_local != 0);
        if (tm) {
            tm->lower_ms = lower_ms;
            tm->host_syncs = RP.syncs;
            read_timings(ev_used, tm);
        }
        if (RP.capturing) { RP.memo->ev_used = ev_used; if (tm) RP.memo->tm_static = *tm; }
        *out = res;
        return RQ_OK;
    } catch (RqError& e) {
        if (capture_open) { cudaGraph_t g = nullptr; cudaStreamEndCapture(E.stream, &g); if (g) cudaGraphDestroy(g); }
        cudaStreamSynchronize(E.stream);
        for (void* p : scratch) dfree(p);
        if (d_strpool && RP.mode != 2) dfree(d_strpool);
        rq_result_free(res);
        return fail(e.code, "%s", e.msg.c_str());
    } catch (ReplayDiverged&) {
        // the recorded script does not fit this execution (only possible on a single rank: sharded
        // replays take identical decisions on every rank): discard, the caller runs the careful way
        if (capture_open) { cudaGraph_t g = nullptr; cudaStreamEndCapture(E.stream, &g); if (g) cudaGraphDestroy(g); }
        cudaStreamSynchronize(E.stream);
        for (void* p : scratch) dfree(p);
        rq_result_free(res);
        if (verdict) *verdict = false;
        *out = nullptr;
        return RQ_OK;
    }
}

static uint64_t plan_signature(const rq_plan& plan) {
    uint64_t h = 0xcbf29ce484222325ULL;
    auto mixin = [&](const void* p, size_t n) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ULL; }
    };
    mixin(&plan.n_tables, sizeof plan.n_tables);
    for (int t = 0; t < plan.n_tables; t++) {
        const rq_table* tb = plan.tables[t];
        if (!tb) continue;
        // table identity: the contents of a table never change after its upload (a new version is a
        // new upload, like the reference's append-only relations), so predictions recorded for this
        // handle stay exact; another table under the same name is another handle
        mixin(&tb->uid, sizeof tb->uid);
        mixin(&tb->n_rows, sizeof tb->n_rows);
        mixin(&tb->borrowed, sizeof tb->borrowed);
        for (auto& c : tb->cols) { mixin(&c.type, sizeof c.type); mixin(&c.width, sizeof c.width); mixin(&c.page_off, sizeof c.page_off); }
    }
    for (int pi = 0; pi < plan.n_pipelines; pi++) {
        const rq_pipeline& pl = plan.pipelines[pi];
        const uint64_t s = pipeline_signature(pl, pl.size_hint);
        mixin(&s, sizeof s);
        mixin(pl.args, sizeof(int32_t) * pl.n_args);
        mixin(&pl.source_id2, sizeof pl.source_id2);
    }
    mixin(plan.order, sizeof(rq_order_key) * plan.n_order);
    mixin(&plan.limit, sizeof plan.limit);
    mixin(&plan.flags, sizeof plan.flags);
    if (plan.strpool_bytes > 0) mixin(plan.strpool, (size_t)plan.strpool_bytes);
    mixin(&E.dist.world, sizeof E.dist.world);
    mixin(&E.opt.split_min_rows, sizeof E.opt.split_min_rows);
    mixin(&E.opt.split_frac, sizeof E.opt.split_frac);
    mixin(&E.opt.zone_skip, sizeof E.opt.zone_skip);
    mixin(&E.opt.share_builds, sizeof E.opt.share_builds);
    mixin(&E.opt.share_min_rows, sizeof E.opt.share_min_rows);
    mixin(&E.opt.stages, sizeof E.opt.stages);
    mixin(&E.opt.warps, sizeof E.opt.warps);
    return h;
}

// device-side copies of what a replay needs: the recorded values (flat, 8-byte aligned), the log, the
// verdict word and the plan's string constants
static bool memo_prepare(PlanMemo& memo, const rq_plan& plan) {
    std::vector<unsigned char> flat;
    for (auto& r : memo.reads) {
        flat.insert(flat.end(), r.begin(), r.end());
        flat.resize((flat.size() + 7) & ~(size_t)7, 0);
    }
    memo.log_cap = flat.size();
    bool ok = true;
    ok &= cudaMalloc(&memo.d_expect, std::max<size_t>(flat.size(), 8)) == cudaSuccess;
    ok &= cudaMalloc(&memo.d_log, std::max<size_t>(flat.size(), 8)) == cudaSuccess;
    ok &= cudaMalloc(&memo.d_ok, 8) == cudaSuccess;
    if (ok && !flat.empty()) ok &= cudaMemcpy(memo.d_expect, flat.data(), flat.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && plan.strpool_bytes > 0) {
        ok &= cudaMalloc(&memo.d_strpool, (size_t)plan.strpool_bytes) == cudaSuccess;
        if (ok) ok &= cudaMemcpy(memo.d_strpool, plan.strpool, (size_t)plan.strpool_bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) { cudaGetLastError(); memo.release(); }
    return ok;
}

static void reset_plan_memos() {
    for (auto& m : g_plan_memo) m.second.release();
    g_plan_memo.clear();
    g_ht_capacity.clear();
    g_emit_rows.clear();
    g_needs_expand.clear();
    g_agg_impl.clear();
    g_no_direct.clear();
    RP = ReplayState();
}

extern "C" int rq_plan_execute(const rq_plan* plan, rq_result** out, rq_timings* tm) {
    if (!E.init) return fail(RQ_ERR_NOT_INIT, "rq_plan_execute before rq_init");
    if (!plan || !out || plan->n_pipelines <= 0 || !plan->pipelines)
        return fail(RQ_ERR_INVALID, "rq_plan_execute: bad arguments");
    *out = nullptr;
    if (!E.opt.replay) {
        RP = ReplayState();
        const int rc = execute_once(plan, out, tm, nullptr);
        RP = ReplayState();
        return rc;
    }
    PlanMemo& memo = g_plan_memo[plan_signature(*plan)];
    if (memo.valid && memo.gexec) {
        // the whole plan is one graph launch
        if (tm) memset(tm, 0, sizeof(*tm));
        cudaError_t ge = cudaGraphLaunch(memo.gexec, E.stream);
        if (ge == cudaSuccess) ge = cudaStreamSynchronize(E.stream);
        if (ge != cudaSuccess) return fail(RQ_ERR_CUDA, "graph launch failed: %s", cudaGetErrorString(ge));
        if (*(const int32_t*)g_pinned != 0) {
            ResultBox* box = (ResultBox*)calloc(1, sizeof(ResultBox));
            rq_result* res = &box->r;
            res->n_rows = memo.res_rows;
            res->n_cols = (int)memo.cols.size();
            res->cols = (rq_result_col*)calloc(memo.cols.size(), sizeof(rq_result_col));
            box->block = (unsigned char*)malloc(std::max<size_t>(memo.pack_bytes, 16));
            if (memo.res_rows > 0) memcpy(box->block, memo.h_pack, memo.pack_bytes);
            for (size_t c = 0; c < memo.cols.size(); c++) {
                const PlanMemo::Col& mc = memo.cols[c];
                rq_result_col& rc = res->cols[c];
                rc.type = mc.type; rc.width = mc.width; rc.sql_type = mc.sql_type; rc.sql_width = mc.sql_width;
                rc.data = box->block + mc.off;
            }
            if (tm) {
                tm->lower_ms = 0;
                tm->kernel_launches = memo.tm_static.kernel_launches;
                tm->host_syncs = 1;
                g_trace = false;
                read_timings(memo.ev_used, tm);
            }
            *out = res;
            return RQ_OK;
        }
        if (E.opt.trace) fprintf(stderr, "[rq] rank %d: graph result discarded (a predicted value was wrong on some rank)\n", E.dist.rank);
        memo.release();
    }
    if (memo.valid) {
        RP = ReplayState();
        RP.mode = 2;
        RP.memo = &memo;
        // the second replay is captured into a graph (the first proved the script right)
        RP.capturing = E.opt.graphs && !memo.no_graph && memo.replays >= 1;
        bool ok = false;
        const int rc = execute_once(plan, out, tm, &ok);
        const bool captured = RP.capturing;
        RP = ReplayState();
        if (rc != RQ_OK) {
            if (captured) {            // capture problems are not plan problems: fall back to plain replays
                cudaGetLastError();
                if (memo.gexec) { cudaGraphExecDestroy(memo.gexec); memo.gexec = nullptr; }
                memo.no_graph = true;
                return rq_plan_execute(plan, out, tm);
            }
            memo.release();
            return rc;
        }
        if (ok) { memo.replays++; return RQ_OK; }
        if (E.opt.trace) fprintf(stderr, "[rq] rank %d: replayed plan discarded (%s); running the careful way\n", E.dist.rank,
                                 *out ? "a predicted value was wrong on some rank" : "the recorded script did not fit");
        // some predicted value was wrong (a full table, a runtime error ...): drop the result and run
        // the careful way, which also reports errors
        if (*out) { rq_result_free(*out); *out = nullptr; }
        memo.release();
    }
    memo.reads.clear();
    RP = ReplayState();
    RP.mode = 1;
    RP.memo = &memo;
    bool clean = false;
    const int rc = execute_once(plan, out, tm, &clean);
    memo.valid = rc == RQ_OK && clean;
    if (memo.valid && !memo_prepare(memo, *plan)) memo.valid = false;
    if (E.opt.trace) fprintf(stderr, "[rq] rank %d: careful run rc=%d clean=%d retries=%d (%s) reads=%zu\n", E.dist.rank, rc, (int)clean, RP.retries, RP.why, memo.reads.size());
    if (!memo.valid) memo.release();
    RP = ReplayState();
    return rc;
}

// ------------------------------------------------------------------------------------------
// debug aid (no GPU needed): lowers one pipeline of a plan against column metadata only and
// prints the device program. tests/test_lowering.py runs the printed program through a Python
// model of the accumulator machine and compares with the plan oracle, so the host-side lowering
// is covered on CPU-only CI. Not part of the public ABI.
// ------------------------------------------------------------------------------------------
extern "C" int rq_debug_lower(const rq_plan* plan, int pi, int impl, const int32_t* col_types,
                              const int32_t* col_widths, const int64_t* col_min, const int64_t* col_max,
                              int n_cols, char* buf, int64_t buflen) {
    try {
        if (!plan || pi < 0 || pi >= plan->n_pipelines) return fail(RQ_ERR_INVALID, "rq_debug_lower: bad pipeline");
        SimplePipe sp;
        simplify_pipeline(plan->pipelines[pi], sp);
        rq_table fake;
        fake.n_rows = 0;
        for (int c = 0; c < n_cols; c++) {
            DevColumn dc;
            dc.type = col_types[c]; dc.width = col_widths[c]; dc.d = nullptr; dc.owned = false;
            dc.tile_stride = (int64_t)kTile * dc.width;
            if (col_min && col_max && dc.type != RQ_STR && col_min[c] <= col_max[c]) {
                dc.has_stats = true; dc.vmin = col_min[c]; dc.vmax = col_max[c];
            }
            fake.cols.push_back(dc);
        }
        std::vector<PipeOut> outs(plan->n_pipelines);
        KParams P;
        memset(&P, 0, sizeof(P));
        P.expand_probe = -1;
        // string constants are lowered to pool base + offset; a recognisable fake base lets the Python
        // model of the VM (tests/vm_model.py) tell them from numeric constants
        Lowerer L(*plan, sp.pl, fake, outs, (const char*)(uintptr_t)(1ULL << 44), P);
        L.prepare();
        AggDedup ad;
        if (sp.pl.sink_kind == RQ_SINK_AGG) ad = dedup_aggs(sp.pl);
        emit_program(L, impl, ad);
        // the device encoding must succeed as well (layout + fusion), even though the model
        // executes the host-level form
        P.na = (int)ad.kind.size();
        KeyUnpack ku;
        const bool packed = pack_group_key(L, P, ku);
        const int gr = impl == IMPL_REGAGG ? (sp.pl.n_keys == 0 ? 1 : kRegGroups) : 0;
        if (!layout_smem(P, L.n_slots, impl == IMPL_LOWAGG ? P.na * 32 * 8 : 0, max_warps_of(gr)))
            raise(RQ_ERR_UNSUPPORTED, "pipeline does not fit in shared memory");
        encode_program(L, P);
        std::string s;
        char line[256];
        snprintf(line, sizeof line, "cols %d strcols %d slots %d insn %d nk %d nout %d uinsn %d warps %d stages %d packed %d\n",
                 P.n_cols, P.n_strcols, L.n_slots, (int)L.prog.size(), (int)L.hkey.size(), (int)L.hout.size(),
                 P.n_insn, P.warps, P.stages, packed ? 1 : 0);
        s += line;
        for (int c = 0; c < P.n_cols; c++) {
            int srccol = -1;
            for (size_t k = 0; k < L.staged_of_col.size(); k++)
                if (fake.cols[k].type != RQ_STR && L.staged_of_col[k] == c) srccol = (int)k;
            snprintf(line, sizeof line, "col %d src %d w %d\n", c, srccol, (int)P.col_w[c]);
            s += line;
            snprintf(line, sizeof line, "coloff %d %u\n", c, P.col_off[c]);
            s += line;
        }
        snprintf(line, sizeof line, "layout %u %u\n", P.stage_bytes, P.slots_rel);
        s += line;
        for (int c = 0; c < P.n_strcols; c++) {
            int srccol = -1;
            for (size_t k = 0; k < L.staged_of_col.size(); k++)
                if (fake.cols[k].type == RQ_STR && L.staged_of_col[k] == c) srccol = (int)k;
            snprintf(line, sizeof line, "strcol %d src %d w %u\n", c, srccol, P.str_w[c]);
            s += line;
        }
        for (size_t i = 0; i < L.prog.size(); i++) {
            const HUnit& h = L.prog[i];
            snprintf(line, sizeof line, "unit %d %d %d %d %lld %d %d %lld %d %d %lld %lld %d %d %d %lld %d\n", h.op, h.gop,
                     h.x.kind, h.x.idx, (long long)h.x.imm, h.y.kind, h.y.idx, (long long)h.y.imm,
                     h.z.kind, h.z.idx, (long long)h.z.imm, (long long)h.imm, h.dst, h.filt ? 1 : 0, h.aux,
                     (long long)h.imm2, h.n32 ? 1 : 0);
            s += line;
        }
        for (int i = 0; i < P.n_insn; i++) {
            const UInsn& u = P.insn[i];
            snprintf(line, sizeof line, "uinsn %d %d %u %d %d %d %d %d %u %u %u %lld\n", u.code, u.flags, u.dstrel, u.aux, u.gop,
                     u.xkind, u.ykind, u.zkind, u.xrel, u.yrel, u.zrel, (long long)u.imm);
            s += line;
        }
        // the device-side value references of the sinks (to_vref): kind, slot / u32 bits, offset >> 4
        for (int k = 0; k < P.nk; k++) { snprintf(line, sizeof line, "vkey %d %d %u\n", P.key[k].kind, P.key[k].slot, P.key[k].off16); s += line; }
        for (int k = 0; k < P.n_out; k++) { snprintf(line, sizeof line, "vout %d %d %u\n", P.out[k].kind, P.out[k].slot, P.out[k].off16); s += line; }
        for (size_t u = 0; u < ad.kind.size(); u++) { snprintf(line, sizeof line, "vagg %d %d %u\n", P.agg_src[u].kind, P.agg_src[u].slot, P.agg_src[u].off16); s += line; }
        for (size_t k = 0; k < L.hkey.size(); k++) { snprintf(line, sizeof line, "key %d %d\n", L.hkey[k].kind, L.hkey[k].idx); s += line; }
        for (size_t k = 0; k < L.hout.size(); k++) { snprintf(line, sizeof line, "out %d %d\n", L.hout[k].kind, L.hout[k].idx); s += line; }
        for (int k = 0; k < kMaxImm; k++) { snprintf(line, sizeof line, "imm %d %lld\n", k, (long long)P.imm[k]); s += line; }
        for (size_t u = 0; u < ad.kind.size(); u++) { snprintf(line, sizeof line, "agg %d %d\n", (int)u, ad.kind[u]); s += line; }
        for (size_t u = 0; u < ad.kind.size(); u++) { snprintf(line, sizeof line, "aggsrc %d %d %d\n", L.hagg_src[u].kind, L.hagg_src[u].idx, L.hagg_src[u].u32); s += line; }
        for (size_t k = 0; k < ad.uniq_of.size(); k++) { snprintf(line, sizeof line, "aggmap %d %d\n", (int)k, ad.uniq_of[k]); s += line; }
        if ((int64_t)s.size() + 1 > buflen) return fail(RQ_ERR_INVALID, "rq_debug_lower: buffer too small");
        memcpy(buf, s.c_str(), s.size() + 1);
        return RQ_OK;
    } catch (RqError& e) {
        return fail(e.code, "%s", e.msg.c_str());
    }
}
