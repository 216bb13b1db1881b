// Fused scan -> filter -> project -> (probe | aggregate | build | materialize) kernel for sm_100a.
//
// One persistent CTA per SM; every WARP is an independent pipeline. A warp owns a ring of TMA
// stages in shared memory (cp.async.bulk + one mbarrier per stage); each stage holds one warp
// tile = 256 tuples of every scanned column. The warp interprets the typed program over a
// register tile of 8 tuples per lane with ONE switch per instruction (opcode and operand form are
// fused on the host, operands are precomputed shared-memory offsets), then re-arms the stage. No
// CTA-wide barrier exists in the steady state, so warps drift out of phase and the copy engine
// always has work.
//
// Aggregation (template parameter GR):
//   GR = 1 / 4  register path: up to GR groups x kNAR aggregates live in registers; a tuple is
//               added to every group accumulator through a 0/1 multiplier (two IMADs per 64-bit
//               add, on the FMA pipe, no shared-memory traffic and no dependent memory chain).
//   GR = 0      generic path: lane-private shared-memory accumulators for up to 8 groups per warp,
//               HBM hash aggregation, hash-join build/probe, materialize.
//
// Semantics restated from the reference (Henning1/resql):
//   arithmetic / compares  src/ExpressionsJitFlounder.h:298-689
//   selection              src/operators/selection.h:52-70
//   aggregation            src/operators/aggregation.h:95-152, :240-295
//   hash join              src/operators/hashjoin.h:118-279
//   string compares        src/qlib/scalar.h:16-120
#pragma once
#include <cuda_runtime.h>
#include "rq_internal.h"
#include "device_util.cuh"
#include "hash_kernels.cuh"

namespace rq {

constexpr unsigned kFull = 0xffffffffu;

// tuple r of lane l sits at row 64*(r/2) + 2*l + (r%2) of the warp tile
__device__ __forceinline__ int row_in_tile(int r, int lane) {
    return (r >> 1) * 64 + 2 * lane + (r & 1);
}

// ---- register tile <-> shared memory ---------------------------------------------------------
__device__ __forceinline__ void ld_m64(const unsigned char* b, int lane, int64_t (&v)[kR]) {
    const longlong2* p = reinterpret_cast<const longlong2*>(b) + lane;
#pragma unroll
    for (int k = 0; k < kR / 2; k++) {
        const longlong2 x = p[k * 32];
        v[2 * k] = x.x; v[2 * k + 1] = x.y;
    }
}
__device__ __forceinline__ void st_m64(unsigned char* b, int lane, const int64_t (&v)[kR]) {
    longlong2* p = reinterpret_cast<longlong2*>(b) + lane;
#pragma unroll
    for (int k = 0; k < kR / 2; k++) p[k * 32] = make_longlong2(v[2 * k], v[2 * k + 1]);
}
__device__ __forceinline__ void ld_m32(const unsigned char* b, int lane, int32_t (&v)[kR]) {
    const int2* p = reinterpret_cast<const int2*>(b) + lane;
#pragma unroll
    for (int k = 0; k < kR / 2; k++) {
        const int2 x = p[k * 32];
        v[2 * k] = x.x; v[2 * k + 1] = x.y;
    }
}
__device__ __forceinline__ void ld_m8(const unsigned char* b, int lane, uint32_t (&v)[kR]) {
    const uchar2* p = reinterpret_cast<const uchar2*>(b) + lane;
#pragma unroll
    for (int k = 0; k < kR / 2; k++) {
        const uchar2 x = p[k * 32];
        v[2 * k] = x.x; v[2 * k + 1] = x.y;
    }
}

struct WarpCtx {
    const unsigned char* stage;   // current stage (columns of the current tile)
    unsigned char*       wbase;   // this warp's shared-memory region
    int64_t              row0;    // first tuple of the tile in the source
    int                  lane;
};

__device__ __forceinline__ const unsigned char* opnd_base(const WarpCtx& c, bool slot, uint32_t off) {
    return (slot ? c.wbase : c.stage) + off;
}

// any operand kind -> 8 int64 values (rare forms, group keys)
__device__ __forceinline__ void fetch(const KParams& P, const WarpCtx& c, int kind,
                                      const unsigned char* ob, int stridx, int64_t imm,
                                      int64_t (&v)[kR]) {
    switch (kind) {
        case K_M64: ld_m64(ob, c.lane, v); break;
        case K_M32: {
            int32_t t[kR]; ld_m32(ob, c.lane, t);
#pragma unroll
            for (int r = 0; r < kR; r++) v[r] = t[r];
            break;
        }
        case K_M8: {
            uint32_t t[kR]; ld_m8(ob, c.lane, t);
#pragma unroll
            for (int r = 0; r < kR; r++) v[r] = t[r];
            break;
        }
        case K_STR:
#pragma unroll
            for (int r = 0; r < kR; r++)
                v[r] = (int64_t)(P.str_ptr[stridx] + (size_t)(c.row0 + row_in_tile(r, c.lane)) * P.str_w[stridx]);
            break;
        default:
#pragma unroll
            for (int r = 0; r < kR; r++) v[r] = imm;
            break;
    }
}

// one tuple of a sink value
__device__ __forceinline__ int64_t ld_row(const KParams& P, const WarpCtx& c, VRef vr, int r) {
    const int row = row_in_tile(r, c.lane);
    const unsigned char* b = (vr.slot ? c.wbase : c.stage) + ((uint32_t)vr.off16 << 4);
    switch (vr.kind) {
        case K_M64: return reinterpret_cast<const int64_t*>(b)[row];
        case K_M32: return reinterpret_cast<const int32_t*>(b)[row];
        case K_M8:  return b[row];
        case K_IMM: return P.imm[vr.off16];
        case K_STR: return (int64_t)(P.str_ptr[vr.off16] + (size_t)(c.row0 + row) * P.str_w[vr.off16]);
        default:    return 0;
    }
}

// acc += v * m for a 0/1 multiplier m (exact mod 2^64): IMAD.WIDE.U32 accumulates the low word
// with carry, the high word is one more IMAD - FMA-pipe work, no shared memory, no branches.
__device__ __forceinline__ void macc(uint64_t& acc, int64_t v, uint32_t m) {
    acc += (uint64_t)v * (uint64_t)m;
}

// ---- low-cardinality global group table (packed key) -----------------------------------------
__device__ __forceinline__ int group_table_slot(const KParams& P, uint64_t key) {
    uint64_t hh = mix64(key ^ 0x9E3779B97F4A7C15ULL);
    uint32_t i = (uint32_t)(hh & (kGroupTableCap - 1));
    for (int tries = 0; tries < kGroupTableCap; tries++) {
        uint32_t st = atomicCAS(&P.g_state[i], 0u, 1u);
        if (st == 0u) {
            ((volatile int64_t*)P.g_keys)[i] = (int64_t)key;
            __threadfence();
            atomicExch(&P.g_state[i], 2u);
            return (int)i;
        }
        while (st == 1u) st = *(volatile uint32_t*)&P.g_state[i];
        __threadfence();
        if ((uint64_t)((volatile int64_t*)P.g_keys)[i] == key) return (int)i;
        i = (i + 1) & (kGroupTableCap - 1);
    }
    return -1;
}

__device__ __forceinline__ int64_t warp_reduce(int64_t v, int kind) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int64_t w = __shfl_xor_sync(kFull, v, o);
        if (kind == 3) v = w < v ? w : v;
        else if (kind == 4) v = w > v ? w : v;
        else v = (int64_t)((uint64_t)v + (uint64_t)w);
    }
    return v;
}

__device__ __forceinline__ void group_table_add(const KParams& P, int slot, int a, int kind, int64_t v) {
    int64_t* dst = &P.g_acc[(size_t)slot * kMaxAggs + a];
    if (kind == 3) atomicMin((long long*)dst, (long long)v);
    else if (kind == 4) atomicMax((long long*)dst, (long long)v);
    else atomicAdd((unsigned long long*)dst, (unsigned long long)v);
}

// ------------------------------------------------------------------------------------------
// the scan kernel
// ------------------------------------------------------------------------------------------
extern __shared__ __align__(128) unsigned char rq_smem[];

#define RQ_EX_ADD(x, y)  ((int64_t)((uint64_t)(x) + (uint64_t)(y)))
#define RQ_EX_SUB(x, y)  ((int64_t)((uint64_t)(x) - (uint64_t)(y)))
#define RQ_EX_RSUB(x, y) ((int64_t)((uint64_t)(y) - (uint64_t)(x)))
#define RQ_EX_MUL(x, y)  ((int64_t)((uint64_t)(x) * (uint64_t)(y)))
#define RQ_EX_AND(x, y)  ((x) & (y))
#define RQ_EX_OR(x, y)   ((x) | (y))
#define RQ_EX_LT(x, y)   ((int64_t)((x) < (y)))
#define RQ_EX_LE(x, y)   ((int64_t)((x) <= (y)))
#define RQ_EX_GT(x, y)   ((int64_t)((x) > (y)))
#define RQ_EX_GE(x, y)   ((int64_t)((x) >= (y)))
#define RQ_EX_EQ(x, y)   ((int64_t)((x) == (y)))
#define RQ_EX_NE(x, y)   ((int64_t)((x) != (y)))

// The aggregate-index switches must stay switches over compile-time register names: an inline
// asm marker that differs per case keeps the compiler from merging the cases into one body that
// indexes the accumulator array dynamically (which would demote it to local memory).
#define RQ_NOMERGE(A) asm volatile("// agg case %0" ::"n"(A))

static_assert(kNAR == 6, "the aggregate-index switches list cases 0..5");
template <int GR> struct ScanCfg;
template <> struct ScanCfg<0> { static constexpr int kThreads = 512; };
template <> struct ScanCfg<1> { static constexpr int kThreads = 512; };
template <> struct ScanCfg<4> { static constexpr int kThreads = 384; };

template <int GR>
__global__ void __launch_bounds__(ScanCfg<GR>::kThreads, 1)
rq_scan_kernel(const __grid_constant__ KParams P) {
    constexpr int NG = GR > 0 ? GR : 1;                 // register groups
    constexpr int ND = GR > 0 ? GR : kLowCardMaxGroups; // dictionary entries
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const int S = P.stages;

    uint64_t* bars = reinterpret_cast<uint64_t*>(rq_smem) + warp * kMaxStages;
    unsigned char* wbase = rq_smem + P.warp_off + (size_t)warp * P.warp_bytes;
    unsigned char* slot_base = wbase + P.slots_rel;
    int64_t* sacc = reinterpret_cast<int64_t*>(wbase + P.acc_rel);   // GR == 0 low-card path

    const int64_t n_rows = P.n_rows_ptr ? *P.n_rows_ptr : P.n_rows;
    const int64_t n_tiles = (n_rows + kTile - 1) / kTile;
    const int64_t stride = (int64_t)gridDim.x * W;
    const int64_t first = (int64_t)blockIdx.x * W + warp;
    const int NA = P.na, NK = P.nk;
    const bool lowagg = (GR > 0) || (P.G > 0);

    if (lane == 0) {
        for (int s = 0; s < S; s++) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    // accumulators
    uint64_t racc[NG][kNAR];
    uint32_t m[NG][kR];
    uint64_t dk[ND];
    int ngroups = 0;
    unsigned seen = 0;     // (no GROUP BY) did this lane aggregate at least one tuple
#pragma unroll
    for (int g = 0; g < NG; g++) {
#pragma unroll
        for (int a = 0; a < kNAR; a++) racc[g][a] = (uint64_t)agg_identity(a < NA ? P.agg_kind[a] : 0);
#pragma unroll
        for (int r = 0; r < kR; r++) m[g][r] = 0;
    }
#pragma unroll
    for (int e = 0; e < ND; e++) dk[e] = 0;
    if (GR == 0 && P.G > 0) {
        for (int i = lane; i < P.G * NA * 32; i += 32) sacc[i] = agg_identity(P.agg_kind[(i >> 5) % NA]);
    }
    __syncwarp();

    // a partial last tile of a borrowed (unpadded) source is staged with guarded plain loads
    auto is_guarded = [&](int64_t tile) -> bool {
        return P.borrowed && (tile + 1) * (int64_t)kTile > n_rows;
    };
    auto issue = [&](int64_t tile, int s) {
        if (is_guarded(tile)) return;
        uint64_t* bar = &bars[s];
        unsigned char* dst = wbase + (size_t)s * P.stage_bytes;
        mbar_expect_tx(bar, P.stage_bytes);
        for (int c = 0; c < P.n_cols; c++) {
            const uint32_t bytes = kTile * P.col_w[c];
            tma_bulk_g2s(dst + P.col_off[c], P.col_ptr[c] + (size_t)tile * bytes, bytes, bar);
        }
    };
    if (lane == 0 && P.n_cols > 0) {
        for (int s = 0; s < S; s++)
            if (first + s * stride < n_tiles) issue(first + s * stride, s);
    }

    int s = 0;
    uint32_t phase = 0;
    for (int64_t tile = first; tile < n_tiles; tile += stride) {
        WarpCtx c;
        c.stage = wbase + (size_t)s * P.stage_bytes;
        c.wbase = wbase;
        c.row0 = tile * (int64_t)kTile;
        c.lane = lane;

        if (P.n_cols > 0) {
            if (is_guarded(tile)) {
                unsigned char* dst = wbase + (size_t)s * P.stage_bytes;
                const int64_t rows = n_rows - c.row0;
                for (int col = 0; col < P.n_cols; col++) {
                    const int w = P.col_w[col];
                    const unsigned char* src = P.col_ptr[col] + (size_t)c.row0 * w;
                    for (int64_t i = lane; i < (int64_t)kTile * w; i += 32)
                        dst[P.col_off[col] + i] = (i < rows * w) ? src[i] : (unsigned char)0;
                }
                __syncwarp();
            } else {
                mbar_wait(&bars[s], phase);
            }
        }

        unsigned valid = 0xffu;
        if (c.row0 + kTile > n_rows) {
            valid = 0;
#pragma unroll
            for (int r = 0; r < kR; r++)
                if (c.row0 + row_in_tile(r, lane) < n_rows) valid |= 1u << r;
        }

        int64_t acc[kR];
#pragma unroll
        for (int r = 0; r < kR; r++) acc[r] = 0;
        unsigned gid = 0;   // GR == 0 low-card path: 4 bits per tuple

        const int n_insn = P.n_insn;
        for (int pc = 0; pc < n_insn; pc++) {
            const UInsn in = P.insn[pc];
            const unsigned char* ob = opnd_base(c, in.flags & UF_SLOT, (uint32_t)in.off16 << 4);
            switch (in.code) {
                case U_LD_M64: ld_m64(ob, lane, acc); break;
                case U_LD_M32: {
                    int32_t t[kR]; ld_m32(ob, lane, t);
#pragma unroll
                    for (int r = 0; r < kR; r++) acc[r] = t[r];
                    break;
                }
                case U_LD_M8: {
                    uint32_t t[kR]; ld_m8(ob, lane, t);
#pragma unroll
                    for (int r = 0; r < kR; r++) acc[r] = t[r];
                    break;
                }
                case U_LD_IMM:
#pragma unroll
                    for (int r = 0; r < kR; r++) acc[r] = in.imm;
                    break;
                case U_LD_STR: fetch(P, c, K_STR, ob, in.off16, 0, acc); break;

#define RQ_CASES(N)                                                                        \
    case U_##N##_AM: {                                                                     \
        int64_t b[kR]; ld_m64(ob, lane, b);                                                \
        _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = RQ_EX_##N(acc[r], b[r]);   \
        break;                                                                             \
    }                                                                                      \
    case U_##N##_AI: {                                                                     \
        const int64_t y = in.imm;                                                          \
        _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = RQ_EX_##N(acc[r], y);      \
        break;                                                                             \
    }                                                                                      \
    case U_##N##_MI: {                                                                     \
        int64_t a[kR]; ld_m64(ob, lane, a);                                                \
        const int64_t y = in.imm;                                                          \
        _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = RQ_EX_##N(a[r], y);        \
        break;                                                                             \
    }                                                                                      \
    case U_##N##_MM: {                                                                     \
        int64_t a[kR], b[kR]; ld_m64(ob, lane, a);                                         \
        ld_m64(opnd_base(c, in.flags & UF_SLOT2, (uint32_t)in.imm), lane, b);              \
        _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = RQ_EX_##N(a[r], b[r]);     \
        break;                                                                             \
    }
                RQ_BINOPS(RQ_CASES)
#undef RQ_CASES

                case U_GEN: {
                    int64_t b[kR];
                    fetch(P, c, in.gsrc, ob, in.off16, in.imm, b);
                    switch (in.gop) {
#define RQ_GBIN(D, N) case D: _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = RQ_EX_##N(acc[r], b[r]); break;
                        case D_LD:
#pragma unroll
                            for (int r = 0; r < kR; r++) acc[r] = b[r];
                            break;
                        RQ_GBIN(D_ADD, ADD) RQ_GBIN(D_SUB, SUB) RQ_GBIN(D_RSUB, RSUB) RQ_GBIN(D_MUL, MUL)
                        RQ_GBIN(D_AND, AND) RQ_GBIN(D_OR, OR) RQ_GBIN(D_LT, LT) RQ_GBIN(D_LE, LE)
                        RQ_GBIN(D_GT, GT) RQ_GBIN(D_GE, GE) RQ_GBIN(D_EQ, EQ) RQ_GBIN(D_NE, NE)
#undef RQ_GBIN
                        case D_DIV:
#pragma unroll
                            for (int r = 0; r < kR; r++)
                                acc[r] = ((valid >> r) & 1) ? div_trunc(acc[r], b[r], P.err) : 0;
                            break;
                        case D_RDIV:
#pragma unroll
                            for (int r = 0; r < kR; r++)
                                acc[r] = ((valid >> r) & 1) ? div_trunc(b[r], acc[r], P.err) : 0;
                            break;
#define RQ_STRBIN(D, EXPR)                                                       \
    case D:                                                                      \
        _Pragma("unroll") for (int r = 0; r < kR; r++) {                         \
            const char* x = reinterpret_cast<const char*>(acc[r]);               \
            const char* y = reinterpret_cast<const char*>(b[r]);                 \
            acc[r] = ((valid >> r) & 1) ? (EXPR) : 0;                            \
        }                                                                        \
        break;
                        RQ_STRBIN(D_EQC, str_eq_char(x, y))
                        RQ_STRBIN(D_EQV, str_eq_varchar(x, y))
                        RQ_STRBIN(D_NEC, 1 - str_eq_char(x, y))
                        RQ_STRBIN(D_NEV, 1 - str_eq_varchar(x, y))
                        RQ_STRBIN(D_LIKE, str_like(x, y))
                        RQ_STRBIN(D_RLIKE, str_like(y, x))
#undef RQ_STRBIN
                        case D_SEL: {
                            int64_t e[kR];
                            if (in.flags & UF_ELSE_IMM) {
#pragma unroll
                                for (int r = 0; r < kR; r++) e[r] = P.imm[in.aux];
                            } else {
                                ld_m64(slot_base + (size_t)in.aux * (kTile * 8), lane, e);
                            }
#pragma unroll
                            for (int r = 0; r < kR; r++) acc[r] = (acc[r] & 0xff) ? b[r] : e[r];
                            break;
                        }
                        default: break;
                    }
                    break;
                }

                case U_FILTER_A:
#pragma unroll
                    for (int r = 0; r < kR; r++)
                        if ((acc[r] & 0xff) == 0) valid &= ~(1u << r);
                    if (!__any_sync(kFull, valid != 0)) pc = n_insn;
                    break;
                case U_FILTER_O: {
                    int64_t b[kR];
                    fetch(P, c, in.gsrc, ob, in.off16, in.imm, b);
#pragma unroll
                    for (int r = 0; r < kR; r++)
                        if ((b[r] & 0xff) == 0) valid &= ~(1u << r);
                    if (!__any_sync(kFull, valid != 0)) pc = n_insn;
                    break;
                }

#define RQ_FCMP(N, OP)                                                                     \
    case U_F##N##_M64: {                                                                   \
        int64_t b[kR]; ld_m64(ob, lane, b);                                                \
        const int64_t y = in.imm;                                                          \
        _Pragma("unroll") for (int r = 0; r < kR; r++)                                     \
            if (!(b[r] OP y)) valid &= ~(1u << r);                                         \
        if (!__any_sync(kFull, valid != 0)) pc = n_insn;                                   \
        break;                                                                             \
    }                                                                                      \
    case U_F##N##_M32: {                                                                   \
        int32_t b[kR]; ld_m32(ob, lane, b);                                                \
        const int32_t y = (int32_t)in.imm;                                                 \
        _Pragma("unroll") for (int r = 0; r < kR; r++)                                     \
            if (!(b[r] OP y)) valid &= ~(1u << r);                                         \
        if (!__any_sync(kFull, valid != 0)) pc = n_insn;                                   \
        break;                                                                             \
    }                                                                                      \
    case U_F##N##_M8: {                                                                    \
        uint32_t b[kR]; ld_m8(ob, lane, b);                                                \
        const int32_t y = (int32_t)in.imm;                                                 \
        _Pragma("unroll") for (int r = 0; r < kR; r++)                                     \
            if (!((int32_t)b[r] OP y)) valid &= ~(1u << r);                                \
        if (!__any_sync(kFull, valid != 0)) pc = n_insn;                                   \
        break;                                                                             \
    }
                RQ_FCMP(LT, <) RQ_FCMP(LE, <=) RQ_FCMP(GT, >) RQ_FCMP(GE, >=) RQ_FCMP(EQ, ==) RQ_FCMP(NE, !=)
#undef RQ_FCMP

                case U_GROUP: {
                    if (!lowagg) break;
                    if (NK == 0) {
                        seen |= valid;
                        if (GR > 0) {
#pragma unroll
                            for (int r = 0; r < kR; r++) m[0][r] = (valid >> r) & 1u;
                        }
                        break;
                    }
                    // packed group key
                    uint64_t key[kR];
#pragma unroll
                    for (int r = 0; r < kR; r++) key[r] = 0;
                    for (int j = 0; j < NK; j++) {
                        const VRef vr = P.key[j];
                        int64_t kv[kR];
                        fetch(P, c, vr.kind, opnd_base(c, vr.slot, (uint32_t)vr.off16 << 4), 0,
                              vr.kind == K_IMM ? P.imm[vr.off16] : 0, kv);
                        const int sh = P.key_shift[j];
                        const uint64_t mask = P.key_bits[j] >= 64 ? ~0ULL : ((1ULL << P.key_bits[j]) - 1);
#pragma unroll
                        for (int r = 0; r < kR; r++) key[r] |= ((uint64_t)kv[r] & mask) << sh;
                    }
                    unsigned unk = 0;
                    if (GR > 0) {
#pragma unroll
                        for (int g = 0; g < NG; g++) {
                            const bool act = g < ngroups;
#pragma unroll
                            for (int r = 0; r < kR; r++) {
                                const bool hit = P.key32 ? ((uint32_t)key[r] == (uint32_t)dk[g]) : (key[r] == dk[g]);
                                m[g][r] = (act && hit && ((valid >> r) & 1)) ? 1u : 0u;
                            }
                        }
#pragma unroll
                        for (int r = 0; r < kR; r++) {
                            uint32_t any = 0;
#pragma unroll
                            for (int g = 0; g < NG; g++) any |= m[g][r];
                            if (((valid >> r) & 1) && !any) unk |= 1u << r;
                        }
                    } else {
                        gid = 0;
#pragma unroll
                        for (int r = 0; r < kR; r++) {
                            unsigned g = 15;
#pragma unroll
                            for (int e = 0; e < ND; e++)
                                if (e < ngroups && key[r] == dk[e]) g = e;
                            if (!((valid >> r) & 1)) g = 0;
                            else if (g == 15) { unk |= 1u << r; g = 0; }
                            gid |= g << (4 * r);
                        }
                    }
                    // slow path: a key this warp has not seen yet joins the dictionary
                    while (__any_sync(kFull, unk != 0)) {
                        const unsigned ball = __ballot_sync(kFull, unk != 0);
                        const int leader = __ffs(ball) - 1;
                        uint64_t lk = 0;
                        const int rr = __ffs(unk) - 1;
#pragma unroll
                        for (int r = 0; r < kR; r++) if (r == rr) lk = key[r];
                        lk = __shfl_sync(kFull, lk, leader);
                        const int cap = GR > 0 ? NG : P.G;
                        if (ngroups >= cap) {
                            if (lane == 0) *P.overflow = 1;
                            unk = 0;
                            break;
                        }
#pragma unroll
                        for (int e = 0; e < ND; e++) if (e == ngroups) dk[e] = lk;
#pragma unroll
                        for (int r = 0; r < kR; r++) {
                            if (((unk >> r) & 1) && key[r] == lk) {
                                unk &= ~(1u << r);
                                if (GR > 0) {
#pragma unroll
                                    for (int g = 0; g < NG; g++) if (g == ngroups) m[g][r] = 1u;
                                } else {
                                    gid |= (unsigned)ngroups << (4 * r);
                                }
                            }
                        }
                        ngroups++;
                    }
                    break;
                }

                case U_AGG_SUM_A:
                case U_AGG_SUM_M: {
                    int64_t v[kR];
                    if (in.code == U_AGG_SUM_M) ld_m64(ob, lane, v);
                    else {
#pragma unroll
                        for (int r = 0; r < kR; r++) v[r] = acc[r];
                    }
                    if (GR > 0) {
                        switch (in.aux) {
#define RQ_SUMCASE(A)                                                                         \
    case A:                                                                                   \
        RQ_NOMERGE(A);                                                                        \
        _Pragma("unroll") for (int g = 0; g < NG; g++)                                        \
            _Pragma("unroll") for (int r = 0; r < kR; r++)                                    \
                macc(racc[g][A], v[r], m[g][r]);                                             \
        break;
                            RQ_SUMCASE(0) RQ_SUMCASE(1) RQ_SUMCASE(2) RQ_SUMCASE(3)
                            RQ_SUMCASE(4) RQ_SUMCASE(5)
#undef RQ_SUMCASE
                            default: break;
                        }
                    } else {
                        int64_t* base = sacc + in.aux * 32 + lane;
#pragma unroll
                        for (int r = 0; r < kR; r++) {
                            if ((valid >> r) & 1) {
                                int64_t* p = base + ((gid >> (4 * r)) & 15) * NA * 32;
                                *p = (int64_t)((uint64_t)*p + (uint64_t)v[r]);
                            }
                        }
                    }
                    break;
                }
                case U_AGG_COUNT: {
                    if (GR > 0) {
                        switch (in.aux) {
#define RQ_CNTCASE(A)                                                                         \
    case A:                                                                                   \
        RQ_NOMERGE(A);                                                                        \
        _Pragma("unroll") for (int g = 0; g < NG; g++) {                                      \
            uint32_t cnt = 0;                                                                 \
            _Pragma("unroll") for (int r = 0; r < kR; r++) cnt += m[g][r];                    \
            racc[g][A] += cnt;                                                                \
        }                                                                                     \
        break;
                            RQ_CNTCASE(0) RQ_CNTCASE(1) RQ_CNTCASE(2) RQ_CNTCASE(3)
                            RQ_CNTCASE(4) RQ_CNTCASE(5)
#undef RQ_CNTCASE
                            default: break;
                        }
                    } else {
                        int64_t* base = sacc + in.aux * 32 + lane;
#pragma unroll
                        for (int r = 0; r < kR; r++) {
                            if ((valid >> r) & 1) {
                                int64_t* p = base + ((gid >> (4 * r)) & 15) * NA * 32;
                                *p = *p + 1;
                            }
                        }
                    }
                    break;
                }
                case U_AGG_GEN: {
                    int64_t v[kR];
                    if (in.gsrc == K_NONE) {
#pragma unroll
                        for (int r = 0; r < kR; r++) v[r] = acc[r];
                    } else {
                        fetch(P, c, in.gsrc, ob, in.off16, in.imm, v);
                    }
                    const int kind = in.gop == D_AGG_SUM ? 1 : (in.gop == D_AGG_MIN ? 3 : 4);
                    if (GR > 0) {
                        // reduce the lane's 8 tuples per group first, then fold into the accumulator
                        int64_t cand[NG];
#pragma unroll
                        for (int g = 0; g < NG; g++) {
                            cand[g] = agg_identity(kind);
#pragma unroll
                            for (int r = 0; r < kR; r++) {
                                if (kind == 1) { uint64_t t = (uint64_t)cand[g]; macc(t, v[r], m[g][r]); cand[g] = (int64_t)t; }
                                else if (m[g][r] && (kind == 3 ? v[r] < cand[g] : v[r] > cand[g])) cand[g] = v[r];
                            }
                        }
                        switch (in.aux) {
#define RQ_GENCASE(A)                                                                         \
    case A:                                                                                   \
        RQ_NOMERGE(A);                                                                        \
        _Pragma("unroll") for (int g = 0; g < NG; g++) {                                      \
            const int64_t cur = (int64_t)racc[g][A];                                          \
            if (kind == 1) racc[g][A] = (uint64_t)cur + (uint64_t)cand[g];                    \
            else if (kind == 3 ? cand[g] < cur : cand[g] > cur) racc[g][A] = (uint64_t)cand[g]; \
        }                                                                                     \
        break;
                            RQ_GENCASE(0) RQ_GENCASE(1) RQ_GENCASE(2) RQ_GENCASE(3)
                            RQ_GENCASE(4) RQ_GENCASE(5)
#undef RQ_GENCASE
                            default: break;
                        }
                    } else {
                        int64_t* base = sacc + in.aux * 32 + lane;
#pragma unroll
                        for (int r = 0; r < kR; r++) {
                            if ((valid >> r) & 1) {
                                int64_t* p = base + ((gid >> (4 * r)) & 15) * NA * 32;
                                const int64_t cur = *p;
                                if (kind == 1) *p = (int64_t)((uint64_t)cur + (uint64_t)v[r]);
                                else if (kind == 3 ? v[r] < cur : v[r] > cur) *p = v[r];
                            }
                        }
                    }
                    break;
                }

                case U_PROBE: {
                    if (GR > 0) break;
                    // hash-join probe (hashjoin.h:118-214): tuples without a match are dropped;
                    // the matching entry's payload words land in value slots. The first tag of
                    // every tuple is fetched up front so that 8 probes are in flight per lane.
                    const DProbe& pr = P.probe[in.aux];
                    const uint64_t cap = pr.ht.cap_mask + 1;
                    const int pnk = pr.ht.nk;
                    uint64_t h[kR], t0[kR];
#pragma unroll
                    for (int r = 0; r < kR; r++) {
                        h[r] = 0; t0[r] = 0;
                        if ((valid >> r) & 1) {
                            int64_t k[kMaxKeys];
                            for (int j = 0; j < pnk; j++) k[j] = ld_row(P, c, pr.key[j], r);
                            h[r] = hash_typed(k, pr.ht.key_kind, pnk);
                            t0[r] = pr.ht.tags[h[r] & pr.ht.cap_mask];
                        }
                    }
#pragma unroll 1
                    for (int r = 0; r < kR; r++) {
                        if (!((valid >> r) & 1)) continue;
                        int64_t k[kMaxKeys];
                        for (int j = 0; j < pnk; j++) k[j] = ld_row(P, c, pr.key[j], r);
                        const uint64_t tag = h[r] | 2ULL;
                        uint64_t i = h[r] & pr.ht.cap_mask;
                        uint64_t t = 0;
#pragma unroll
                        for (int q = 0; q < kR; q++) if (q == r) t = t0[q];
                        int64_t found = -1;
                        unsigned matches = 0;
                        for (uint64_t tries = 0; tries < cap; tries++) {
                            if (t == 0ULL) break;
                            if (t == tag && slot_keys_equal(pr.ht, i, k)) {
                                if (found < 0) found = (int64_t)i;
                                matches++;
                                if (pr.single) break;
                            }
                            i = (i + 1) & pr.ht.cap_mask;
                            t = pr.ht.tags[i];
                        }
                        if (found < 0) { valid &= ~(1u << r); continue; }
                        if (matches > 1) atomicAdd(pr.dup_counter, (unsigned long long)(matches - 1));
                        const int row = row_in_tile(r, lane);
                        for (int q = 0; q < pr.n_out; q++)
                            if (pr.out_slot[q] != 0xff)
                                reinterpret_cast<int64_t*>(slot_base + (size_t)pr.out_slot[q] * (kTile * 8))[row] =
                                    pr.ht.vals[(size_t)q * cap + found];
                    }
                    __syncwarp();
                    if (!__any_sync(kFull, valid != 0)) pc = n_insn;
                    break;
                }
                case U_BUILD: {
                    if (GR > 0) break;
                    const uint64_t cap = P.ht.cap_mask + 1;
#pragma unroll 1
                    for (int r = 0; r < kR; r++) {
                        if (!((valid >> r) & 1)) continue;
                        int64_t k[kMaxKeys];
                        for (int j = 0; j < P.ht.nk; j++) k[j] = ld_row(P, c, P.key[j], r);
                        const uint64_t hh = hash_typed(k, P.ht.key_kind, P.ht.nk);
                        uint64_t slot;
                        if (!ht_insert_dup(P.ht, k, hh, &slot)) { *P.ht_full = 1; continue; }
                        for (int q = 0; q < P.n_out; q++)
                            P.ht.vals[(size_t)q * cap + slot] = ld_row(P, c, P.out[q], r);
                    }
                    break;
                }
                case U_HAGG: {
                    if (GR > 0) break;
                    const uint64_t cap = P.ht.cap_mask + 1;
#pragma unroll 1
                    for (int r = 0; r < kR; r++) {
                        if (!((valid >> r) & 1)) continue;
                        int64_t k[kMaxKeys];
                        for (int j = 0; j < P.ht.nk; j++) k[j] = ld_row(P, c, P.key[j], r);
                        const uint64_t hh = hash_typed(k, P.ht.key_kind, P.ht.nk);
                        uint64_t slot;
                        if (!ht_find_or_insert(P.ht, k, hh, &slot)) { *P.ht_full = 1; continue; }
                        for (int a = 0; a < P.na; a++) {
                            int64_t* dst = &P.ht.vals[(size_t)a * cap + slot];
                            const int kind = P.agg_kind[a];
                            if (kind == 2) { atomicAdd((unsigned long long*)dst, 1ULL); continue; }
                            const int64_t v = ld_row(P, c, P.agg_src[a], r);
                            if (kind == 1) atomicAdd((unsigned long long*)dst, (unsigned long long)v);
                            else if (kind == 3) atomicMin((long long*)dst, (long long)v);
                            else atomicMax((long long*)dst, (long long)v);
                        }
                    }
                    break;
                }
                case U_EMIT: {
                    if (GR > 0) break;
                    // one atomic per warp tile: lanes take consecutive output ranges
                    const int cnt = __popc(valid);
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const int total = __shfl_sync(kFull, incl, 31);
                    if (total == 0) break;
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(P.out_count, (unsigned long long)total);
                    base = __shfl_sync(kFull, base, 0);
                    int64_t pos = (int64_t)base + (incl - cnt);
#pragma unroll 1
                    for (int r = 0; r < kR; r++) {
                        if (!((valid >> r) & 1)) continue;
                        if (pos < P.out_cap)
                            for (int k = 0; k < P.n_out; k++) P.out_col[k][pos] = ld_row(P, c, P.out[k], r);
                        pos++;
                    }
                    break;
                }
                default: break;
            }
            if (in.flags & UF_STORE) st_m64(slot_base + (size_t)in.dst * (kTile * 8), lane, acc);
        }

        // everyone is done with stage s (and the slots) before it is refilled
        __syncwarp();
        if (lane == 0 && P.n_cols > 0) {
            const int64_t nt = tile + (int64_t)S * stride;
            if (nt < n_tiles) issue(nt, s);
        }
        if (++s == S) { s = 0; phase ^= 1u; }
    }

    // ---- flush the per-warp accumulators of the low-cardinality aggregate -------------------
    if (lowagg) {
        __syncwarp();
        int n = ngroups;
        if (NK == 0) n = __any_sync(kFull, seen != 0) ? 1 : 0;
        if (GR > 0) {
#pragma unroll
            for (int g = 0; g < NG; g++) {
                if (g >= n) continue;
                int slot = -1;
                if (lane == 0) {
                    slot = group_table_slot(P, NK == 0 ? 0ULL : dk[g]);
                    if (slot < 0) *P.overflow = 1;
                }
                slot = __shfl_sync(kFull, slot, 0);
#pragma unroll
                for (int a = 0; a < kNAR; a++) {
                    if (a >= NA) continue;
                    const int kind = P.agg_kind[a];
                    const int64_t v = warp_reduce((int64_t)racc[g][a], kind);
                    if (lane == 0 && slot >= 0) group_table_add(P, slot, a, kind, v);
                }
            }
        } else {
            const int cap = P.G;
            if (n > cap) n = cap;
            for (int e = 0; e < n; e++) {
                uint64_t kk = 0;
#pragma unroll
                for (int q = 0; q < ND; q++) if (q == e) kk = dk[q];
                int slot = -1;
                if (lane == 0) {
                    slot = group_table_slot(P, NK == 0 ? 0ULL : kk);
                    if (slot < 0) *P.overflow = 1;
                }
                slot = __shfl_sync(kFull, slot, 0);
                for (int a = 0; a < NA; a++) {
                    const int kind = P.agg_kind[a];
                    const int64_t v = warp_reduce(sacc[(e * NA + a) * 32 + lane], kind);
                    if (lane == 0 && slot >= 0) group_table_add(P, slot, a, kind, v);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------
// reset the global group table: state 0, accumulators to their identities
__global__ void rq_group_table_init(uint32_t* state, int64_t* acc, const uint8_t* kinds_dev,
                                    int na) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kGroupTableCap) {
        state[i] = 0;
        for (int a = 0; a < na; a++) acc[(size_t)i * kMaxAggs + a] = agg_identity(kinds_dev[a]);
    }
}

struct KeyUnpack {
    int32_t nk;
    uint8_t shift[kMaxKeys];
    uint8_t bits[kMaxKeys];
    uint8_t sign[kMaxKeys];    // sign-extend the field (INT/DATE/BIGINT/DECIMAL keys)
};

// group table -> dense int64 columns (keys unpacked first, then aggregates); order is
// unspecified, as in the reference where it is hash-slot order (aggregation.h:298-343)
__global__ void rq_group_table_compact(const uint32_t* state, const int64_t* keys,
                                       const int64_t* acc, KeyUnpack ku, int na,
                                       int64_t* const* out_cols, int64_t* out_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kGroupTableCap && state[i] == 2u) {
        const long long pos = atomicAdd((unsigned long long*)out_count, 1ULL);
        const uint64_t k = (uint64_t)keys[i];
        for (int j = 0; j < ku.nk; j++) {
            const int b = ku.bits[j];
            uint64_t f = (b >= 64) ? k : ((k >> ku.shift[j]) & ((1ULL << b) - 1));
            if (ku.sign[j] && b < 64 && ((f >> (b - 1)) & 1)) f |= ~((1ULL << b) - 1);
            out_cols[j][pos] = (int64_t)f;
        }
        for (int a = 0; a < na; a++) out_cols[ku.nk + a][pos] = acc[(size_t)i * kMaxAggs + a];
    }
}

// result columns: int64 values -> physical width, strings by value
__global__ void rq_narrow_i32(const int64_t* in, int32_t* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)in[i];
}
__global__ void rq_narrow_i8(const int64_t* in, uint8_t* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint8_t)in[i];
}
__global__ void rq_gather_str(const int64_t* addrs, unsigned char* out, int width, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned char* s = reinterpret_cast<const unsigned char*>(addrs[i]);
        unsigned char* d = out + (size_t)i * width;
        int k = 0;
        for (; k < width - 1 && s[k] != 0; k++) d[k] = s[k];
        for (; k < width; k++) d[k] = 0;
    }
}

// row store (reference DataBlocks, dbdata.h:23-102) -> columns
__global__ void rq_transpose_rows(const unsigned char* rows, int64_t n, int tuple_size, int offset,
                                  int width, unsigned char* col, int64_t col_row0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned char* s = rows + (size_t)i * tuple_size + offset;
        unsigned char* d = col + (size_t)(col_row0 + i) * width;
        for (int k = 0; k < width; k++) d[k] = s[k];
    }
}

}  // namespace rq
