// ORDER BY / LIMIT on the device. Replaces src/qlib/sort.h (Quicksorter, :21-173) and
// Relation::applyLimit (src/dbdata.h:407-425). Typed multi-key compare follows
// src/types.h:264-353: integers signed, DATE as int, strings byte-wise (strcmp).
#pragma once
#include <cuda_runtime.h>
#include "rq_internal.h"

namespace rq {

constexpr int kMaxSortKeys = 24;     // = kMaxOut: every output column may be an ORDER BY key
constexpr int kBitonicMax = 4096;

struct SortKeys {
    int32_t n_keys;
    const int64_t* col[kMaxSortKeys];    // int64 values, or device addresses of strings
    uint8_t is_str[kMaxSortKeys];
    uint8_t desc[kMaxSortKeys];
};

__device__ __forceinline__ int cmp_str(const unsigned char* a, const unsigned char* b) {
    while (*a && *a == *b) { a++; b++; }
    return (int)*a - (int)*b;
}

// true if row i sorts strictly before row j
__device__ __forceinline__ bool row_less(const SortKeys& K, uint32_t i, uint32_t j) {
    for (int k = 0; k < K.n_keys; k++) {
        int c;
        if (K.is_str[k]) {
            c = cmp_str(reinterpret_cast<const unsigned char*>(K.col[k][i]),
                        reinterpret_cast<const unsigned char*>(K.col[k][j]));
        } else {
            const int64_t x = K.col[k][i], y = K.col[k][j];
            c = (x < y) ? -1 : (x > y ? 1 : 0);
        }
        if (c != 0) return K.desc[k] ? (c > 0) : (c < 0);
    }
    return i < j;   // total order; makes the network deterministic
}

// single-CTA bitonic sort of row indices (n <= kBitonicMax), perm[] receives the order
__global__ void __launch_bounds__(1024)
rq_sort_small(const __grid_constant__ SortKeys K, const int64_t* n_ptr, int64_t n_cap, uint32_t* perm, const uint32_t* cand) {
    __shared__ uint32_t idx[kBitonicMax];
    // (the device-side count is clamped to what the host sized the buffers for: a replayed plan runs
    // on predicted sizes, and a wrong prediction must stay inside its allocations)
    int64_t n64 = *n_ptr;
    if (n64 > n_cap) n64 = n_cap;
    if (n64 > kBitonicMax) n64 = kBitonicMax;
    const int n = (int)(n64 < 0 ? 0 : n64);
    int m = 1;
    while (m < n) m <<= 1;
    // rows to sort: all of 0..n-1, or the n candidate rows a top-k pre-selection left over
    for (int i = threadIdx.x; i < m; i += blockDim.x) idx[i] = (i < n) ? (cand ? cand[i] : (uint32_t)i) : 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const uint32_t a = idx[i], b = idx[l];
                    // padding (0xffffffff) sorts last
                    bool a_lt_b = (b == 0xffffffffu) ? (a != 0xffffffffu)
                                : (a == 0xffffffffu) ? false : row_less(K, a, b);
                    const bool up = ((i & k) == 0);
                    if (up ? !a_lt_b : a_lt_b) { idx[i] = b; idx[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = idx[i];
}

// out[i] = in[perm[i]] for i < min(n, limit)
__global__ void rq_apply_perm(const int64_t* in, int64_t* out, const uint32_t* perm,
                              const int64_t* n_ptr, int64_t n_cap, int64_t limit) {
    int64_t n = *n_ptr;
    if (n > n_cap) n = n_cap;
    if (limit >= 0 && limit < n) n = limit;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

// ---- result path: ORDER BY permutation, LIMIT, narrowing to the reference's physical widths and
// strings by value for ALL output columns in one launch, into one packed buffer (one D2H copy per
// query instead of a permutation, a narrowing kernel and a copy per column: a result of 10 columns
// cost 30 graph nodes of a few microseconds each, which is what a 0.4 ms shard of Q6 notices)
struct FinishCols {
    int32_t ncols;
    int32_t pad_;
    const int64_t* in[kMaxOut];
    uint64_t off[kMaxOut];          // byte offset of the column in the packed buffer (16-byte aligned)
    uint16_t width[kMaxOut];        // bytes per row in the result
    uint8_t kind[kMaxOut];          // 0 int64, 1 int32, 2 byte, 3 string (value = address of NUL-terminated bytes)
};
__global__ void rq_finish_result(FinishCols F, const uint32_t* perm, const int64_t* n_ptr, int64_t n_cap, int64_t limit,
                                 unsigned char* out) {
    int64_t n = n_ptr ? *n_ptr : n_cap;
    if (n > n_cap) n = n_cap;
    if (limit >= 0 && limit < n) n = limit;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = perm ? (int64_t)perm[i] : i;
        for (int c = 0; c < F.ncols; c++) {
            const int64_t v = F.in[c][j];
            unsigned char* d = out + F.off[c] + (size_t)i * F.width[c];
            switch (F.kind[c]) {
                case 0: *reinterpret_cast<int64_t*>(d) = v; break;
                case 1: *reinterpret_cast<int32_t*>(d) = (int32_t)v; break;
                case 2: *d = (unsigned char)v; break;
                default: {
                    const unsigned char* sp = reinterpret_cast<const unsigned char*>(v);
                    const int w = F.width[c];
                    int k = 0;
                    for (; k < w - 1 && sp[k] != 0; k++) d[k] = sp[k];
                    for (; k < w; k++) d[k] = 0;
                }
            }
        }
    }
}
// column-wise copy of a few small relations' worth of int64 columns in one launch
struct CopyCols { int32_t ncols; int32_t rows; const int64_t* in[kMaxOut + 1]; int64_t* out[kMaxOut + 1]; int32_t len[kMaxOut + 1]; };
__global__ void rq_copy_cols(CopyCols C) {
    for (int c = blockIdx.y; c < C.ncols; c += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C.len[c]; i += gridDim.x * blockDim.x) C.out[c][i] = C.in[c][i];
}

// ---- ORDER BY ... LIMIT k over many rows: radix select on the first key, then a small sort ------
// state[0] = key prefix decided so far, state[1] = rank still to find inside that prefix.
constexpr int kSelBits = 11;
constexpr int kSelBins = 1 << kSelBits;
__global__ void __launch_bounds__(256)
rq_topk_hist(const uint64_t* keys, int64_t n, const unsigned long long* state, int shift, int bits,
             uint32_t* hist) {
    // block-private histogram in shared memory; lanes of a warp that hit the same bin (the usual case
    // in the first passes, where the high bits of all keys agree) are combined before the atomic
    __shared__ uint32_t sh[kSelBins];
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint64_t prefix = state[0];
    const int hs = shift + bits;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n + stride - 1) / stride * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        bool in = false;
        uint32_t bin = 0;
        if (i < n) {
            const uint64_t k = keys[i];
            in = hs >= 64 ? true : (k >> hs) == prefix;
            bin = (uint32_t)(k >> shift) & ((1u << bits) - 1);
        }
        const unsigned act = __ballot_sync(0xffffffffu, in);
        if (in) {
            const unsigned peers = __match_any_sync(act, bin);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[bin], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
__global__ void __launch_bounds__(1024) rq_topk_pick(uint32_t* hist, unsigned long long* state, int bits) {
    // block-wide prefix sum over the bins (two per thread); the thread whose pair of bins contains
    // the wanted rank publishes the digit and the rank inside that bin
    __shared__ unsigned long long warp_tot[32];
    const int nb = 1 << bits;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned long long a = (2 * t < nb) ? hist[2 * t] : 0ULL;
    const unsigned long long b = (2 * t + 1 < nb) ? hist[2 * t + 1] : 0ULL;
    const unsigned long long s = a + b;
    unsigned long long inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_tot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        warp_tot[lane] = wi - w;      // exclusive offset of the warp
    }
    __syncthreads();
    const unsigned long long want = state[1];
    const unsigned long long excl = warp_tot[warp] + inc - s;
    __syncthreads();
    if (excl < want && want <= excl + s) {
        const bool first = want <= excl + a;
        state[0] = (state[0] << bits) | (unsigned long long)(first ? 2 * t : 2 * t + 1);
        state[1] = want - (first ? excl : excl + a);
    }
    for (int i = t; i < nb; i += blockDim.x) hist[i] = 0;
}
// rows whose first-key value is not larger than the k-th smallest one (ties included)
__global__ void rq_topk_compact(const uint64_t* keys, int64_t n, const unsigned long long* state, uint32_t* cand,
                                int64_t cap, unsigned long long* count) {
    const uint64_t thr = state[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (keys[i] <= thr) {
            const unsigned long long pos = atomicAdd(count, 1ULL);
            if ((int64_t)pos < cap) cand[pos] = (uint32_t)i;
        }
    }
}

// ---- LSD radix sort (n > kBitonicMax): 4-bit digits over order-preserving 64-bit keys -------
constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 8;
constexpr int kRadixChunk = kRadixThreads * kRadixItems;

// order-preserving key of one ORDER BY column (word w of a string = 8 bytes, big endian)
__global__ void rq_sort_make_keys(const int64_t* col, const uint32_t* perm, uint64_t* keys, int64_t n,
                                  int is_str, int word, int desc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t v = col[perm[i]];
    uint64_t k;
    if (is_str) {
        const unsigned char* s = reinterpret_cast<const unsigned char*>(v);
        int len = 0;
        while (s[len] != 0) len++;
        k = 0;
        for (int b = 0; b < 8; b++) {
            const int p = word * 8 + b;
            k = (k << 8) | (uint64_t)(p < len ? s[p] : 0);
        }
    } else {
        k = (uint64_t)v ^ 0x8000000000000000ULL;
    }
    keys[i] = desc ? ~k : k;
}

__global__ void rq_sort_iota(uint32_t* perm, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = (uint32_t)i;
}

// OR / AND of all keys: digits where both agree are constant and their pass is skipped
__global__ void rq_sort_key_bits(const uint64_t* keys, int64_t n, unsigned long long* or_and) {
    uint64_t o = 0, a = ~0ULL;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        o |= keys[i]; a &= keys[i];
    }
    for (int s = 16; s > 0; s >>= 1) {
        o |= __shfl_xor_sync(0xffffffffu, o, s);
        a &= __shfl_xor_sync(0xffffffffu, a, s);
    }
    if ((threadIdx.x & 31) == 0) { atomicOr(&or_and[0], o); atomicAnd(&or_and[1], a); }
}

// per-thread digit counts of its kRadixItems consecutive items -> exclusive offsets inside the
// block (stable: items keep their order) and the block total per digit
__device__ __forceinline__ void radix_block_offsets(const uint64_t* keys, int64_t n, int shift,
                                                    uint32_t (&mine)[16], uint32_t* s_cnt /*[16][kRadixThreads]*/,
                                                    uint32_t* s_tot /*[16]*/) {
    const int t = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kRadixChunk + (int64_t)t * kRadixItems;
#pragma unroll
    for (int d = 0; d < 16; d++) mine[d] = 0;
    for (int q = 0; q < kRadixItems; q++)
        if (base + q < n) mine[(keys[base + q] >> shift) & 15]++;
#pragma unroll
    for (int d = 0; d < 16; d++) s_cnt[d * kRadixThreads + t] = mine[d];
    __syncthreads();
    // one warp per two digits: exclusive scan over the 256 thread counts
    const int warp = t >> 5, lane = t & 31;
    for (int d = warp; d < 16; d += kRadixThreads / 32) {
        uint32_t run = 0;
        for (int c0 = 0; c0 < kRadixThreads; c0 += 32) {
            const uint32_t v = s_cnt[d * kRadixThreads + c0 + lane];
            uint32_t x = v;
            for (int s = 1; s < 32; s <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, s);
                if (lane >= s) x += y;
            }
            s_cnt[d * kRadixThreads + c0 + lane] = run + x - v;
            run += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) s_tot[d] = run;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kRadixThreads)
rq_radix_hist(const uint64_t* keys, int64_t n, int shift, uint32_t* hist /*[16][n_blocks]*/) {
    __shared__ uint32_t s_cnt[16 * kRadixThreads];
    __shared__ uint32_t s_tot[16];
    uint32_t mine[16];
    radix_block_offsets(keys, n, shift, mine, s_cnt, s_tot);
    if (threadIdx.x < 16) hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s_tot[threadIdx.x];
}

// exclusive scan over hist in digit-major order (single CTA)
__global__ void __launch_bounds__(1024) rq_radix_scan(uint32_t* hist, int64_t m) {
    __shared__ uint32_t s_part[1024];
    const int t = threadIdx.x;
    const int64_t per = (m + 1023) / 1024;
    const int64_t lo = t * per, hi = (lo + per < m) ? lo + per : m;
    uint32_t sum = 0;
    for (int64_t i = lo; i < hi; i++) sum += hist[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; i++) { const uint32_t v = s_part[i]; s_part[i] = run; run += v; }
    }
    __syncthreads();
    uint32_t run = s_part[t];
    for (int64_t i = lo; i < hi; i++) { const uint32_t v = hist[i]; hist[i] = run; run += v; }
}

__global__ void __launch_bounds__(kRadixThreads)
rq_radix_scatter(const uint64_t* keys_in, const uint32_t* perm_in, uint64_t* keys_out, uint32_t* perm_out,
                 int64_t n, int shift, const uint32_t* hist) {
    __shared__ uint32_t s_cnt[16 * kRadixThreads];
    __shared__ uint32_t s_tot[16];
    uint32_t mine[16];
    radix_block_offsets(keys_in, n, shift, mine, s_cnt, s_tot);
    const int t = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kRadixChunk + (int64_t)t * kRadixItems;
    uint32_t run[16];
#pragma unroll
    for (int d = 0; d < 16; d++) run[d] = hist[(size_t)d * gridDim.x + blockIdx.x] + s_cnt[d * kRadixThreads + t];
    for (int q = 0; q < kRadixItems; q++) {
        if (base + q >= n) break;
        const uint64_t k = keys_in[base + q];
        const int d = (int)((k >> shift) & 15);
        uint32_t pos = 0;
#pragma unroll
        for (int e = 0; e < 16; e++) if (e == d) pos = run[e]++;
        keys_out[pos] = k;
        perm_out[pos] = perm_in[base + q];
    }
}

}  // namespace rq
