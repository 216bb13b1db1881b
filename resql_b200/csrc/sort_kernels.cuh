// ORDER BY / LIMIT on the device. Replaces src/qlib/sort.h (Quicksorter, :21-173) and
// Relation::applyLimit (src/dbdata.h:407-425). Typed multi-key compare follows
// src/types.h:264-353: integers signed, DATE as int, strings byte-wise (strcmp).
#pragma once
#include <cuda_runtime.h>
#include "rq_internal.h"

namespace rq {

constexpr int kMaxSortKeys = 8;
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
rq_sort_small(const __grid_constant__ SortKeys K, const int64_t* n_ptr, uint32_t* perm) {
    __shared__ uint32_t idx[kBitonicMax];
    const int n = (int)*n_ptr;
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) idx[i] = (i < n) ? (uint32_t)i : 0xffffffffu;
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
                              const int64_t* n_ptr, int64_t limit) {
    int64_t n = *n_ptr;
    if (limit >= 0 && limit < n) n = limit;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

}  // namespace rq
