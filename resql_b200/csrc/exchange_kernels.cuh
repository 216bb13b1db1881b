// Hash-partitioned exchange of a relation between the ranks (one process per GPU): every row goes
// to rank hash(key columns) mod world. Used by plans executed with RQ_PLAN_PARTITIONED for joins
// whose two sides are both sharded (build rows and probe rows with equal keys meet on one rank) and
// for GROUP BY (the partial groups of one key are merged on one rank) - the reference computes
// these with one shared hash table (hashjoin.h:226-279, aggregation.h:240-295); partitioning by the
// key hash gives every rank a private share of that table.
//   1. rq_ex_count    destination rank per row (kept as one byte) + rows per destination
//   2. rq_ex_offsets  prefix sums
//   3. rq_ex_scatter  rows grouped by destination: send = [dest][column][rows of dest]
//   4. ncclSend/ncclRecv inside one group (engine_exec.inl exchange_relation)
//   5. rq_ex_unpack   received [source][column][rows from source] -> dense columns
#pragma once
#include <cuda_runtime.h>
#include "rq_internal.h"
#include "hash_kernels.cuh"

namespace rq {

constexpr int kMaxRanks = 64;
constexpr int kExThreads = 256;
constexpr int kExRowsPerThread = 8;
constexpr int kExBlockRows = kExThreads * kExRowsPerThread;

struct ExCols {
    const int64_t* in[kMaxOut];
    int32_t ncols;
    int32_t nkeys;
    int32_t key_col[kMaxKeys];
    uint8_t key_kind[kMaxKeys];     // 0 integer, 1 CHAR, 2 VARCHAR (values are device addresses)
    int32_t world;
};

__device__ __forceinline__ uint32_t ex_dest(const ExCols& X, int64_t row) {
    int64_t k[kMaxKeys];
    for (int j = 0; j < X.nkeys; j++) k[j] = X.in[X.key_col[j]][row];
    // (not the home-slot bits of the local hash tables: those use the high bits of key * golden ratio)
    const uint64_t h = mix64(hash_typed(k, X.key_kind, X.nkeys) ^ 0x51ed270b9f3c1a27ULL);
    return (uint32_t)(((h >> 32) * (uint64_t)X.world) >> 32);
}

__global__ void __launch_bounds__(kExThreads)
rq_ex_count(ExCols X, const int64_t* n_ptr, int64_t n_host, int64_t n_cap, uint8_t* dest, unsigned long long* cnt) {
    __shared__ unsigned int hist[kMaxRanks];
    int64_t n = n_ptr ? *n_ptr : n_host;
    if (n > n_cap) n = n_cap;
    if (threadIdx.x < kMaxRanks) hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kExBlockRows;
    for (int k = 0; k < kExRowsPerThread; k++) {
        const int64_t i = base + k * kExThreads + threadIdx.x;
        if (i < n) {
            const uint32_t d = X.nkeys > 0 ? ex_dest(X, i) : 0u;
            dest[i] = (uint8_t)d;
            atomicAdd(&hist[d], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < X.world && hist[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

// cnt[W] -> off[W] (exclusive prefix sums), cursor[W] = 0
__global__ void rq_ex_offsets(const unsigned long long* cnt, unsigned long long* off, unsigned long long* cursor, int world) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long s = 0;
        for (int r = 0; r < world; r++) { off[r] = s; s += cnt[r]; cursor[r] = 0; }
    }
}

__global__ void __launch_bounds__(kExThreads)
rq_ex_scatter(ExCols X, const int64_t* n_ptr, int64_t n_host, int64_t n_cap, const uint8_t* dest, const unsigned long long* cnt,
              const unsigned long long* off, unsigned long long* cursor, int64_t* send) {
    __shared__ unsigned int hist[kMaxRanks];
    __shared__ unsigned long long base_of[kMaxRanks];
    int64_t n = n_ptr ? *n_ptr : n_host;
    if (n > n_cap) n = n_cap;
    if (threadIdx.x < kMaxRanks) hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kExBlockRows;
    unsigned int my_pos[kExRowsPerThread];
    uint8_t my_dest[kExRowsPerThread];
    for (int k = 0; k < kExRowsPerThread; k++) {
        const int64_t i = base + k * kExThreads + threadIdx.x;
        my_dest[k] = 0; my_pos[k] = 0;
        if (i < n) {
            my_dest[k] = dest[i];
            my_pos[k] = atomicAdd(&hist[my_dest[k]], 1u);      // position inside this block's share
        }
    }
    __syncthreads();
    if (threadIdx.x < X.world)
        base_of[threadIdx.x] = hist[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (unsigned long long)hist[threadIdx.x]) : 0ULL;
    __syncthreads();
    for (int k = 0; k < kExRowsPerThread; k++) {
        const int64_t i = base + k * kExThreads + threadIdx.x;
        if (i < n) {
            const int d = my_dest[k];
            const unsigned long long j = base_of[d] + my_pos[k];                   // row inside the block of destination d
            int64_t* blk = send + off[d] * (unsigned long long)X.ncols;
            for (int c = 0; c < X.ncols; c++) blk[(unsigned long long)c * cnt[d] + j] = X.in[c][i];
        }
    }
}

struct ExUnpack {
    int64_t* out[kMaxOut];
    int64_t count[kMaxRanks];     // rows received from source r
    int64_t off[kMaxRanks];       // their first row in the output
    int32_t ncols;
    int32_t world;
};
__global__ void rq_ex_unpack(ExUnpack U, const int64_t* recv, int64_t total) {
    // recv = [source][column][rows from source]; element index e -> (source, column, row)
    const int64_t all = total * U.ncols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < all; e += (int64_t)gridDim.x * blockDim.x) {
        int r = 0;
        while (r + 1 < U.world && e >= (U.off[r + 1]) * U.ncols) r++;
        const int64_t w = e - U.off[r] * U.ncols;
        const int64_t c = w / U.count[r], j = w - c * U.count[r];
        U.out[c][U.off[r] + j] = recv[e];
    }
}

}  // namespace rq
