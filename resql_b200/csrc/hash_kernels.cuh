// Hash tables in HBM: join build/probe and high-cardinality GROUP BY.
// Replaces src/qlib/hash.h of the reference (linear probing with a stored hash per entry,
// :385-478) and its users hashjoin.h:118-279 / aggregation.h:240-295. Layout is columnar:
// one 64-bit tag word per slot (0 = empty, 1 = being written, else hash|2), key words and
// payload/accumulator words in separate arrays indexed by slot.
#pragma once
#include <cuda_runtime.h>
#include "rq_internal.h"
#include "device_util.cuh"

namespace rq {

struct HashTableDev {
    DHashTable d{};
    uint64_t capacity = 0;
    ~HashTableDev() {
        if (d.tags) cudaFree(d.tags);
        if (d.keys) cudaFree(d.keys);
        if (d.vals) cudaFree(d.vals);
    }
};

constexpr uint64_t kTagLocked = 1ULL;
constexpr uint64_t kMaxProbeLen = 2048;   // longer runs mean the table is (nearly) full: report and regrow

// key kinds: 0 integer word, 1 CHAR (equality ignores trailing blanks), 2 VARCHAR (exact)
__device__ __forceinline__ uint64_t hash_str(const unsigned char* s, bool strip) {
    // FNV-1a over the bytes; for CHAR the trailing blanks are left out so that values equal
    // under compareChar hash alike
    int len = 0, last = 0;
    while (s[len] != 0) { if (s[len] != ' ') last = len + 1; len++; }
    const int n = strip ? last : len;
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int i = 0; i < n; i++) { h ^= s[i]; h *= 0x100000001b3ULL; }
    return h;
}

__device__ __forceinline__ uint64_t hash_typed(const int64_t* k, const uint8_t* kind, int nk) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (int j = 0; j < nk; j++) {
        uint64_t w = (kind[j] == 0) ? (uint64_t)k[j]
                                    : hash_str(reinterpret_cast<const unsigned char*>(k[j]), kind[j] == 1);
        h = mix64(h ^ w) + 0x9E3779B97F4A7C15ULL;
    }
    return h;
}

__device__ __forceinline__ bool key_word_equal(int64_t a, int64_t b, int kind) {
    if (kind == 0) return a == b;
    if (kind == 1) return str_eq_char(reinterpret_cast<const char*>(a), reinterpret_cast<const char*>(b)) != 0;
    return str_eq_varchar(reinterpret_cast<const char*>(a), reinterpret_cast<const char*>(b)) != 0;
}

__device__ __forceinline__ bool slot_keys_equal(const DHashTable& ht, uint64_t i, const int64_t* k) {
    const uint64_t cap = ht.cap_mask + 1;
    for (int j = 0; j < ht.nk; j++)
        if (!key_word_equal(((volatile int64_t*)ht.keys)[(size_t)j * cap + i], k[j], ht.key_kind[j])) return false;
    return true;
}

// join build: every tuple gets its own slot (duplicates are kept, hashjoin.h:226-256)
__device__ __forceinline__ bool ht_insert_dup(const DHashTable& ht, const int64_t* k, uint64_t h,
                                              uint64_t* slot_out) {
    const uint64_t tag = h | 2ULL;
    const uint64_t cap = ht.cap_mask + 1;
    uint64_t i = h & ht.cap_mask;
    const uint64_t lim = cap < kMaxProbeLen ? cap : kMaxProbeLen;
    for (uint64_t tries = 0; tries < lim; tries++) {
        const unsigned long long old = atomicCAS((unsigned long long*)&ht.tags[i], 0ULL, (unsigned long long)tag);
        if (old == 0ULL) {
            for (int j = 0; j < ht.nk; j++) ht.keys[(size_t)j * cap + i] = k[j];
            *slot_out = i;
            return true;
        }
        i = (i + 1) & ht.cap_mask;
    }
    return false;
}

// GROUP BY: find the slot of the key or claim a new one (aggregation.h:262-279)
__device__ __forceinline__ bool ht_find_or_insert(const DHashTable& ht, const int64_t* k, uint64_t h,
                                                  uint64_t* slot_out) {
    const uint64_t tag = h | 2ULL;
    const uint64_t cap = ht.cap_mask + 1;
    uint64_t i = h & ht.cap_mask;
    const uint64_t lim = cap < kMaxProbeLen ? cap : kMaxProbeLen;
    for (uint64_t tries = 0; tries < lim; tries++) {
        uint64_t t = *(volatile uint64_t*)&ht.tags[i];
        if (t == 0ULL) {
            t = atomicCAS((unsigned long long*)&ht.tags[i], 0ULL, (unsigned long long)kTagLocked);
            if (t == 0ULL) {
                for (int j = 0; j < ht.nk; j++) ((volatile int64_t*)ht.keys)[(size_t)j * cap + i] = k[j];
                __threadfence();
                atomicExch((unsigned long long*)&ht.tags[i], (unsigned long long)tag);
                *slot_out = i;
                return true;
            }
        }
        while (t == kTagLocked) t = *(volatile uint64_t*)&ht.tags[i];
        if (t == tag) {
            __threadfence();
            if (slot_keys_equal(ht, i, k)) { *slot_out = i; return true; }
        }
        i = (i + 1) & ht.cap_mask;
    }
    return false;
}

// ---- helper kernels ------------------------------------------------------------------------
__global__ void rq_ht_init_vals(int64_t* vals, uint64_t cap, int nv, const uint8_t* kinds) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap)
        for (int a = 0; a < nv; a++) vals[(size_t)a * cap + i] = agg_identity(kinds[a]);
}

__global__ void rq_ht_count(const uint64_t* tags, uint64_t cap, unsigned long long* count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool used = i < cap && tags[i] != 0ULL;
    const unsigned bal = __ballot_sync(0xffffffffu, used);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count, (unsigned long long)__popc(bal));
}

// occupied slots -> dense int64 columns. colmap[c] < nk selects key word colmap[c], otherwise
// accumulator colmap[c]-nk (duplicate aggregates share one accumulator).
__global__ void rq_ht_compact(DHashTable ht, const int* colmap, int n_out, int64_t* const* out_cols,
                              unsigned long long* count) {
    const uint64_t cap = ht.cap_mask + 1;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool used = i < cap && ht.tags[i] != 0ULL;
    const unsigned bal = __ballot_sync(0xffffffffu, used);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (used) {
        const unsigned long long pos = base + __popc(bal & ((1u << lane) - 1));
        for (int c = 0; c < n_out; c++) {
            const int m = colmap[c];
            out_cols[c][pos] = (m < ht.nk) ? ht.keys[(size_t)m * cap + i] : ht.vals[(size_t)(m - ht.nk) * cap + i];
        }
    }
}

}  // namespace rq
