// Hash tables in HBM: join build/probe and high-cardinality GROUP BY.
// Replaces src/qlib/hash.h (linear probing, :385-478) and its users hashjoin.h / aggregation.h.
#pragma once
#include <cuda_runtime.h>
#include "rq_internal.h"

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

}  // namespace rq
