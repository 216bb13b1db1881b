// Host side of an upload from HOST buffers (rq_table_upload without RQ_DEVICE_PTR): 8-byte integer
// columns are re-encoded to the narrowest exact width (1 or 4 bytes) by the host cores BEFORE they
// cross PCIe and widened again on the device when a chunk lands, so a DECIMAL column whose values fit
// one byte (TPC-H's l_quantity, l_discount, l_tax) costs one byte per row on the bus instead of eight.
// The link (about 52 GB/s measured) is what bounds a load; the conversion runs on all host cores and
// overlaps the DMA of the chunks already converted.
//
// Replaces nothing in the reference (its tables never leave host memory); the counterpart is the row
// store fill of executeBulkInsert (execute.h:332-388).
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace rq {
namespace hostnarrow {

// out[i] = (narrow) in[i]; *acc |= test word of every value. The value fits the narrow type iff the
// OR of all test words is below 2^8 (uint8: a negative value sets the sign bit) resp. below 2^32
// (int32: test word = value + 2^31).
#define RQ_HN_BODY8                                                               \
    uint64_t a = 0;                                                               \
    for (size_t i = 0; i < n; i++) { const int64_t v = in[i]; a |= (uint64_t)v; out[i] = (uint8_t)v; } \
    *acc |= a;
#define RQ_HN_BODY32                                                              \
    uint64_t a = 0;                                                               \
    for (size_t i = 0; i < n; i++) { const int64_t v = in[i]; a |= (uint64_t)v + 0x80000000ULL; out[i] = (int32_t)v; } \
    *acc |= a;
#define RQ_HN_MM8                                                                 \
    uint8_t l = 255, h = 0;                                                       \
    for (size_t i = 0; i < n; i++) { const uint8_t v = p[i]; l = v < l ? v : l; h = v > h ? v : h; } \
    if ((int64_t)l < *lo) *lo = l; if ((int64_t)h > *hi) *hi = h;
#define RQ_HN_MM32                                                                \
    int32_t l = INT32_MAX, h = INT32_MIN;                                         \
    for (size_t i = 0; i < n; i++) { const int32_t v = p[i]; l = v < l ? v : l; h = v > h ? v : h; } \
    if ((int64_t)l < *lo) *lo = l; if ((int64_t)h > *hi) *hi = h;

__attribute__((target("avx2"))) static void conv8_avx2(const int64_t* in, uint8_t* out, size_t n, uint64_t* acc) { RQ_HN_BODY8 }
__attribute__((target("avx2"))) static void conv32_avx2(const int64_t* in, int32_t* out, size_t n, uint64_t* acc) { RQ_HN_BODY32 }
__attribute__((target("avx2"))) static void mm8_avx2(const uint8_t* p, size_t n, int64_t* lo, int64_t* hi) { RQ_HN_MM8 }
__attribute__((target("avx2"))) static void mm32_avx2(const int32_t* p, size_t n, int64_t* lo, int64_t* hi) { RQ_HN_MM32 }
static void conv8_plain(const int64_t* in, uint8_t* out, size_t n, uint64_t* acc) { RQ_HN_BODY8 }
static void conv32_plain(const int64_t* in, int32_t* out, size_t n, uint64_t* acc) { RQ_HN_BODY32 }
static void mm8_plain(const uint8_t* p, size_t n, int64_t* lo, int64_t* hi) { RQ_HN_MM8 }
static void mm32_plain(const int32_t* p, size_t n, int64_t* lo, int64_t* hi) { RQ_HN_MM32 }
#undef RQ_HN_BODY8
#undef RQ_HN_BODY32
#undef RQ_HN_MM8
#undef RQ_HN_MM32

// the same from 4-byte values (INT columns whose values fit one byte: flags, small enumerations)
#define RQ_HN_BODY8_32                                                            \
    uint32_t a = 0;                                                               \
    for (size_t i = 0; i < n; i++) { const int32_t v = in[i]; a |= (uint32_t)v; out[i] = (uint8_t)v; } \
    *acc |= a;
__attribute__((target("avx2"))) static void conv8_from32_avx2(const int32_t* in, uint8_t* out, size_t n, uint64_t* acc) { RQ_HN_BODY8_32 }
static void conv8_from32_plain(const int32_t* in, uint8_t* out, size_t n, uint64_t* acc) { RQ_HN_BODY8_32 }
#undef RQ_HN_BODY8_32

static bool has_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}

// one chunk of an int64 column -> `w`-byte values (w = 1 or 4); returns false when a value does
// not fit; lo / hi are widened to the chunk's exact value range
static bool convert_chunk(const int64_t* in, void* out, size_t n, int w, int64_t* lo, int64_t* hi) {
    uint64_t acc = 0;
    const bool vx = has_avx2();
    if (w == 1) {
        if (vx) conv8_avx2(in, (uint8_t*)out, n, &acc); else conv8_plain(in, (uint8_t*)out, n, &acc);
        if (acc >= 256) return false;
        if (vx) mm8_avx2((const uint8_t*)out, n, lo, hi); else mm8_plain((const uint8_t*)out, n, lo, hi);
    } else {
        if (vx) conv32_avx2(in, (int32_t*)out, n, &acc); else conv32_plain(in, (int32_t*)out, n, &acc);
        if (acc >= (1ULL << 32)) return false;
        if (vx) mm32_avx2((const int32_t*)out, n, lo, hi); else mm32_plain((const int32_t*)out, n, lo, hi);
    }
    return true;
}

// one chunk of an int32 column -> bytes; false when a value does not fit [0, 255]
static bool convert_chunk32(const int32_t* in, void* out, size_t n, int64_t* lo, int64_t* hi) {
    uint64_t acc = 0;
    const bool vx = has_avx2();
    if (vx) conv8_from32_avx2(in, (uint8_t*)out, n, &acc); else conv8_from32_plain(in, (uint8_t*)out, n, &acc);
    if (acc >= 256) return false;
    if (vx) mm8_avx2((const uint8_t*)out, n, lo, hi); else mm8_plain((const uint8_t*)out, n, lo, hi);
    return true;
}

}  // namespace hostnarrow
}  // namespace rq
