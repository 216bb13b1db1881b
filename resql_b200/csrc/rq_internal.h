// Internal (host+device) definitions of the resql_b200 engine. Not part of the ABI.
#pragma once
#include <stdint.h>

namespace rq {

// ---- tile geometry of the pipeline kernel ------------------------------------------------
constexpr int kThreads       = 256;                 // threads per CTA (8 warps)
constexpr int kRowsPerThread = 4;                   // register tile: 4 tuples per thread
constexpr int kTileRows      = kThreads * kRowsPerThread;   // 1024 tuples per staged tile
constexpr int kStages        = 2;                   // TMA double buffering
constexpr int kWarps         = kThreads / 32;

constexpr int kMaxStagedCols = 16;
constexpr int kMaxStrCols    = 8;
constexpr int kMaxInsn       = 160;
constexpr int kMaxKeys       = 8;
constexpr int kMaxAggs       = 16;
constexpr int kMaxOut        = 24;
constexpr int kMaxImm        = 32;
constexpr int kMaxProbes     = 4;
constexpr int kMaxSlots      = 12;
constexpr int kLowCardMaxGroups = 8;    // groups a warp can track with lane-private accumulators
constexpr int kGroupTableCap = 2048;    // global table of the low-cardinality aggregate path

// ---- device instruction: accumulator machine ----------------------------------------------
// The ABI-level postfix program (rq_node) is linearised on the host into instructions of the
// form   acc = acc OP operand   over a register tile of kRowsPerThread tuples per thread.
// Values used more than once or not consumed by the next instruction are kept in shared-memory
// slots (one int64 per tuple of the tile).
enum DOp : uint8_t {
    D_LD = 1, D_ADD, D_SUB, D_RSUB, D_MUL, D_DIV, D_RDIV, D_AND, D_OR,
    D_LT, D_LE, D_GT, D_GE, D_EQ, D_NE,
    D_EQC, D_EQV, D_NEC, D_NEV, D_LIKE, D_RLIKE,
    D_SEL,            // acc = (acc&0xff) ? operand : slot[aux]
    D_FILTER,         // valid &= (acc & 0xff) != 0
    D_GROUP,          // low-cardinality group lookup (keys via KParams::key)
    D_AGG_SUM, D_AGG_COUNT, D_AGG_MIN, D_AGG_MAX,   // aux = aggregate index
    D_PROBE,          // aux = probe index
    D_HAGG,           // hash aggregate sink
    D_BUILD,          // hash-join build sink
    D_EMIT,           // materialize sink
    D_NOP
};

enum DSrc : uint8_t { S_NONE = 0, S_COL = 1, S_SLOT = 2, S_IMM = 3, S_STR = 4 };

struct DInsn {
    uint8_t  op;
    uint8_t  src;       // DSrc
    uint8_t  flags;     // bit0: store acc to slot `dst` after the op
    uint8_t  dst;
    uint16_t idx;       // column / slot index of the operand
    uint16_t aux;
    int64_t  imm;
};
static_assert(sizeof(DInsn) == 16, "DInsn must be 16 bytes");

struct VRef {           // value reference used by sinks (keys, payloads, outputs)
    uint8_t  kind;      // DSrc
    uint8_t  pad;
    uint16_t idx;       // S_IMM: index into KParams::imm
};

// hash table used by joins (build then probe in separate kernels) and by hash aggregation
struct DHashTable {
    uint64_t  cap_mask;     // capacity - 1 (power of two)
    uint64_t* tags;         // 0 = empty; else fingerprint | 1
    int64_t*  keys;         // [nk][capacity]
    int64_t*  vals;         // [nv][capacity]  payload / accumulators
    int32_t   nk, nv;
    uint8_t   key_kind[kMaxKeys];   // 0 integer, 1 CHAR, 2 VARCHAR
};

struct DProbe {
    DHashTable ht;
    VRef       key[kMaxKeys];
    uint8_t    out_slot[kMaxOut];   // payload column k -> slot (0xff = not needed)
    int32_t    n_out;
    int32_t    single;
    unsigned long long* dup_counter;   // counts tuples with more than one match (multi-match mode)
};

struct KParams {
    // source
    int64_t        n_rows;
    const int64_t* n_rows_ptr;          // if non-null the row count is read on the device
    int32_t        borrowed;            // source buffers may end exactly at n_rows (no padding)
    int32_t        n_cols;              // staged (TMA) columns
    const unsigned char* col_ptr[kMaxStagedCols];
    uint32_t       col_off[kMaxStagedCols];   // byte offset inside a stage
    uint8_t        col_w[kMaxStagedCols];     // 1, 4, 8
    int32_t        n_strcols;
    const unsigned char* str_ptr[kMaxStrCols];
    uint32_t       str_w[kMaxStrCols];
    uint32_t       stage_bytes;
    int32_t        stages;              // 2 = double buffered, 1 when shared memory is short
    // shared memory carve-up (byte offsets)
    uint32_t       slots_off, acc_off, dict_off, smem_bytes;
    int32_t        n_slots;
    // program
    int32_t        n_insn;
    DInsn          insn[kMaxInsn];
    int64_t        imm[kMaxImm];
    // grouping / aggregation
    int32_t        nk;
    VRef           key[kMaxKeys];
    int32_t        na;
    uint8_t        agg_kind[kMaxAggs];
    VRef           agg_src[kMaxAggs];    // hash aggregate only
    int32_t        G;                    // lane-private groups per warp (low-card path)
    // low-card global table
    uint32_t*      g_state;              // [kGroupTableCap]
    int64_t*       g_keys;               // [kGroupTableCap][nk]
    int64_t*       g_acc;                // [kGroupTableCap][na]
    int32_t*       overflow;             // set when a warp meets more than G groups
    // hash aggregate / build
    DHashTable     ht;
    int32_t*       ht_full;              // set when the table is full
    // probes
    int32_t        n_probes;
    DProbe         probe[kMaxProbes];
    // materialize
    int32_t        n_out;
    VRef           out[kMaxOut];
    int64_t*       out_col[kMaxOut];
    unsigned long long* out_count;
    int64_t        out_cap;
    // runtime error flag (division by zero)
    int32_t*       err;
};

}  // namespace rq
