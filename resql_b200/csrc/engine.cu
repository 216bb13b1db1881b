// resql_b200 engine: C ABI (include/resql_b200.h), device-resident columnar tables, lowering of
// the typed postfix programs to the accumulator-machine encoding, kernel orchestration.
//
// Replaces, on the reference side (Henning1/resql): JitContextFlounder::{compile,execute}
// (src/JitContextFlounder.h:410-487), the produceFlounder/consumeFlounder bodies of
// src/operators/*.h, qlib/hash.h and qlib/sort.h. No CPU fallback exists: every entry point fails
// with RQ_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <algorithm>
#include <chrono>
#include <functional>

#include "../../include/resql_b200.h"
#include "rq_internal.h"
#include "scan_kernel.cuh"
#include "sort_kernels.cuh"
#include "hash_kernels.cuh"
#include "dist.h"

using namespace rq;

// ------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------
static std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
struct RqError {
    int code;
    std::string msg;
};
[[noreturn]] static void raise(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw RqError{code, buf};
}
#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            raise(RQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                  __FILE__, __LINE__);                                                    \
    } while (0)

// ------------------------------------------------------------------------------------------
// engine state
// ------------------------------------------------------------------------------------------
struct DevColumn {
    int type = 0, width = 0;
    unsigned char* d = nullptr;
    bool owned = true;
};
struct rq_table {
    std::string name;
    int64_t n_rows = 0;       // host-known row count (-1: only on device)
    int64_t cap_rows = 0;     // allocated rows (multiple of kTileRows for owned tables)
    int64_t* d_n_rows = nullptr;
    bool borrowed = false;
    std::vector<DevColumn> cols;
    // logical type info for intermediates (per column)
    std::vector<int> sql_type, sql_width;
    ~rq_table() {
        for (auto& c : cols)
            if (c.owned && c.d) cudaFree(c.d);
        if (d_n_rows) cudaFree(d_n_rows);
    }
};

struct Engine {
    bool init = false;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[8];
    // scratch for the low-card aggregate path
    uint32_t* g_state = nullptr;
    int64_t* g_keys = nullptr;
    int64_t* g_acc = nullptr;
    uint8_t* g_kinds = nullptr;
    int32_t* flags = nullptr;   // [0]=overflow [1]=ht_full [2]=err
    int32_t* h_flags = nullptr; // pinned
    Dist dist;
};
static Engine E;

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static constexpr int kSmemMax = 227 * 1024;

// ------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------
extern "C" int rq_init(int device) {
    try {
        if (E.init) return RQ_OK;
        int n = 0;
        CK(cudaGetDeviceCount(&n));
        if (device < 0 || device >= n) return fail(RQ_ERR_CUDA, "rq_init: device %d of %d", device, n);
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(RQ_ERR_CUDA, "rq_init: device %s is sm_%d%d, this engine is built for sm_100a",
                        prop.name, prop.major, prop.minor);
        E.device = device;
        E.sm_count = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&E.stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&E.copy_stream, cudaStreamNonBlocking));
        for (auto& e : E.ev) CK(cudaEventCreate(&e));
        CK(cudaMalloc(&E.g_state, sizeof(uint32_t) * kGroupTableCap));
        CK(cudaMalloc(&E.g_keys, sizeof(int64_t) * kGroupTableCap));
        CK(cudaMalloc(&E.g_acc, sizeof(int64_t) * kGroupTableCap * kMaxAggs));
        CK(cudaMalloc(&E.g_kinds, kMaxAggs));
        CK(cudaMalloc(&E.flags, 64));
        CK(cudaMallocHost(&E.h_flags, 64));
        CK(cudaFuncSetAttribute(rq_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        CK(cudaFuncSetAttribute(rq_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        CK(cudaFuncSetAttribute(rq_scan_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        E.init = true;
        return RQ_OK;
    } catch (RqError& e) {
        return fail(e.code, "%s", e.msg.c_str());
    }
}

extern "C" int rq_shutdown(void) {
    if (!E.init) return RQ_OK;
    cudaDeviceSynchronize();
    dist_shutdown(E.dist);
    cudaFree(E.g_state); cudaFree(E.g_keys); cudaFree(E.g_acc); cudaFree(E.g_kinds);
    cudaFree(E.flags); cudaFreeHost(E.h_flags);
    for (auto& e : E.ev) cudaEventDestroy(e);
    cudaStreamDestroy(E.stream);
    cudaStreamDestroy(E.copy_stream);
    E = Engine();
    return RQ_OK;
}

extern "C" const char* rq_last_error(void) { return g_err.c_str(); }
extern "C" void* rq_stream(void) { return (void*)E.stream; }

extern "C" int rq_dist_unique_id(uint8_t out_id[128]) {
    std::string err;
    if (!dist_unique_id(out_id, err)) return fail(RQ_ERR_NCCL, "%s", err.c_str());
    return RQ_OK;
}
extern "C" int rq_dist_init(int rank, int world, const uint8_t id[128]) {
    if (!E.init) return fail(RQ_ERR_NOT_INIT, "rq_dist_init before rq_init");
    std::string err;
    if (!dist_init(E.dist, rank, world, id, err)) return fail(RQ_ERR_NCCL, "%s", err.c_str());
    return RQ_OK;
}

// ------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------
static bool valid_col(int type, int width) {
    if (type == RQ_I8) return width == 1;
    if (type == RQ_I32) return width == 4;
    if (type == RQ_I64) return width == 8;
    if (type == RQ_STR) return width >= 2;
    return false;
}

extern "C" int rq_table_upload(const char* name, int32_t n_cols, const rq_column* cols,
                               int64_t n_rows, int32_t flags, rq_table** out) {
    if (!E.init) return fail(RQ_ERR_NOT_INIT, "rq_table_upload before rq_init");
    if (!out || n_cols <= 0 || !cols || n_rows < 0) return fail(RQ_ERR_INVALID, "rq_table_upload: bad arguments");
    std::unique_ptr<rq_table> t(new rq_table());
    try {
        t->name = name ? name : "";
        t->n_rows = n_rows;
        const bool dev = flags & RQ_DEVICE_PTR;
        const bool borrow = dev && (flags & RQ_BORROW);
        t->borrowed = borrow;
        t->cap_rows = borrow ? n_rows : round_up(std::max<int64_t>(n_rows, 1), kPadRows);
        for (int c = 0; c < n_cols; c++) {
            if (!valid_col(cols[c].type, cols[c].width))
                raise(RQ_ERR_INVALID, "rq_table_upload: column %d has bad type/width %d/%d", c, cols[c].type, cols[c].width);
            DevColumn dc;
            dc.type = cols[c].type;
            dc.width = cols[c].width;
            if (borrow) {
                if (((uintptr_t)cols[c].data & 15) != 0)
                    raise(RQ_ERR_INVALID, "rq_table_upload: borrowed column %d is not 16-byte aligned", c);
                dc.d = (unsigned char*)cols[c].data;
                dc.owned = false;
            } else {
                const size_t bytes = (size_t)t->cap_rows * dc.width;
                CK(cudaMalloc(&dc.d, bytes));
                const size_t used = (size_t)n_rows * dc.width;
                if (used)
                    CK(cudaMemcpyAsync(dc.d, cols[c].data, used,
                                       dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, E.stream));
                if (bytes > used) CK(cudaMemsetAsync(dc.d + used, 0, bytes - used, E.stream));
            }
            t->cols.push_back(dc);
        }
        CK(cudaStreamSynchronize(E.stream));
    } catch (RqError& e) {
        return fail(e.code, "%s", e.msg.c_str());
    }
    *out = t.release();
    return RQ_OK;
}

extern "C" int rq_table_upload_rows(const char* name, int32_t n_cols, const int32_t* types,
                                    const int32_t* widths, const int32_t* offsets,
                                    int32_t tuple_size, int32_t n_blocks,
                                    const uint8_t* const* blocks, const size_t* block_bytes,
                                    rq_table** out) {
    if (!E.init) return fail(RQ_ERR_NOT_INIT, "rq_table_upload_rows before rq_init");
    if (!out || n_cols <= 0 || tuple_size <= 0 || n_blocks < 0) return fail(RQ_ERR_INVALID, "rq_table_upload_rows: bad arguments");
    std::unique_ptr<rq_table> t(new rq_table());
    unsigned char* d_rows[2] = {nullptr, nullptr};
    try {
        int64_t n_rows = 0;
        size_t max_block = 0;
        for (int b = 0; b < n_blocks; b++) {
            if (block_bytes[b] % tuple_size) raise(RQ_ERR_INVALID, "block %d holds a partial tuple", b);
            n_rows += block_bytes[b] / tuple_size;
            max_block = std::max(max_block, block_bytes[b]);
        }
        t->name = name ? name : "";
        t->n_rows = n_rows;
        t->cap_rows = round_up(std::max<int64_t>(n_rows, 1), kPadRows);
        for (int c = 0; c < n_cols; c++) {
            if (!valid_col(types[c], widths[c]) || offsets[c] < 0 || offsets[c] + widths[c] > tuple_size)
                raise(RQ_ERR_INVALID, "rq_table_upload_rows: column %d bad type/width/offset", c);
            DevColumn dc;
            dc.type = types[c];
            dc.width = widths[c];
            CK(cudaMalloc(&dc.d, (size_t)t->cap_rows * dc.width));
            CK(cudaMemsetAsync(dc.d, 0, (size_t)t->cap_rows * dc.width, E.stream));
            t->cols.push_back(dc);
        }
        if (max_block) {
            CK(cudaMalloc(&d_rows[0], max_block));
            CK(cudaMalloc(&d_rows[1], max_block));
        }
        int64_t row0 = 0;
        cudaEvent_t done[2];
        CK(cudaEventCreate(&done[0]));
        CK(cudaEventCreate(&done[1]));
        for (int b = 0; b < n_blocks; b++) {
            const int s = b & 1;
            const int64_t n = block_bytes[b] / tuple_size;
            if (n == 0) continue;
            if (b >= 2) CK(cudaEventSynchronize(done[s]));
            CK(cudaMemcpyAsync(d_rows[s], blocks[b], block_bytes[b], cudaMemcpyHostToDevice, E.stream));
            for (int c = 0; c < n_cols; c++) {
                rq_transpose_rows<<<(unsigned)((n + 255) / 256), 256, 0, E.stream>>>(
                    d_rows[s], n, tuple_size, offsets[c], widths[c], t->cols[c].d, row0);
            }
            CK(cudaEventRecord(done[s], E.stream));
            row0 += n;
        }
        CK(cudaStreamSynchronize(E.stream));
        CK(cudaGetLastError());
        cudaEventDestroy(done[0]);
        cudaEventDestroy(done[1]);
        cudaFree(d_rows[0]);
        cudaFree(d_rows[1]);
    } catch (RqError& e) {
        cudaFree(d_rows[0]);
        cudaFree(d_rows[1]);
        return fail(e.code, "%s", e.msg.c_str());
    }
    *out = t.release();
    return RQ_OK;
}

extern "C" int64_t rq_table_rows(const rq_table* t) { return t ? t->n_rows : -1; }
extern "C" int rq_table_free(rq_table* t) {
    if (E.init) cudaStreamSynchronize(E.stream);
    delete t;
    return RQ_OK;
}

// ------------------------------------------------------------------------------------------
// lowering: ABI postfix program -> accumulator-machine instructions (no code generation)
// ------------------------------------------------------------------------------------------
namespace {

struct PipeOut {                    // what a finished pipeline left on the device
    std::unique_ptr<rq_table> table;   // AGG / MATERIALIZE output (int64 columns)
    std::unique_ptr<HashTableDev> ht;  // BUILD output
    std::vector<int> payload_sql_type, payload_sql_width;
};

bool is_leaf(int op) { return op == RQ_OP_COL || op == RQ_OP_CONST || op == RQ_OP_CONST_STR; }
bool is_binary(int op) {
    switch (op) {
        case RQ_OP_ADD: case RQ_OP_SUB: case RQ_OP_MUL: case RQ_OP_DIV: case RQ_OP_AND: case RQ_OP_OR:
        case RQ_OP_LT: case RQ_OP_LE: case RQ_OP_GT: case RQ_OP_GE: case RQ_OP_EQ: case RQ_OP_NEQ:
        case RQ_OP_EQ_CHAR: case RQ_OP_EQ_VARCHAR: case RQ_OP_NEQ_CHAR: case RQ_OP_NEQ_VARCHAR:
        case RQ_OP_LIKE:
            return true;
        default:
            return false;
    }
}
// opcode when the LEFT operand is in the accumulator / when the RIGHT operand is
uint8_t dop_left(int op) {
    switch (op) {
        case RQ_OP_ADD: return D_ADD; case RQ_OP_SUB: return D_SUB; case RQ_OP_MUL: return D_MUL;
        case RQ_OP_DIV: return D_DIV; case RQ_OP_AND: return D_AND; case RQ_OP_OR: return D_OR;
        case RQ_OP_LT: return D_LT; case RQ_OP_LE: return D_LE; case RQ_OP_GT: return D_GT;
        case RQ_OP_GE: return D_GE; case RQ_OP_EQ: return D_EQ; case RQ_OP_NEQ: return D_NE;
        case RQ_OP_EQ_CHAR: return D_EQC; case RQ_OP_EQ_VARCHAR: return D_EQV;
        case RQ_OP_NEQ_CHAR: return D_NEC; case RQ_OP_NEQ_VARCHAR: return D_NEV;
        case RQ_OP_LIKE: return D_LIKE;
    }
    return D_NOP;
}
uint8_t dop_right(int op) {
    switch (op) {
        case RQ_OP_ADD: return D_ADD; case RQ_OP_SUB: return D_RSUB; case RQ_OP_MUL: return D_MUL;
        case RQ_OP_DIV: return D_RDIV; case RQ_OP_AND: return D_AND; case RQ_OP_OR: return D_OR;
        case RQ_OP_LT: return D_GT; case RQ_OP_LE: return D_GE; case RQ_OP_GT: return D_LT;
        case RQ_OP_GE: return D_LE; case RQ_OP_EQ: return D_EQ; case RQ_OP_NEQ: return D_NE;
        case RQ_OP_EQ_CHAR: return D_EQC; case RQ_OP_EQ_VARCHAR: return D_EQV;
        case RQ_OP_NEQ_CHAR: return D_NEC; case RQ_OP_NEQ_VARCHAR: return D_NEV;
        case RQ_OP_LIKE: return D_RLIKE;
    }
    return D_NOP;
}
// selection-fused compare: column CMP constant (constant on the right / on the left)
uint8_t fcmp_left(int op) {
    switch (op) {
        case RQ_OP_LT: return D_FLT; case RQ_OP_LE: return D_FLE; case RQ_OP_GT: return D_FGT;
        case RQ_OP_GE: return D_FGE; case RQ_OP_EQ: return D_FEQ; case RQ_OP_NEQ: return D_FNE;
    }
    return 0;
}
uint8_t fcmp_right(int op) {
    switch (op) {
        case RQ_OP_LT: return D_FGT; case RQ_OP_LE: return D_FGE; case RQ_OP_GT: return D_FLT;
        case RQ_OP_GE: return D_FLE; case RQ_OP_EQ: return D_FEQ; case RQ_OP_NEQ: return D_FNE;
    }
    return 0;
}

struct Operand {
    uint8_t src = S_NONE;
    uint16_t idx = 0;
    int64_t imm = 0;
};

struct HRef {            // host-level value reference of a sink (resolved to VRef by encode)
    uint8_t kind = S_NONE;   // DSrc
    uint16_t idx = 0;        // staged column / slot / imm-table index / string column
};

// Host-level program of one pipeline (what tests/vm_model.py executes) plus the pieces of
// KParams that do not depend on the shared-memory layout.
struct Lowerer {
    const rq_plan& plan;
    const rq_pipeline& pl;
    const rq_table& src;
    const std::vector<PipeOut>& outs;
    const char* d_strpool;
    KParams& P;

    int n;
    std::vector<DInsn> prog;
    std::vector<HRef> hkey, hout;
    HRef hagg_src[kMaxAggs];
    HRef hprobe_key[kMaxProbes][kMaxKeys];
    int n_slots = 0;

    std::vector<int> uses;            // consumers per node
    std::vector<int> last_use;        // last consuming node index (n = sink)
    std::vector<char> sink_ref;       // referenced by the sink (needs a slot unless leaf)
    std::vector<int> slot;            // assigned slot or -1
    std::vector<Operand> leaf_op;     // operand descriptor of leaves
    std::vector<int> staged_of_col;   // source column -> staged index / str index
    std::vector<int> free_slots;
    std::vector<char> fused;          // FILTER nodes folded into a selection-fused compare
    int acc_node = -1;
    int n_imm = 0;

    Lowerer(const rq_plan& plan, const rq_pipeline& pl, const rq_table& src,
            const std::vector<PipeOut>& outs, const char* d_strpool, KParams& P)
        : plan(plan), pl(pl), src(src), outs(outs), d_strpool(d_strpool), P(P), n(pl.n_nodes) {}

    void check_ref(int i, int ref) {
        if (ref < 0 || ref >= i) raise(RQ_ERR_INVALID, "node %d refers to node %d (must be an earlier node)", i, ref);
    }

    void emit(uint8_t op, Operand o = Operand(), uint16_t aux = 0) {
        if ((int)prog.size() >= kMaxInsn - 1) raise(RQ_ERR_UNSUPPORTED, "program longer than %d instructions", kMaxInsn - 1);
        DInsn in;
        in.op = op; in.src = o.src; in.flags = 0; in.dst = 0; in.idx = o.idx; in.aux = aux; in.imm = o.imm;
        prog.push_back(in);
    }

    Operand operand_of(int node) {
        if (is_leaf(pl.nodes[node].op)) return leaf_op[node];
        if (slot[node] < 0) raise(RQ_ERR_INVALID, "internal: node %d has no slot", node);
        Operand o; o.src = S_SLOT; o.idx = (uint16_t)slot[node];
        return o;
    }
    HRef href_of(int node) {
        HRef v;
        if (is_leaf(pl.nodes[node].op)) {
            Operand o = leaf_op[node];
            if (o.src == S_IMM) {
                if (n_imm >= kMaxImm) raise(RQ_ERR_UNSUPPORTED, "too many constants in sink");
                P.imm[n_imm] = o.imm;
                v.kind = S_IMM; v.idx = (uint16_t)n_imm++;
            } else { v.kind = o.src; v.idx = o.idx; }
            return v;
        }
        if (slot[node] < 0) raise(RQ_ERR_INVALID, "internal: sink node %d has no slot", node);
        v.kind = S_SLOT; v.idx = (uint16_t)slot[node];
        return v;
    }

    int alloc_slot() {
        if (!free_slots.empty()) { int s = free_slots.back(); free_slots.pop_back(); return s; }
        if (n_slots >= kMaxSlots) raise(RQ_ERR_UNSUPPORTED, "expression needs more than %d live temporaries", kMaxSlots);
        return n_slots++;
    }
    void release_dead(int at) {   // free slots of nodes whose last use is `at`
        for (int i = 0; i < n; i++)
            if (slot[i] >= 0 && last_use[i] == at) { free_slots.push_back(slot[i]); slot[i] = -2 - slot[i]; }
    }

    void prepare() {
        uses.assign(n, 0); last_use.assign(n, -1); sink_ref.assign(n, 0); slot.assign(n, -1);
        leaf_op.assign(n, Operand()); fused.assign(n, 0);
        staged_of_col.assign(src.cols.size(), -1);
        auto use = [&](int i, int ref) { check_ref(i, ref); uses[ref]++; last_use[ref] = std::max(last_use[ref], i); };
        for (int i = 0; i < n; i++) {
            const rq_node& nd = pl.nodes[i];
            switch (nd.op) {
                case RQ_OP_COL: {
                    if (nd.a < 0 || nd.a >= (int)src.cols.size()) raise(RQ_ERR_INVALID, "node %d: column %d out of range", i, nd.a);
                    const DevColumn& dc = src.cols[nd.a];
                    Operand o;
                    if (dc.type == RQ_STR) {
                        if (staged_of_col[nd.a] < 0) {
                            if (P.n_strcols >= kMaxStrCols) raise(RQ_ERR_UNSUPPORTED, "too many string columns");
                            P.str_ptr[P.n_strcols] = dc.d; P.str_w[P.n_strcols] = dc.width;
                            staged_of_col[nd.a] = P.n_strcols++;
                        }
                        o.src = S_STR;
                    } else {
                        if (staged_of_col[nd.a] < 0) {
                            if (P.n_cols >= kMaxStagedCols) raise(RQ_ERR_UNSUPPORTED, "more than %d columns in one pipeline", kMaxStagedCols);
                            P.col_ptr[P.n_cols] = dc.d; P.col_w[P.n_cols] = (uint8_t)dc.width;
                            staged_of_col[nd.a] = P.n_cols++;
                        }
                        o.src = S_COL;
                    }
                    o.idx = (uint16_t)staged_of_col[nd.a];
                    leaf_op[i] = o;
                    break;
                }
                case RQ_OP_CONST: { Operand o; o.src = S_IMM; o.imm = nd.imm; leaf_op[i] = o; break; }
                case RQ_OP_CONST_STR: {
                    if (nd.imm < 0 || nd.imm >= plan.strpool_bytes) raise(RQ_ERR_INVALID, "node %d: string offset out of range", i);
                    Operand o; o.src = S_IMM; o.imm = (int64_t)(d_strpool + nd.imm); leaf_op[i] = o; break;
                }
                case RQ_OP_FILTER: use(i, nd.a); break;
                case RQ_OP_SELECT: use(i, nd.a); use(i, nd.b); use(i, nd.c); break;
                case RQ_OP_PROBE: {
                    if (nd.b < 0 || nd.c < 0 || nd.b + nd.c > pl.n_args) raise(RQ_ERR_INVALID, "node %d: probe args out of range", i);
                    for (int k = 0; k < nd.c; k++) { use(i, pl.args[nd.b + k]); sink_ref[pl.args[nd.b + k]] = 1; }
                    break;
                }
                case RQ_OP_PAYLOAD: check_ref(i, nd.a); break;
                default:
                    if (is_binary(nd.op)) { use(i, nd.a); use(i, nd.b); }
                    else raise(RQ_ERR_INVALID, "node %d: unknown op %d", i, nd.op);
            }
        }
        auto sink_use = [&](int ref) {
            if (ref < 0 || ref >= n) raise(RQ_ERR_INVALID, "sink refers to node %d", ref);
            uses[ref]++; last_use[ref] = n; sink_ref[ref] = 1;
        };
        for (int k = 0; k < pl.n_keys; k++) sink_use(pl.keys[k].node);
        for (int k = 0; k < pl.n_vals; k++)
            if (!(pl.sink_kind == RQ_SINK_AGG && pl.vals[k].kind == RQ_AGG_COUNT)) sink_use(pl.vals[k].node);
        // stage layout: one warp tile of every staged column
        uint32_t off = 0;
        for (int c = 0; c < P.n_cols; c++) { P.col_off[c] = off; off += kTile * P.col_w[c]; }
        P.stage_bytes = off;
        mark_fused_filters();
    }

    // FILTER(COL cmp CONST) where the compare has no other consumer becomes one instruction that
    // leaves the accumulator alone. 4- and 1-byte columns compare in 32 bits, so the constant
    // must fit (it always does for dates / flags; otherwise the generic form is used).
    int fcmp_of(int f, Operand* col, int64_t* imm) const {
        const rq_node& fl = pl.nodes[f];
        if (fl.op != RQ_OP_FILTER) return 0;
        const rq_node& cm = pl.nodes[fl.a];
        if (uses[fl.a] != 1 || sink_ref[fl.a]) return 0;
        if (!fcmp_left(cm.op)) return 0;
        const int xo = pl.nodes[cm.a].op, yo = pl.nodes[cm.b].op;
        int code = 0, cn = -1, kn = -1;
        if (xo == RQ_OP_COL && yo == RQ_OP_CONST) { code = fcmp_left(cm.op); cn = cm.a; kn = cm.b; }
        else if (xo == RQ_OP_CONST && yo == RQ_OP_COL) { code = fcmp_right(cm.op); cn = cm.b; kn = cm.a; }
        else return 0;
        const Operand o = leaf_op[cn];
        if (o.src != S_COL) return 0;
        const int64_t k = leaf_op[kn].imm;
        if (P.col_w[o.idx] != 8 && (k < INT32_MIN || k > INT32_MAX)) return 0;
        *col = o; *imm = k;
        return code;
    }
    void mark_fused_filters() {
        for (int f = 0; f < n; f++) {
            Operand o; int64_t k;
            if (fcmp_of(f, &o, &k)) { fused[f] = 1; fused[pl.nodes[f].a] = 1; }
        }
    }

    // ---- slot decision -------------------------------------------------------------------
    // A computed value sits in the accumulator until the next clobbering instruction. A consumer
    // can take it from there if it is a FILTER inside that window or the clobbering node itself
    // (using it as exactly one operand), or an aggregate fused right behind its input.
    bool clobbers(int j) const {
        const int op = pl.nodes[j].op;
        return !is_leaf(op) && op != RQ_OP_FILTER && op != RQ_OP_PAYLOAD && !fused[j];
    }
    int gpos = -1;                 // GROUP is emitted right after node gpos
    std::vector<std::vector<int>> aggs_of;   // node -> aggregate indices fed by it
    bool lowagg = false;

    void decide_slots() {
        std::vector<std::vector<int>> cons(n);
        for (int j = 0; j < n; j++) {
            const rq_node& nd = pl.nodes[j];
            if (is_binary(nd.op)) { cons[nd.a].push_back(j); cons[nd.b].push_back(j); }
            else if (nd.op == RQ_OP_FILTER) cons[nd.a].push_back(j);
            else if (nd.op == RQ_OP_SELECT) { cons[nd.a].push_back(j); cons[nd.b].push_back(j); cons[nd.c].push_back(j); }
        }
        for (int i = 0; i < n; i++) {
            const int op = pl.nodes[i].op;
            if (is_leaf(op) || op == RQ_OP_FILTER || op == RQ_OP_PROBE || fused[i]) continue;
            if (op == RQ_OP_PAYLOAD) continue;   // slot handed out when the PROBE is emitted
            bool need = false;
            if (sink_ref[i]) {
                // only a low-card aggregate fused behind its input reads the accumulator
                bool all_fused = lowagg && i > gpos && !aggs_of[i].empty();
                for (int k = 0; k < pl.n_keys; k++) if (pl.keys[k].node == i) all_fused = false;
                for (int j = 0; j < n; j++)
                    if (pl.nodes[j].op == RQ_OP_PROBE)
                        for (int k = 0; k < pl.nodes[j].c; k++) if (pl.args[pl.nodes[j].b + k] == i) all_fused = false;
                if (!all_fused) need = true;
            }
            int wend = -1;
            for (int j = i + 1; j < n; j++) if (clobbers(j)) { wend = j; break; }
            for (int j : cons[i]) {
                if (wend >= 0 && j > wend) need = true;
                if (j == wend) {
                    const rq_node& nd = pl.nodes[j];
                    int cnt = 0;
                    if (is_binary(nd.op)) cnt = (nd.a == i) + (nd.b == i);
                    else if (nd.op == RQ_OP_SELECT) {
                        cnt = (nd.a == i) ? 1 : 2;
                        if (nd.b == i || nd.c == i) cnt = 2;
                        // a non-constant leaf else is staged through the accumulator first
                        const int co = pl.nodes[nd.c].op;
                        if (is_leaf(co) && co != RQ_OP_CONST && co != RQ_OP_CONST_STR) cnt = 2;
                    }
                    else cnt = 2;
                    if (cnt != 1) need = true;
                }
            }
            if (need) slot[i] = -3;   // marker: allocate at emission
        }
    }
};

}  // namespace

#include "engine_exec.inl"
