// Included by engine.cu (after engine_exec.inl): parallel `.tbl` text loader, rq_table_load_tbl.
//
// Replaces executeBulkInsert (execute.h:332-388) for loads that go straight to the GPU: the reference
// reads the file line by line on one thread, builds an Expr per field (never freed: ~0.1-0.2 kB per
// field) and fills its row store. Here the file is cut into segments at line ends,
// every segment crosses PCIe as text, and the GPU finds the line starts (count / scan / scatter) and
// parses every row with one thread, writing the typed values into the tile-major table. Field
// semantics are the reference's parse*Constant functions (expressions.h:369-455):
//   INT      std::stoi                     BIGINT  (int32_t) std::stoll  - yes, truncated to 32 bits
//   DECIMAL  the first '.' removed, std::stoll of the rest (no rescale to the column's scale)
//   DATE     %4d-%2d-%2d or %4d/%2d/%2d -> y * 10000 + m * 100 + d
//   BOOL     "true" / anything else false  CHAR(1) first byte
//   CHAR(n) / VARCHAR(n)  the token's first n bytes, NUL padded (values.h:151-198)
// A line is the text up to '\n'; fields end at the terminator; a terminator at the end of the line
// starts no further field (std::getline); fewer or more fields than columns is an error, as there.

namespace {

constexpr int kTblSliceBytes = 16384;          // text bytes per counting / scattering thread block

struct TblCol {
    int32_t sql_type;        // RQ_SQL_*
    int32_t phys;            // RQ_I8 / RQ_I32 / RQ_I64 / RQ_STR
    int32_t width;           // physical bytes per row
    int32_t pad_;
    unsigned char* dst;      // tile-major chunk of page 0, or the plain string array
    int64_t tile_stride;
};
struct TblSchema {
    int32_t n_cols;
    char terminator;
    TblCol col[kMaxStagedCols * 2];
};

// line ends per slice
__global__ void rq_tbl_count(const char* text, int64_t n, int32_t* counts) {
    const int64_t lo = (int64_t)blockIdx.x * kTblSliceBytes, hi = min(lo + (int64_t)kTblSliceBytes, n);
    int c = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) c += text[i] == '\n';
    c = (int)warp_reduce((int64_t)c, 1);
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = s;
}
// exclusive scan of the slice counts (one block; the slices of a segment number a few thousand)
__global__ void rq_tbl_scan(int32_t* counts, int n) {
    __shared__ int carry;
    __shared__ int warp_sums[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n ? counts[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = threadIdx.x < (blockDim.x >> 5) ? warp_sums[threadIdx.x] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = carry + ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < n) counts[i] = before;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
}
// starts[r + 1] = offset behind the r-th line end of the segment (starts[0] = 0 is written by the host)
__global__ void rq_tbl_starts(const char* text, int64_t n, const int32_t* prefix, int64_t* starts) {
    const int64_t lo = (int64_t)blockIdx.x * kTblSliceBytes, hi = min(lo + (int64_t)kTblSliceBytes, n);
    __shared__ int base;
    if (threadIdx.x == 0) base = prefix[blockIdx.x];
    __syncthreads();
    // ordered within the slice: 256-byte strips, ballot per warp of 32 bytes
    for (int64_t strip = lo; strip < hi; strip += blockDim.x) {
        const int64_t i = strip + threadIdx.x;
        const bool nl = i < hi && text[i] == '\n';
        const unsigned m = __ballot_sync(0xffffffffu, nl);
        __shared__ int wcount[32];
        if ((threadIdx.x & 31) == 0) wcount[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        int before = base;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += wcount[w];
        if (nl) starts[before + __popc(m & ((1u << (threadIdx.x & 31)) - 1)) + 1] = i + 1;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += wcount[w]; base += t; }
        __syncthreads();
    }
}

__device__ __forceinline__ long long tbl_stoll(const char* p, const char* end, bool skip_dot) {
    while (p < end && (*p == ' ' || *p == '\t')) p++;
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
    unsigned long long v = 0;
    bool dot_seen = false;
    for (; p < end; p++) {
        if (*p >= '0' && *p <= '9') v = v * 10ULL + (unsigned long long)(*p - '0');
        else if (skip_dot && *p == '.' && !dot_seen) dot_seen = true;
        else break;
    }
    return neg ? (long long)(0ULL - v) : (long long)v;
}

// one thread per line: err[0] = 1 + first bad line (relative to row0), err[1] = 1 missing / 2 extra attributes
__global__ void rq_tbl_parse(const char* text, int64_t n_bytes, const int64_t* starts, int64_t n_rows, int64_t row0,
                             TblSchema S, unsigned long long* err) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const char* p = text + starts[r];
    const char* end = text + (r + 1 < n_rows || starts[r + 1] > 0 ? starts[r + 1] : n_bytes);
    if (end > p && end[-1] == '\n') end--;
    const int64_t row = row0 + r;
    int c = 0;
    while (p < end) {
        const char* q = p;
        while (q < end && *q != S.terminator) q++;
        if (c >= S.n_cols) { atomicMin(&err[0], (unsigned long long)r * 4 + 2); return; }
        const TblCol& col = S.col[c];
        if (col.dst) {
            unsigned char* d = col.dst + (col.phys == RQ_STR ? (size_t)row * col.width
                                                             : (size_t)(row / kTile) * col.tile_stride + (size_t)(row % kTile) * col.width);
            switch (col.sql_type) {
                case RQ_SQL_INT: *reinterpret_cast<int32_t*>(d) = (int32_t)tbl_stoll(p, q, false); break;
                case RQ_SQL_BIGINT: *reinterpret_cast<int64_t*>(d) = (int64_t)(int32_t)tbl_stoll(p, q, false); break;
                case RQ_SQL_DECIMAL: *reinterpret_cast<int64_t*>(d) = tbl_stoll(p, q, true); break;
                case RQ_SQL_DATE: {
                    int f[3] = {0, 0, 0}, k = 0, digits = 0;
                    const int lim[3] = {4, 2, 2};
                    for (const char* t = p; t < q && k < 3; t++) {
                        if (*t >= '0' && *t <= '9' && digits < lim[k]) { f[k] = f[k] * 10 + (*t - '0'); digits++; }
                        else if (*t == '-' || *t == '/') { k++; digits = 0; }
                        else break;
                    }
                    *reinterpret_cast<int32_t*>(d) = f[0] * 10000 + f[1] * 100 + f[2];
                    break;
                }
                case RQ_SQL_BOOL: *d = (q - p == 4 && p[0] == 't' && p[1] == 'r' && p[2] == 'u' && p[3] == 'e') ? 1 : 0; break;
                default:          // CHAR / VARCHAR
                    if (col.phys == RQ_I8) { *d = q > p ? (unsigned char)*p : 0; break; }
                    {
                        const int nmax = col.width - 1;
                        int k = 0;
                        for (; k < nmax && p + k < q; k++) d[k] = (unsigned char)p[k];
                        for (; k < col.width; k++) d[k] = 0;
                    }
                    break;
            }
        }
        c++;
        p = q < end ? q + 1 : end;
    }
    if (c < S.n_cols) atomicMin(&err[0], (unsigned long long)r * 4 + 1);
}

}  // namespace

extern "C" int rq_table_load_tbl(const char* name, const char* path, char terminator, int32_t n_cols,
                                 const int32_t* sql_types, const int32_t* sql_widths, rq_table** out) {
    if (!E.init) return fail(RQ_ERR_NOT_INIT, "rq_table_load_tbl before rq_init");
    if (!out || !path || n_cols <= 0 || n_cols > kMaxStagedCols * 2 || !sql_types || !sql_widths)
        return fail(RQ_ERR_INVALID, "rq_table_load_tbl: bad arguments");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(RQ_ERR_INVALID, "Could not open file %s", path);
    std::unique_ptr<rq_table> t(new rq_table());
    char* h_buf[2] = {nullptr, nullptr};
    char* d_text[2] = {nullptr, nullptr};
    int32_t* d_counts = nullptr;
    int64_t* d_starts = nullptr;
    unsigned long long* d_err = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
        if (f) fclose(f);
        for (int b = 0; b < 2; b++) { if (h_buf[b]) cudaFreeHost(h_buf[b]); dfree(d_text[b]); if (done[b]) cudaEventDestroy(done[b]); }
        dfree(d_counts); dfree(d_starts); dfree(d_err);
    };
    try {
        // ---- pass 1 (host): file size, line count per segment; segments end at line ends ------------
        fseek(f, 0, SEEK_END);
        const int64_t file_bytes = ftell(f);
        fseek(f, 0, SEEK_SET);
        const int64_t kSeg = 64LL << 20;
        struct Seg { int64_t off, bytes, rows; };
        std::vector<Seg> segs;
        {
            std::vector<char> buf((size_t)kSeg);
            int64_t off = 0, seg_off = 0, pending = 0, total_rows = 0;
            while (off < file_bytes) {
                const int64_t n = (int64_t)fread(buf.data(), 1, (size_t)std::min<int64_t>(kSeg, file_bytes - off), f);
                if (n <= 0) raise(RQ_ERR_INVALID, "read error in %s", path);
                int64_t last_nl = -1, rows = 0;
                for (int64_t i = 0; i < n; i++) if (buf[(size_t)i] == '\n') { rows++; last_nl = i; }
                if (last_nl >= 0) {
                    // [seg_off, off + last_nl + 1) is a whole number of lines
                    segs.push_back({seg_off, off + last_nl + 1 - seg_off, rows});
                    total_rows += rows;
                    seg_off = off + last_nl + 1;
                }
                pending = off + n - seg_off;
                off += n;
            }
            if (pending > 0) { segs.push_back({seg_off, pending, 1}); total_rows += 1; }   // last line without '\n'
            t->n_rows = total_rows;
        }
        int64_t max_seg = 1, max_rows = 1;
        for (auto& s : segs) { max_seg = std::max(max_seg, s.bytes); max_rows = std::max(max_rows, s.rows); }
        t->name = name ? name : "";
        t->cap_rows = round_up(std::max<int64_t>(t->n_rows, 1), kPadRows);
        std::vector<int> pt(n_cols), pw(n_cols);
        for (int c = 0; c < n_cols; c++) pt[c] = phys_type(sql_types[c], sql_widths[c], &pw[c]);
        alloc_tile_major(*t, n_cols, [&](int c) { return pt[c]; }, [&](int c) { return pw[c]; });
        TblSchema S;
        memset(&S, 0, sizeof(S));
        S.n_cols = n_cols;
        S.terminator = terminator;
        for (int c = 0; c < n_cols; c++) {
            S.col[c].sql_type = sql_types[c]; S.col[c].phys = pt[c]; S.col[c].width = pw[c];
            S.col[c].dst = t->cols[c].d; S.col[c].tile_stride = t->cols[c].tile_stride;
        }
        // ---- pass 2: stream the segments (double-buffered pinned staging), mark and parse on the GPU ----
        const int max_slices = (int)((max_seg + kTblSliceBytes - 1) / kTblSliceBytes);
        for (int b = 0; b < 2; b++) {
            CK(cudaMallocHost(&h_buf[b], (size_t)max_seg));
            CK(dmalloc(&d_text[b], (size_t)max_seg));
            CK(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
        }
        CK(dmalloc(&d_counts, sizeof(int32_t) * (size_t)max_slices));
        CK(dmalloc(&d_starts, sizeof(int64_t) * (size_t)(max_rows + 2)));
        CK(dmalloc(&d_err, 16));
        CK(cudaMemsetAsync(d_err, 0xff, 16, E.stream));
        int64_t row0 = 0;
        std::vector<unsigned long long> h_err(2);
        for (size_t k = 0; k < segs.size(); k++) {
            const int b = (int)(k & 1);
            const Seg& s = segs[k];
            if (k >= 2) CK(cudaEventSynchronize(done[b]));
            fseek(f, (long)s.off, SEEK_SET);
            if ((int64_t)fread(h_buf[b], 1, (size_t)s.bytes, f) != s.bytes) raise(RQ_ERR_INVALID, "read error in %s", path);
            CK(cudaMemcpyAsync(d_text[b], h_buf[b], (size_t)s.bytes, cudaMemcpyHostToDevice, E.stream));
            const int slices = (int)((s.bytes + kTblSliceBytes - 1) / kTblSliceBytes);
            CK(cudaMemsetAsync(d_starts, 0, sizeof(int64_t) * (size_t)(s.rows + 2), E.stream));
            rq_tbl_count<<<slices, 256, 0, E.stream>>>(d_text[b], s.bytes, d_counts);
            rq_tbl_scan<<<1, 1024, 0, E.stream>>>(d_counts, slices);
            rq_tbl_starts<<<slices, 256, 0, E.stream>>>(d_text[b], s.bytes, d_counts, d_starts);
            rq_tbl_parse<<<(unsigned)((s.rows + 127) / 128), 128, 0, E.stream>>>(d_text[b], s.bytes, d_starts, s.rows, row0, S, d_err);
            CK(cudaGetLastError());
            CK(cudaEventRecord(done[b], E.stream));
            // errors are looked at per segment so that the line number can be reported
            CK(cudaMemcpyAsync(h_err.data(), d_err, 16, cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            if (h_err[0] != ~0ULL) {
                const long long line = (long long)(row0 + (int64_t)(h_err[0] / 4));
                if ((h_err[0] & 3) == 2) raise(RQ_ERR_INVALID, "Line %lld in %s contains extra attributes.", line, path);
                raise(RQ_ERR_INVALID, "Line %lld in %s is missing attributes.", line, path);
            }
            row0 += s.rows;
        }
        compute_stats(*t);
        CK(cudaStreamSynchronize(E.stream));
    } catch (RqError& e) {
        cudaStreamSynchronize(E.stream);
        cleanup();
        return fail(e.code, "%s", e.msg.c_str());
    }
    cleanup();
    *out = t.release();
    return RQ_OK;
}
