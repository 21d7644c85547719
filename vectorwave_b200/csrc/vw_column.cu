// vw_column.cu -- single-level MODWT kernels for DEEP levels (dilation d >= 32), where the dilated halo
// (L-1)*d no longer fits beside a tile in shared memory.
//
// At dilation d the transform splits into d independent "columns" (positions p = q*d + phi share phi): level j on
// the signal is the undilated filter on each column.  A thread owns one column phi and walks a chunk of rows q
// with a register-resident sliding window: per block of R rows it loads R new samples (lanes hold consecutive
// phi, so every warp load / store is one contiguous 256-byte segment), produces R outputs x 2 filters of FP64 FMAs
// with constant-bank taps, then shifts the last L-1 samples down.  Every input is read once (plus L-1 warm-up rows
// per chunk), every output written once: 24 B/sample/level, no shared memory, no recompute.
//
// Analysis:   V_j[p] = sum_k hs[k] X[ext(p - k d)],  W_j[p] likewise with gs       (ScalarOps.java:700-723,790-835)
// Synthesis:  out[p] = sum_k hs[k] V[ext(p + sh (k d - th))] + sum_k gs[k] W[ext(p + sg (k d - tg))]
//             (MultiLevelMODWTTransform.java:554-645; sigma = -1 streams are run with reversed taps)
#include "vw_internal.cuh"

namespace {

constexpr int kCR = 9;         // rows per register block (synthesis; analysis uses col_rows<L>)
template <int L> struct col_rows { static constexpr int value = L >= 24 ? 7 : 9; };
constexpr int kCThreads = 128;

__device__ __forceinline__ int64_t wrap_mod(int64_t i, int64_t n) { i %= n; return i < 0 ? i + n : i; }
// EDGE == false: the caller has proven pos is inside [0, n) (interior chunks, the overwhelming majority)
template <bool EDGE>
__device__ __forceinline__ double ext_load(const double *__restrict__ row, int64_t pos, int64_t n, int mode) {
    if (!EDGE) return __ldg(row + pos);
    if (pos >= 0 && pos < n) return __ldg(row + pos);
    if (mode == VW_PERIODIC) return __ldg(row + wrap_mod(pos, n));
    if (mode == VW_SYMMETRIC) { int64_t m = wrap_mod(pos, 2 * n); return __ldg(row + (m < n ? m : 2 * n - 1 - m)); }
    return 0.0;
}

struct ColArgs {
    const double *x; long long ldx;          // analysis input / synthesis V (may be null => zeros)
    const double *w_in; long long ldw_in;    // synthesis W (may be null => zeros)
    double *v; long long ldv;                // analysis V out / synthesis out
    double *w; long long ldw;                // analysis W out
    long long n_in, t0, n_out, batch, d;     // outputs cover positions [t0, t0+n_out) of the input coordinate system
    long long off_h, off_g;                  // synthesis stream offsets (see below)
    int rows_per_chunk, chunks, mode;
    VwFilt32 f;                              // synthesis: taps already reversed for sigma = -1 streams
};

// ---- analysis ------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kCThreads, (L >= 24) ? 2 : ((L >= 16) ? 3 : 4)) k_column_analysis(const __grid_constant__ ColArgs a) {
    constexpr int R = col_rows<L>::value;
    // flattened (chunk, column) index, column fastest: 32 | d keeps every warp inside one chunk => coalesced rows
    const long long gid = (long long)blockIdx.x * kCThreads + threadIdx.x;
    const long long col = gid & (a.d - 1);   // phase phi in [0, d)
    const long long chunk = gid / a.d;
    if (chunk >= a.chunks) return;
    for (long long b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const double *x = a.x + b * a.ldx;
        double *vo = a.v + b * a.ldv, *wo = a.w + b * a.ldw;
        // rows of this column inside the output range: positions t0 + col + q*d < t0 + n_out
        const long long rows = (a.n_out - col + a.d - 1) / a.d;
        long long q = chunk * a.rows_per_chunk;
        const long long qend = q + a.rows_per_chunk < rows ? q + a.rows_per_chunk : rows;
        if (q >= qend) continue;
        long long p = a.t0 + col + q * a.d;                // input-coordinate position of the current row
        // seq[0 .. L-2] = the L-1 rows before the block (oldest first), seq[L-1 + r] = the block's new rows
        double seq[L - 1 + R];
        if (p - (long long)(L - 1) * a.d >= 0) {
#pragma unroll
            for (int i = 0; i < L - 1; i++) seq[i] = ext_load<false>(x, p - (long long)(L - 1 - i) * a.d, a.n_in, a.mode);
        } else {
#pragma unroll
            for (int i = 0; i < L - 1; i++) seq[i] = ext_load<true>(x, p - (long long)(L - 1 - i) * a.d, a.n_in, a.mode);
        }
        // software pipeline: the next block's rows are in flight while this block's FMAs run
        double nxt[R];
#pragma unroll
        for (int r = 0; r < R; r++) nxt[r] = (q + r < qend) ? __ldg(x + p + r * a.d) : 0.0;   // output rows lie inside [0, n_in)
        for (; q < qend; q += R) {
#pragma unroll
            for (int r = 0; r < R; r++) seq[L - 1 + r] = nxt[r];
            {
                const long long qn = q + R, pn = p + (long long)R * a.d;
#pragma unroll
                for (int r = 0; r < R; r++) nxt[r] = (qn + r < qend) ? __ldg(x + pn + r * a.d) : 0.0;
            }
            double ah[R], ag[R];
#pragma unroll
            for (int r = 0; r < R; r++) { ah[r] = 0.0; ag[r] = 0.0; }
            // out[r] = sum_k f[k] seq[L-1+r-k]; walk seq downwards so every output meets its taps in ascending order
#pragma unroll
            for (int m = R - 1; m >= -(L - 1); m--) {
                const double xv = seq[L - 1 + m];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int k = r - m;
                    if (k >= 0 && k < L) { ah[r] = fma(a.f.h[k], xv, ah[r]); ag[r] = fma(a.f.g[k], xv, ag[r]); }
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (q + r < qend) {
                    const long long o = p - a.t0 + r * a.d;
                    vo[o] = ah[r];
                    wo[o] = ag[r];
                }
            }
#pragma unroll
            for (int i = 0; i < L - 1; i++) seq[i] = seq[i + R];   // slide: keep the last L-1 rows
            p += (long long)R * a.d;
        }
    }
}

// ---- synthesis -----------------------------------------------------------------------------------------------
// out[o] = sum_k th[k] V[pos(o+k) + off_h] + sum_k tg[k] W[pos(o+k) + off_g],  pos(m) = t0 + col + m d
// with (taps, off) = (f, -tau) for sigma=+1 and (reversed f, tau - (L-1) d) for sigma=-1.
// Transposed FIR: the thread keeps the L-1 unfinished output sums instead of two L-1 deep input windows (half the
// registers): input row m adds th[k] V_m + tg[k] W_m to output m-k; an output is complete after input row o+L-1.
template <int L, bool EDGE>
__device__ __forceinline__ void col_synth_chunk(const ColArgs &a, const double *__restrict__ v, const double *__restrict__ w,
                                                double *__restrict__ out, long long col, long long o_start, long long o_end,
                                                long long pin) {
    constexpr int R = col_rows<L>::value;
    // acc[j] <-> output row  m0 - (L-1) + j ; rows below o_start are never emitted
    double acc[L - 1 + R];
#pragma unroll
    for (int j = 0; j < L - 1 + R; j++) acc[j] = 0.0;
    const long long m_end = o_end + (L - 1);           // one past the last input row this chunk consumes
    // software pipeline: block i+1's rows are in flight while block i's FMAs run
    double nv[R], nw[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const bool live = o_start + r < m_end;
        const long long pos = pin + r * a.d;
        nv[r] = (live && v) ? ext_load<EDGE>(v, pos + a.off_h, a.n_in, a.mode) : 0.0;
        nw[r] = (live && w) ? ext_load<EDGE>(w, pos + a.off_g, a.n_in, a.mode) : 0.0;
    }
    for (long long m0 = o_start; m0 < m_end; m0 += R) {
        double cv[R], cw[R];
#pragma unroll
        for (int r = 0; r < R; r++) { cv[r] = nv[r]; cw[r] = nw[r]; }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const bool live = m0 + R + r < m_end;
            const long long pos = pin + (long long)(R + r) * a.d;
            nv[r] = (live && v) ? ext_load<EDGE>(v, pos + a.off_h, a.n_in, a.mode) : 0.0;
            nw[r] = (live && w) ? ext_load<EDGE>(w, pos + a.off_g, a.n_in, a.mode) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int k = 0; k < L; k++) {
                const int j = r + L - 1 - k;
                acc[j] = fma(a.f.h[k], cv[r], acc[j]);
                acc[j] = fma(a.f.g[k], cw[r], acc[j]);
            }
        }
        // outputs j in [0, R) are complete: rows m0-(L-1) .. m0-(L-1)+R-1
#pragma unroll
        for (int j = 0; j < R; j++) {
            const long long o = m0 - (L - 1) + j;
            if (o >= o_start && o < o_end) out[col + o * a.d] = acc[j];
        }
#pragma unroll
        for (int j = 0; j < L - 1; j++) acc[j] = acc[j + R];
#pragma unroll
        for (int j = L - 1; j < L - 1 + R; j++) acc[j] = 0.0;
        pin += (long long)R * a.d;
    }
}

template <int L>
__global__ void __launch_bounds__(kCThreads, (L >= 24) ? 2 : ((L >= 16) ? 3 : 4)) k_column_synthesis(const __grid_constant__ ColArgs a) {
    constexpr int R = col_rows<L>::value;
    const long long gid = (long long)blockIdx.x * kCThreads + threadIdx.x;
    const long long col = gid & (a.d - 1);
    const long long chunk = gid / a.d;
    if (chunk >= a.chunks) return;
    for (long long b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const double *v = a.x ? a.x + b * a.ldx : nullptr;
        const double *w = a.w_in ? a.w_in + b * a.ldw_in : nullptr;
        double *out = a.v + b * a.ldv;
        const long long rows = (a.n_out - col + a.d - 1) / a.d;
        const long long o_start = chunk * a.rows_per_chunk;
        const long long o_end = o_start + a.rows_per_chunk < rows ? o_start + a.rows_per_chunk : rows;
        if (o_start >= o_end) continue;
        const long long pin0 = a.t0 + col + o_start * a.d;  // position of input row o_start (before stream offsets)
        const long long lo_off = a.off_h < a.off_g ? a.off_h : a.off_g, hi_off = a.off_h < a.off_g ? a.off_g : a.off_h;
        const long long last_pos = pin0 + (o_end - o_start + L - 2) * a.d;
        if (pin0 + lo_off >= 0 && last_pos + hi_off < a.n_in) col_synth_chunk<L, false>(a, v, w, out, col, o_start, o_end, pin0);
        else col_synth_chunk<L, true>(a, v, w, out, col, o_start, o_end, pin0);
    }
}

#define VW_DISPATCH_CL(L, CALL)           \
    switch (L) {                          \
        case 2: CALL(2); break;           \
        case 4: CALL(4); break;           \
        case 6: CALL(6); break;           \
        case 8: CALL(8); break;           \
        case 10: CALL(10); break;         \
        case 12: CALL(12); break;         \
        case 16: CALL(16); break;         \
        case 18: CALL(18); break;         \
        case 20: CALL(20); break;         \
        case 30: CALL(30); break;         \
        default: return VW_EUNSUPPORTED;  \
    }

int geometry(const vw_ctx *ctx, int64_t n_out, int64_t d, int64_t batch, dim3 &grid, int &rows_per_chunk, int &chunks_out) {
    if (d < 32) return VW_EUNSUPPORTED;
    const int64_t rows = (n_out + d - 1) / d;
    // enough (chunk, column) threads to fill the machine several times over, but chunks long enough to amortise the
    // L-1 warm-up rows each chunk re-reads
    const int64_t want_threads = (int64_t)ctx->sm_count * 512 * 6;
    int64_t chunks = (want_threads + d * batch - 1) / (d * batch);
    if (chunks < 1) chunks = 1;
    int64_t rpc = (rows + chunks - 1) / chunks;
    if (rpc < 32 * kCR) rpc = 32 * kCR;
    rpc = ((rpc + kCR - 1) / kCR) * kCR;
    chunks = (rows + rpc - 1) / rpc;
    const int64_t blocks = (chunks * d + kCThreads - 1) / kCThreads;
    if (blocks > 0x7fffffffll) return VW_EUNSUPPORTED;
    rows_per_chunk = (int)rpc;
    chunks_out = (int)chunks;
    grid = dim3((unsigned)blocks, (unsigned)(batch < 65535 ? batch : 65535));
    return VW_OK;
}

}  // namespace

int vw_column_analysis(vw_ctx *ctx, const double *x, int64_t ldx, double *v, int64_t ldv, double *w, int64_t ldw,
                       int64_t n_in, int64_t t0, int64_t n_out, int64_t batch, const VwFilt &f, int l, int64_t d, int mode) {
    if (l < 2 || l > VW_FUSED_MAX_L || n_out < 1 || batch < 1) return VW_EUNSUPPORTED;
    ColArgs a;
    dim3 grid;
    int rc = geometry(ctx, n_out, d, batch, grid, a.rows_per_chunk, a.chunks);
    if (rc) return rc;
    a.x = x; a.ldx = ldx; a.w_in = nullptr; a.ldw_in = 0; a.v = v; a.ldv = ldv; a.w = w; a.ldw = ldw;
    a.n_in = n_in; a.t0 = t0; a.n_out = n_out; a.batch = batch; a.d = d; a.off_h = a.off_g = 0;
    a.mode = mode;
    for (int k = 0; k < VW_FUSED_MAX_L; k++) { a.f.h[k] = k < l ? f.h[k] : 0.0; a.f.g[k] = k < l ? f.g[k] : 0.0; }
#define VW_CA(LL) k_column_analysis<LL><<<grid, kCThreads, 0, ctx->stream>>>(a)
    VW_DISPATCH_CL(l, VW_CA)
#undef VW_CA
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "column analysis launch");
}

int vw_column_synthesis(vw_ctx *ctx, const double *v, int64_t ldv, const double *w, int64_t ldw, double *out, int64_t ldo,
                        int64_t n_in, int64_t t0, int64_t n_out, int64_t batch, const VwFilt &f, int l, int64_t d, int mode,
                        vw_align al) {
    if (l < 2 || l > VW_FUSED_MAX_L || n_out < 1 || batch < 1) return VW_EUNSUPPORTED;
    ColArgs a;
    dim3 grid;
    int rc = geometry(ctx, n_out, d, batch, grid, a.rows_per_chunk, a.chunks);
    if (rc) return rc;
    a.x = v; a.ldx = ldv; a.w_in = w; a.ldw_in = ldw; a.v = out; a.ldv = ldo; a.w = nullptr; a.ldw = 0;
    a.n_in = n_in; a.t0 = t0; a.n_out = n_out; a.batch = batch; a.d = d;
    a.mode = mode;
    // sigma=+1: sum_k f[k] S[p - tau + k d];  sigma=-1: sum_k f[k] S[p + tau - k d] = sum_k' f[L-1-k'] S[p + tau - (L-1)d + k' d]
    a.off_h = al.sigma_h > 0 ? -(int64_t)al.tau_h : (int64_t)al.tau_h - (int64_t)(l - 1) * d;
    a.off_g = al.sigma_g > 0 ? -(int64_t)al.tau_g : (int64_t)al.tau_g - (int64_t)(l - 1) * d;
    for (int k = 0; k < VW_FUSED_MAX_L; k++) {
        a.f.h[k] = k < l ? (al.sigma_h > 0 ? f.h[k] : f.h[l - 1 - k]) : 0.0;
        a.f.g[k] = k < l ? (al.sigma_g > 0 ? f.g[k] : f.g[l - 1 - k]) : 0.0;
    }
#define VW_CS(LL) k_column_synthesis<LL><<<grid, kCThreads, 0, ctx->stream>>>(a)
    VW_DISPATCH_CL(l, VW_CS)
#undef VW_CS
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "column synthesis launch");
}
