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

constexpr int kCR = 9;         // rows per register block
constexpr int kCThreads = 256;

__device__ __forceinline__ int64_t wrap_mod(int64_t i, int64_t n) { i %= n; return i < 0 ? i + n : i; }
__device__ __forceinline__ double ext_load(const double *__restrict__ row, int64_t pos, int64_t n, int mode) {
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
__global__ void __launch_bounds__(kCThreads, (L <= 8) ? 2 : 1) k_column_analysis(const __grid_constant__ ColArgs a) {
    const long long col = (long long)blockIdx.x * kCThreads + threadIdx.x;  // phase phi in [0, d)
    if (col >= a.d) return;
    const int chunk = blockIdx.y;
    for (long long b = blockIdx.z; b < a.batch; b += gridDim.z) {
        const double *x = a.x + b * a.ldx;
        double *vo = a.v + b * a.ldv, *wo = a.w + b * a.ldw;
        // rows of this column inside the output range: positions t0 + col + q*d < t0 + n_out
        const long long rows = (a.n_out - col + a.d - 1) / a.d;
        long long q = (long long)chunk * a.rows_per_chunk;
        const long long qend = q + a.rows_per_chunk < rows ? q + a.rows_per_chunk : rows;
        if (q >= qend) continue;
        long long p = a.t0 + col + q * a.d;                // input-coordinate position of the current row
        double win[L - 1 > 0 ? L - 1 : 1];                 // win[i] = X[p - (L-1-i) d]  (oldest first)
#pragma unroll
        for (int i = 0; i < L - 1; i++) win[i] = ext_load(x, p - (long long)(L - 1 - i) * a.d, a.n_in, a.mode);
        for (; q < qend; q += kCR) {
            double nw[kCR];
#pragma unroll
            for (int r = 0; r < kCR; r++) {
                const long long pr = p + r * a.d;
                nw[r] = (q + r < qend) ? __ldg(x + pr) : 0.0;   // output positions are always inside [0, n_in)
            }
            double ah[kCR], ag[kCR];
#pragma unroll
            for (int r = 0; r < kCR; r++) { ah[r] = 0.0; ag[r] = 0.0; }
            // sequence s[0..L-2] = win, s[L-1+r] = nw[r]; out[r] = sum_k f[k] s[L-1+r-k]; walk s downwards so every
            // output sees its taps in ascending order
#pragma unroll
            for (int m = kCR - 1; m >= -(L - 1); m--) {
                const double xv = m >= 0 ? nw[m] : win[L - 1 + m];
#pragma unroll
                for (int r = 0; r < kCR; r++) {
                    const int k = r - m;
                    if (k >= 0 && k < L) { ah[r] = fma(a.f.h[k], xv, ah[r]); ag[r] = fma(a.f.g[k], xv, ag[r]); }
                }
            }
#pragma unroll
            for (int r = 0; r < kCR; r++) {
                if (q + r < qend) {
                    const long long o = p - a.t0 + r * a.d;
                    vo[o] = ah[r];
                    wo[o] = ag[r];
                }
            }
            // slide: keep the last L-1 samples of the sequence
#pragma unroll
            for (int i = 0; i < L - 1; i++) {
                const int src = i + kCR;                   // index into the L-1+R sequence
                win[i] = src < L - 1 ? win[src] : nw[src - (L - 1)];
            }
            p += (long long)kCR * a.d;
        }
    }
}

// ---- synthesis -----------------------------------------------------------------------------------------------
// out[p] = sum_k th[k] V[ext(p + off_h + k d)] + sum_k tg[k] W[ext(p + off_g + k d)]
// with (taps, off) = (f, -tau) for sigma=+1 and (reversed f, tau - (L-1) d) for sigma=-1.
template <int L>
__device__ __forceinline__ void col_stream(const double *__restrict__ src, long long p0, long long d, long long n, int mode,
                                           long long qcount, const double (&taps)[VW_FUSED_MAX_L], double (&win)[L > 1 ? L - 1 : 1],
                                           double (&acc)[kCR], bool first) {
    // window convention: win[i] = S[p0 + i d] for i in [0, L-1); new samples S[p0 + (L-1+r) d]
    if (first) {
#pragma unroll
        for (int i = 0; i < L - 1; i++) win[i] = ext_load(src, p0 + (long long)i * d, n, mode);
    }
    double nw[kCR];
#pragma unroll
    for (int r = 0; r < kCR; r++) nw[r] = (r < qcount) ? ext_load(src, p0 + (long long)(L - 1 + r) * d, n, mode) : 0.0;
#pragma unroll
    for (int m = 0; m <= kCR + L - 2; m++) {
        const double xv = m < L - 1 ? win[m] : nw[m - (L - 1)];
#pragma unroll
        for (int r = 0; r < kCR; r++) {
            const int k = m - r;
            if (k >= 0 && k < L) acc[r] = fma(taps[k], xv, acc[r]);
        }
    }
#pragma unroll
    for (int i = 0; i < L - 1; i++) {
        const int s = i + kCR;
        win[i] = s < L - 1 ? win[s] : nw[s - (L - 1)];
    }
}

template <int L>
__global__ void __launch_bounds__(kCThreads, (L <= 8) ? 2 : 1) k_column_synthesis(const __grid_constant__ ColArgs a) {
    const long long col = (long long)blockIdx.x * kCThreads + threadIdx.x;
    if (col >= a.d) return;
    const int chunk = blockIdx.y;
    for (long long b = blockIdx.z; b < a.batch; b += gridDim.z) {
        const double *v = a.x ? a.x + b * a.ldx : nullptr;
        const double *w = a.w_in ? a.w_in + b * a.ldw_in : nullptr;
        double *out = a.v + b * a.ldv;
        const long long rows = (a.n_out - col + a.d - 1) / a.d;
        long long q = (long long)chunk * a.rows_per_chunk;
        const long long qend = q + a.rows_per_chunk < rows ? q + a.rows_per_chunk : rows;
        if (q >= qend) continue;
        long long p = a.t0 + col + q * a.d;
        double winv[L > 1 ? L - 1 : 1], winw[L > 1 ? L - 1 : 1];
        bool first = true;
        for (; q < qend; q += kCR) {
            double acc[kCR];
#pragma unroll
            for (int r = 0; r < kCR; r++) acc[r] = 0.0;
            const long long cnt = qend - q;
            if (v) col_stream<L>(v, p + a.off_h, a.d, a.n_in, a.mode, cnt, a.f.h, winv, acc, first);
            if (w) col_stream<L>(w, p + a.off_g, a.d, a.n_in, a.mode, cnt, a.f.g, winw, acc, first);
            first = false;
#pragma unroll
            for (int r = 0; r < kCR; r++)
                if (q + r < qend) out[p - a.t0 + r * a.d] = acc[r];
            p += (long long)kCR * a.d;
        }
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

int geometry(const vw_ctx *ctx, int64_t n_out, int64_t d, int64_t batch, dim3 &grid, int &rows_per_chunk) {
    if (d < 32) return VW_EUNSUPPORTED;
    const int64_t rows = (n_out + d - 1) / d;
    // enough chunks to fill the machine a few times over, but chunks long enough to amortise the L-1 warm-up rows
    const int64_t col_blocks = (d + kCThreads - 1) / kCThreads;
    int64_t want = (int64_t)ctx->sm_count * 16;
    int64_t chunks = (want + col_blocks * batch - 1) / (col_blocks * batch);
    int64_t rpc = (rows + chunks - 1) / chunks;
    if (rpc < 8 * kCR) rpc = 8 * kCR;
    rpc = ((rpc + kCR - 1) / kCR) * kCR;
    chunks = (rows + rpc - 1) / rpc;
    if (chunks > 65535) { rpc = ((rows + 65534) / 65535 + kCR - 1) / kCR * kCR; chunks = (rows + rpc - 1) / rpc; }
    rows_per_chunk = (int)rpc;
    grid = dim3((unsigned)col_blocks, (unsigned)chunks, (unsigned)(batch < 65535 ? batch : 65535));
    return VW_OK;
}

}  // namespace

int vw_column_analysis(vw_ctx *ctx, const double *x, int64_t ldx, double *v, int64_t ldv, double *w, int64_t ldw,
                       int64_t n_in, int64_t t0, int64_t n_out, int64_t batch, const VwFilt &f, int l, int64_t d, int mode) {
    if (l < 2 || l > VW_FUSED_MAX_L || n_out < 1 || batch < 1) return VW_EUNSUPPORTED;
    ColArgs a;
    dim3 grid;
    int rc = geometry(ctx, n_out, d, batch, grid, a.rows_per_chunk);
    if (rc) return rc;
    a.x = x; a.ldx = ldx; a.w_in = nullptr; a.ldw_in = 0; a.v = v; a.ldv = ldv; a.w = w; a.ldw = ldw;
    a.n_in = n_in; a.t0 = t0; a.n_out = n_out; a.batch = batch; a.d = d; a.off_h = a.off_g = 0;
    a.chunks = (int)grid.y; a.mode = mode;
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
    int rc = geometry(ctx, n_out, d, batch, grid, a.rows_per_chunk);
    if (rc) return rc;
    a.x = v; a.ldx = ldv; a.w_in = w; a.ldw_in = ldw; a.v = out; a.ldv = ldo; a.w = nullptr; a.ldw = 0;
    a.n_in = n_in; a.t0 = t0; a.n_out = n_out; a.batch = batch; a.d = d;
    a.chunks = (int)grid.y; a.mode = mode;
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
