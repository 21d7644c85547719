// vw_generic.cu -- per-level MODWT kernels that accept every shape the reference API accepts.
//
// One thread per output sample, taps read from the kernel-parameter constant bank, inputs through
// L1/L2.  These are (1) the path for shapes the fused tile kernels do not take (odd n, n shorter than
// the fused halo, l > 32, L_j > n single-level multi-wrap), (2) the VW_FLAG_BITEXACT path: with
// EXACT the products and sums are rounded separately in the reference's tap order, so results are
// bit-identical to the JVM loops they replace:
//   analysis   CORE/internal/ScalarOps.java:700-723 (periodic), :790-808 (zero), :818-835 (symmetric)
//   synthesis  CORE/modwt/MultiLevelMODWTTransform.java:554-645, CORE/modwt/MODWTTransform.java:244-296,672-684
//   mirror     CORE/util/MathUtils.java:30-51
//   threshold  CORE/modwt/MutableMultiLevelMODWTResult.java:97-118
#include "vw_internal.cuh"

namespace {

__device__ __forceinline__ int64_t ext_periodic(int64_t i, int64_t n) {
    if (i < 0) { i += n; if (i < 0) { i %= n; if (i < 0) i += n; } }
    else if (i >= n) { i -= n; if (i >= n) i %= n; }
    return i;
}
__device__ __forceinline__ int64_t ext_mirror(int64_t i, int64_t n) {
    if (i >= 0 && i < n) return i;
    int64_t p = 2 * n;
    i %= p; if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}
// returns false when the term is skipped (zero padding / linear span)
__device__ __forceinline__ bool ext_index(int64_t &i, int64_t n, int mode) {
    if (mode == VW_PERIODIC) { i = ext_periodic(i, n); return true; }
    if (mode == VW_SYMMETRIC) { i = ext_mirror(i, n); return true; }
    return i >= 0 && i < n;
}

template <bool EXACT>
__device__ __forceinline__ double mac(double acc, double a, double b) {
    if (EXACT) return __dadd_rn(acc, __dmul_rn(a, b));
    return fma(a, b, acc);
}

template <bool EXACT>
__global__ void __launch_bounds__(256)
k_analysis_level(const double *__restrict__ in, int64_t ld_in, double *__restrict__ v_out, int64_t ld_v,
                 double *__restrict__ w_out, int64_t ld_w, int64_t n_in, int64_t t0, int64_t n_out, int64_t batch,
                 const __grid_constant__ VwFilt f, int l, int64_t d, int mode) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
        const double *x = in + b * ld_in;
        double a = 0.0, w = 0.0;
        int64_t tt = t + t0;
        for (int k = 0; k < l; k++) {
            int64_t i = tt - (int64_t)k * d;
            if (!ext_index(i, n_in, mode)) continue;
            double xv = __ldg(x + i);
            a = mac<EXACT>(a, xv, f.h[k]);
            w = mac<EXACT>(w, xv, f.g[k]);
        }
        if (v_out) v_out[b * ld_v + t] = a;
        if (w_out) w_out[b * ld_w + t] = w;
    }
}

template <bool EXACT, bool PAIR>
__global__ void __launch_bounds__(256)
k_synthesis_level(const double *__restrict__ v, int64_t ld_v, const double *__restrict__ w, int64_t ld_w,
                  double *__restrict__ out, int64_t ld_o, int64_t n_in, int64_t t0, int64_t n_out, int64_t batch,
                  const __grid_constant__ VwFilt f, int l, int64_t d, int mode, vw_align al) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
        const double *vv = v ? v + b * ld_v : nullptr;
        const double *ww = w ? w + b * ld_w : nullptr;
        int64_t tt = t + t0;
        double acc = 0.0;
        if (PAIR) {
            for (int k = 0; k < l; k++) {
                int64_t i = tt + (int64_t)al.sigma_h * ((int64_t)k * d - al.tau_h);
                if (!ext_index(i, n_in, mode)) continue;
                double a = vv ? __ldg(vv + i) : 0.0, c = ww ? __ldg(ww + i) : 0.0;
                if (EXACT) acc = __dadd_rn(acc, __dadd_rn(__dmul_rn(f.h[k], a), __dmul_rn(f.g[k], c)));
                else acc = fma(f.h[k], a, fma(f.g[k], c, acc));
            }
        } else {
            if (vv)
                for (int k = 0; k < l; k++) {
                    int64_t i = tt + (int64_t)al.sigma_h * ((int64_t)k * d - al.tau_h);
                    if (!ext_index(i, n_in, mode)) continue;
                    acc = mac<EXACT>(acc, f.h[k], __ldg(vv + i));
                }
            if (ww)
                for (int k = 0; k < l; k++) {
                    int64_t i = tt + (int64_t)al.sigma_g * ((int64_t)k * d - al.tau_g);
                    if (!ext_index(i, n_in, mode)) continue;
                    acc = mac<EXACT>(acc, f.g[k], __ldg(ww + i));
                }
        }
        out[b * ld_o + t] = acc;
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(256)
k_conv_dense(const double *__restrict__ x, int64_t n, const double *__restrict__ filt, int64_t lf, int mode,
             double *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double acc = 0.0;
    for (int64_t q = 0; q < lf; q++) {
        int64_t i = t - q;
        if (!ext_index(i, n, mode)) continue;
        acc = mac<EXACT>(acc, __ldg(x + i), __ldg(filt + q));
    }
    out[t] = acc;
}

__global__ void __launch_bounds__(256)
k_threshold(double *__restrict__ c, int64_t batch, int64_t n, int64_t ld, const double *__restrict__ thr,
            int per_row, int soft) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
        double lam = thr[per_row ? b : 0];
        const double r = vw_threshold_value(c[b * ld + t], lam, soft);
        c[b * ld + t] = r;
    }
}

__global__ void __launch_bounds__(256)
k_nonfinite_count(const double *__restrict__ x, int64_t batch, int64_t n, int64_t ld, unsigned long long *count) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int bad = 0;
    if (t < n)
        for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) bad += !isfinite(x[b * ld + t]);
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, (unsigned long long)bad);
}

// one CTA per row; deterministic: fixed per-thread strides, fixed tree
__global__ void __launch_bounds__(512) k_energy(const double *__restrict__ c, int64_t n, int64_t ld, double *out) {
    __shared__ double part[16];
    const double *row = c + (int64_t)blockIdx.x * ld;
    double s = 0.0;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) { double x = row[t]; s = fma(x, x, s); }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
        for (int o = 8; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[blockIdx.x] = s;
    }
}

// one CTA per row: sum of (c - shift)^POW; deterministic tree.  shift = nullptr: 0
template <int POW>
__global__ void __launch_bounds__(512) k_row_moment(const double *__restrict__ c, int64_t n, int64_t ld, const double *shift,
                                                    double scale, double *out) {
    __shared__ double part[16];
    const double *row = c + (int64_t)blockIdx.x * ld;
    const double m = shift ? shift[blockIdx.x] : 0.0;
    double s = 0.0;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) {
        const double x = row[t] - m;
        s = POW == 1 ? s + x : fma(x, x, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
        for (int o = 8; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[blockIdx.x] = s * scale;
    }
}

dim3 grid2d(int64_t n, int64_t batch) {
    return dim3((unsigned)((n + 255) / 256), (unsigned)(batch < 32768 ? batch : 32768));
}

}  // namespace

int vw_launch_analysis_level(vw_ctx *ctx, const double *in, int64_t ld_in, double *v_out, int64_t ld_v,
                             double *w_out, int64_t ld_w, int64_t n_in, int64_t t0, int64_t n_out, int64_t batch,
                             const VwFilt &f, int l, int64_t d, int mode, bool exact) {
    if (n_out <= 0 || batch <= 0) return VW_OK;
    dim3 g = grid2d(n_out, batch);
    if (exact)
        k_analysis_level<true><<<g, 256, 0, ctx->stream>>>(in, ld_in, v_out, ld_v, w_out, ld_w, n_in, t0, n_out,
                                                           batch, f, l, d, mode);
    else
        k_analysis_level<false><<<g, 256, 0, ctx->stream>>>(in, ld_in, v_out, ld_v, w_out, ld_w, n_in, t0, n_out,
                                                            batch, f, l, d, mode);
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "analysis level launch");
}

int vw_launch_synthesis_level(vw_ctx *ctx, const double *v, int64_t ld_v, const double *w, int64_t ld_w, double *out,
                              int64_t ld_o, int64_t n_in, int64_t t0, int64_t n_out, int64_t batch, const VwFilt &f,
                              int l, int64_t d, int mode, vw_align al, bool pair, bool exact) {
    if (n_out <= 0 || batch <= 0) return VW_OK;
    dim3 g = grid2d(n_out, batch);
#define VW_SYN(E, P)                                                                                              \
    k_synthesis_level<E, P><<<g, 256, 0, ctx->stream>>>(v, ld_v, w, ld_w, out, ld_o, n_in, t0, n_out, batch, f, l, \
                                                        d, mode, al)
    if (exact) { if (pair) VW_SYN(true, true); else VW_SYN(true, false); }
    else { if (pair) VW_SYN(false, true); else VW_SYN(false, false); }
#undef VW_SYN
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "synthesis level launch");
}

int vw_launch_conv_dense(vw_ctx *ctx, const double *x, int64_t n, const double *filter, int64_t lf, int mode,
                         double *out, bool exact) {
    dim3 g((unsigned)((n + 255) / 256));
    if (exact) k_conv_dense<true><<<g, 256, 0, ctx->stream>>>(x, n, filter, lf, mode, out);
    else k_conv_dense<false><<<g, 256, 0, ctx->stream>>>(x, n, filter, lf, mode, out);
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "dense conv launch");
}

int vw_launch_threshold(vw_ctx *ctx, double *c, int64_t batch, int64_t n, int64_t ld, const double *thr_dev,
                        int per_row, int soft) {
    if (n <= 0 || batch <= 0) return VW_OK;
    k_threshold<<<grid2d(n, batch), 256, 0, ctx->stream>>>(c, batch, n, ld, thr_dev, per_row, soft);
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "threshold launch");
}

int vw_launch_nonfinite_count(vw_ctx *ctx, const double *x, int64_t batch, int64_t n, int64_t ld,
                              unsigned long long *count_dev) {
    if (n <= 0 || batch <= 0) return VW_OK;
    k_nonfinite_count<<<grid2d(n, batch), 256, 0, ctx->stream>>>(x, batch, n, ld, count_dev);
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "finite check launch");
}

int vw_launch_mean_variance(vw_ctx *ctx, const double *c, int64_t batch, int64_t n, int64_t ld, double *mean_dev,
                            double *var_dev) {
    if (batch <= 0 || n <= 0) return VW_OK;
    // mean = sum / n, then variance = sum (c - mean)^2 / n: the two loops of calculateBayesThreshold (:525-537)
    k_row_moment<1><<<(unsigned)batch, 512, 0, ctx->stream>>>(c, n, ld, nullptr, 1.0 / (double)n, mean_dev);
    k_row_moment<2><<<(unsigned)batch, 512, 0, ctx->stream>>>(c, n, ld, mean_dev, 1.0 / (double)n, var_dev);
    ctx->launches += 2;
    return vw_cuda_check(ctx, cudaGetLastError(), "mean/variance launch");
}

int vw_launch_energy(vw_ctx *ctx, const double *c, int64_t batch, int64_t n, int64_t ld, double *out_dev) {
    if (batch <= 0) return VW_OK;
    k_energy<<<(unsigned)batch, 512, 0, ctx->stream>>>(c, n, ld, out_dev);
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "energy launch");
}
