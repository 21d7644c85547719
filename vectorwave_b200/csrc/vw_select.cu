// vw_select.cu -- exact per-signal median of |W_1| for the universal threshold.
//
// Replaces VectorWaveSwtAdapter.estimateNoiseSigma's full Arrays.sort
// (CORE/swt/VectorWaveSwtAdapter.java:627-645) with an MSB-first radix select on the IEEE bit
// pattern of |x| (order-isomorphic to the value for finite non-negative doubles).  Both middle order
// statistics of an even-length row are tracked in the same passes.  Result per row:
//   thr = (median / 0.6745) * sqrt(2 ln n)        (CORE/swt/VectorWaveSwtAdapter.java:505-520)
#include "vw_internal.cuh"

namespace {

struct SelState {           // per row
    unsigned long long prefix[2];  // key bits decided so far (high bits), for rank lo / hi
    unsigned long long rank[2];    // remaining 0-based rank inside the current prefix bucket
};

__global__ void k_select_init(SelState *st, int64_t batch, int64_t n) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    st[b].prefix[0] = st[b].prefix[1] = 0ull;
    st[b].rank[1] = (unsigned long long)(n / 2);
    st[b].rank[0] = (n % 2 == 0) ? (unsigned long long)(n / 2 - 1) : (unsigned long long)(n / 2);
}

// histogram of the next 8-bit digit among keys that match each tracked prefix
__global__ void __launch_bounds__(256)
k_select_hist(const double *__restrict__ w, int64_t n, int64_t ld, const SelState *__restrict__ st, int shift,
              unsigned int *__restrict__ hist /*[batch][2][256]*/) {
    __shared__ unsigned int sh[2][256];
    int64_t b = blockIdx.y;
    sh[0][threadIdx.x] = 0; sh[1][threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long p0 = st[b].prefix[0], p1 = st[b].prefix[1];
    const bool same = p0 == p1;
    const double *row = w + b * ld;
    const int hs = shift + 8;  // bits above the current digit
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long key = (unsigned long long)__double_as_longlong(fabs(row[t]));
        unsigned long long hi = hs >= 64 ? 0ull : (key >> hs);
        unsigned int dg = (unsigned int)(key >> shift) & 255u;
        if (hi == p0) atomicAdd(&sh[0][dg], 1u);
        if (!same && hi == p1) atomicAdd(&sh[1][dg], 1u);
    }
    __syncthreads();
    unsigned int c0 = sh[0][threadIdx.x], c1 = sh[1][threadIdx.x];
    if (c0) atomicAdd(&hist[(b * 2 + 0) * 256 + threadIdx.x], c0);
    if (same) { if (c0) atomicAdd(&hist[(b * 2 + 1) * 256 + threadIdx.x], c0); }
    else if (c1) atomicAdd(&hist[(b * 2 + 1) * 256 + threadIdx.x], c1);
}

// one thread per (row, which): walk the 256 bins, pick the digit holding the rank, clear the bins
__global__ void k_select_pick(SelState *st, unsigned int *hist, int64_t batch) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * 2) return;
    int64_t b = i >> 1; int which = (int)(i & 1);
    unsigned int *h = hist + i * 256;
    unsigned long long r = st[b].rank[which], cum = 0; int dg = 255;
    for (int q = 0; q < 256; q++) {
        unsigned long long c = h[q];
        if (r < cum + c) { dg = q; break; }
        cum += c;
    }
    for (int q = 0; q < 256; q++) h[q] = 0;
    st[b].rank[which] = r - cum;
    st[b].prefix[which] = (st[b].prefix[which] << 8) | (unsigned long long)dg;
}

__global__ void k_select_finish(const SelState *st, int64_t batch, int64_t n, double factor, double *thr) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double lo = __longlong_as_double((long long)st[b].prefix[0]);
    double hi = __longlong_as_double((long long)st[b].prefix[1]);
    double med = (n % 2 == 0) ? __ddiv_rn(__dadd_rn(lo, hi), 2.0) : hi;
    thr[b] = __dmul_rn(__ddiv_rn(med, 0.6745), factor);
}

}  // namespace

int vw_launch_universal_threshold(vw_ctx *ctx, const double *w1, int64_t batch, int64_t n, int64_t ld,
                                  double *thr_dev) {
    if (batch <= 0 || n <= 0) return VW_OK;
    void *ws = nullptr;
    size_t st_bytes = (size_t)batch * sizeof(SelState);
    size_t hist_bytes = (size_t)batch * 2 * 256 * sizeof(unsigned int);
    int rc = vw_scratch(ctx, 4, st_bytes + hist_bytes, &ws);
    if (rc) return rc;
    SelState *st = (SelState *)ws;
    unsigned int *hist = (unsigned int *)((char *)ws + st_bytes);
    rc = vw_cuda_check(ctx, cudaMemsetAsync(hist, 0, hist_bytes, ctx->stream), "select memset");
    if (rc) return rc;
    unsigned nb = (unsigned)((batch + 127) / 128);
    k_select_init<<<nb, 128, 0, ctx->stream>>>(st, batch, n);
    ctx->launches++;
    int64_t chunks = (n + 256 * 16 - 1) / (256 * 16);
    int64_t cap = ((int64_t)ctx->sm_count * 16 + batch - 1) / batch;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    for (int shift = 56; shift >= 0; shift -= 8) {
        k_select_hist<<<dim3((unsigned)chunks, (unsigned)batch), 256, 0, ctx->stream>>>(w1, n, ld, st, shift, hist);
        k_select_pick<<<(unsigned)((batch * 2 + 127) / 128), 128, 0, ctx->stream>>>(st, hist, batch);
        ctx->launches += 2;
    }
    double factor = sqrt(2.0 * log((double)n));  // Math.sqrt(2 * Math.log(n)), CORE/swt/VectorWaveSwtAdapter.java:512
    k_select_finish<<<nb, 128, 0, ctx->stream>>>(st, batch, n, factor, thr_dev);
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "universal threshold launch");
}
