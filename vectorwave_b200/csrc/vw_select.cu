// vw_select.cu -- exact per-signal median of |W_1| for the universal threshold.
//
// Replaces VectorWaveSwtAdapter.estimateNoiseSigma's full Arrays.sort
// (CORE/swt/VectorWaveSwtAdapter.java:627-645) with an MSB-first radix select on the IEEE bit
// pattern of |x| (order-isomorphic to the value for finite non-negative doubles).  Both middle order
// statistics of an even-length row are tracked in the same passes.  Result per row:
//   thr = (median / 0.6745) * sqrt(2 ln n)        (CORE/swt/VectorWaveSwtAdapter.java:505-520)
#include "vw_internal.cuh"

namespace {

// Passes over the data: two 12-bit digits (exponent, then the top 12 mantissa bits) on the full row, then one pass
// that compacts the keys sharing the 24 decided bits (a few dozen per million N(0,1) samples) into a candidate
// buffer; the remaining 40 bits are resolved by five 8-bit passes over the candidates only.  24 B/sample instead of
// the 64 B/sample of eight full passes.  A row whose bucket overflows the buffer (heavy ties, e.g. a constant
// signal) keeps running the 8-bit passes over the full row -- still exact, just slower.
struct SelState {           // per row
    unsigned long long prefix[2];  // key bits decided so far (high bits), for rank lo / hi
    unsigned long long rank[2];    // remaining 0-based rank inside the current prefix bucket
    unsigned int count[2];         // candidates collected for lo / hi
    unsigned int pad[2];
};

constexpr int kBig = 12;                 // bits per digit of the two full-row passes
constexpr int kBigBins = 1 << kBig;

__global__ void k_select_init(SelState *st, int64_t batch, int64_t n) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    st[b].prefix[0] = st[b].prefix[1] = 0ull;
    st[b].rank[1] = (unsigned long long)(n / 2);
    st[b].rank[0] = (n % 2 == 0) ? (unsigned long long)(n / 2 - 1) : (unsigned long long)(n / 2);
    st[b].count[0] = st[b].count[1] = 0u;
}

__device__ __forceinline__ unsigned long long key_of(double v) { return (unsigned long long)__double_as_longlong(fabs(v)); }

// histogram of the next BITS-bit digit among keys that match each tracked prefix.
// cand != nullptr: rows whose candidate buffers did not overflow read their candidates instead of the full row.
template <int BITS>
__global__ void __launch_bounds__(256)
k_select_hist(const double *__restrict__ w, int64_t n, int64_t ld, const SelState *__restrict__ st, int shift,
              unsigned int *__restrict__ hist /*[batch][2][1 << BITS]*/, const double *__restrict__ cand, unsigned int cap) {
    constexpr int BINS = 1 << BITS;
    __shared__ unsigned int sh[2][BINS];
    const int64_t b = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const unsigned long long p0 = st[b].prefix[0], p1 = st[b].prefix[1];
    const bool same = p0 == p1;
    const int hs = shift + BITS;  // bits above the current digit
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool use_cand = cand != nullptr && st[b].count[0] <= cap && st[b].count[1] <= cap;
    if (!use_cand) {
        const double *row = w + b * ld;
        for (int64_t t = t0; t < n; t += stride) {
            const unsigned long long key = key_of(row[t]);
            const unsigned long long hi = hs >= 64 ? 0ull : (key >> hs);
            const unsigned int dg = (unsigned int)(key >> shift) & (BINS - 1);
            if (hi == p0) atomicAdd(&sh[0][dg], 1u);
            if (!same && hi == p1) atomicAdd(&sh[1][dg], 1u);
        }
    } else {
        // candidates of rank lo and of rank hi sit in separate buffers (identical content when the prefixes agree)
        for (int which = 0; which < (same ? 1 : 2); which++) {
            const double *src = cand + ((size_t)b * 2 + which) * cap;
            const unsigned long long pw = which ? p1 : p0;
            const int64_t len = st[b].count[which];
            for (int64_t t = t0; t < len; t += stride) {
                const unsigned long long key = key_of(src[t]);
                const unsigned long long hi = hs >= 64 ? 0ull : (key >> hs);
                if (hi == pw) atomicAdd(&sh[which][(unsigned int)(key >> shift) & (BINS - 1)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BINS; i += blockDim.x) {
        const unsigned int c0 = sh[0][i], c1 = same ? c0 : sh[1][i];
        if (c0) atomicAdd(&hist[((size_t)b * 2 + 0) * BINS + i], c0);
        if (c1) atomicAdd(&hist[((size_t)b * 2 + 1) * BINS + i], c1);
    }
}

// one block per (row, which): find the bin holding the rank (parallel prefix over the bins), clear the bins
template <int BITS>
__global__ void __launch_bounds__(256) k_select_pick(SelState *st, unsigned int *hist) {
    constexpr int BINS = 1 << BITS, PER = BINS / 256 > 0 ? BINS / 256 : 1;
    __shared__ unsigned long long part[256];
    const int64_t i = blockIdx.x;
    const int64_t b = i >> 1; const int which = (int)(i & 1);
    unsigned int *h = hist + (size_t)i * BINS;
    const int tid = threadIdx.x;
    unsigned long long mine = 0;
    unsigned int loc[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) { const int bin = tid * PER + q; loc[q] = bin < BINS ? h[bin] : 0u; mine += loc[q]; }
    part[tid] = mine;
    __syncthreads();
    if (tid == 0) {   // 256-entry exclusive scan: tiny
        unsigned long long run = 0;
        for (int q = 0; q < 256; q++) { const unsigned long long c = part[q]; part[q] = run; run += c; }
    }
    __syncthreads();
    const unsigned long long r = st[b].rank[which];
    unsigned long long cum = part[tid];
    __syncthreads();
    if (r >= cum && r < cum + mine) {
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (r < cum + loc[q]) {
                st[b].rank[which] = r - cum;
                st[b].prefix[which] = (st[b].prefix[which] << BITS) | (unsigned long long)(tid * PER + q);
                break;
            }
            cum += loc[q];
        }
    }
#pragma unroll
    for (int q = 0; q < PER; q++) { const int bin = tid * PER + q; if (bin < BINS) h[bin] = 0; }
}

// compact the keys that share the decided high bits into the per-(row, which) candidate buffers
__global__ void __launch_bounds__(256)
k_select_collect(const double *__restrict__ w, int64_t n, int64_t ld, SelState *st, int decided_bits, double *cand,
                 unsigned int cap) {
    const int64_t b = blockIdx.y;
    const unsigned long long p0 = st[b].prefix[0], p1 = st[b].prefix[1];
    const bool same = p0 == p1;
    const int hs = 64 - decided_bits;
    const double *row = w + b * ld;
    double *c0 = cand + ((size_t)b * 2 + 0) * cap, *c1 = cand + ((size_t)b * 2 + 1) * cap;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double v = fabs(row[t]);
        const unsigned long long hi = key_of(v) >> hs;
        if (hi == p0) {   // one bucket for both ranks: fill both buffers, the prefixes may part ways in a later digit
            const unsigned int k = atomicAdd(&st[b].count[0], 1u);
            if (k < cap) { c0[k] = v; if (same) c1[k] = v; }
        }
        if (!same && hi == p1) { const unsigned int k = atomicAdd(&st[b].count[1], 1u); if (k < cap) c1[k] = v; }
    }
}
__global__ void k_select_mirror_count(SelState *st, int64_t batch) {   // same prefix: hi shares lo's candidates
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (st[b].prefix[0] == st[b].prefix[1]) st[b].count[1] = st[b].count[0];
}

// factor >= 0: universal threshold (median / 0.6745) * factor (0 for n = 1); factor < 0: the raw median
__global__ void k_select_finish(const SelState *st, int64_t batch, int64_t n, double factor, double *thr) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double lo = __longlong_as_double((long long)st[b].prefix[0]);
    double hi = __longlong_as_double((long long)st[b].prefix[1]);
    double med = (n % 2 == 0) ? __ddiv_rn(__dadd_rn(lo, hi), 2.0) : hi;
    thr[b] = factor >= 0.0 ? __dmul_rn(__ddiv_rn(med, 0.6745), factor) : med;
}

}  // namespace

static int select_median(vw_ctx *ctx, const double *w1, int64_t batch, int64_t n, int64_t ld, double *thr_dev, double factor);

int vw_launch_universal_threshold(vw_ctx *ctx, const double *w1, int64_t batch, int64_t n, int64_t ld,
                                  double *thr_dev) {
    // Math.sqrt(2 * Math.log(n)), CORE/swt/VectorWaveSwtAdapter.java:512
    return select_median(ctx, w1, batch, n, ld, thr_dev, sqrt(2.0 * log((double)n)));
}

int vw_launch_median_abs(vw_ctx *ctx, const double *c, int64_t batch, int64_t n, int64_t ld, double *med_dev) {
    return select_median(ctx, c, batch, n, ld, med_dev, -1.0);
}

static int select_median(vw_ctx *ctx, const double *w1, int64_t batch, int64_t n, int64_t ld, double *thr_dev, double factor) {
    if (batch <= 0 || n <= 0) return VW_OK;
    void *ws = nullptr;
    int64_t cap64 = n / 64;
    if (cap64 < 1024) cap64 = 1024;
    if (cap64 > (1 << 20)) cap64 = 1 << 20;
    const unsigned int cap = (unsigned int)cap64;
    const size_t st_bytes = ((size_t)batch * sizeof(SelState) + 255) & ~(size_t)255;
    const size_t hist_bytes = (size_t)batch * 2 * kBigBins * sizeof(unsigned int);
    const size_t cand_bytes = (size_t)batch * 2 * cap * sizeof(double);
    int rc = vw_scratch(ctx, 4, st_bytes + hist_bytes + cand_bytes, &ws);
    if (rc) return rc;
    SelState *st = (SelState *)ws;
    unsigned int *hist = (unsigned int *)((char *)ws + st_bytes);
    double *cand = (double *)((char *)ws + st_bytes + hist_bytes);
    rc = vw_cuda_check(ctx, cudaMemsetAsync(hist, 0, hist_bytes, ctx->stream), "select memset");
    if (rc) return rc;
    const unsigned nb = (unsigned)((batch + 127) / 128);
    k_select_init<<<nb, 128, 0, ctx->stream>>>(st, batch, n);
    ctx->launches++;
    int64_t chunks = (n + 256 * 16 - 1) / (256 * 16);
    const int64_t capc = ((int64_t)ctx->sm_count * 16 + batch - 1) / batch;
    if (chunks > capc) chunks = capc;
    if (chunks < 1) chunks = 1;
    const dim3 grid((unsigned)chunks, (unsigned)batch);
    // two 12-bit digits over the full rows: bits 63..52 (sign 0 + exponent), bits 51..40
    for (int shift = 64 - kBig; shift >= 64 - 2 * kBig; shift -= kBig) {
        k_select_hist<kBig><<<grid, 256, 0, ctx->stream>>>(w1, n, ld, st, shift, hist, nullptr, 0u);
        k_select_pick<kBig><<<(unsigned)(batch * 2), 256, 0, ctx->stream>>>(st, hist);
        ctx->launches += 2;
    }
    k_select_collect<<<grid, 256, 0, ctx->stream>>>(w1, n, ld, st, 2 * kBig, cand, cap);
    k_select_mirror_count<<<nb, 128, 0, ctx->stream>>>(st, batch);
    ctx->launches += 2;
    // remaining 40 bits, 8 at a time, over the candidates (or the full row where they overflowed)
    for (int shift = 64 - 2 * kBig - 8; shift >= 0; shift -= 8) {
        k_select_hist<8><<<grid, 256, 0, ctx->stream>>>(w1, n, ld, st, shift, hist, cand, cap);
        k_select_pick<8><<<(unsigned)(batch * 2), 256, 0, ctx->stream>>>(st, hist);
        ctx->launches += 2;
    }
    k_select_finish<<<nb, 128, 0, ctx->stream>>>(st, batch, n, factor, thr_dev);
    ctx->launches++;
    return vw_cuda_check(ctx, cudaGetLastError(), "universal threshold launch");
}

// ---- SURE threshold (CORE/denoising/WaveletDenoiser.java:441-492) -------------------------------------------------
// The reference scores every candidate threshold t = |c_i| with a full pass over the coefficients (an O(n^2) double
// loop) and keeps the first minimum in ascending order of t.  Here one thread owns kSureCand candidates and walks the
// coefficients in the reference's order with the reference's roundings (separate multiplies and adds, no contraction),
// so every risk is bit-identical to the JVM's; (|c|, c*c) pairs are staged through shared memory and broadcast.  The
// sort disappears: "first minimum in ascending t" = the smallest t among the candidates of minimal risk.
namespace {
constexpr int kSureThreads = 256, kSureCand = 4, kSureChunk = 2048;

struct SureBest { double risk, t; };

__device__ __forceinline__ SureBest sure_better(SureBest a, SureBest b) {
    if (a.risk < b.risk) return a;
    if (b.risk < a.risk) return b;
    a.t = a.t < b.t ? a.t : b.t;
    return a;
}

__global__ void __launch_bounds__(kSureThreads) k_sure_scan(const double *__restrict__ c, int64_t batch, int64_t n, int64_t ld,
                                                            const double *__restrict__ sigma, SureBest *partial) {
    __shared__ double2 sm[kSureChunk];
    __shared__ SureBest red[kSureThreads / 32];
    for (int64_t row = blockIdx.y; row < batch; row += gridDim.y) {
        const double *r = c + row * ld;
        const double sigma2 = __dmul_rn(sigma[row], sigma[row]);             // :479
        double t[kSureCand], risk[kSureCand];
        const int64_t i0 = ((int64_t)blockIdx.x * kSureThreads + threadIdx.x) * kSureCand;
#pragma unroll
        for (int q = 0; q < kSureCand; q++) {
            t[q] = i0 + q < n ? fabs(r[i0 + q]) : 0.0;
            risk[q] = __dmul_rn((double)(-(long long)n), sigma2);            // :480  -n * sigma2
        }
        for (int64_t j0 = 0; j0 < n; j0 += kSureChunk) {
            const int m = (int)(n - j0 < kSureChunk ? n - j0 : kSureChunk);
            __syncthreads();
            for (int j = threadIdx.x; j < m; j += kSureThreads) {
                const double x = r[j0 + j];
                sm[j] = make_double2(fabs(x), __dmul_rn(x, x));
            }
            __syncthreads();
#pragma unroll 4
            for (int j = 0; j < m; j++) {
                const double2 p = sm[j];
#pragma unroll
                for (int q = 0; q < kSureCand; q++) {
                    const double d = __dsub_rn(p.x, t[q]);
                    const double over = __dadd_rn(sigma2, __dmul_rn(d, d));  // :487
                    risk[q] = __dadd_rn(risk[q], p.x <= t[q] ? p.y : over);  // :484-488 (NaN compares false: else branch)
                }
            }
        }
        SureBest best{INFINITY, 0.0};                                        // :452-453 minRisk = +inf, bestThreshold = 0
#pragma unroll
        for (int q = 0; q < kSureCand; q++) {
            const double rk = __ddiv_rn(risk[q], (double)n);                 // :491
            if (i0 + q < n && rk < INFINITY) best = sure_better(best, SureBest{rk, t[q]});   // `risk < minRisk` never takes NaN / +inf
        }
        for (int o = 16; o > 0; o >>= 1) {
            SureBest other{__shfl_down_sync(0xffffffffu, best.risk, o), __shfl_down_sync(0xffffffffu, best.t, o)};
            best = sure_better(best, other);
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kSureThreads / 32; w++) best = sure_better(best, red[w]);
            partial[row * gridDim.x + blockIdx.x] = best;
        }
    }
}

__global__ void __launch_bounds__(256) k_sure_finish(const SureBest *partial, int per_row, double *thr_out, double *risk_out) {
    __shared__ SureBest red[8];
    SureBest best{INFINITY, 0.0};
    for (int i = threadIdx.x; i < per_row; i += 256) best = sure_better(best, partial[(int64_t)blockIdx.x * per_row + i]);
    for (int o = 16; o > 0; o >>= 1) {
        SureBest other{__shfl_down_sync(0xffffffffu, best.risk, o), __shfl_down_sync(0xffffffffu, best.t, o)};
        best = sure_better(best, other);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) best = sure_better(best, red[w]);
        thr_out[blockIdx.x] = best.t;
        risk_out[blockIdx.x] = best.risk;
    }
}
}  // namespace

size_t vw_sure_workspace(int64_t batch, int64_t n) {
    const int64_t per_row = (n + kSureThreads * kSureCand - 1) / (kSureThreads * kSureCand);
    return (size_t)batch * (size_t)per_row * sizeof(SureBest);
}

int vw_launch_sure(vw_ctx *ctx, const double *c, int64_t batch, int64_t n, int64_t ld, const double *sigma_dev, void *ws,
                   double *thr_dev, double *risk_dev) {
    if (batch <= 0 || n <= 0) return VW_OK;
    const int64_t per_row = (n + kSureThreads * kSureCand - 1) / (kSureThreads * kSureCand);
    const dim3 grid((unsigned)per_row, (unsigned)(batch < 65535 ? batch : 65535));
    k_sure_scan<<<grid, kSureThreads, 0, ctx->stream>>>(c, batch, n, ld, sigma_dev, (SureBest *)ws);
    k_sure_finish<<<(unsigned)batch, 256, 0, ctx->stream>>>((const SureBest *)ws, (int)per_row, thr_dev, risk_dev);
    ctx->launches += 2;
    return vw_cuda_check(ctx, cudaGetLastError(), "sure launch");
}
