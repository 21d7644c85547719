/*
 * modwt_oracle.c -- CPU restatement of VectorWave's MODWT / SWT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vectorwave_b200/ may link, import or
 * call this file; it is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED by the reference's in-tree known answers and formulas
 * (tests/test_oracle_kat.py): Haar {1,2,3,4} -> V={2.5,1.5,2.5,3.5}
 * (CTEST/modwt/MODWTPercivalWaldenValidationTest.java:40-73), Haar on {1..8}
 * (ETEST/modwt/TimeReversedFilterTest.java:19-51), {0.5,-0.5} on {1,2,3,4}
 * (CTEST/modwt/MODWTMathematicalValidationTest.java:296-315), the level cap 9
 * (CTEST/modwt/MultiLevelMODWTTransformTest.java:271-305), perfect
 * reconstruction / energy / linearity / shift properties.  The reference is
 * pure Java and no JVM exists in the build image, so it cannot be executed
 * here; every function below cites the Java lines it restates.
 *
 * Arithmetic contract: Java never contracts a*b+c, so
 * every sum is  sum = RN(sum + RN(a*b))  in ascending tap order from +0.0.
 * Build with -O2 -ffp-contract=off (see oracle/Makefile) to keep that.
 *
 * Short path names: CORE = vectorwave-core/src/main/java/com/morphiqlabs/wavelet
 *                   EXT  = vectorwave-extensions/src/main/java/com/morphiqlabs/wavelet
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VWO_PERIODIC 0
#define VWO_ZERO 1
#define VWO_SYMMETRIC 2

/* wavelet identities used by SymmetricAlignmentStrategy's `==` tests */
#define VWO_W_OTHER 0
#define VWO_W_DB6 1
#define VWO_W_DB8 2
#define VWO_W_SYM4 3
#define VWO_W_SYM8 4
#define VWO_W_COIF2 5
#define VWO_W_COIF3 6

/* CORE/util/MathUtils.java:30-51  symmetricBoundaryExtension */
static inline int64_t mirror(int64_t idx, int64_t n) {
    if (idx >= 0 && idx < n) return idx;
    int64_t period = 2 * n;
    idx = ((idx % period) + period) % period;
    if (idx >= n) idx = period - idx - 1;
    return idx;
}

/* CORE/internal/ScalarOps.java:700-723  circularConvolveMODWTScalar */
void vwo_conv_periodic(const double *x, int64_t n, const double *f, int64_t lf, double *out) {
    for (int64_t t = 0; t < n; t++) {
        double sum = 0.0;
        for (int64_t l = 0; l < lf; l++) {
            int64_t idx = t - l, si;
            if (idx >= 0 && idx < n) si = idx;
            else if (idx < 0 && idx >= -n) si = idx + n;
            else si = ((idx % n) + n) % n;
            sum += x[si] * f[l];
        }
        out[t] = sum;
    }
}

/* CORE/internal/ScalarOps.java:790-808  zeroPaddingConvolveMODWT */
void vwo_conv_zero(const double *x, int64_t n, const double *f, int64_t lf, double *out) {
    for (int64_t t = 0; t < n; t++) {
        double sum = 0.0;
        for (int64_t l = 0; l < lf; l++) {
            int64_t si = t - l;
            if (si >= 0 && si < n) sum += x[si] * f[l];
        }
        out[t] = sum;
    }
}

/* CORE/internal/ScalarOps.java:818-835  symmetricConvolveMODWT */
void vwo_conv_symmetric(const double *x, int64_t n, const double *f, int64_t lf, double *out) {
    for (int64_t t = 0; t < n; t++) {
        double sum = 0.0;
        for (int64_t l = 0; l < lf; l++) sum += x[mirror(t - l, n)] * f[l];
        out[t] = sum;
    }
}

void vwo_conv(const double *x, int64_t n, const double *f, int64_t lf, int mode, double *out) {
    if (mode == VWO_PERIODIC) vwo_conv_periodic(x, n, f, lf, out);
    else if (mode == VWO_ZERO) vwo_conv_zero(x, n, f, lf, out);
    else vwo_conv_symmetric(x, n, f, lf, out);
}

/* CORE/internal/ScalarOps.java:909-916  upsampleAndScaleForIMODWTSynthesis
 * out has (L-1)*2^(level-1)+1 entries, zero except out[i*up] = base[i]*(1/sqrt 2). */
int64_t vwo_upsampled_len(int64_t l, int level) {
    int64_t up = level <= 1 ? 1 : ((int64_t)1 << (level - 1));
    return (l - 1) * up + 1;
}
void vwo_upsample_scale(const double *base, int64_t l, int level, double *out) {
    int64_t up = level <= 1 ? 1 : ((int64_t)1 << (level - 1));
    double scale = 1.0 / sqrt(2.0);
    int64_t lj = (l - 1) * up + 1;
    memset(out, 0, (size_t)lj * sizeof(double));
    for (int64_t i = 0; i < l; i++) out[i * up] = base[i] * scale;
}

/* CORE/modwt/MultiLevelMODWTTransform.java:455-501  calculateMaxLevels
 * (loop `while (maxLevel < 10)` then `return maxLevel - 1` => cap 9).
 * cap<=0 selects the uncapped variant the engine needs for J=10. */
int vwo_max_levels(int64_t n, int64_t l, int cap) {
    if (n <= l) return 0;
    int limit = cap > 0 ? cap : 62;
    int max_level = 1;
    while (max_level < limit) {
        if (max_level - 1 >= 31 && cap > 0) break;
        int64_t lj = (l - 1) * ((int64_t)1 << (max_level - 1)) + 1;
        if (lj > n) break;
        max_level++;
    }
    return max_level - 1;
}

/* CORE/modwt/MODWTTransform.java:131-189  forward (single level, filters scaled by
 * 1/sqrt 2 once: lines 139-150). */
void vwo_forward_single(const double *x, int64_t n, const double *h, const double *g, int64_t l,
                        int mode, double *v, double *w) {
    double scale = 1.0 / sqrt(2.0);
    double *hs = (double *)malloc(sizeof(double) * (size_t)l * 2), *gs = hs + l;
    for (int64_t i = 0; i < l; i++) { hs[i] = h[i] * scale; gs[i] = g[i] * scale; }
    vwo_conv(x, n, hs, l, mode, v);
    vwo_conv(x, n, gs, l, mode, w);
    free(hs);
}

/* CORE/modwt/MODWTTransform.java:203-299 inverse (pair-added products);
 * batch_variant!=0 selects inverseBatchOptimized's SYMMETRIC rule (t+l), :672-684. */
void vwo_inverse_single(const double *v, const double *w, int64_t n, const double *hr,
                        const double *gr, int64_t l, int mode, int batch_variant, double *out) {
    double scale = 1.0 / sqrt(2.0);
    double *hs = (double *)malloc(sizeof(double) * (size_t)l * 2), *gs = hs + l;
    for (int64_t i = 0; i < l; i++) { hs[i] = hr[i] * scale; gs[i] = gr[i] * scale; }
    for (int64_t t = 0; t < n; t++) {
        double sum = 0.0;
        for (int64_t k = 0; k < l; k++) {
            int64_t ci;
            if (mode == VWO_PERIODIC) ci = (t + k) % n;
            else if (mode == VWO_ZERO) { ci = t + k; if (ci >= n) continue; }
            else if (batch_variant) { int64_t m = (t + k) % (2 * n); ci = m < n ? m : 2 * n - m - 1; }
            else ci = mirror(t - k, n);
            sum += hs[k] * v[ci] + gs[k] * w[ci];
        }
        out[t] = sum;
    }
    free(hs);
}

/* CORE/modwt/MultiLevelMODWTTransform.java:209-255 decompose + :710-757 applyScaledMODWT.
 * w is [J][n]; v is [n].  dense!=0 walks the zero-upsampled filter exactly as the
 * reference does (cost sum_j 2*L_j MACs/sample -- the CPU baseline); dense==0 walks
 * only the L non-zero taps (bit-identical: zero taps add +-0.0). Returns 0, or
 * -1 when L_j > n (VAL_TOO_LARGE, :717-729). */
int vwo_decompose(const double *x, int64_t n, const double *h, const double *g, int64_t l, int levels,
                  int mode, int dense, double *w, double *v) {
    double scale = 1.0 / sqrt(2.0);
    double *cur = (double *)malloc(sizeof(double) * (size_t)n);
    double *nxt = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(cur, x, sizeof(double) * (size_t)n);
    int rc = 0;
    for (int level = 1; level <= levels; level++) {
        int64_t lj = vwo_upsampled_len(l, level);
        if (lj > n) { rc = -1; break; }
        double *wj = w + (size_t)(level - 1) * (size_t)n;
        if (dense) {
            double *hf = (double *)malloc(sizeof(double) * (size_t)lj * 2), *gf = hf + lj;
            vwo_upsample_scale(h, l, level, hf);
            vwo_upsample_scale(g, l, level, gf);
            vwo_conv(cur, n, hf, lj, mode, nxt);
            vwo_conv(cur, n, gf, lj, mode, wj);
            free(hf);
        } else {
            int64_t d = (int64_t)1 << (level - 1);
            for (int64_t t = 0; t < n; t++) {
                double a = 0.0, b = 0.0;
                for (int64_t k = 0; k < l; k++) {
                    int64_t idx = t - k * d;
                    if (mode == VWO_PERIODIC) idx = ((idx % n) + n) % n;
                    else if (mode == VWO_ZERO) { if (idx < 0) continue; }
                    else idx = mirror(idx, n);
                    a += cur[idx] * (h[k] * scale);
                    b += cur[idx] * (g[k] * scale);
                }
                nxt[t] = a; wj[t] = b;
            }
        }
        double *tmp = cur; cur = nxt; nxt = tmp;
    }
    memcpy(v, cur, sizeof(double) * (size_t)n);
    free(cur); free(nxt);
    return rc;
}

/* CORE/modwt/SymmetricAlignmentStrategy.java:43-117 decide().
 * out = {approxPlus, deltaApprox, detailPlus, deltaDetail}. */
void vwo_alignment(int wavelet_id, int64_t l0, int level, int *out) {
    int detail_plus = 1, approx_plus, dh, dg;
    if (l0 <= 2) { approx_plus = 1; dg = 0; dh = level <= 1 ? 0 : -1; }
    else {
        approx_plus = 0;
        if (wavelet_id == VWO_W_DB6) { dh = level <= 1 ? 0 : -1; dg = level >= 3 ? 1 : 0; }
        else if (wavelet_id == VWO_W_DB8) { dh = level <= 1 ? 0 : 1; dg = level >= 2 ? 1 : 0; }
        else if (wavelet_id == VWO_W_SYM4) { approx_plus = 1; detail_plus = 0; dh = 0; dg = 0; }
        else if (wavelet_id == VWO_W_SYM8) {
            if (level <= 1) { dh = 0; dg = 0; } else if (level == 2) { dh = 1; dg = 0; } else { dh = 1; dg = 1; }
        } else if (wavelet_id == VWO_W_COIF2) { approx_plus = 1; dh = level <= 1 ? 0 : 1; detail_plus = 0; dg = 0; }
        else if (wavelet_id == VWO_W_COIF3) {
            detail_plus = 0;
            if (level <= 1) { dh = 0; dg = 0; } else { dh = -1; dg = 1; }
        } else if (l0 >= 12) {
            if (level <= 1) { dh = 0; dg = 0; } else { int even = level % 2 == 0; dh = even ? 0 : -1; dg = even ? 0 : -1; }
        } else { if (level <= 1) { dh = 0; dg = 0; } else { dh = -1; dg = 0; } }
    }
    out[0] = approx_plus; out[1] = dh; out[2] = detail_plus; out[3] = dg;
}

/* CORE/modwt/MultiLevelMODWTTransform.java:795-806 computeTauJ */
static int64_t tau_j(int64_t l0, int level) {
    int64_t lm1 = l0 - 1;
    if (level <= 1) return lm1 / 2 > 0 ? lm1 / 2 : 0;
    int64_t lj = lm1 * ((int64_t)1 << (level - 1)) + 1;
    return (lj - 1) / 2;
}

/* One synthesis stage: CORE/modwt/MultiLevelMODWTTransform.java:554-645
 * applyScaledInverseMODWT.  PERIODIC: H fully then G (578-589; identical loop in
 * CORE/swt/VectorWaveSwtAdapter.java:459-470).  ZERO: pair-added, t+l>=N dropped
 * (591-601).  SYMMETRIC: per-branch orientation/tau (602-642).  dense as above. */
static void synth_level(const double *a, const double *dcoef, int64_t n, const double *hr,
                        const double *gr, int64_t l, int level, int mode, int wavelet_id, int dense,
                        double *out) {
    double scale = 1.0 / sqrt(2.0);
    int64_t d = (int64_t)1 << (level - 1);
    int64_t lj = (l - 1) * d + 1;
    double *hf = (double *)malloc(sizeof(double) * (size_t)lj * 2), *gf = hf + lj;
    vwo_upsample_scale(hr, l, level, hf);
    vwo_upsample_scale(gr, l, level, gf);
    int64_t step = dense ? 1 : d; /* sparse walk visits only the non-zero taps */
    (void)scale;
    if (mode == VWO_PERIODIC) {
        for (int64_t t = 0; t < n; t++) {
            double sum = 0.0;
            for (int64_t q = 0; q < lj; q += step) sum += hf[q] * a[(t + q) % n];
            for (int64_t q = 0; q < lj; q += step) sum += gf[q] * dcoef[(t + q) % n];
            out[t] = sum;
        }
    } else if (mode == VWO_ZERO) {
        for (int64_t t = 0; t < n; t++) {
            double sum = 0.0;
            for (int64_t q = 0; q < lj; q += step) {
                int64_t idx = t + q;
                if (idx < n) sum += hf[q] * a[idx] + gf[q] * dcoef[idx];
            }
            out[t] = sum;
        }
    } else {
        int dec[4];
        vwo_alignment(wavelet_id, l, level, dec);
        int64_t tau_h = tau_j(l, level) + dec[1], tau_g = tau_j(l, level) + dec[3];
        for (int64_t t = 0; t < n; t++) {
            double sum = 0.0;
            for (int64_t q = 0; q < lj; q += step) {
                int64_t idx = dec[0] ? t + q - tau_h : t - q + tau_h;
                sum += hf[q] * a[mirror(idx, n)];
            }
            for (int64_t q = 0; q < lj; q += step) {
                int64_t idx = dec[2] ? t + q - tau_g : t - q + tau_g;
                sum += gf[q] * dcoef[mirror(idx, n)];
            }
            out[t] = sum;
        }
    }
    free(hf);
}

/* CORE/modwt/MultiLevelMODWTTransform.java:339-349 reconstruct, :361-386
 * reconstructFromLevel, :398-446 reconstructLevels.  detail_mask bit (j-1) set =>
 * use W_j, else a zero array; use_approx==0 => zero approximation (the
 * "highest level excluded" branch of reconstructLevels).  w is [J][n]. */
void vwo_reconstruct(const double *w, const double *v, int64_t n, const double *hr, const double *gr,
                     int64_t l, int levels, int mode, int wavelet_id, int dense, uint64_t detail_mask,
                     int use_approx, double *out) {
    double *cur = (double *)calloc((size_t)n, sizeof(double));
    double *nxt = (double *)malloc(sizeof(double) * (size_t)n);
    double *zero = (double *)calloc((size_t)n, sizeof(double));
    if (use_approx) memcpy(cur, v, sizeof(double) * (size_t)n);
    for (int level = levels; level >= 1; level--) {
        const double *dj = ((detail_mask >> (level - 1)) & 1u) ? w + (size_t)(level - 1) * (size_t)n : zero;
        synth_level(cur, dj, n, hr, gr, l, level, mode, wavelet_id, dense, nxt);
        double *tmp = cur; cur = nxt; nxt = tmp;
    }
    memcpy(out, cur, sizeof(double) * (size_t)n);
    free(cur); free(nxt); free(zero);
}

/* CORE/modwt/MutableMultiLevelMODWTResult.java:97-118 applyThresholdToArray */
void vwo_threshold(double *c, int64_t n, double thr, int soft) {
    for (int64_t i = 0; i < n; i++) {
        double a = fabs(c[i]);
        if (soft) {
            if (a > thr) { double sg = c[i] > 0 ? 1.0 : (c[i] < 0 ? -1.0 : c[i]); c[i] = sg * (a - thr); }
            else c[i] = 0.0;
        } else if (a <= thr) c[i] = 0.0;
    }
}

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* CORE/swt/VectorWaveSwtAdapter.java:627-645 estimateNoiseSigma + :505-520
 * universal threshold = sigma * sqrt(2 ln N). */
double vwo_universal_threshold(const double *w1, int64_t n) {
    double *a = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; i++) a[i] = fabs(w1[i]);
    qsort(a, (size_t)n, sizeof(double), cmp_double);
    double med = (n % 2 == 0) ? (a[n / 2 - 1] + a[n / 2]) / 2.0 : a[n / 2];
    free(a);
    double sigma = med / 0.6745;
    return sigma * sqrt(2 * log((double)n));
}

/* CORE/denoising/WaveletDenoiser.java:477-492 calculateSURERisk */
static double sure_risk(const double *c, int64_t n, double threshold, double sigma) {
    double sigma2 = sigma * sigma;
    double risk = (double)(-n) * sigma2;
    for (int64_t i = 0; i < n; i++) {
        double a = fabs(c[i]);
        if (a <= threshold) risk += c[i] * c[i];
        else risk += sigma2 + (a - threshold) * (a - threshold);
    }
    return risk / (double)n;
}

/* CORE/denoising/WaveletDenoiser.java:441-472 calculateSUREThreshold: every sorted |c| is a candidate (the O(n^2) double
 * loop), first strict minimum wins, capped at the universal threshold.  *min_risk_out (may be NULL) = the minimum. */
double vwo_sure_threshold(const double *c, int64_t n, double sigma, double *min_risk_out) {
    double *a = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; i++) a[i] = fabs(c[i]);
    qsort(a, (size_t)n, sizeof(double), cmp_double);
    double min_risk = INFINITY, best = 0.0;
    for (int64_t k = 0; k < n; k++) {
        double risk = sure_risk(c, n, a[k], sigma);
        if (risk < min_risk) { min_risk = risk; best = a[k]; }
    }
    free(a);
    double universal = sigma * sqrt(2.0 * log((double)n));
    if (best > universal) best = universal;
    if (min_risk_out) *min_risk_out = min_risk;
    return best;
}

/* CORE/swt/VectorWaveSwtAdapter.java:546-562 denoise: decompose, threshold all detail
 * levels (thr<0 => universal), reconstruct.  Returns the threshold used. */
double vwo_swt_denoise(const double *x, int64_t n, const double *h, const double *g, int64_t l,
                       int levels, int mode, int wavelet_id, double thr, int soft, int dense, double *out) {
    double *w = (double *)malloc(sizeof(double) * (size_t)n * (size_t)(levels + 1));
    double *v = w + (size_t)n * (size_t)levels;
    vwo_decompose(x, n, h, g, l, levels, mode, dense, w, v);
    if (thr < 0) thr = vwo_universal_threshold(w, n);
    for (int j = 0; j < levels; j++) vwo_threshold(w + (size_t)j * (size_t)n, n, thr, soft);
    uint64_t mask = levels >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << levels) - 1);
    vwo_reconstruct(w, v, n, h, g, l, levels, mode, wavelet_id, dense, mask, 1, out);
    free(w);
    return thr;
}

/* ---- batch forms, used as CPU baselines --------------------------
 * core scalar: loop of decompose/reconstruct per signal (CORE/modwt/MODWTTransform.java:589-605);
 * structured concurrency: one signal per task over all host threads
 * (EXT/extensions/parallel/StructuredParallelTransform.java:185-218) -- OpenMP here. */
typedef struct {
    const double *x; int64_t b, n; const double *h, *g; int64_t l; int levels, mode, wavelet_id, dense;
    double *w, *v, *xr; int64_t *next; pthread_mutex_t *mu;
} vwo_job;

static void *vwo_worker(void *arg) {
    vwo_job *j = (vwo_job *)arg;
    uint64_t mask = j->levels >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << j->levels) - 1);
    for (;;) {
        pthread_mutex_lock(j->mu);
        int64_t i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->b) break;
        double *wi = j->w + (size_t)i * (size_t)j->levels * (size_t)j->n; /* [b][J][n] scratch layout */
        double *vi = j->v + (size_t)i * (size_t)j->n;
        vwo_decompose(j->x + (size_t)i * (size_t)j->n, j->n, j->h, j->g, j->l, j->levels, j->mode, j->dense, wi, vi);
        if (j->xr)
            vwo_reconstruct(wi, vi, j->n, j->h, j->g, j->l, j->levels, j->mode, j->wavelet_id, j->dense, mask, 1,
                            j->xr + (size_t)i * (size_t)j->n);
    }
    return NULL;
}

/* xr==NULL => forward only.  threads<=1 runs inline on the caller's thread. */
void vwo_batch_fwd_inv(const double *x, int64_t b, int64_t n, const double *h, const double *g, int64_t l,
                       int levels, int mode, int wavelet_id, int dense, int threads, double *w, double *v,
                       double *xr) {
    int64_t next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    vwo_job job = {x, b, n, h, g, l, levels, mode, wavelet_id, dense, w, v, xr, &next, &mu};
    if (threads <= 1) { vwo_worker(&job); return; }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&tid[t], NULL, vwo_worker, &job);
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid);
}

/* EXT/extensions/modwt/BatchSIMDMODWT.java:343-424 batchMultiLevelMODWTSoA: SoA index
 * t*B+b, lanes = signals, dense upsampled taps, srcT=(t-l+N)%N, non-fused
 * `sum.add(samples.mul(c))`.  soa_w is [J][n*b], soa_v is [n*b]. */
void vwo_batch_soa_decompose(const double *soa_x, int64_t b, int64_t n, const double *h, const double *g,
                             int64_t l, int levels, double *soa_w, double *soa_v) {
    size_t tot = (size_t)b * (size_t)n;
    double *cur = (double *)malloc(sizeof(double) * tot), *nxt = (double *)malloc(sizeof(double) * tot);
    memcpy(cur, soa_x, sizeof(double) * tot);
    for (int level = 1; level <= levels; level++) {
        int64_t lj = vwo_upsampled_len(l, level);
        double *hf = (double *)malloc(sizeof(double) * (size_t)lj * 2), *gf = hf + lj;
        vwo_upsample_scale(h, l, level, hf);
        vwo_upsample_scale(g, l, level, gf);
        double *wj = soa_w + (size_t)(level - 1) * tot;
        for (int64_t t = 0; t < n; t++) {
            double *ao = nxt + (size_t)t * (size_t)b, *dd = wj + (size_t)t * (size_t)b;
            for (int64_t i = 0; i < b; i++) { ao[i] = 0.0; dd[i] = 0.0; }
            for (int64_t q = 0; q < lj; q++) {
                int64_t st = (t - q + n) % n;
                const double *s = cur + (size_t)st * (size_t)b;
                double ch = hf[q], cg = gf[q];
                for (int64_t i = 0; i < b; i++) { ao[i] = ao[i] + s[i] * ch; dd[i] = dd[i] + s[i] * cg; }
            }
        }
        free(hf);
        double *tmp = cur; cur = nxt; nxt = tmp;
    }
    memcpy(soa_v, cur, sizeof(double) * tot);
    free(cur); free(nxt);
}

/* EXT/extensions/modwt/BatchSIMDMODWT.java:86-140 haarBatchMODWTSoA: literal +-0.5 taps. */
void vwo_batch_soa_haar_single(const double *soa_x, int64_t b, int64_t n, double *soa_v, double *soa_w) {
    for (int64_t t = 0; t < n; t++) {
        int64_t tm1 = (t - 1 + n) % n;
        for (int64_t i = 0; i < b; i++) {
            double s0 = soa_x[t * b + i], s1 = soa_x[tm1 * b + i];
            soa_v[t * b + i] = s0 * 0.5 + s1 * 0.5;
            soa_w[t * b + i] = s0 * 0.5 + s1 * -0.5;
        }
    }
}
