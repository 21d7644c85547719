// vw_internal.cuh -- shared declarations of the MODWT engine's translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "vw_modwt.h"

#define VW_MODE_LINEAR 4  // internal: span calls -- indices outside the provided buffer are skipped

// Pre-scaled base filters passed by value in kernel parameter space (constant bank).
struct VwFilt {
    double h[VW_MAX_FILTER_TAPS];
    double g[VW_MAX_FILTER_TAPS];
};

// Compact filter block of the fused kernels (L <= 32): fully unrolled tap loops read these
// straight from the constant bank as DFMA operands.
#define VW_FUSED_MAX_L 32
struct VwFilt32 {
    double h[VW_FUSED_MAX_L];
    double g[VW_FUSED_MAX_L];
};
#ifdef __CUDACC__
// high-pass tap k with compile-time k: f.g[k], or +-f.h[L-1-k] when the pair is a quadrature mirror
template <int L, bool QMF>
__device__ __forceinline__ double vw_tap_g(const VwFilt32 &f, int k) {
    return QMF ? ((k & 1) ? -f.h[L - 1 - k] : f.h[L - 1 - k]) : f.g[k];
}
#endif

// Orthogonal wavelets: g[k] = (-1)^k h[L-1-k] (quadrature mirror, CORE/api/Daubechies.java:323-330 and the Symlet /
// Coiflet twins); scaling by 1/sqrt(2) keeps the relation bit-exact.  When it holds, kernels for long filters read
// only the L low-pass taps -- 2L doubles no longer fit the 63 uniform registers of an SM sub-partition, L do -- and
// negate on the fly (free operand modifier).  Anything else (biorthogonal pairs, reversed streams) keeps both arrays.
inline bool vw_is_qmf(const double *h, const double *g, int l) {
    for (int k = 0; k < l; k++)
        if (g[k] != ((k & 1) ? -h[l - 1 - k] : h[l - 1 - k])) return false;
    return true;
}
struct VwPlanGroup { int first, nlev; int64_t tile; double cost; };

struct vw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    // grow-only device scratch (ping-pong V buffers, staged host I/O, selection workspaces)
    static const int kScratch = 6;         // per set; set 1 belongs to the second stream of the pipelined host path
    void *scratch[2 * kScratch] = {};
    size_t scratch_bytes[2 * kScratch] = {};
    int scratch_set = 0;
    cudaStream_t pipe_stream[2] = {nullptr, nullptr};   // created on first use by the pipelined host path
    int64_t opt_pipe_min = 64ll << 20;                  // staged bytes from which host calls are chunked and overlapped
    void *pinned = nullptr;  // small pinned mailbox for D2H scalars
    size_t pinned_bytes = 0;
    struct PlanEntry { bool forward; int l, levels; int64_t n; std::vector<VwPlanGroup> groups; };
    mutable std::vector<PlanEntry> plan_cache;   // launch plans by shape (the planner costs 2-35 us); cleared by vw_set_option
    std::vector<std::pair<const void *, size_t>> smem_set;   // dynamic shared memory already opted in, per kernel
    struct OccEntry { const void *func; int nthreads; size_t smem; int per_sm; };
    std::vector<OccEntry> occ_cache;   // occupancy queries of the tile kernels (vw_fused.cu: prefetch_distance)
    int64_t opt_l2pf = 1;    // tile kernels prefetch the successor CTA's input tile into L2 (x resident CTAs ahead); 0 = off
    std::recursive_mutex mu;   // every public entry point holds it: calls on one ctx from several host threads serialise
    int64_t opt_tile = 0, opt_fuse = 0, opt_threads = 0, opt_poly = 1;
    int64_t opt_wave = 1;    // column kernels: size single-signal grids to whole waves
    int64_t opt_colmin = 0;  // first level the column kernels may take (0 = auto: see vw_column_min_level)
};

// MutableMultiLevelMODWTResult.applyThresholdToArray (CORE/modwt/MutableMultiLevelMODWTResult.java:97-118):
// soft: |c| > t ? Math.signum(c) * (|c| - t) : 0;  hard: |c| <= t ? 0 : c.  One definition for every kernel that
// thresholds on the fly.
#ifdef __CUDACC__
__device__ __forceinline__ double vw_threshold_value(double c, double lam, int soft) {
    const double a = fabs(c), m = a - lam;
    if (soft) return a > lam ? (c > 0.0 ? m : (c < 0.0 ? -m : c * m)) : 0.0;
    return a <= lam ? 0.0 : c;
}
// same value for lam >= 0 (or NaN): |c| > lam then implies c != 0 and not NaN, so signum(c) * m == copysign(m, c) --
// two FP64-pipe instructions instead of five
__device__ __forceinline__ double vw_threshold_nonneg(double c, double lam, int soft) {
    const double a = fabs(c);
    if (soft) return a > lam ? copysign(a - lam, c) : 0.0;
    return a <= lam ? 0.0 : c;
}
#endif

int vw_fail(vw_ctx *ctx, int status, const char *fmt, ...);
int vw_cuda_check(vw_ctx *ctx, cudaError_t e, const char *what);
int vw_scratch(vw_ctx *ctx, int slot, size_t bytes, void **out);

// ---- generic per-level kernels (vw_generic.cu): any n, l, dilation, mode ------------------
int vw_launch_analysis_level(vw_ctx *ctx, const double *in, int64_t ld_in, double *v_out, int64_t ld_v,
                             double *w_out, int64_t ld_w, int64_t n_in, int64_t t0, int64_t n_out, int64_t batch,
                             const VwFilt &f, int l, int64_t d, int mode, bool exact);
int vw_launch_synthesis_level(vw_ctx *ctx, const double *v, int64_t ld_v, const double *w, int64_t ld_w, double *out,
                              int64_t ld_o, int64_t n_in, int64_t t0, int64_t n_out, int64_t batch, const VwFilt &f,
                              int l, int64_t d, int mode, vw_align al, bool pair, bool exact);
int vw_launch_conv_dense(vw_ctx *ctx, const double *x, int64_t n, const double *filter, int64_t lf, int mode,
                         double *out, bool exact);
int vw_launch_threshold(vw_ctx *ctx, double *c, int64_t batch, int64_t n, int64_t ld, const double *thr_dev,
                        int per_row, int soft);
int vw_launch_nonfinite_count(vw_ctx *ctx, const double *x, int64_t batch, int64_t n, int64_t ld,
                              unsigned long long *count_dev);
int vw_launch_energy(vw_ctx *ctx, const double *c, int64_t batch, int64_t n, int64_t ld, double *out_dev);

// ---- fused multi-level tile kernels (vw_fused.cu) ------------------------------------------
// Return VW_OK when the fused path ran, VW_EUNSUPPORTED when the shape is not eligible (caller
// falls back to the per-level kernels), anything else is an error.
struct VwFusedFwd {
    const double *x; int64_t ldx;            // V_{first-1}: [batch][n_in]
    double *w; int64_t ldw, lsw;             // W_j rows for j = first .. first+nlev-1
    double *v; int64_t ldv;                  // V_{first+nlev-1}
    int64_t batch, n_in, t0, n_out;          // outputs cover input positions [t0, t0+n_out)
    int l, first_level, nlevels, mode;
    int64_t tile;                            // 0 = choose here
};
int vw_fused_forward(vw_ctx *ctx, const VwFusedFwd &p, const VwFilt &f);
struct VwFusedInv {
    const double *v; int64_t ldv;            // V_top: [batch][n_in]  (nullptr => zeros)
    const double *w; int64_t ldw, lsw;       // W_j rows, j = first .. first+nlev-1
    uint64_t detail_mask;                    // bit (j-first) clear => zeros
    double *out; int64_t ldo;                // V_{first-1}
    int64_t batch, n_in, n_out;              // outputs cover positions [0, n_out) of the input buffers
    int l, first_level, nlevels, mode;
    const double *thr_dev; int thr_per_row; int thr_soft;  // optional fused thresholding of W on load
    int64_t tile;                            // 0 = choose here
    bool has_align = false;                  // single-level stage with a general (sigma, tau) alignment (any mode)
    vw_align align = {1, 0, 1, 0};
};

// Level schedule: which consecutive levels share one fused launch and with which tile, from a small cost model
// (FP64-pipe cycles incl. halo recompute and ragged rounds vs HBM bytes incl. halo re-reads, overlapped across the
// CTAs that fit one SM).  tile < 0 marks a level that only the per-level kernels can take.
#include <vector>
int vw_plan_levels(const vw_ctx *ctx, bool forward, int l, int levels, int64_t n, std::vector<VwPlanGroup> &out);
int vw_fused_inverse(vw_ctx *ctx, const VwFusedInv &p, const VwFilt &f);

// first level (1-based) the column kernels take for filter length l: dilation >= 32 normally; FP64-bound filters
// (l >= 24) already from dilation 4 (analysis) / 16 (synthesis, whose two input streams suffer more from 32-byte row pieces)
int vw_column_min_level(const vw_ctx *ctx, int l, bool forward = true);
// ---- deep single levels, dilation >= 32 (vw_column.cu): register sliding window per dilation column ---------
int vw_column_analysis(vw_ctx *ctx, const double *x, int64_t ldx, double *v, int64_t ldv, double *w, int64_t ldw,
                       int64_t n_in, int64_t t0, int64_t n_out, int64_t batch, const VwFilt &f, int l, int64_t d, int mode);
int vw_column_synthesis(vw_ctx *ctx, const double *v, int64_t ldv, const double *w, int64_t ldw, double *out, int64_t ldo,
                        int64_t n_in, int64_t t0, int64_t n_out, int64_t batch, const VwFilt &f, int l, int64_t d, int mode,
                        vw_align al, const double *thr_dev = nullptr, int thr_per_row = 0, int thr_soft = 0);

// ---- exact order statistics (vw_select.cu) ---------------------------------------------------
// median(|w1|) per row -> universal thresholds written to thr_dev[batch] (device).
int vw_launch_universal_threshold(vw_ctx *ctx, const double *w1, int64_t batch, int64_t n, int64_t ld,
                                  double *thr_dev);
// same selection, raw median(|c|) per row
int vw_launch_median_abs(vw_ctx *ctx, const double *c, int64_t batch, int64_t n, int64_t ld, double *med_dev);
size_t vw_sure_workspace(int64_t batch, int64_t n);
int vw_launch_sure(vw_ctx *ctx, const double *c, int64_t batch, int64_t n, int64_t ld, const double *sigma_dev, void *ws,
                   double *thr_dev, double *risk_dev);
int vw_launch_mean_variance(vw_ctx *ctx, const double *c, int64_t batch, int64_t n, int64_t ld, double *mean_dev,
                            double *var_dev);
