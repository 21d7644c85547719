// tools/latency.cpp -- call latency of the small shapes through the C ABI, from a compiled host (what a JVM caller sees;
// the Python binding adds ~10 us of ctypes/numpy work per call).  Prints one JSON object.  Built by __graft_entry__.build()
// (g++ only: the program uses nothing but include/vw_modwt.h), run by bench.py for `extra.c1_latency_us`.
//   c1: 1 x 4096, db4, J = 1, forward, HOST (pinned) buffers -- H2D + kernel + 2 D2H + sync per call; the reference's JVM
//       takes 358 us (core) / 117 us (extensions) for this call (docs/BENCHMARK-RESULTS.md:26)
//   c2s: 16 x 4096, db4, J = 4, forward, device-resident
// each as plain synchronous calls, as back-to-back VW_FLAG_NO_SYNC calls (device shapes) and as CUDA-graph replays.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <vector>

#include "vw_modwt.h"

static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define CK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s failed: %d %s\n", #x, rc_, vw_last_error(ctx)); return 1; } } while (0)

int main(int argc, char **argv) {
    const int device = argc > 1 ? atoi(argv[1]) : 0;
    const int reps = argc > 2 ? atoi(argv[2]) : 2000;
    vw_ctx *ctx = nullptr;
    if (vw_init(device, &ctx)) { fprintf(stderr, "vw_init failed (no CUDA device?)\n"); return 2; }
    const double s = 1.0 / sqrt(2.0);
    const double h0[8] = {0.2303778133088964, 0.7148465705529154, 0.6308807679298587, -0.0279837693982488,
                          -0.1870348117190931, 0.0308413818355607, 0.0328830116668852, -0.0105974017850690};
    double hs[8], gs[8];
    for (int k = 0; k < 8; k++) hs[k] = h0[k] * s;
    for (int k = 0; k < 8; k++) gs[k] = ((k & 1) ? -1.0 : 1.0) * h0[7 - k] * s;

    // ---- c1: host buffers ---------------------------------------------------------------------------------
    const int64_t n = 4096;
    double *x = (double *)vw_alloc_pinned(n * 8), *w = (double *)vw_alloc_pinned(n * 8), *v = (double *)vw_alloc_pinned(n * 8);
    for (int64_t i = 0; i < n; i++) x[i] = sin(0.01 * i) + 0.1 * ((i * 2654435761u) % 1000) * 1e-3;
    for (int i = 0; i < 20; i++) CK(vw_modwt_forward(ctx, x, 1, n, n, hs, gs, 8, 1, VW_PERIODIC, w, n, n, v, n, 0));
    double t0 = now_us();
    for (int i = 0; i < reps; i++) CK(vw_modwt_forward(ctx, x, 1, n, n, hs, gs, 8, 1, VW_PERIODIC, w, n, n, v, n, 0));
    const double c1_call = (now_us() - t0) / reps;
    vw_graph *g1 = nullptr;
    CK(vw_graph_begin(ctx));
    CK(vw_modwt_forward(ctx, x, 1, n, n, hs, gs, 8, 1, VW_PERIODIC, w, n, n, v, n, 0));
    CK(vw_graph_end(ctx, &g1));
    for (int i = 0; i < 20; i++) CK(vw_graph_launch(ctx, g1, 0));
    t0 = now_us();
    for (int i = 0; i < reps; i++) CK(vw_graph_launch(ctx, g1, 0));
    const double c1_graph = (now_us() - t0) / reps;
    const double w_check = w[17];
    // the same call with the pinned buffers addressed in place by the kernels (vw_set_option "zero_copy"): run with a third
    // argument "zc"; results must be identical to the staged call's
    double c1_zc = -1.0, zc_diff = -1.0;
    if (argc > 3) {
        std::vector<double> wref(w, w + n), vref(v, v + n);
        CK(vw_set_option(ctx, "zero_copy", 1 << 20));
        for (int i = 0; i < 20; i++) CK(vw_modwt_forward(ctx, x, 1, n, n, hs, gs, 8, 1, VW_PERIODIC, w, n, n, v, n, 0));
        t0 = now_us();
        for (int i = 0; i < reps; i++) CK(vw_modwt_forward(ctx, x, 1, n, n, hs, gs, 8, 1, VW_PERIODIC, w, n, n, v, n, 0));
        c1_zc = (now_us() - t0) / reps;
        zc_diff = 0.0;
        for (int64_t i = 0; i < n; i++) zc_diff = fmax(zc_diff, fmax(fabs(w[i] - wref[i]), fabs(v[i] - vref[i])));
        CK(vw_set_option(ctx, "zero_copy", 0));
    }

    // ---- c2s: 16 x 4096, J = 4, device-resident -------------------------------------------------------------
    const int64_t b = 16, levels = 4;
    void *xd, *wd, *vd;
    CK(vw_device_alloc(ctx, b * n * 8, &xd));
    CK(vw_device_alloc(ctx, levels * b * n * 8, &wd));
    CK(vw_device_alloc(ctx, b * n * 8, &vd));
    std::vector<double> xh(b * n);
    for (size_t i = 0; i < xh.size(); i++) xh[i] = cos(0.003 * i);
    CK(vw_copy_h2d(ctx, xd, xh.data(), xh.size() * 8));
    const uint32_t D = VW_FLAG_DEVICE_PTRS;
    auto fwd = [&](uint32_t fl) {
        return vw_modwt_forward(ctx, (const double *)xd, b, n, n, hs, gs, 8, (int32_t)levels, VW_PERIODIC, (double *)wd, n, b * n,
                                (double *)vd, n, D | fl);
    };
    for (int i = 0; i < 20; i++) CK(fwd(0));
    t0 = now_us();
    for (int i = 0; i < reps; i++) CK(fwd(0));
    const double c2_sync = (now_us() - t0) / reps;
    t0 = now_us();
    for (int i = 0; i < reps; i++) CK(fwd(VW_FLAG_NO_SYNC));
    CK(vw_synchronize(ctx));
    const double c2_nosync = (now_us() - t0) / reps;
    vw_graph *g2 = nullptr;
    CK(vw_graph_begin(ctx));
    CK(fwd(VW_FLAG_NO_SYNC));
    CK(vw_graph_end(ctx, &g2));
    t0 = now_us();
    for (int i = 0; i < reps; i++) CK(vw_graph_launch(ctx, g2, VW_FLAG_NO_SYNC));
    CK(vw_synchronize(ctx));
    const double c2_graph = (now_us() - t0) / reps;
    t0 = now_us();
    for (int i = 0; i < reps; i++) CK(vw_graph_launch(ctx, g2, 0));
    const double c2_graph_sync = (now_us() - t0) / reps;

    printf("{\"reps\": %d, \"c1_1x4096_db4_J1_forward_host_buffers_us\": %.2f, \"c1_graph_replay_us\": %.2f, "
           "\"c2s_16x4096_db4_J4_forward_device_sync_us\": %.2f, \"c2s_no_sync_back_to_back_us\": %.2f, "
           "\"c2s_graph_replay_back_to_back_us\": %.2f, \"c2s_graph_replay_sync_us\": %.2f, \"check\": %.17g, "
           "\"c1_zero_copy_us\": %.2f, \"c1_zero_copy_max_abs_diff\": %g, "
           "\"reference_jvm_us\": {\"core\": 358, \"extensions\": 117, \"source\": \"docs/BENCHMARK-RESULTS.md:26\"}}\n",
           reps, c1_call, c1_graph, c2_sync, c2_nosync, c2_graph, c2_graph_sync, w_check, c1_zc, zc_diff);
    vw_graph_destroy(ctx, g1);
    vw_graph_destroy(ctx, g2);
    vw_device_free(ctx, xd); vw_device_free(ctx, wd); vw_device_free(ctx, vd);
    vw_free_pinned(x); vw_free_pinned(w); vw_free_pinned(v);
    vw_destroy(ctx);
    return 0;
}
