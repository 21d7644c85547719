// vw_probe.cu -- the FP64 roofline denominator, measured on the device the engine runs on.
// MEASURED_PEAKS.json carries the HBM copy rate and the bf16 GEMM rate of the pool, but no FP64 figure, and the MODWT
// path's second roof is the FP64 FMA pipe (4*L FLOP per sample per level).  vw_probe_fp64 runs the instruction the tile
// kernels are made of -- DFMA with one uniform-register operand, eight independent chains per thread, 64 warps per SM --
// for a few milliseconds and reports what the pipe sustains.  Not on the product path; bench.py calls it once per run.
#include "vw_internal.cuh"

namespace {
struct ProbeTaps { double t[8]; };

__global__ void __launch_bounds__(256) k_probe_dfma(double *out, int iters, const __grid_constant__ ProbeTaps tp) {
    double acc[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = threadIdx.x * 1e-9 + i; x[i] = 1.0 + 1e-12 * (threadIdx.x + i); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = fma(x[(r + i) & 7], tp.t[r], acc[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    if (s == 123.456) out[0] = s;   // keeps the chains alive; never true
}
}  // namespace

extern "C" int vw_probe_fp64(vw_ctx *ctx, double *tflops_out, double *sm_mhz_out) {
    if (!ctx || !tflops_out) return VW_ENULL;
    vwshim::DeviceGuard g(ctx, "vw_probe_fp64");
    if (int rc = vwshim::no_capture(ctx, "vw_probe_fp64")) return rc;
    void *scratch;
    int rc = vw_scratch(ctx, 5, 64, &scratch);
    if (rc) return rc;
    ProbeTaps tp;
    for (int i = 0; i < 8; i++) tp.t[i] = 1.0 + 1e-9 * i;
    const int blocks = ctx->sm_count * 8, iters = 4096;
    cudaEvent_t e0, e1;
    if ((rc = vw_cuda_check(ctx, cudaEventCreate(&e0), "probe event"))) return rc;
    if ((rc = vw_cuda_check(ctx, cudaEventCreate(&e1), "probe event"))) { cudaEventDestroy(e0); return rc; }
    k_probe_dfma<<<blocks, 256, 0, ctx->stream>>>((double *)scratch, 64, tp);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5 && rc == VW_OK; rep++) {
        cudaEventRecord(e0, ctx->stream);
        k_probe_dfma<<<blocks, 256, 0, ctx->stream>>>((double *)scratch, iters, tp);
        cudaEventRecord(e1, ctx->stream);
        rc = vw_cuda_check(ctx, cudaEventSynchronize(e1), "probe");
        float ms = 0.f;
        if (!rc && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms > 0.f && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    const double dfma = (double)blocks * 256.0 * iters * 64.0;
    *tflops_out = 2.0 * dfma / (best * 1e-3) * 1e-12;
    if (sm_mhz_out) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        *sm_mhz_out = khz * 1e-3;   // the driver's maximum SM clock; the achieved rate above already contains the real one
    }
    return VW_OK;
}
