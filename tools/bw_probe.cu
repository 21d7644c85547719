// Developer probe: achievable HBM bandwidth on B200 by read:write mix (streaming, 16-byte accesses, buffers >> L2).
#include <cstdio>
#include <cuda_runtime.h>

template <int NR, int NW>
__global__ void k_mix(const double2 *__restrict__ a, const double2 *__restrict__ b, double2 *__restrict__ o0,
                      double2 *__restrict__ o1, size_t n, double *sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    double acc = 0;
    for (; i < n; i += st) {
        double2 v = make_double2(1.0, 2.0);
        if (NR >= 1) v = a[i];
        if (NR >= 2) { double2 u = b[i]; v.x += u.x; v.y += u.y; }
        if (NW >= 1) o0[i] = v;
        if (NW >= 2) o1[i] = make_double2(v.y, v.x);
        if (NW == 0) acc += v.x + v.y;
    }
    if (NW == 0 && acc == 123.456) *sink = acc;
}

template <int NR, int NW>
void run(const char *name, double2 *a, double2 *b, double2 *c, double2 *d, size_t n, double *sink, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks_per_sm : {8, 16, 32}) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            k_mix<NR, NW><<<sms * blocks_per_sm, 256>>>(a, b, c, d, n, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        printf("{\"mix\": \"%s\", \"blocks_per_sm\": %d, \"gbs\": %.0f}\n", name, blocks_per_sm, (NR + NW) * n * 16.0 / best * 1e-6);
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    size_t n = (size_t)1 << 27;   // 2 GiB per buffer
    double2 *a, *b, *c, *d; double *sink;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&c, n * 16); cudaMalloc(&d, n * 16); cudaMalloc(&sink, 8);
    cudaMemset(a, 0, n * 16); cudaMemset(b, 0, n * 16);
    run<1, 0>("1R0W", a, b, c, d, n, sink, p.multiProcessorCount);
    run<2, 0>("2R0W", a, b, c, d, n, sink, p.multiProcessorCount);
    run<0, 1>("0R1W", a, b, c, d, n, sink, p.multiProcessorCount);
    run<0, 2>("0R2W", a, b, c, d, n, sink, p.multiProcessorCount);
    run<1, 1>("1R1W", a, b, c, d, n, sink, p.multiProcessorCount);
    run<1, 2>("1R2W", a, b, c, d, n, sink, p.multiProcessorCount);
    run<2, 1>("2R1W", a, b, c, d, n, sink, p.multiProcessorCount);
    return 0;
}
