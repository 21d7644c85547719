// Measures the FP64 FMA peak (the second roofline denominator; MEASURED_PEAKS.json has no FP64 entry)
// and a device copy bandwidth, with CUDA events.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(1024) k_dfma(double *out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < CHAINS; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__global__ void k_copy(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = in[i];
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *d; cudaMalloc(&d, 1 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, chains = 8;
    for (int threads : {256, 512, 1024}) {
        int blocks = sms * (2048 / threads);
        k_dfma<chains><<<blocks, threads>>>(d, 64, 1.0000001, 1e-9);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            k_dfma<chains><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double fl = 2.0 * (double)blocks * threads * iters * 8 * chains;
        printf("{\"fp64_fma_tflops\": %.2f, \"threads\": %d, \"blocks\": %d, \"ms\": %.3f, \"sms\": %d}\n",
               fl / best * 1e-9, threads, blocks, best, sms);
    }
    // sustained: 2 s back to back
    {
        int threads = 1024, blocks = sms * 2;
        cudaEventRecord(e0);
        int launches = 0; float ms = 0;
        while (ms < 2000.f) {
            for (int i = 0; i < 10; i++) k_dfma<chains><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            launches += 10;
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double fl = 2.0 * (double)blocks * threads * iters * 8 * chains * launches;
        printf("{\"fp64_fma_tflops_sustained\": %.2f, \"seconds\": %.2f}\n", fl / ms * 1e-9, ms * 1e-3);
    }
    size_t n = (size_t)1 << 30;  // 1 GiB each way
    double2 *a, *b; cudaMalloc(&a, n); cudaMalloc(&b, n); cudaMemset(a, 1, n);
    k_copy<<<sms * 8, 1024>>>(a, b, n / 16); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 10; rep++) {
        cudaEventRecord(e0); k_copy<<<sms * 8, 1024>>>(a, b, n / 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("{\"copy_gbs\": %.1f}\n", 2.0 * n / best * 1e-6);
    float bestm = 1e30f;
    for (int rep = 0; rep < 10; rep++) {
        cudaEventRecord(e0); cudaMemcpyAsync(b, a, n, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < bestm) bestm = ms;
    }
    printf("{\"memcpy_d2d_gbs\": %.1f}\n", 2.0 * n / bestm * 1e-6);
    return 0;
}
