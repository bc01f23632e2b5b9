// FP64 pipe microbenchmark: independent DFMA / DADD / DMUL chains per thread, enough warps to
// saturate the pipe.  Prints achieved warp-instructions per clock per SM and TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            if (OP == 1) x[i] = x[i] + b;
            if (OP == 2) x[i] = x[i] * a;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int sms, double mhz) {
    const int blocks = sms * 4, threads = 512, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = double(blocks) * threads * iters * 8.0;
    const double per_clk_sm = ops / (ms * 1e-3) / (mhz * 1e6) / sms;
    printf("%s: %.3f ms, %.1f lane-ops/clk/SM (at %.0f MHz), %.2f T%s/s\n", name, ms, per_clk_sm, mhz,
           ops * (OP == 0 ? 2 : 1) / (ms * 1e-3) / 1e12, OP == 0 ? "FLOP" : "OP");
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, khz);
    run<0>("DFMA", p.multiProcessorCount, khz / 1e3);
    run<1>("DADD", p.multiProcessorCount, khz / 1e3);
    run<2>("DMUL", p.multiProcessorCount, khz / 1e3);
    return 0;
}
