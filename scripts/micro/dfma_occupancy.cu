// Micro-benchmark: DFMA issue rate per SM as a function of resident warps per scheduler and independent chains per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_occupancy dfma_occupancy.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(int iters, double *sink, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) sink[0] = s;
}

template <int ILP>
void run(int warps_per_sm, int sms, double clock_ghz)
{
    double *sink;
    cudaMalloc(&sink, 8);
    const int iters = 20000;
    const int threads = 32 * warps_per_sm;   // one CTA per SM
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<ILP><<<sms, threads>>>(100, sink, 0.999999, 1e-7);
    cudaEventRecord(e0);
    k<ILP><<<sms, threads>>>(iters, sink, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma_per_s = (double)sms * threads * ILP * iters / (ms * 1e-3);
    printf("warps/SM %2d (%.1f per scheduler)  ILP %d : %.3e DFMA/s  = %.1f lanes/clk/SM at %.3f GHz\n", warps_per_sm, warps_per_sm / 4.0, ILP,
           fma_per_s, fma_per_s / sms / (clock_ghz * 1e9), clock_ghz);
    cudaFree(sink);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const double ghz = 1.965;
    for (int w : {4, 8, 16, 24, 32, 64}) {
        run<1>(w, sms, ghz);
        run<2>(w, sms, ghz);
        run<4>(w, sms, ghz);
        run<8>(w, sms, ghz);
    }
    return 0;
}
