// FP64 issue rate on sm_100a: K independent DFMA chains per thread, 1..8 warps of one CTA (clock64 per warp-instruction).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_tput fp64_tput.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void k(double *out, long long *cyc, double x0)
{
    double x[K];
#pragma unroll
    for (int j = 0; j < K; ++j) x[j] = x0 + threadIdx.x * 1e-3 + j;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < K; ++j) x[j] = __fma_rn(x[j], 1.0000001, 1e-9);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) s += x[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int K>
void run(double *out, long long *cyc)
{
    for (int warps = 1; warps <= 16; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) k<K><<<1, 32 * warps>>>(out, cyc, 1.1);
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("K=%2d independent chains, %2d warps: %6.2f cycles per warp-DFMA (per warp), %6.2f warp-DFMA / cycle / SM\n", K, warps,
               (double)h / (64.0 * 4 * K), warps * 64.0 * 4 * K / (double)h);
    }
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    run<1>(out, cyc); run<4>(out, cyc); run<16>(out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
