// Dependent-issue latencies on sm_100a of the instructions on k_lba_solve's per-step critical path (one warp, clock64 around a
// chain of N dependent operations).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 256
__global__ void k(double *out, long long *cyc, double x0, int two_warps)
{
    const int lane = threadIdx.x & 31;
    double x = x0 + lane * 1e-3;
    long long t0, t1;
    int c = 0;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __fma_rn(x, 1.0000001, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // DMUL + DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N / 2; ++i) { x = __dmul_rn(x, 1.0000001); x = __dadd_rn(x, 1e-9); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // __drcp_rn chain
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N / 4; ++i) x = __drcp_rn(x) + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // REDUX chain (u32 max)
    unsigned u = (unsigned)__double_as_longlong(x) ^ lane;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) u = __reduce_max_sync(0xffffffffu, u + lane) ^ 0x5u;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // SHFL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) u = __shfl_sync(0xffffffffu, u, (lane + 1) & 31) + 1u;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // ballot + ffs chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) u += __ffs(__ballot_sync(0xffffffffu, ((u >> (lane & 7)) & 1u) != 0u));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // named barrier over the launched warps
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory");
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // STS -> bar -> LDS round trip (publish / consume), dependent
    __shared__ double sh[256];
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        sh[threadIdx.x] = x;
        asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory");
        x = sh[(threadIdx.x + 1) % blockDim.x] + 1e-9;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // integer ALU dependent chain (IADD3 / LOP3)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) u = (u ^ (u >> 3)) + 0x9e3779b9u;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    // 64-bit select chain (FSEL pairs) on a predicate from DSETP
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = (x > 1.25) ? x - 0.25 : x + 0.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[c] = t1 - t0; ++c;
    out[threadIdx.x] = x + u;
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 16 * 8);
    const char *names[] = {"DFMA", "DMUL+DADD (per op)", "__drcp_rn + DADD (x4 fewer)", "REDUX.max + 2 ALU", "SHFL + IADD", "VOTE + FLO + IADD", "bar.sync named",
                           "STS -> bar -> LDS -> DADD", "2 ALU (shift-xor-add)", "DSETP + 2 DADD + FSEL"};
    for (int warps = 1; warps <= 8; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) k<<<1, 32 * warps>>>(out, cyc, 1.1, warps);
        long long h[16];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("warps %d (err %s)\n", warps, cudaGetErrorString(cudaGetLastError()));
        for (int i = 0; i < 10; ++i) printf("  %-34s %8.1f cycles / iteration\n", names[i], (double)h[i] / (i == 2 ? N / 4 : N));
    }
    return 0;
}
