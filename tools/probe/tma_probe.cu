// tma_probe.cu -- micro-tests of the mbarrier / TMA plumbing used by klt.cu, one test per process invocation
// (a faulting kernel poisons the context).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
// Usage: ./tma_probe <test-id>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t phase)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ int mbar_wait(uint32_t bar, uint32_t &phase, int *timeout_flag)
{
    int spins = 0;
    while (!mbar_try_wait(bar, phase)) { if (++spins > (1 << 20)) { if (timeout_flag) *timeout_flag = 1; return spins; } }
    phase ^= 1u;
    return spins;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Out { int timeout; int spins; unsigned checksum; unsigned first; };

// test 1: mbarrier alone
__global__ void k_mbar(Out *o)
{
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = smem_addr(&bar);
    if (threadIdx.x == 0) { mbar_init(b, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (threadIdx.x == 0) mbar_arrive(b);
    uint32_t ph = 0;
    int to = 0;
    const int s = mbar_wait(b, ph, &to);
    if (threadIdx.x == 0) { o->timeout = to; o->spins = s; o->checksum = 1; }
}

// tests 2-6: TMA box load; per-warp staging like klt.cu (4 warps, dynamic smem manually aligned to 128 B)
template <int RANK>
__global__ void k_tma(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int box_bytes, Out *o)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bars[4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint8_t *base = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
    uint8_t *wbuf = base + (size_t)wib * 8192;
    const uint32_t b = smem_addr(bars + wib);
    if (lane == 0) { mbar_init(b, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(b, box_bytes);
        if (RANK == 2) tma_load_2d(smem_addr(wbuf), &tm, b, c0 + wib, c1);
        else tma_load_3d(smem_addr(wbuf), &tm, b, c0 + wib, c1, c2);
    }
    uint32_t ph = 0;
    int to = 0;
    const int s = mbar_wait(b, ph, &to);
    unsigned cs = 0;
    if (!to) for (int i = lane; i < box_bytes; i += 32) cs += wbuf[i] * (unsigned)(i + 1);
    for (int off = 16; off > 0; off >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, off);
    if (lane == 0) { o[wib].timeout = to; o[wib].spins = s; o[wib].checksum = cs; o[wib].first = to ? 0u : *(unsigned *)wbuf; }
}

// test 8/9: descriptor in global memory (without / with the proxy fence)
__global__ void k_tma_gmem(const CUtensorMap *tm, int fence, int c0, int c1, int box_bytes, Out *o)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    uint8_t *base = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
    const uint32_t b = smem_addr(&bar);
    if (threadIdx.x == 0) { mbar_init(b, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (threadIdx.x == 0) {
        if (fence) asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"((unsigned long long)tm) : "memory");
        mbar_expect_tx(b, box_bytes);
        tma_load_2d(smem_addr(base), tm, b, c0, c1);
    }
    uint32_t ph = 0;
    int to = 0;
    const int s = mbar_wait(b, ph, &to);
    unsigned cs = 0;
    if (!to) for (int i = threadIdx.x; i < box_bytes; i += 32) cs += base[i] * (unsigned)(i + 1);
    for (int off = 16; off > 0; off >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, off);
    if (threadIdx.x == 0) { o->timeout = to; o->spins = s; o->checksum = cs; o->first = to ? 0u : *(unsigned *)base; }
}

__global__ void k_trap() { if (threadIdx.x == 0) __trap(); }

typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("  CUDA error at %s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)

int main(int argc, char **argv)
{
    const int test = argc > 1 ? atoi(argv[1]) : 1;
    printf("test %d: ", test);
    fflush(stdout);
    Out *o_d; Out o_h[4];
    CK(cudaMalloc(&o_d, sizeof(o_h)));
    CK(cudaMemset(o_d, 0xff, sizeof(o_h)));
    if (test == 1) {
        k_mbar<<<1, 32>>>(o_d);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(o_h, o_d, sizeof(Out), cudaMemcpyDeviceToHost));
        printf("mbarrier alone: timeout %d spins %d -> %s\n", o_h[0].timeout, o_h[0].spins, o_h[0].timeout == 0 ? "PASS" : "FAIL");
        return 0;
    }
    if (test == 7) {
        k_trap<<<1, 32>>>();
        cudaError_t e = cudaDeviceSynchronize();
        printf("__trap() is reported as: %s\n", cudaGetErrorString(e));
        return 0;
    }
    // tensor: 3 "slots" of a 256-pitch x 100-row plane; u8 value = (x + 3 y + 7 slot) & 255; u32 view for test 6
    const int pitch = 256, rows = 100, nslot = 3;
    const size_t slot_stride = (size_t)pitch * rows * 4;     // room for the u32 view too
    std::vector<uint8_t> h(slot_stride * nslot);
    for (int s = 0; s < nslot; ++s)
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < pitch * 4; ++x) h[s * slot_stride + (size_t)y * pitch * 4 + x] = (uint8_t)(x + 3 * y + 7 * s);
    uint8_t *g;
    CK(cudaMalloc(&g, h.size()));
    CK(cudaMemcpy(g, h.data(), h.size(), cudaMemcpyHostToDevice));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    if (!fp || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    PFN_enc enc = (PFN_enc)fp;
    CUtensorMap tm;
    const cuuint32_t estr[3] = {1, 1, 1};
    int rank = 2, c0 = 0, c1 = 0, c2 = 0, box_bytes = 0;
    CUresult r;
    // row pitch in BYTES of the u8 view is pitch*4 (the plane is pitch*4 bytes wide)
    if (test == 2 || test == 3 || test == 4 || test == 8 || test == 9) {
        const cuuint64_t dims[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)rows}, str[1] = {(cuuint64_t)pitch * 4};
        const cuuint32_t box[2] = {48, 32};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, g, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        box_bytes = 48 * 32;
        if (test == 3) { c0 = 5; c1 = 3; }
        if (test == 4) { c0 = -3; c1 = -2; }
        if (test >= 8) { c0 = 5; c1 = 3; }
    } else if (test == 5) {
        rank = 3;
        const cuuint64_t dims[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)rows, (cuuint64_t)nslot}, str[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)slot_stride};
        const cuuint32_t box[3] = {48, 32, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, g, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        box_bytes = 48 * 32; c0 = 5; c1 = 3; c2 = 2;
    } else if (test == 6) {
        rank = 3;
        const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)nslot}, str[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)slot_stride};
        const cuuint32_t box[3] = {28, 22, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, g, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        box_bytes = 28 * 22 * 4; c0 = 5; c1 = 3; c2 = 1;
    } else { printf("unknown test\n"); return 2; }
    if (r != CUDA_SUCCESS) { printf("encode failed: %d\n", (int)r); return 2; }
    if (test == 8 || test == 9) {
        CUtensorMap *tm_d;
        CK(cudaMalloc(&tm_d, sizeof(tm)));
        CK(cudaMemcpy(tm_d, &tm, sizeof(tm), cudaMemcpyHostToDevice));
        k_tma_gmem<<<1, 32, 8192 + 128>>>(tm_d, test == 9, c0, c1, box_bytes, o_d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("descriptor in global memory (%s fence): kernel error: %s\n", test == 9 ? "with" : "no", cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(o_h, o_d, sizeof(Out), cudaMemcpyDeviceToHost));
        printf("descriptor in global memory (%s fence): timeout %d spins %d first %08x\n", test == 9 ? "with" : "no", o_h[0].timeout, o_h[0].spins, o_h[0].first);
        return 0;
    }
    if (rank == 2) k_tma<2><<<1, 128, 4 * 8192 + 128>>>(tm, c0, c1, c2, box_bytes, o_d);
    else k_tma<3><<<1, 128, 4 * 8192 + 128>>>(tm, c0, c1, c2, box_bytes, o_d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(o_h, o_d, sizeof(o_h), cudaMemcpyDeviceToHost));
    // expected first word of warp w: bytes at (c0 + w .. +3, c1, c2) of the u8 pattern (u32 view: element (c0+w, c1))
    bool ok = true;
    for (int w = 0; w < 4; ++w) {
        unsigned exp = 0;
        const int x0 = (test == 6 ? 4 * (c0 + w) : c0 + w);
        for (int k = 0; k < 4; ++k) {
            const int x = x0 + k;
            const uint8_t v = (x < 0 || c1 < 0) ? 0 : (uint8_t)(x + 3 * c1 + 7 * c2);
            exp |= (unsigned)v << (8 * k);
        }
        printf("[w%d timeout %d spins %d first %08x expect %08x] ", w, o_h[w].timeout, o_h[w].spins, o_h[w].first, exp);
        ok &= o_h[w].timeout == 0 && o_h[w].first == exp;
    }
    printf("-> %s\n", ok ? "PASS" : "FAIL");
    return ok ? 0 : 1;
}
