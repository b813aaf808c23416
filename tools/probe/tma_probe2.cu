// tma_probe2.cu -- the CUDA programming guide's own TMA example (libcu++ wrappers), a 1-D bulk copy, and expect_tx alone.
// Usage: ./tma_probe2 <test-id>   (one test per process)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

constexpr int SM_W = 64, SM_H = 16;     // box: 64 ints x 16 rows

// test 1: programming-guide example, verbatim structure
__global__ void k_guide(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) int smem_buffer[SM_H][SM_W];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    if (threadIdx.x == 0) { out[0] = smem_buffer[0][0]; out[1] = smem_buffer[1][2]; }
}

// test 2: 1-D bulk copy (cp.async.bulk, no tensor map)
__global__ void k_bulk1d(const int *src, int *out)
{
    __shared__ alignas(128) int buf[256];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cuda::memcpy_async(buf, src, cuda::aligned_size_t<16>(sizeof(buf)), bar);
        token = bar.arrive();
    } else token = bar.arrive();
    bar.wait(std::move(token));
    if (threadIdx.x == 0) { out[0] = buf[0]; out[1] = buf[255]; }
}

// test 3: expect_tx + manual complete_tx (no copy engine involved)
__global__ void k_expect(int *out)
{
    __shared__ alignas(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 64;" ::"r"(b) : "memory");
        asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], 64;" ::"r"(b) : "memory");
        unsigned ok = 0;
        int spins = 0;
        while (!ok && spins < 100000) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b) : "memory");
            ++spins;
        }
        out[0] = (int)ok; out[1] = spins;
    }
}

typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("  CUDA error at %s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)

int main(int argc, char **argv)
{
    const int test = argc > 1 ? atoi(argv[1]) : 1;
    printf("probe2 test %d: ", test);
    fflush(stdout);
    int *out_d, out_h[2] = {-1, -1};
    CK(cudaMalloc(&out_d, 8));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    if (test == 0) {
        int drv = 0, rt = 0;
        cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
        printf("%s cc %d.%d driver %d runtime %d MIG? multiGpuBoard %d\n", pr.name, pr.major, pr.minor, drv, rt, pr.isMultiGpuBoard);
        return 0;
    }
    const int GW = 256, GH = 64;
    std::vector<int> h(GW * GH);
    for (int i = 0; i < GW * GH; ++i) h[i] = i;
    int *g;
    CK(cudaMalloc(&g, h.size() * 4));
    CK(cudaMemcpy(g, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    if (test == 1 || test == 4) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (test == 1) CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
        else CK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fp, 12000, cudaEnableDefault, &q));
        CUtensorMap tm;
        const cuuint64_t size[2] = {GW, GH}, stride[1] = {GW * sizeof(int)};
        const cuuint32_t box[2] = {SM_W, SM_H}, es[2] = {1, 1};
        CUresult r = ((PFN_enc)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, g, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
        k_guide<<<1, 128>>>(tm, 64, 16, out_d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("programming-guide TMA example: kernel error: %s\n", cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(out_h, out_d, 8, cudaMemcpyDeviceToHost));
        printf("programming-guide TMA example: got %d %d expect %d %d\n", out_h[0], out_h[1], 16 * GW + 64, 17 * GW + 66);
        return 0;
    }
    if (test == 2) {
        k_bulk1d<<<1, 128>>>(g, out_d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("1-D cp.async.bulk: kernel error: %s\n", cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(out_h, out_d, 8, cudaMemcpyDeviceToHost));
        printf("1-D cp.async.bulk: got %d %d expect 0 255\n", out_h[0], out_h[1]);
        return 0;
    }
    if (test == 3) {
        k_expect<<<1, 32>>>(out_d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("expect_tx/complete_tx: kernel error: %s\n", cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(out_h, out_d, 8, cudaMemcpyDeviceToHost));
        printf("expect_tx/complete_tx: ok %d spins %d\n", out_h[0], out_h[1]);
        return 0;
    }
    return 0;
}
