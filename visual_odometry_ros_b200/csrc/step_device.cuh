// step_device.cuh -- device helpers shared by the stereo and mono frame steps (include only from translation units
// compiled with -fmad=false).
#pragma once
#include "vo_internal.cuh"

static __device__ __forceinline__ void xform(const float *T, const float *X, float *Y)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) Y[r] = ((T[r * 4 + 0] * X[0] + T[r * 4 + 1] * X[1]) + T[r * 4 + 2] * X[2]) + T[r * 4 + 3];
}

// Single-CTA stable compaction helper: returns, for every thread's element of the current chunk, its
// output position (or -1), and advances the running base in shared memory.
static __device__ __forceinline__ int scan_chunk(bool keep, int *s_warp, int *s_base)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int within = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {
        int v = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        s_warp[lane] = v;
    }
    __syncthreads();
    const int pos = keep ? (*s_base + (wid ? s_warp[wid - 1] : 0) + within) : -1;
    __syncthreads();
    if (tid == 0) *s_base += s_warp[31];
    __syncthreads();
    return pos;
}

static __global__ void __launch_bounds__(1024) k_step_count(const uint8_t *mask, int n, int *out)
{
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < n; i += 1024) c += mask[i] ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) *out = s;
}


// host-side 4x4 float helpers (row-major), in the reference's operation order
static inline void step_inv_se3_f(const float *T, float *O)     // geometry::inverseSE3_f
{
    float R[16];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i * 4 + j] = T[j * 4 + i];
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
        R[i * 4 + 3] = -s;
    }
    R[12] = R[13] = R[14] = 0.f; R[15] = 1.f;
    for (int i = 0; i < 16; ++i) O[i] = R[i];
}
static inline void step_mul4_f(const float *A, const float *B, float *C)
{
    float T[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j]; T[i * 4 + j] = s; }
    for (int i = 0; i < 16; ++i) C[i] = T[i];
}
static inline size_t step_a16(size_t v) { return (v + 15) / 16 * 16; }

int vo_detect_launch_d(vo_ctx *ctx, int slot, const float *occ_d, const int *n_occ_d, int n_occ, int n_bins_u, int n_bins_v,
                       int edge, long long min_score, float *out_d, uint8_t *out_mask_d, int *n_out_d, int max_out);
