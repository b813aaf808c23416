// vo_host_util.h -- small host-side helpers shared by the StereoVO / MonoVO glue (row-major 4x4 float / double algebra in
// the reference's operation order, timing, status -> exception mapping).
#pragma once
#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/vo_b200.h"
#include "vo_shim_types.h"

namespace vo_host {
using Clock = std::chrono::steady_clock;
inline float ms_since(Clock::time_point t0) { return std::chrono::duration<float, std::milli>(Clock::now() - t0).count(); }
static const float D2R = 3.14159265358979323846f / 180.0f;

inline void ident(float *T) { for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f; }
inline void mul4_f(const float *A, const float *B, float *C)
{
    float T[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j]; T[i * 4 + j] = s; }
    memcpy(C, T, sizeof(T));
}
// geometry::inverseSE3_f (core/util/geometry_library.cpp:554-560)
inline void inv_se3_f(const float *T, float *O)
{
    float R[16];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i * 4 + j] = T[j * 4 + i];
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += T[k * 4 + i] * T[k * 4 + 3];
        R[i * 4 + 3] = -s;
    }
    R[12] = R[13] = R[14] = 0.f; R[15] = 1.f;
    memcpy(O, R, sizeof(R));
}
// Eigen::Matrix4f::inverse() (stereo_vo.cpp:643) restated as the adjugate formula (third-party, unpinned)
inline void inv4_f(const float *m, float *out)
{
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    const float id = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * id;
}
inline void mul4_d(const double *A, const double *B, double *C)
{
    double T[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j]; T[i * 4 + j] = s; }
    memcpy(C, T, sizeof(T));
}
inline void rowmajor_to_pose(const float *o, PoseSE3 &T) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T(r, c) = o[r * 4 + c]; }

[[noreturn]] inline void fail(vo_ctx *ctx, int rc)
{
    std::string msg = ctx ? vo_last_error(ctx) : "";
    if (rc == VO_ERR_NAN && !msg.empty()) throw std::runtime_error(msg);            // the reference's own texts
    throw std::runtime_error(std::string("vo_b200: ") + vo_status_string(rc) + (msg.empty() ? "" : " (" + msg + ")"));
}
}  // namespace vo_host
