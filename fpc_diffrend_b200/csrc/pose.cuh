// Pose math shared by project.cu and geometry.cu: 4x4 helpers, the rigid transform of an XYZW quaternion
// (roma.unitquat_to_rotmat semantics), the MVP chain of reference fit.py:546-553 and its backward.
#pragma once
#include "common.cuh"

namespace {

struct M4 { float m[4][4]; };

__device__ __forceinline__ M4 load_m4(const float* p)
{
    M4 r;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) r.m[i][j] = __ldg(p + 4 * i + j);
    return r;
}

__device__ __forceinline__ M4 mul(const M4& a, const M4& b)
{
    M4 r;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}

// [R(q) t; 0 0 0 1], q = XYZW, not normalised (roma.unitquat_to_rotmat semantics)
__device__ __forceinline__ M4 rigid(const float* t, const float* q)
{
    float x = q[0], y = q[1], z = q[2], w = q[3];
    M4 r;
    r.m[0][0] = x * x - y * y - z * z + w * w; r.m[0][1] = 2.f * (x * y - z * w); r.m[0][2] = 2.f * (x * z + y * w); r.m[0][3] = t[0];
    r.m[1][0] = 2.f * (x * y + z * w); r.m[1][1] = -x * x + y * y - z * z + w * w; r.m[1][2] = 2.f * (y * z - x * w); r.m[1][3] = t[1];
    r.m[2][0] = 2.f * (x * z - y * w); r.m[2][1] = 2.f * (y * z + x * w); r.m[2][2] = -x * x - y * y + z * z + w * w; r.m[2][3] = t[2];
    r.m[3][0] = 0.f; r.m[3][1] = 0.f; r.m[3][2] = 0.f; r.m[3][3] = 1.f;
    return r;
}

__device__ __forceinline__ M4 cam_base(const float* A, const float* t_cam, const float* q_cam, int c)
{
    M4 a = load_m4(A + 16 * c);
    if (t_cam && q_cam) {
        float t[3] = {t_cam[3 * c], t_cam[3 * c + 1], t_cam[3 * c + 2]};
        float q[4] = {q_cam[4 * c], q_cam[4 * c + 1], q_cam[4 * c + 2], q_cam[4 * c + 3]};
        a = mul(rigid(t, q), a);
    }
    return a;
}

// mvp of (frame f, camera c): same association as the reference, P @ (T_frame @ (T_cam @ A))
__device__ __forceinline__ M4 frame_camera_mvp(const float* P, const float* A, const float* t, const float* q,
                                               const float* t_cam, const float* q_cam, int f, int c)
{
    float tf[3] = {t[3 * f], t[3 * f + 1], t[3 * f + 2]};
    float qf[4] = {q[4 * f], q[4 * f + 1], q[4 * f + 2], q[4 * f + 3]};
    return mul(load_m4(P + 16 * c), mul(rigid(tf, qf), cam_base(A, t_cam, q_cam, c)));
}

// One camera's contribution to the gradient of the frame's rigid transform: out[0..8] = d R (row-major), out[9..11] = d t.
// d_mvp_c = 16 floats (any address space).
__device__ __forceinline__ void pose_backward_camera(const float* P, const float* A, const float* t_cam, const float* q_cam,
                                                     const float* d_mvp_c, int c, float* out)
{
    M4 p = load_m4(P + 16 * c);
    M4 b = cam_base(A, t_cam, q_cam, c);
    M4 g;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) g.m[i][j] = d_mvp_c[4 * i + j];
    // dRig = P^T g B^T ; only the top 3 rows are needed
    M4 pg;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) s += p.m[k][i] * g.m[k][j];
            pg.m[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) s += pg.m[i][k] * b.m[j][k];
            out[3 * i + j] = s;
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) s += pg.m[i][k] * b.m[3][k];
        out[9 + i] = s;
    }
}

// Contribution of (frame f, camera c) to the gradient of the CAMERA's rigid correction Rigid(t_cam_c, q_cam_c)
// (fit.py:443-448 t_opt / q_opt):  mvp = P Rf Rc A  ->  d Rc = (P Rf)^T d_mvp A^T.  out[0..8] = d R, out[9..11] = d t.
__device__ __forceinline__ void cam_pose_backward_one(const float* P, const float* A, const float* t, const float* q,
                                                      const float* d_mvp_fc, int f, int c, float* out)
{
    float tf[3] = {t[3 * f], t[3 * f + 1], t[3 * f + 2]};
    float qf[4] = {q[4 * f], q[4 * f + 1], q[4 * f + 2], q[4 * f + 3]};
    M4 pr = mul(load_m4(P + 16 * c), rigid(tf, qf));
    M4 a = load_m4(A + 16 * c);
    M4 g;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) g.m[i][j] = d_mvp_fc[4 * i + j];
    M4 pg;      // (P Rf)^T g
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) s += pr.m[k][i] * g.m[k][j];
            pg.m[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) s += pg.m[i][k] * a.m[j][k];
            out[3 * i + j] = s;
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) s += pg.m[i][k] * a.m[3][k];
        out[9 + i] = s;
    }
}

// summed contributions (d R [9], d t [3]) -> d_t[f], d_q[f]
__device__ __forceinline__ void pose_backward_finish(const float* q, int f, const float* g, float* d_t, float* d_q)
{
    float x = q[4 * f], y = q[4 * f + 1], z = q[4 * f + 2], w = q[4 * f + 3];
    const float (*gR)[3] = reinterpret_cast<const float (*)[3]>(g);
    d_t[3 * f] = g[9]; d_t[3 * f + 1] = g[10]; d_t[3 * f + 2] = g[11];
    d_q[4 * f + 0] = 2.f * (gR[0][0] * x + gR[0][1] * y + gR[0][2] * z + gR[1][0] * y - gR[1][1] * x - gR[1][2] * w + gR[2][0] * z + gR[2][1] * w - gR[2][2] * x);
    d_q[4 * f + 1] = 2.f * (-gR[0][0] * y + gR[0][1] * x + gR[0][2] * w + gR[1][0] * x + gR[1][1] * y + gR[1][2] * z - gR[2][0] * w + gR[2][1] * z - gR[2][2] * y);
    d_q[4 * f + 2] = 2.f * (-gR[0][0] * z - gR[0][1] * w + gR[0][2] * x + gR[1][0] * w - gR[1][1] * z + gR[1][2] * y + gR[2][0] * x + gR[2][1] * y + gR[2][2] * z);
    d_q[4 * f + 3] = 2.f * (gR[0][0] * w - gR[0][1] * z + gR[0][2] * y + gR[1][0] * z + gR[1][1] * w - gR[1][2] * x - gR[2][0] * y + gR[2][1] * x + gR[2][2] * w);
}

// d_mvp_f [C][16] (any address space) -> d_t[f], d_q[f]; cameras summed in index order (deterministic)
__device__ __forceinline__ void pose_backward_frame(const float* P, const float* A, const float* t, const float* q,
                                                    const float* t_cam, const float* q_cam, const float* d_mvp_f,
                                                    int f, int C, float* d_t, float* d_q)
{
    (void)t;
    float g[12] = {};
    for (int c = 0; c < C; c++) {
        float o[12];
        pose_backward_camera(P, A, t_cam, q_cam, d_mvp_f + 16 * c, c, o);
#pragma unroll
        for (int i = 0; i < 12; i++) g[i] += o[i];
    }
    pose_backward_finish(q, f, g, d_t, d_q);
}

}  // namespace
