// Pose + projection: the MVP chain and the clip-space transform, forward and backward
// (replaces reference fit.py:546-553 with camera.rigid_grad camera.py:128-132, roma.unitquat_to_rotmat,
//  and camera.transform_clip camera.py:11-23).
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

__global__ void k_pose_mvp_fwd(const float* __restrict__ P, const float* __restrict__ A, const float* __restrict__ t,
                               const float* __restrict__ q, const float* __restrict__ t_cam, const float* __restrict__ q_cam,
                               int F, int C, float* __restrict__ mvp)
{
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= F * C) return;
    int f = gid / C, c = gid - f * C;
    float tf[3] = {t[3 * f], t[3 * f + 1], t[3 * f + 2]};
    float qf[4] = {q[4 * f], q[4 * f + 1], q[4 * f + 2], q[4 * f + 3]};
    // same association as the reference: P @ (T_frame @ (T_cam @ A))
    M4 m = mul(load_m4(P + 16 * c), mul(rigid(tf, qf), cam_base(A, t_cam, q_cam, c)));
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) mvp[(size_t)gid * 16 + 4 * i + j] = m.m[i][j];
}

// one thread per frame, cameras summed in index order (deterministic)
__global__ void k_pose_mvp_bwd(const float* __restrict__ P, const float* __restrict__ A, const float* __restrict__ t,
                               const float* __restrict__ q, const float* __restrict__ t_cam, const float* __restrict__ q_cam,
                               const float* __restrict__ d_mvp, int F, int C, float* __restrict__ d_t, float* __restrict__ d_q)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float x = q[4 * f], y = q[4 * f + 1], z = q[4 * f + 2], w = q[4 * f + 3];
    float gt[3] = {0.f, 0.f, 0.f};
    float gR[3][3] = {};
    for (int c = 0; c < C; c++) {
        M4 p = load_m4(P + 16 * c);
        M4 b = cam_base(A, t_cam, q_cam, c);
        M4 g = load_m4(d_mvp + ((size_t)f * C + c) * 16);
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
                gR[i][j] += s;
            }
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) s += pg.m[i][k] * b.m[3][k];
            gt[i] += s;
        }
    }
    d_t[3 * f] = gt[0]; d_t[3 * f + 1] = gt[1]; d_t[3 * f + 2] = gt[2];
    d_q[4 * f + 0] = 2.f * (gR[0][0] * x + gR[0][1] * y + gR[0][2] * z + gR[1][0] * y - gR[1][1] * x - gR[1][2] * w + gR[2][0] * z + gR[2][1] * w - gR[2][2] * x);
    d_q[4 * f + 1] = 2.f * (-gR[0][0] * y + gR[0][1] * x + gR[0][2] * w + gR[1][0] * x + gR[1][1] * y + gR[1][2] * z - gR[2][0] * w + gR[2][1] * z - gR[2][2] * y);
    d_q[4 * f + 2] = 2.f * (-gR[0][0] * z - gR[0][1] * w + gR[0][2] * x + gR[1][0] * w - gR[1][1] * z + gR[1][2] * y + gR[2][0] * x + gR[2][1] * y + gR[2][2] * z);
    d_q[4 * f + 3] = 2.f * (gR[0][0] * w - gR[0][1] * z + gR[0][2] * y + gR[1][0] * z + gR[1][1] * w - gR[1][2] * x - gR[2][0] * y + gR[2][1] * x + gR[2][2] * w);
}

// pos_clip[f*C+c, v, :] = mvp[f*C+c] @ (verts[f,v], 1)
__global__ void __launch_bounds__(256) k_project_fwd(const float* __restrict__ verts, const float* __restrict__ mvp,
                                                     int F, int C, int V, float* __restrict__ pos_clip)
{
    extern __shared__ float sm[];      // [C][16]
    int f = blockIdx.y;
    for (int i = threadIdx.x; i < C * 16; i += blockDim.x) sm[i] = mvp[(size_t)f * C * 16 + i];
    __syncthreads();
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float* p = verts + ((size_t)f * V + v) * 3;
    float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    for (int c = 0; c < C; c++) {
        const float* m = sm + 16 * c;
        float4 o;
        // posw @ mvp^T with a homogeneous 1, accumulated in k order like a row-vector x matrix product
        o.x = m[0] * x + m[1] * y + m[2] * z + m[3];
        o.y = m[4] * x + m[5] * y + m[6] * z + m[7];
        o.z = m[8] * x + m[9] * y + m[10] * z + m[11];
        o.w = m[12] * x + m[13] * y + m[14] * z + m[15];
        reinterpret_cast<float4*>(pos_clip)[((size_t)f * C + c) * V + v] = o;
    }
}

// d_verts[f,v] = sum_c mvp[:, :3]^T g ;  partial d_mvp[blk, f*C+c, i, j] = sum_{v in blk} g_i * (x,y,z,1)_j
__global__ void __launch_bounds__(256) k_project_bwd(const float* __restrict__ verts, const float* __restrict__ mvp,
                                                     const float* __restrict__ d_pos_clip, int F, int C, int V,
                                                     float* __restrict__ d_verts, float* __restrict__ partial)
{
    extern __shared__ float sm[];      // [C][16] mvp, then [8 warps][16] reduction buffer
    float* red = sm + C * 16;
    int f = blockIdx.y;
    for (int i = threadIdx.x; i < C * 16; i += blockDim.x) sm[i] = mvp[(size_t)f * C * 16 + i];
    __syncthreads();
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = v < V;
    float vh[4] = {0.f, 0.f, 0.f, live ? 1.f : 0.f};
    if (live) {
        const float* p = verts + ((size_t)f * V + v) * 3;
        vh[0] = __ldg(p); vh[1] = __ldg(p + 1); vh[2] = __ldg(p + 2);
    }
    float gx = 0.f, gy = 0.f, gz = 0.f;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < C; c++) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) g = ldg4(d_pos_clip + (((size_t)f * C + c) * V + v) * 4);
        const float* m = sm + 16 * c;
        gx += m[0] * g.x + m[4] * g.y + m[8] * g.z + m[12] * g.w;
        gy += m[1] * g.x + m[5] * g.y + m[9] * g.z + m[13] * g.w;
        gz += m[2] * g.x + m[6] * g.y + m[10] * g.z + m[14] * g.w;
        float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float s = gg[i] * vh[j];
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) red[warp * 16 + 4 * i + j] = s;
            }
        __syncthreads();
        if (threadIdx.x < 16) {
            float s = 0.f;
            for (int k = 0; k < 8; k++) s += red[k * 16 + threadIdx.x];
            partial[((size_t)blockIdx.x * F * C + (size_t)f * C + c) * 16 + threadIdx.x] = s;
        }
        __syncthreads();
    }
    if (live) {
        float* o = d_verts + ((size_t)f * V + v) * 3;
        o[0] = gx; o[1] = gy; o[2] = gz;
    }
}

__global__ void k_project_bwd_reduce(const float* __restrict__ partial, int nblk, int n16, float* __restrict__ d_mvp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    float s = 0.f;
    for (int k = 0; k < nblk; k++) s += partial[(size_t)k * n16 + i];
    d_mvp[i] = s;
}

}  // namespace

extern "C" int fpc_pose_mvp_fwd(const float* P, const float* A, const float* t, const float* q,
                                const float* t_cam, const float* q_cam, int F, int C, float* mvp, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(P && A && t && q && mvp, "pose_mvp_fwd: null pointer argument");
    FPC_CHECK_ARG((t_cam == nullptr) == (q_cam == nullptr), "pose_mvp_fwd: t_cam and q_cam must both be given or both be NULL");
    FPC_CHECK_ARG(F > 0 && C > 0, "pose_mvp_fwd: F and C must be positive");
    k_pose_mvp_fwd<<<fpc_div_up((long long)F * C, 128), 128, 0, stream>>>(P, A, t, q, t_cam, q_cam, F, C, mvp);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_pose_mvp_bwd(const float* P, const float* A, const float* t, const float* q,
                                const float* t_cam, const float* q_cam, const float* d_mvp, int F, int C,
                                float* d_t, float* d_q, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(P && A && t && q && d_mvp && d_t && d_q, "pose_mvp_bwd: null pointer argument");
    FPC_CHECK_ARG((t_cam == nullptr) == (q_cam == nullptr), "pose_mvp_bwd: t_cam and q_cam must both be given or both be NULL");
    FPC_CHECK_ARG(F > 0 && C > 0, "pose_mvp_bwd: F and C must be positive");
    k_pose_mvp_bwd<<<fpc_div_up(F, 64), 64, 0, stream>>>(P, A, t, q, t_cam, q_cam, d_mvp, F, C, d_t, d_q);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_project_fwd(const float* verts, const float* mvp, int F, int C, int V, float* pos_clip, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(verts && mvp && pos_clip, "project_fwd: null pointer argument");
    FPC_CHECK_ARG(F > 0 && F <= 65535 && C > 0 && C <= 512 && V > 0, "project_fwd: need 0 < F <= 65535, 0 < C <= 512, V > 0");
    k_project_fwd<<<dim3(fpc_div_up(V, 256), F), 256, (size_t)C * 16 * sizeof(float), stream>>>(verts, mvp, F, C, V, pos_clip);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" size_t fpc_project_bwd_scratch_bytes(int F, int C, int V)
{
    if (F <= 0 || C <= 0 || V <= 0) return 256;
    return (size_t)fpc_div_up(V, 256) * F * C * 16 * sizeof(float) + 256;
}

extern "C" int fpc_project_bwd(const float* verts, const float* mvp, const float* d_pos_clip, int F, int C, int V,
                               float* d_verts, float* d_mvp, void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(verts && mvp && d_pos_clip && d_verts && d_mvp, "project_bwd: null pointer argument");
    FPC_CHECK_ARG(F > 0 && F <= 65535 && C > 0 && C <= 512 && V > 0, "project_bwd: need 0 < F <= 65535, 0 < C <= 512, V > 0");
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_project_bwd_scratch_bytes(F, C, V), "project_bwd: scratch too small");
    int nblk = fpc_div_up(V, 256);
    k_project_bwd<<<dim3(nblk, F), 256, (size_t)(C * 16 + 8 * 16) * sizeof(float), stream>>>(verts, mvp, d_pos_clip, F, C, V, d_verts, (float*)scratch);
    FPC_LAUNCH_CHECK();
    int n16 = F * C * 16;
    k_project_bwd_reduce<<<fpc_div_up(n16, 256), 256, 0, stream>>>((const float*)scratch, nblk, n16, d_mvp);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
