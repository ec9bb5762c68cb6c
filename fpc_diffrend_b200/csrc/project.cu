// Pose + projection: the MVP chain and the clip-space transform, forward and backward
// (replaces reference fit.py:546-553 with camera.rigid_grad camera.py:128-132, roma.unitquat_to_rotmat,
//  and camera.transform_clip camera.py:11-23).
#include "pose.cuh"

namespace {

__global__ void k_pose_mvp_fwd(const float* __restrict__ P, const float* __restrict__ A, const float* __restrict__ t,
                               const float* __restrict__ q, const float* __restrict__ t_cam, const float* __restrict__ q_cam,
                               int F, int C, float* __restrict__ mvp)
{
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= F * C) return;
    int f = gid / C, c = gid - f * C;
    M4 m = frame_camera_mvp(P, A, t, q, t_cam, q_cam, f, c);
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
    pose_backward_frame(P, A, t, q, t_cam, q_cam, d_mvp + (size_t)f * C * 16, f, C, d_t, d_q);
}

// one thread per camera, frames summed in index order (deterministic): gradient of the per-camera pose corrections
__global__ void k_pose_cam_bwd(const float* __restrict__ P, const float* __restrict__ A, const float* __restrict__ t,
                               const float* __restrict__ q, const float* __restrict__ q_cam, const float* __restrict__ d_mvp,
                               int F, int C, float* __restrict__ d_t_cam, float* __restrict__ d_q_cam)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float g[12] = {};
    for (int f = 0; f < F; f++) {
        float o[12];
        cam_pose_backward_one(P, A, t, q, d_mvp + ((size_t)f * C + c) * 16, f, c, o);
#pragma unroll
        for (int i = 0; i < 12; i++) g[i] += o[i];
    }
    pose_backward_finish(q_cam, c, g, d_t_cam, d_q_cam);
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

extern "C" int fpc_pose_cam_bwd(const float* P, const float* A, const float* t, const float* q, const float* t_cam, const float* q_cam,
                                const float* d_mvp, int F, int C, float* d_t_cam, float* d_q_cam, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(P && A && t && q && t_cam && q_cam && d_mvp && d_t_cam && d_q_cam, "pose_cam_bwd: null pointer argument");
    FPC_CHECK_ARG(F > 0 && C > 0, "pose_cam_bwd: F and C must be positive");
    k_pose_cam_bwd<<<fpc_div_up(C, 32), 32, 0, stream>>>(P, A, t, q, q_cam, d_mvp, F, C, d_t_cam, d_q_cam);
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
