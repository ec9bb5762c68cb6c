// Fused geometry stages of the fit iteration for small frame batches (one pass over D per direction):
//
//   fpc_geometry_fwd : pose -> MVP chain (fit.py:546-553)  +  blend V = base + D w (fit.py:103-129)
//                      +  clip transform [V 1] mvp^T (camera.py:11-23)                       -> ONE kernel
//   fpc_geometry_bwd : d pos_clip -> d V (transform_clip bwd) -> d w = D^T d V (blend bwd), d mvp -> d t, d q
//                      (pose bwd) [-> Adam step of the packed parameters]                     -> ONE kernel: the per-CTA
//                      partials are combined by the CTAs that arrive last (groups of 8, then the groups, then the frames:
//                      three counters, fixed summation order), which also run the pose backward and the optimiser
//   fpc_adam_fused   : Adam + LambdaLR for the packed [w | t | q] vector, quaternion renorm and the step
//                      counter advance (fit.py:493-505,610-618)                              -> ONE kernel (when not folded)
//
// A warp owns one vertex at a time: its three rows of D are 3B contiguous floats (HBM/L2-bound float4 stream,
// the only large operand), reduced with warp shuffles; lanes 0..C-1 then act as the cameras of that vertex.
// Everything is deterministic: per-warp partials are combined in a fixed order (no float atomics).
#include <stdlib.h>

#include "pose.cuh"
#include "tma.cuh"

namespace {

constexpr int GEO_THREADS = 256;
constexpr int GEO_WARPS = GEO_THREADS / 32;

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
// row loader: the 3B floats of vertex v as K float4 per lane (flat index i = lane + 32 k)
template <int K>
__device__ __forceinline__ void load_rows(const float* __restrict__ D, int v, int B, int n4, int lane, float4 (&d)[K])
{
    const float4* row4 = reinterpret_cast<const float4*>(D + (size_t)v * 3 * B);
#pragma unroll
    for (int k = 0; k < K; k++) {
        int i = lane + 32 * k;
        d[k] = (i < n4) ? __ldg(row4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}


// ---------------------------------------------------------------------------------------------------------
// Adam for the packed parameter vector [w (F*B) | t (F*3) | q (F*4)] by ONE CTA (n is small: F*(B+7)); the body of
// k_adam_fused and of the last CTA of k_geom_bwd
// ---------------------------------------------------------------------------------------------------------
struct AdamArgs {
    float* p; float* m; float* v; float* step_count;      // p == nullptr: no optimiser step
    int pose, quat_mode;
    float lr_w, lr_t, lr_q, b1, b2, eps, lr_ramp, max_iter;
};

__device__ __forceinline__ void adam_packed_cta(const AdamArgs& a, const float* __restrict__ g, int nw, int F, float* red /* [32] shared */)
{
    const float step0 = a.step_count[0];
    const float tstep = step0 + 1.f;
    const float ramp = powf(a.lr_ramp, step0 / a.max_iter);
    const float bc1 = 1.f - powf(a.b1, tstep), bc2 = 1.f - powf(a.b2, tstep);
    const int n = nw + (a.pose ? 7 * F : 0);
    float* p = a.p;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float lr = (i < nw) ? a.lr_w : ((i < nw + 3 * F) ? a.lr_t : a.lr_q);
        float gi = __ldcg(g + i);                           // (written by other CTAs of the same launch in the folded case)
        float mi = a.b1 * a.m[i] + (1.f - a.b1) * gi;
        float vi = a.b2 * a.v[i] + (1.f - a.b2) * gi * gi;
        a.m[i] = mi;
        a.v[i] = vi;
        float denom = sqrtf(vi) / sqrtf(bc2) + a.eps;      // torch.optim.Adam op order (k_adam, loss_adam.cu)
        p[i] -= (lr * ramp / bc1) * (mi / denom);
    }
    __syncthreads();
    if (a.pose) {
        float* qq = p + nw + 3 * F;
        if (a.quat_mode == 0) {
            for (int i = threadIdx.x; i < F; i += blockDim.x) {
                float x = qq[4 * i], y = qq[4 * i + 1], z = qq[4 * i + 2], w = qq[4 * i + 3];
                float s = 1.f / sqrtf(x * x + y * y + z * z + w * w);
                qq[4 * i] = x * s; qq[4 * i + 1] = y * s; qq[4 * i + 2] = z * s; qq[4 * i + 3] = w * s;
            }
        } else {
            float s = 0.f;
            for (int i = threadIdx.x; i < 4 * F; i += blockDim.x) s += qq[i] * qq[i];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
            __syncthreads();
            float tot = 0.f;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++) tot += red[k];
            float inv = 1.f / sqrtf(tot);
            for (int i = threadIdx.x; i < 4 * F; i += blockDim.x) qq[i] *= inv;
        }
    }
    if (threadIdx.x == 0) a.step_count[0] = tstep;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// sum of n values `stride` floats apart IN INDEX ORDER, loaded U at a time (independent L2 reads in flight; written by other CTAs
// of this launch, hence ld.cg)
template <int U>
__device__ __forceinline__ float ordered_sum(const float* __restrict__ src, size_t stride, int n)
{
    float s = 0.f;
    for (int k0 = 0; k0 < n; k0 += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = (k0 + u < n) ? __ldcg(src + (size_t)(k0 + u) * stride) : 0.f;
#pragma unroll
        for (int u = 0; u < U; u++) s += v[u];
    }
    return s;
}

// what the last-arriving CTAs of k_geom_bwd need: second-level partials, the arrival counters (zero on entry, left zero), the
// pose chain, the outputs, and the optional optimiser step
constexpr int GEO_GROUP = 16;         // 296 CTAs -> 19 groups: each level sums its operands with ONE batch of loads in flight
struct GeomTail {
    int* counters;                  // [F * ng] groups | [F] frames | [1] all
    float* part2_w;                 // [ng][F][B]
    float* part2_mvp;               // [ng][F][C*16]
    const float *P, *A, *t, *q, *t_cam, *q_cam;
    float *d_w, *d_mvp, *d_t, *d_q;
    AdamArgs adam;
};

// ---------------------------------------------------------------------------------------------------------
// backward: per-CTA partials of d w [F,B] and d mvp [F*C,16], then the tail (see GeomTail)
// ---------------------------------------------------------------------------------------------------------
template <int K>   // K = ceil(3B/4 / 32): float4 accumulators per lane
__global__ void __launch_bounds__(GEO_THREADS) k_geom_bwd(const float* __restrict__ D, const float* __restrict__ verts,
                                                          const float* __restrict__ mvp, const float* __restrict__ g_pos,
                                                          const float* __restrict__ d_verts_add, int V, int B, int F, int C,
                                                          float* __restrict__ d_verts, float* __restrict__ part_w,
                                                          float* __restrict__ part_mvp, GeomTail tl)
{
    extern __shared__ float sm[];
    __shared__ int s_last;
    __shared__ float s_red[32];
    float* s_mvp = sm;                                  // [C][16]
    float* s_w = s_mvp + C * 16;                        // [GEO_WARPS][3B]
    float* s_m = s_w + GEO_WARPS * 3 * B;               // [GEO_WARPS][C][16]
    const int f = blockIdx.y;
    for (int i = threadIdx.x; i < C * 16; i += GEO_THREADS) s_mvp[i] = mvp[(size_t)f * C * 16 + i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * GEO_WARPS + warp, nw = gridDim.x * GEO_WARPS;
    const int n4 = (3 * B) >> 2;
    float4 acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    float am[16];                                       // d mvp of camera `lane` (cameras beyond 32: see host check)
#pragma unroll
    for (int i = 0; i < 16; i++) am[i] = 0.f;

    // Vertices that no view sees (the back of a head seen from the front: close to half of them) carry a zero gradient, and
    // D^T d V does not need their rows of D: the clip-space gradient of a vertex is fetched TWO vertices ahead, so that by the
    // time the rows of the next vertex would be requested the warp knows whether it needs them.  (With mesh regularisers every
    // vertex has a gradient and nothing is skipped.)
    const bool dense = d_verts_add != nullptr;
    auto load_g = [&](int v) -> float4 {
        return (lane < C && v < V) ? ldg4(g_pos + (((size_t)f * C + lane) * V + v) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto nonzero = [&](const float4& g) -> bool {
        return dense || __any_sync(0xffffffffu, g.x != 0.f || g.y != 0.f || g.z != 0.f || g.w != 0.f);
    };
    float4 g0 = load_g(gw), g1 = load_g(gw + nw);
    bool nz0 = gw < V && nonzero(g0);
    float4 cur[K];
#pragma unroll
    for (int k = 0; k < K; k++) cur[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nz0) load_rows<K>(D, gw, B, n4, lane, cur);
    for (int v = gw; v < V; v += nw) {
        const float4 g2 = load_g(v + 2 * nw);                         // decided on in the next iteration
        const bool nz1 = (v + nw < V) && nonzero(g1);
        float4 nxt[K];
#pragma unroll
        for (int k = 0; k < K; k++) nxt[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nz1) load_rows<K>(D, v + nw, B, n4, lane, nxt);            // in flight while this vertex is processed
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (nz0) {
            if (lane < C) {
                const float4 g = g0;
                const float* m = s_mvp + 16 * lane;
                gx = m[0] * g.x + m[4] * g.y + m[8] * g.z + m[12] * g.w;
                gy = m[1] * g.x + m[5] * g.y + m[9] * g.z + m[13] * g.w;
                gz = m[2] * g.x + m[6] * g.y + m[10] * g.z + m[14] * g.w;
                const float* p = verts + ((size_t)f * V + v) * 3;
                float vh[4] = {__ldg(p), __ldg(p + 1), __ldg(p + 2), 1.f};
                float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) am[4 * i + j] += gg[i] * vh[j];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                gx += __shfl_xor_sync(0xffffffffu, gx, o);
                gy += __shfl_xor_sync(0xffffffffu, gy, o);
                gz += __shfl_xor_sync(0xffffffffu, gz, o);
            }
            if (d_verts_add) {
                const float* e = d_verts_add + ((size_t)f * V + v) * 3;
                gx += __ldg(e); gy += __ldg(e + 1); gz += __ldg(e + 2);
            }
        }
        if (d_verts && lane == 0) {
            float* o = d_verts + ((size_t)f * V + v) * 3;
            o[0] = gx; o[1] = gy; o[2] = gz;
        }
        if (nz0) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                int e = 4 * (lane + 32 * k);
                int r = (e >= B) + (e >= 2 * B);
                float gv = (r == 0) ? gx : ((r == 1) ? gy : gz);       // rows beyond 3B were loaded as zeros
                acc[k].x += cur[k].x * gv; acc[k].y += cur[k].y * gv; acc[k].z += cur[k].z * gv; acc[k].w += cur[k].w * gv;
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) cur[k] = nxt[k];
        g0 = g1; g1 = g2; nz0 = nz1;
    }
    // per-warp partials -> shared memory, then a fixed-order sum over warps (and over the 3 rows of a vertex)
    float4* mine = reinterpret_cast<float4*>(s_w + (size_t)warp * 3 * B);
#pragma unroll
    for (int k = 0; k < K; k++) {
        int i = lane + 32 * k;
        if (i < n4) mine[i] = acc[k];
    }
    if (lane < C) {
#pragma unroll
        for (int i = 0; i < 16; i++) s_m[((size_t)warp * C + lane) * 16 + i] = am[i];
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += GEO_THREADS) {
        float s = 0.f;
        for (int wv = 0; wv < GEO_WARPS; wv++) {
            const float* r = s_w + (size_t)wv * 3 * B;
            s += r[b] + r[B + b] + r[2 * B + b];
        }
        part_w[((size_t)blockIdx.x * F + f) * B + b] = s;
    }
    for (int i = threadIdx.x; i < C * 16; i += GEO_THREADS) {
        float s = 0.f;
        for (int wv = 0; wv < GEO_WARPS; wv++) s += s_m[(size_t)wv * C * 16 + i];
        part_mvp[((size_t)blockIdx.x * F + f) * C * 16 + i] = s;
    }

    // ---- tail.  Level 1: the CTA that arrives last in its group of GEO_GROUP sums the group's partials (in block order) ----
    const int nval = C * 16;
    const int ng = (gridDim.x + GEO_GROUP - 1) / GEO_GROUP, grp = blockIdx.x / GEO_GROUP;
    const int gsz = min(GEO_GROUP, (int)gridDim.x - grp * GEO_GROUP);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(tl.counters + f * ng + grp, 1) == gsz - 1;
    __syncthreads();
    if (!s_last) return;            // (the partials are read with ld.cg below: L2 is the point of coherence, as in CUDA's threadFenceReduction sample)
    for (int i = threadIdx.x; i < B + nval; i += GEO_THREADS) {
        const bool isw = i < B;
        const float* src = isw ? part_w + ((size_t)(grp * GEO_GROUP) * F + f) * B + i : part_mvp + ((size_t)(grp * GEO_GROUP) * F + f) * nval + (i - B);
        const size_t stride = (size_t)F * (isw ? B : nval);
        const float s = ordered_sum<GEO_GROUP>(src, stride, gsz);
        if (isw) tl.part2_w[((size_t)grp * F + f) * B + i] = s;
        else tl.part2_mvp[((size_t)grp * F + f) * nval + (i - B)] = s;
    }
    // ---- level 2: the group that arrives last for this frame sums the groups (in group order), runs the pose backward ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(tl.counters + F * ng + f, 1) == ng - 1;
    __syncthreads();
    if (!s_last) return;
    {
        // what the rest of the tail reads does not depend on the sums: request it now (L1 prefetch), so that the pose backward
        // and the optimiser step do not each wait for their own round trip to L2 after the sums
        if (threadIdx.x < C) {
            prefetch_l1(tl.P + 16 * threadIdx.x); prefetch_l1(tl.P + 16 * threadIdx.x + 8);
            prefetch_l1(tl.A + 16 * threadIdx.x); prefetch_l1(tl.A + 16 * threadIdx.x + 8);
        }
        if (threadIdx.x == 32) prefetch_l1(tl.q + 4 * f);
        if (tl.adam.p && F == 1) {
            const int n = B + 7;
            if (threadIdx.x == 33) prefetch_l1(tl.adam.step_count);
            for (int i = 8 * threadIdx.x; i < n; i += 8 * GEO_THREADS) { prefetch_l1(tl.adam.m + i); prefetch_l1(tl.adam.v + i); prefetch_l1(tl.adam.p + i); }
        }
    }
    float* s_dm = sm;                                   // [nval]   (the main loop's shared memory is free now)
    float* s_cam = sm + nval;                           // [C][12]
    for (int i = threadIdx.x; i < B + nval; i += GEO_THREADS) {
        const bool isw = i < B;
        const float* src = isw ? tl.part2_w + (size_t)f * B + i : tl.part2_mvp + (size_t)f * nval + (i - B);
        const float s = ordered_sum<20>(src, (size_t)F * (isw ? B : nval), ng);
        if (isw) tl.d_w[(size_t)f * B + i] = s;
        else {
            s_dm[i - B] = s;
            if (tl.d_mvp) tl.d_mvp[(size_t)f * nval + (i - B)] = s;
        }
    }
    for (int i = threadIdx.x; i < ng; i += GEO_THREADS) tl.counters[f * ng + i] = 0;        // left as found
    if (threadIdx.x == 0) tl.counters[F * ng + f] = 0;
    __syncthreads();
    if (threadIdx.x < C) pose_backward_camera(tl.P, tl.A, tl.t_cam, tl.q_cam, s_dm + 16 * threadIdx.x, threadIdx.x, s_cam + 12 * threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        float g[12] = {};
        for (int c = 0; c < C; c++)
#pragma unroll
            for (int i = 0; i < 12; i++) g[i] += s_cam[12 * c + i];
        pose_backward_finish(tl.q, f, g, tl.d_t, tl.d_q);
    }
    if (!tl.adam.p) return;
    // ---- level 3: the frame that finishes last steps the optimiser on the packed [w | t | q] vector (d_w is its gradient) ----
    __syncthreads();                                   // (d_t, d_q of this frame are written)
    if (F > 1) {
        __threadfence();
        if (threadIdx.x == 0) s_last = atomicAdd(tl.counters + F * ng + F, 1) == F - 1;
        __syncthreads();
        if (!s_last) return;
        if (threadIdx.x == 0) tl.counters[F * ng + F] = 0;
    }
    adam_packed_cta(tl.adam, tl.d_w, F * B, F, s_red);
}

// ---------------------------------------------------------------------------------------------------------
// TMA variants (the ones the host functions launch): the rows of D are streamed by the copy engine
// (cp.async.bulk, tma.cuh) into a ring of shared-memory stages — one stage = the 3B-float rows of GT_VB = 8 consecutive
// vertices, one vertex per consumer warp — so the bytes in flight per SM are set by the ring depth (>= 2 CTAs x
// stages x 8 x 12 B bytes) and not by how many registers a warp can spare for prefetching.  Warp 8 is the producer.
// The arithmetic of a vertex (order of every sum) is unchanged from the register-prefetch kernels above.
// ---------------------------------------------------------------------------------------------------------
constexpr int GT_VB = 8;                            // vertices per stage = consumer warps
constexpr int GT_THREADS = (GT_VB + 1) * 32;

__device__ __forceinline__ size_t gt_ring_floats(int B, int stages) { return (size_t)stages * GT_VB * 3 * B; }

template <int K>
__device__ __forceinline__ void lds_rows(const float* __restrict__ rows, int n4, int lane, float4 (&d)[K])
{
    const float4* r4 = reinterpret_cast<const float4*>(rows);
#pragma unroll
    for (int k = 0; k < K; k++) {
        int i = lane + 32 * k;
        d[k] = (i < n4) ? r4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int K>
__global__ void __launch_bounds__(GT_THREADS) k_geom_fwd_tma(const float* __restrict__ P, const float* __restrict__ A,
                                                             const float* __restrict__ t, const float* __restrict__ q,
                                                             const float* __restrict__ t_cam, const float* __restrict__ q_cam,
                                                             const float* __restrict__ D, const float* __restrict__ base,
                                                             const float* __restrict__ w, int V, int B, int F, int C, int stages,
                                                             float* __restrict__ mvp_out, float* __restrict__ verts,
                                                             float* __restrict__ pos_clip)
{
    extern __shared__ __align__(128) float sm[];
    const int rowf = 3 * B;
    float* ring = sm;                                   // [stages][GT_VB][3B]
    float* s_mvp = ring + gt_ring_floats(B, stages);    // [C][16]
    float* s_w = s_mvp + C * 16;                        // [B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + B);      // full [stages] | empty [stages]   (8-byte aligned: B % 4 == 0)
    const uint32_t full0 = fpc::smem_u32(bars), empty0 = full0 + 8 * stages;
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunk = (V + GT_VB - 1) / GT_VB;
    const int n4 = rowf >> 2;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { fpc::mbar_init(full0 + 8 * s, 1); fpc::mbar_init(empty0 + 8 * s, GT_VB); }
        fpc::mbar_init_fence();
    }
    __syncthreads();
    if (warp == GT_VB) {
        // ---- producer: D is a constant of the fit, nothing to wait for: the stream starts before the pose chain below ----
        if (lane == 0) {
            int it = 0;
            for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
                const int s = it % stages;
                if (it >= stages) fpc::mbar_wait(empty0 + 8 * s, ((it / stages) & 1) ^ 1);
                const int nv = min(GT_VB, V - c * GT_VB);
                const uint32_t bytes = (uint32_t)nv * rowf * 4u;
                fpc::mbar_expect_tx(full0 + 8 * s, bytes);
                fpc::bulk_load(fpc::smem_u32(ring + (size_t)s * GT_VB * rowf), D + (size_t)c * GT_VB * rowf, bytes, full0 + 8 * s);
            }
        }
        return;
    }
    // ---- consumers (256 threads; they synchronise among themselves on named barrier 1) ----
    for (int c = threadIdx.x; c < C; c += GT_VB * 32) {
        M4 m = frame_camera_mvp(P, A, t, q, t_cam, q_cam, f, c);
#pragma unroll
        for (int i = 0; i < 16; i++) s_mvp[16 * c + i] = m.m[i >> 2][i & 3];
        if (blockIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) mvp_out[((size_t)f * C + c) * 16 + i] = m.m[i >> 2][i & 3];
        }
    }
    for (int i = threadIdx.x; i < B; i += GT_VB * 32) s_w[i] = w[(size_t)f * B + i];
    asm volatile("bar.sync 1, %0;" ::"n"(GT_VB * 32) : "memory");

    // a lane always meets the same columns of D: its activations live in registers for the whole vertex loop
    float4 wr[K];
    int rsel[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        int e = 4 * (lane + 32 * k);
        rsel[k] = (e >= B) + (e >= 2 * B);
        wr[k] = (e < 3 * B) ? *reinterpret_cast<const float4*>(s_w + (e - rsel[k] * B)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int it = 0;
    for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
        const int s = it % stages;
        const int v = c * GT_VB + warp;
        float bx = 0.f, by = 0.f, bz = 0.f;
        if (v < V) { bx = __ldg(base + 3 * (size_t)v); by = __ldg(base + 3 * (size_t)v + 1); bz = __ldg(base + 3 * (size_t)v + 2); }
        fpc::mbar_wait(full0 + 8 * s, (it / stages) & 1);
        if (v < V) {
            float4 cur[K];
            lds_rows<K>(ring + ((size_t)s * GT_VB + warp) * rowf, n4, lane, cur);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int k = 0; k < K; k++) {
                // rows beyond 3B were loaded as zeros (and carry zero activations)
                float sdot = cur[k].x * wr[k].x + cur[k].y * wr[k].y + cur[k].z * wr[k].z + cur[k].w * wr[k].w;
                a0 += (rsel[k] == 0) ? sdot : 0.f;
                a1 += (rsel[k] == 1) ? sdot : 0.f;
                a2 += (rsel[k] == 2) ? sdot : 0.f;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            const float x = bx + a0, y = by + a1, z = bz + a2;
            if (lane == 0) {
                float* o = verts + ((size_t)f * V + v) * 3;
                o[0] = x; o[1] = y; o[2] = z;
            }
            for (int cc = lane; cc < C; cc += 32) {
                const float* m = s_mvp + 16 * cc;
                float4 o;      // same op order as k_project_fwd (project.cu)
                o.x = m[0] * x + m[1] * y + m[2] * z + m[3];
                o.y = m[4] * x + m[5] * y + m[6] * z + m[7];
                o.z = m[8] * x + m[9] * y + m[10] * z + m[11];
                o.w = m[12] * x + m[13] * y + m[14] * z + m[15];
                reinterpret_cast<float4*>(pos_clip)[((size_t)f * C + cc) * V + v] = o;
            }
        }
        __syncwarp();
        if (lane == 0) fpc::mbar_arrive(empty0 + 8 * s);     // this warp's rows of the stage are consumed
    }
}

// Adam as a kernel of its own (fpc_adam_fused): a single CTA
__global__ void __launch_bounds__(1024) k_adam_fused(AdamArgs a, const float* __restrict__ g, int nw, int F)
{
    __shared__ float red[32];
    adam_packed_cta(a, g, nw, F, red);
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int geom_blocks(int V)
{
    int want = fpc_div_up(V, GEO_WARPS);
    return want < 296 ? want : 296;           // 2 CTAs per SM on 148 SMs; each warp strides over the vertices
}

// ring depth / shared memory / grid of the TMA kernels.  Two CTAs per SM while a stage is small enough for that to leave at
// least 3 stages each; one CTA per SM with the whole shared memory otherwise (B > ~380).
struct GtPlan { int stages, blocks; size_t smem; };

GtPlan gt_plan(int V, int B, int F, int C, bool bwd)
{
    const size_t stage = (size_t)GT_VB * 3 * B * sizeof(float);
    const size_t fixed = (bwd ? (size_t)(C * 16 + GT_VB * C * 16) : (size_t)(C * 16 + B)) * sizeof(float) + 8 * 2 * 16 + 4 * 16 + 128;
    GtPlan p;
    int per_sm = 2;
    long long room = 110 * 1024 - (long long)fixed;
    p.stages = (int)(room / (long long)stage);
    if (p.stages < 3) {
        per_sm = 1;
        room = 224 * 1024 - (long long)fixed;
        p.stages = (int)(room / (long long)stage);
    }
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) p.stages = 2;          // B <= 1024 (fused_supported) keeps 2 stages within 227 KB
    p.smem = (size_t)p.stages * stage + fixed;
    const int nchunk = fpc_div_up(V, GT_VB);
    int want = (148 * per_sm) / (F < 1 ? 1 : F);
    if (want < 37) want = 37;
    p.blocks = nchunk < want ? nchunk : want;
    return p;
}

template <int K>
int launch_fwd_tma(const GtPlan& pl, cudaStream_t stream, const float* P, const float* A, const float* t, const float* q, const float* t_cam,
                   const float* q_cam, const float* D, const float* base, const float* w, int V, int B, int F, int C, float* mvp, float* verts,
                   float* pos_clip)
{
    static FpcPerDeviceOnce attr_set;
    if (attr_set.need()) {
        FPC_CUDA(cudaFuncSetAttribute(k_geom_fwd_tma<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set.done();
    }
    k_geom_fwd_tma<K><<<dim3(pl.blocks, F), GT_THREADS, pl.smem, stream>>>(P, A, t, q, t_cam, q_cam, D, base, w, V, B, F, C, pl.stages, mvp, verts, pos_clip);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

template <int K>
int launch_bwd(dim3 grid, size_t smem, cudaStream_t stream, const float* D, const float* verts, const float* mvp, const float* g_pos,
               const float* d_verts_add, int V, int B, int F, int C, float* d_verts, float* part_w, float* part_mvp, const GeomTail& tl)
{
    static FpcPerDeviceOnce attr_set;
    if (attr_set.need()) {
        FPC_CUDA(cudaFuncSetAttribute(k_geom_bwd<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set.done();
    }
    k_geom_bwd<K><<<grid, GEO_THREADS, smem, stream>>>(D, verts, mvp, g_pos, d_verts_add, V, B, F, C, d_verts, part_w, part_mvp, tl);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

}  // namespace

extern "C" int fpc_geometry_fused_supported(int V, int B, int F, int C)
{
    return V > 0 && B > 0 && (B & 3) == 0 && B <= 1024 && F > 0 && F <= 65535 && C > 0 && C <= 32;
}

extern "C" int fpc_geometry_fwd(const float* P, const float* A, const float* t, const float* q, const float* t_cam, const float* q_cam,
                                const float* D, const float* base, const float* w, int V, int B, int F, int C,
                                float* mvp, float* verts, float* pos_clip, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(P && A && t && q && D && base && w && mvp && verts && pos_clip, "geometry_fwd: null pointer argument");
    FPC_CHECK_ARG((t_cam == nullptr) == (q_cam == nullptr), "geometry_fwd: t_cam and q_cam must both be given or both be NULL");
    FPC_CHECK_ARG(fpc_geometry_fused_supported(V, B, F, C),
                  "geometry_fwd: needs B %% 4 == 0, B <= 1024, F <= 65535, C <= 32 (got V=%d B=%d F=%d C=%d); use blend_fwd + project_fwd", V, B, F, C);
    const GtPlan pl = gt_plan(V, B, F, C, false);
    const int K = fpc_div_up((3 * B) >> 2, 32);
    int st = FPC_OK;
#define FPC_FWD_CASE(k) case k: st = launch_fwd_tma<k>(pl, stream, P, A, t, q, t_cam, q_cam, D, base, w, V, B, F, C, mvp, verts, pos_clip); break
    switch (K <= 8 ? K : (K <= 10 ? 10 : (K <= 12 ? 12 : (K <= 16 ? 16 : 24)))) {
        FPC_FWD_CASE(1); FPC_FWD_CASE(2); FPC_FWD_CASE(3); FPC_FWD_CASE(4); FPC_FWD_CASE(5); FPC_FWD_CASE(6);
        FPC_FWD_CASE(7); FPC_FWD_CASE(8); FPC_FWD_CASE(10); FPC_FWD_CASE(12); FPC_FWD_CASE(16); FPC_FWD_CASE(24);
    }
#undef FPC_FWD_CASE
    return st;
}

extern "C" size_t fpc_geometry_bwd_scratch_bytes(int V, int B, int F, int C)
{
    if (V <= 0 || B <= 0 || F <= 0 || C <= 0) return 256;
    const size_t nblk = geom_blocks(V), ng = fpc_div_up((int)nblk, GEO_GROUP);
    return align256(nblk * F * B * sizeof(float)) + align256(nblk * F * C * 16 * sizeof(float)) +
           align256(ng * F * B * sizeof(float)) + align256(ng * F * C * 16 * sizeof(float));
}

extern "C" size_t fpc_geometry_bwd_counter_bytes(int V, int F)
{
    if (V <= 0 || F <= 0) return 256;
    return align256(((size_t)F * fpc_div_up(geom_blocks(V), GEO_GROUP) + F + 1) * sizeof(int));
}

extern "C" int fpc_geometry_bwd(const float* P, const float* A, const float* t, const float* q, const float* t_cam, const float* q_cam,
                                const float* D, const float* verts, const float* mvp, const float* g_pos, const float* d_verts_add,
                                int V, int B, int F, int C, float* d_w, float* d_t, float* d_q, float* d_verts, float* d_mvp,
                                int32_t* counters, const fpc_adam_fused_args* adam,
                                void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(P && A && t && q && D && verts && mvp && g_pos && d_w && d_t && d_q && counters, "geometry_bwd: null pointer argument");
    FPC_CHECK_ARG((t_cam == nullptr) == (q_cam == nullptr), "geometry_bwd: t_cam and q_cam must both be given or both be NULL");
    FPC_CHECK_ARG(fpc_geometry_fused_supported(V, B, F, C),
                  "geometry_bwd: needs B %% 4 == 0, B <= 1024, F <= 65535, C <= 32 (got V=%d B=%d F=%d C=%d); use project_bwd + blend_bwd + pose_mvp_bwd", V, B, F, C);
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_geometry_bwd_scratch_bytes(V, B, F, C), "geometry_bwd: scratch too small");
    const int nblk = geom_blocks(V), ng = fpc_div_up(nblk, GEO_GROUP);
    char* sp = (char*)scratch;
    float* part_w = (float*)sp;      sp += align256((size_t)nblk * F * B * sizeof(float));
    float* part_mvp = (float*)sp;    sp += align256((size_t)nblk * F * C * 16 * sizeof(float));
    GeomTail tl;
    tl.part2_w = (float*)sp;         sp += align256((size_t)ng * F * B * sizeof(float));
    tl.part2_mvp = (float*)sp;
    tl.counters = counters;
    tl.P = P; tl.A = A; tl.t = t; tl.q = q; tl.t_cam = t_cam; tl.q_cam = q_cam;
    tl.d_w = d_w; tl.d_mvp = d_mvp; tl.d_t = d_t; tl.d_q = d_q;
    tl.adam = AdamArgs{};
    if (adam) {
        // the optimiser step rides in the last CTA: the gradient must be the packed vector [d_w | d_t | d_q] the step reads
        FPC_CHECK_ARG(adam->params && adam->m && adam->v && adam->step_count, "geometry_bwd: adam: null pointer member");
        FPC_CHECK_ARG(d_t == d_w + (size_t)F * B && d_q == d_t + (size_t)F * 3, "geometry_bwd: adam needs d_w, d_t, d_q packed as [d_w (F*B) | d_t (F*3) | d_q (F*4)]");
        FPC_CHECK_ARG(adam->max_iter > 0.f && adam->lr_ramp > 0.f, "geometry_bwd: adam: max_iter and lr_ramp must be positive");
        FPC_CHECK_ARG(adam->quat_mode == 0 || adam->quat_mode == 1, "geometry_bwd: adam: quat_mode must be 0 (per row) or 1 (Frobenius)");
        FPC_CHECK_ARG((long long)F * (B + 7) <= (1 << 22), "geometry_bwd: adam: packed parameter vector too long (%lld)", (long long)F * (B + 7));
        tl.adam = AdamArgs{adam->params, adam->m, adam->v, adam->step_count, adam->optimize_pose ? 1 : 0, adam->quat_mode,
                           adam->lr_w, adam->lr_t, adam->lr_q, adam->b1, adam->b2, adam->eps, adam->lr_ramp, adam->max_iter};
    }
    const int K = fpc_div_up((3 * B) >> 2, 32);
    int st;
    const size_t lsmem = (size_t)(C * 16 + GEO_WARPS * 3 * B + GEO_WARPS * C * 16) * sizeof(float);
#define FPC_BWD_CASE(k) case k: st = launch_bwd<k>(dim3(nblk, F), lsmem, stream, D, verts, mvp, g_pos, d_verts_add, V, B, F, C, d_verts, part_w, part_mvp, tl); break
    switch (K <= 8 ? K : (K <= 10 ? 10 : (K <= 12 ? 12 : (K <= 16 ? 16 : 24)))) {
        FPC_BWD_CASE(1); FPC_BWD_CASE(2); FPC_BWD_CASE(3); FPC_BWD_CASE(4); FPC_BWD_CASE(5); FPC_BWD_CASE(6);
        FPC_BWD_CASE(7); FPC_BWD_CASE(8); FPC_BWD_CASE(10); FPC_BWD_CASE(12); FPC_BWD_CASE(16); FPC_BWD_CASE(24);
        default: st = FPC_ERR_UNSUPPORTED; fpc_set_error("geometry_bwd: unsupported B=%d", B);
    }
#undef FPC_BWD_CASE
    return st;
}

extern "C" int fpc_adam_fused(float* params, const float* grads, float* m, float* v, int B, int F, int optimize_pose,
                              float lr_w, float lr_t, float lr_q, float b1, float b2, float eps, float lr_ramp, float max_iter,
                              int quat_mode, float* step_count, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(params && grads && m && v && step_count, "adam_fused: null pointer argument");
    FPC_CHECK_ARG(B > 0 && F > 0 && max_iter > 0.f && lr_ramp > 0.f, "adam_fused: B, F, max_iter and lr_ramp must be positive");
    FPC_CHECK_ARG((long long)F * (B + 7) <= (1 << 22), "adam_fused: packed parameter vector too long for the single-CTA kernel (%lld)", (long long)F * (B + 7));
    FPC_CHECK_ARG(quat_mode == 0 || quat_mode == 1, "adam_fused: quat_mode must be 0 (per row) or 1 (Frobenius)");
    const AdamArgs a{params, m, v, step_count, optimize_pose ? 1 : 0, quat_mode, lr_w, lr_t, lr_q, b1, b2, eps, lr_ramp, max_iter};
    k_adam_fused<<<1, 1024, 0, stream>>>(a, grads, F * B, F);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
