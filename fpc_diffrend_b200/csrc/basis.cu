// Learned vertex basis of the reference's "free" and "combined" optimisation modes:
//   blend_free      (fit.py:47-62)   V_f = base + m3 (m2 (m1 e_f))
//   blend_combined  (fit.py:66-99)   V_f = base + D w_f + c * m3 (m2 (m1 e_f)),  c = learned_coefficient (0.5, fit.py:562)
// with m1, m2 [Fn,Fn] (identity at start) and m3 [R,Fn] (zero at start) shared by all Fn frames of the take
// (fit.py:166-179), e_f the one-hot vector of frame f.  For a batch of Fb frames with take-wide ids `frame_ids`:
//   x1[b,:] = m1[:, id_b]            (the one-hot product is a column gather)
//   x2[b,:] = m2 x1[b,:]             -> the "activations" of the learned basis: fpc_blend_fwd_ex(m3, ..., x2, ...)
// and the backward pass
//   d_x2 = d_verts m3                (fpc_blend_bwd with D := m3)
//   d_m3 = c * d_verts^T x2          (fpc_basis_grad)
//   d_m2 = c * d_x2^T x1,  d_x1 = c * d_x2 m2,  d_m1[:, id_b] = d_x1[b,:]   (fpc_basis_code_bwd)
// plus the two optional L2 terms of the loop (fit.py:584-595): mean(deformations^2), mean(activations^2).
// All sums run in a fixed order (no atomics): results are deterministic.
#include "common.cuh"

namespace {

// one CTA per batch frame: x1 = column id of m1 (kept in shared memory), then one warp per row of m2
__global__ void __launch_bounds__(256) k_code_fwd(const float* __restrict__ m1, const float* __restrict__ m2,
                                                  const int32_t* __restrict__ frame_ids, int Fn,
                                                  float* __restrict__ x1, float* __restrict__ x2)
{
    extern __shared__ float sx[];
    const int b = blockIdx.x, id = frame_ids[b];
    for (int j = threadIdx.x; j < Fn; j += blockDim.x) {
        float v = __ldg(m1 + (size_t)j * Fn + id);
        sx[j] = v;
        x1[(size_t)b * Fn + j] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; i < Fn; i += nw) {
        const float* row = m2 + (size_t)i * Fn;
        float acc = 0.f;
        for (int j = lane; j < Fn; j += 32) acc += __ldg(row + j) * sx[j];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) x2[(size_t)b * Fn + i] = acc;
    }
}

constexpr int BG_ROWS = 32;     // rows of d_m3 per CTA
constexpr int BG_FCHUNK = 64;   // batch frames staged in shared memory per pass

// d_m3[r,k] = coef * sum_b d_verts[b,r] x2[b,k]: consecutive threads own consecutive k (coalesced stores, coalesced
// x2 loads through L1), the d_verts column block of the CTA's rows sits in shared memory
__global__ void __launch_bounds__(256) k_basis_grad(const float* __restrict__ d_verts, const float* __restrict__ x2,
                                                    int R, int Fn, int Fb, float coef, float* __restrict__ d_m3)
{
    __shared__ float sdv[BG_FCHUNK][BG_ROWS + 1];
    const int r0 = blockIdx.x * BG_ROWS, nrows = min(BG_ROWS, R - r0);
    for (int b0 = 0; b0 < Fb; b0 += BG_FCHUNK) {
        const int nb = min(BG_FCHUNK, Fb - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * BG_ROWS; i += blockDim.x) {
            int bb = i / BG_ROWS, rr = i - bb * BG_ROWS;
            sdv[bb][rr] = (rr < nrows) ? __ldg(d_verts + (size_t)(b0 + bb) * R + r0 + rr) : 0.f;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < nrows * Fn; idx += blockDim.x) {
            int rr = idx / Fn, k = idx - rr * Fn;
            float acc = 0.f;
            for (int bb = 0; bb < nb; bb++) acc += sdv[bb][rr] * __ldg(x2 + (size_t)(b0 + bb) * Fn + k);
            float* dst = d_m3 + (size_t)(r0 + rr) * Fn + k;
            *dst = (b0 == 0) ? coef * acc : *dst + coef * acc;
        }
    }
}

// d_m2[i,j] = coef * sum_b d_x2[b,i] x1[b,j]
__global__ void __launch_bounds__(256) k_code_bwd_m2(const float* __restrict__ x1, const float* __restrict__ d_x2, int Fn, int Fb,
                                                     float coef, float* __restrict__ d_m2)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)Fn * Fn) return;
    int i = (int)(gid / Fn), j = (int)(gid - (long long)i * Fn);
    float acc = 0.f;
    for (int b = 0; b < Fb; b++) acc += __ldg(d_x2 + (size_t)b * Fn + i) * __ldg(x1 + (size_t)b * Fn + j);
    d_m2[gid] = coef * acc;
}

// d_m1[j, id_b] = coef * sum_i d_x2[b,i] m2[i,j]   (d_m1 is cleared by the host function; frame ids of a batch are distinct)
__global__ void __launch_bounds__(256) k_code_bwd_m1(const float* __restrict__ m2, const float* __restrict__ d_x2,
                                                     const int32_t* __restrict__ frame_ids, int Fn, float coef, float* __restrict__ d_m1)
{
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Fn) return;
    float acc = 0.f;
    for (int i = 0; i < Fn; i++) acc += __ldg(d_x2 + (size_t)b * Fn + i) * __ldg(m2 + (size_t)i * Fn + j);
    d_m1[(size_t)j * Fn + frame_ids[b]] = coef * acc;
}

constexpr int L2_THREADS = 256, L2_PER_THREAD = 8;

// partial sums of x^2 and  g_out = g_scale * g_in + (2 weight / n) x
__global__ void __launch_bounds__(L2_THREADS) k_l2_reg(const float* __restrict__ x, long long total, float k2, const float* g_in, float g_scale,
                                                       float* g_out, double* __restrict__ partial)
{
    __shared__ double red[L2_THREADS / 32];
    long long base = (long long)blockIdx.x * L2_THREADS * L2_PER_THREAD;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < L2_PER_THREAD; u++) {
        long long i = base + (long long)u * L2_THREADS + threadIdx.x;
        if (i < total) {
            float v = __ldg(x + i);
            acc += (double)(v * v);
            if (g_out) g_out[i] = (g_in ? g_scale * g_in[i] : 0.f) + k2 * v;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < L2_THREADS / 32; w++) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) k_l2_reduce(const double* __restrict__ partial, int n, double k, float* __restrict__ loss_accum,
                                                   float* __restrict__ term)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float t = (float)(red[0] * k);
        if (term) term[0] = t;
        if (loss_accum) loss_accum[0] += t;
    }
}

}  // namespace

extern "C" int fpc_basis_code_fwd(const float* m1, const float* m2, const int32_t* frame_ids, int Fn, int Fb, float* x1, float* x2,
                                  fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(m1 && m2 && frame_ids && x1 && x2, "basis_code_fwd: null pointer argument");
    FPC_CHECK_ARG(Fn > 0 && Fb > 0 && Fn <= 8192, "basis_code_fwd: need 0 < Fn <= 8192 and Fb > 0 (got %d %d)", Fn, Fb);
    k_code_fwd<<<Fb, 256, (size_t)Fn * sizeof(float), stream>>>(m1, m2, frame_ids, Fn, x1, x2);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_basis_grad(const float* d_verts, const float* x2, int R, int Fn, int Fb, float coef, float* d_m3, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(d_verts && x2 && d_m3, "basis_grad: null pointer argument");
    FPC_CHECK_ARG(R > 0 && Fn > 0 && Fb > 0, "basis_grad: R, Fn, Fb must be positive (got %d %d %d)", R, Fn, Fb);
    k_basis_grad<<<fpc_div_up(R, BG_ROWS), 256, 0, stream>>>(d_verts, x2, R, Fn, Fb, coef, d_m3);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_basis_code_bwd(const float* m2, const float* x1, const float* d_x2, const int32_t* frame_ids, int Fn, int Fb, float coef,
                                  float* d_m1, float* d_m2, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(m2 && x1 && d_x2 && frame_ids && d_m1 && d_m2, "basis_code_bwd: null pointer argument");
    FPC_CHECK_ARG(Fn > 0 && Fb > 0 && Fb <= 65535, "basis_code_bwd: need Fn > 0 and 0 < Fb <= 65535 (got %d %d)", Fn, Fb);
    FPC_CUDA(cudaMemsetAsync(d_m1, 0, (size_t)Fn * Fn * sizeof(float), stream));
    k_code_bwd_m2<<<fpc_div_up((long long)Fn * Fn, 256), 256, 0, stream>>>(x1, d_x2, Fn, Fb, coef, d_m2);
    FPC_LAUNCH_CHECK();
    k_code_bwd_m1<<<dim3(fpc_div_up(Fn, 256), Fb), 256, 0, stream>>>(m2, d_x2, frame_ids, Fn, coef, d_m1);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" size_t fpc_l2_reg_scratch_bytes(long long total)
{
    if (total <= 0) return 256;
    return (size_t)fpc_div_up(total, L2_THREADS * L2_PER_THREAD) * sizeof(double) + 256;
}

extern "C" int fpc_l2_reg_fwd_bwd(const float* x, int F, long long n, float weight, float* loss_accum, float* term,
                                  const float* g_in, float g_scale, float* g_out, void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(x, "l2_reg_fwd_bwd: x must be non-null");
    FPC_CHECK_ARG(F > 0 && n > 0, "l2_reg_fwd_bwd: F and n must be positive");
    const long long total = (long long)F * n;
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_l2_reg_scratch_bytes(total), "l2_reg_fwd_bwd: scratch too small");
    const int nblk = fpc_div_up(total, L2_THREADS * L2_PER_THREAD);
    k_l2_reg<<<nblk, L2_THREADS, 0, stream>>>(x, total, 2.f * weight / (float)n, g_in, g_scale, g_out, (double*)scratch);
    FPC_LAUNCH_CHECK();
    k_l2_reduce<<<1, 256, 0, stream>>>((const double*)scratch, nblk, (double)weight / (double)n, loss_accum, term);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
