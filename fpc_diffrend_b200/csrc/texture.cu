// Bilinear texture fetch fwd/bwd, boundary 'wrap' (replaces dr.texture(filter_mode='linear'),
// reference fit.py:158).
#include "common.cuh"

namespace {

struct TexFetch { int i00, i10, i01, i11; float fx, fy; };

__device__ __forceinline__ TexFetch tex_index(float u, float v, int Wt, int Ht)
{
    TexFetch f;
    u = u - floorf(u);
    v = v - floorf(v);
    float x = xsub(xmul(u, (float)Wt), 0.5f), y = xsub(xmul(v, (float)Ht), 0.5f);
    float x0f = floorf(x), y0f = floorf(y);
    int ix0 = (int)x0f, iy0 = (int)y0f, ix1 = ix0 + 1, iy1 = iy0 + 1;
    f.fx = x - x0f; f.fy = y - y0f;
    if (ix0 < 0) ix0 += Wt;
    if (iy0 < 0) iy0 += Ht;
    if (ix1 >= Wt) ix1 -= Wt;
    if (iy1 >= Ht) iy1 -= Ht;
    f.i00 = iy0 * Wt + ix0; f.i10 = iy0 * Wt + ix1; f.i01 = iy1 * Wt + ix0; f.i11 = iy1 * Wt + ix1;
    return f;
}

template <int C_STATIC>
__global__ void __launch_bounds__(256) k_tex_fwd(const float* __restrict__ tex, size_t tex_stride, int Ht, int Wt, int C_dyn,
                                                 const float* __restrict__ uv, long long npx_total, long long npx_inst,
                                                 float* __restrict__ out)
{
    const int C = C_STATIC > 0 ? C_STATIC : C_dyn;
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= npx_total) return;
    float2 q = __ldg(reinterpret_cast<const float2*>(uv) + pi);
    TexFetch f = tex_index(q.x, q.y, Wt, Ht);
    const float* tx = tex + (size_t)(pi / npx_inst) * tex_stride;
    float* o = out + pi * C;
#pragma unroll
    for (int c = 0; c < C; c++) {
        float t00 = __ldg(tx + (size_t)f.i00 * C + c), t10 = __ldg(tx + (size_t)f.i10 * C + c);
        float t01 = __ldg(tx + (size_t)f.i01 * C + c), t11 = __ldg(tx + (size_t)f.i11 * C + c);
        float a = t00 + (t10 - t00) * f.fx, b = t01 + (t11 - t01) * f.fx;
        o[c] = a + (b - a) * f.fy;
    }
}

template <int C_STATIC>
__global__ void __launch_bounds__(256) k_tex_bwd(const float* __restrict__ tex, size_t tex_stride, int Ht, int Wt, int C_dyn,
                                                 const float* __restrict__ uv, const float* __restrict__ dy,
                                                 long long npx_total, long long npx_inst,
                                                 float* __restrict__ g_tex, float* __restrict__ g_uv)
{
    const int C = C_STATIC > 0 ? C_STATIC : C_dyn;
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= npx_total) return;
    float2 q = __ldg(reinterpret_cast<const float2*>(uv) + pi);
    TexFetch f = tex_index(q.x, q.y, Wt, Ht);
    size_t to = (size_t)(pi / npx_inst) * tex_stride;
    const float* tx = tex + to;
    const float* d = dy + pi * C;
    float gu = 0.f, gv = 0.f;
    float w00 = (1.f - f.fx) * (1.f - f.fy), w10 = f.fx * (1.f - f.fy), w01 = (1.f - f.fx) * f.fy, w11 = f.fx * f.fy;
#pragma unroll
    for (int c = 0; c < C; c++) {
        float g = __ldg(d + c);
        float t00 = __ldg(tx + (size_t)f.i00 * C + c), t10 = __ldg(tx + (size_t)f.i10 * C + c);
        float t01 = __ldg(tx + (size_t)f.i01 * C + c), t11 = __ldg(tx + (size_t)f.i11 * C + c);
        if (g_tex && g != 0.f) {
            float* gt = g_tex + to;
            atomicAdd(gt + (size_t)f.i00 * C + c, g * w00);
            atomicAdd(gt + (size_t)f.i10 * C + c, g * w10);
            atomicAdd(gt + (size_t)f.i01 * C + c, g * w01);
            atomicAdd(gt + (size_t)f.i11 * C + c, g * w11);
        }
        gu += g * ((t10 - t00) * (1.f - f.fy) + (t11 - t01) * f.fy);
        gv += g * ((t01 - t00) * (1.f - f.fx) + (t11 - t10) * f.fx);
    }
    reinterpret_cast<float2*>(g_uv)[pi] = make_float2(gu * (float)Wt, gv * (float)Ht);
}

}  // namespace

extern "C" int fpc_texture_linear_fwd(const float* tex, int Nt, int Ht, int Wt, int C, const float* uv,
                                      int N, int H, int W, float* out, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(tex && uv && out, "texture_linear_fwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && H > 0 && W > 0 && Ht > 0 && Wt > 0 && C > 0, "texture_linear_fwd: sizes must be positive");
    FPC_CHECK_ARG(Nt == 1 || Nt == N, "texture_linear_fwd: texture batch must be 1 or N (got %d, N=%d)", Nt, N);
    long long npx_inst = (long long)H * W, npx = npx_inst * N;
    size_t stride = Nt == 1 ? 0 : (size_t)Ht * Wt * C;
    int grid = fpc_div_up(npx, 256);
    switch (C) {
    case 1: k_tex_fwd<1><<<grid, 256, 0, stream>>>(tex, stride, Ht, Wt, C, uv, npx, npx_inst, out); break;
    case 3: k_tex_fwd<3><<<grid, 256, 0, stream>>>(tex, stride, Ht, Wt, C, uv, npx, npx_inst, out); break;
    case 4: k_tex_fwd<4><<<grid, 256, 0, stream>>>(tex, stride, Ht, Wt, C, uv, npx, npx_inst, out); break;
    default: k_tex_fwd<0><<<grid, 256, 0, stream>>>(tex, stride, Ht, Wt, C, uv, npx, npx_inst, out); break;
    }
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_texture_linear_bwd(const float* tex, int Nt, int Ht, int Wt, int C, const float* uv, const float* dy,
                                      int N, int H, int W, float* grad_tex, float* grad_uv, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(tex && uv && dy && grad_uv, "texture_linear_bwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && H > 0 && W > 0 && Ht > 0 && Wt > 0 && C > 0, "texture_linear_bwd: sizes must be positive");
    FPC_CHECK_ARG(Nt == 1 || Nt == N, "texture_linear_bwd: texture batch must be 1 or N (got %d, N=%d)", Nt, N);
    long long npx_inst = (long long)H * W, npx = npx_inst * N;
    size_t stride = Nt == 1 ? 0 : (size_t)Ht * Wt * C;
    if (grad_tex) FPC_CUDA(cudaMemsetAsync(grad_tex, 0, (size_t)Nt * Ht * Wt * C * sizeof(float), stream));
    int grid = fpc_div_up(npx, 256);
    switch (C) {
    case 1: k_tex_bwd<1><<<grid, 256, 0, stream>>>(tex, stride, Ht, Wt, C, uv, dy, npx, npx_inst, grad_tex, grad_uv); break;
    case 3: k_tex_bwd<3><<<grid, 256, 0, stream>>>(tex, stride, Ht, Wt, C, uv, dy, npx, npx_inst, grad_tex, grad_uv); break;
    case 4: k_tex_bwd<4><<<grid, 256, 0, stream>>>(tex, stride, Ht, Wt, C, uv, dy, npx, npx_inst, grad_tex, grad_uv); break;
    default: k_tex_bwd<0><<<grid, 256, 0, stream>>>(tex, stride, Ht, Wt, C, uv, dy, npx, npx_inst, grad_tex, grad_uv); break;
    }
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
