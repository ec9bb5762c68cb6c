// Attribute interpolation fwd/bwd (replaces dr.interpolate, reference fit.py:157).
#include "common.cuh"

namespace {

template <int A_STATIC>
__global__ void __launch_bounds__(256) k_interp_fwd(const float* __restrict__ attr, int attr_stride, int Vt, int A_dyn,
                                                    const float* __restrict__ rast, const int32_t* __restrict__ tri,
                                                    long long npx_total, long long npx_inst, int T, float* __restrict__ out)
{
    const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= npx_total) return;
    float4 r = ldg4(rast + 4 * pi);
    int t = rast_tri(r.w);
    float* o = out + pi * A;
    bool ok = t >= 0 && t < T;
    int i0 = 0, i1 = 0, i2 = 0;
    if (ok) {
        i0 = __ldg(tri + 3 * t); i1 = __ldg(tri + 3 * t + 1); i2 = __ldg(tri + 3 * t + 2);
        ok = (unsigned)i0 < (unsigned)Vt && (unsigned)i1 < (unsigned)Vt && (unsigned)i2 < (unsigned)Vt;
    }
    if (!ok) {
#pragma unroll
        for (int c = 0; c < A; c++) o[c] = 0.f;
        return;
    }
    const float* at = attr + (size_t)(pi / npx_inst) * attr_stride;
    const float* a0 = at + (size_t)i0 * A;
    const float* a1 = at + (size_t)i1 * A;
    const float* a2 = at + (size_t)i2 * A;
    float b0 = r.x, b1 = r.y, b2 = 1.f - r.x - r.y;
#pragma unroll
    for (int c = 0; c < A; c++) o[c] = b0 * __ldg(a0 + c) + b1 * __ldg(a1 + c) + b2 * __ldg(a2 + c);
}

template <int A_STATIC>
__global__ void __launch_bounds__(256) k_interp_bwd(const float* __restrict__ attr, int attr_stride, int Vt, int A_dyn,
                                                    const float* __restrict__ rast, const int32_t* __restrict__ tri,
                                                    const float* __restrict__ dy, long long npx_total, long long npx_inst,
                                                    int T, float* __restrict__ g_attr, float* __restrict__ g_rast)
{
    const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= npx_total) return;
    float4 r = ldg4(rast + 4 * pi);
    int t = rast_tri(r.w);
    float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
    bool ok = t >= 0 && t < T;
    int i0 = 0, i1 = 0, i2 = 0;
    if (ok) {
        i0 = __ldg(tri + 3 * t); i1 = __ldg(tri + 3 * t + 1); i2 = __ldg(tri + 3 * t + 2);
        ok = (unsigned)i0 < (unsigned)Vt && (unsigned)i1 < (unsigned)Vt && (unsigned)i2 < (unsigned)Vt;
    }
    if (ok) {
        size_t ao = (size_t)(pi / npx_inst) * attr_stride;
        const float* a0 = attr + ao + (size_t)i0 * A;
        const float* a1 = attr + ao + (size_t)i1 * A;
        const float* a2 = attr + ao + (size_t)i2 * A;
        float* ga = g_attr + ao;
        const float* d = dy + pi * A;
        float b0 = r.x, b1 = r.y, b2 = 1.f - r.x - r.y;
        float gu = 0.f, gv = 0.f;
#pragma unroll
        for (int c = 0; c < A; c++) {
            float g = __ldg(d + c);
            if (g != 0.f) {
                atomicAdd(ga + (size_t)i0 * A + c, b0 * g);
                atomicAdd(ga + (size_t)i1 * A + c, b1 * g);
                atomicAdd(ga + (size_t)i2 * A + c, b2 * g);
            }
            float v2 = __ldg(a2 + c);
            gu += g * (__ldg(a0 + c) - v2);
            gv += g * (__ldg(a1 + c) - v2);
        }
        gr.x = gu; gr.y = gv;
    }
    reinterpret_cast<float4*>(g_rast)[pi] = gr;
}

}  // namespace

extern "C" int fpc_interpolate_fwd(const float* attr, int Na, int Vt, int A, const float* rast, const int32_t* tri,
                                   int N, int T, int H, int W, float* out, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(attr && rast && tri && out, "interpolate_fwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && T > 0 && H > 0 && W > 0 && Vt > 0 && A > 0, "interpolate_fwd: sizes must be positive");
    FPC_CHECK_ARG(Na == 1 || Na == N, "interpolate_fwd: attr batch must be 1 or N (got %d, N=%d)", Na, N);
    long long npx_inst = (long long)H * W, npx = npx_inst * N;
    int stride = Na == 1 ? 0 : Vt * A;
    int grid = fpc_div_up(npx, 256);
    switch (A) {
    case 1: k_interp_fwd<1><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, npx, npx_inst, T, out); break;
    case 2: k_interp_fwd<2><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, npx, npx_inst, T, out); break;
    case 3: k_interp_fwd<3><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, npx, npx_inst, T, out); break;
    case 4: k_interp_fwd<4><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, npx, npx_inst, T, out); break;
    default: k_interp_fwd<0><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, npx, npx_inst, T, out); break;
    }
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_interpolate_bwd(const float* attr, int Na, int Vt, int A, const float* rast, const int32_t* tri,
                                   const float* dy, int N, int T, int H, int W, float* grad_attr, float* grad_rast,
                                   fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(attr && rast && tri && dy && grad_attr && grad_rast, "interpolate_bwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && T > 0 && H > 0 && W > 0 && Vt > 0 && A > 0, "interpolate_bwd: sizes must be positive");
    FPC_CHECK_ARG(Na == 1 || Na == N, "interpolate_bwd: attr batch must be 1 or N (got %d, N=%d)", Na, N);
    long long npx_inst = (long long)H * W, npx = npx_inst * N;
    int stride = Na == 1 ? 0 : Vt * A;
    FPC_CUDA(cudaMemsetAsync(grad_attr, 0, (size_t)Na * Vt * A * sizeof(float), stream));
    int grid = fpc_div_up(npx, 256);
    switch (A) {
    case 1: k_interp_bwd<1><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, dy, npx, npx_inst, T, grad_attr, grad_rast); break;
    case 2: k_interp_bwd<2><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, dy, npx, npx_inst, T, grad_attr, grad_rast); break;
    case 3: k_interp_bwd<3><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, dy, npx, npx_inst, T, grad_attr, grad_rast); break;
    case 4: k_interp_bwd<4><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, dy, npx, npx_inst, T, grad_attr, grad_rast); break;
    default: k_interp_bwd<0><<<grid, 256, 0, stream>>>(attr, stride, Vt, A, rast, tri, dy, npx, npx_inst, T, grad_attr, grad_rast); break;
    }
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
