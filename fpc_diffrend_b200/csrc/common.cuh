// Shared helpers for the sm_100a kernels of the fit hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/fpc_b200.h"

// ---- error plumbing (C-ABI never throws; see include/fpc_b200.h) -------------------------------------
void fpc_set_error(const char* fmt, ...);

#define FPC_CHECK_ARG(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) { fpc_set_error(__VA_ARGS__); return FPC_ERR_INVALID_ARGUMENT; } \
    } while (0)

#define FPC_CUDA(call)                                             \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) {                                  \
            fpc_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return FPC_ERR_CUDA;                                   \
        }                                                          \
    } while (0)

#define FPC_LAUNCH_CHECK()                                         \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) {                                  \
            fpc_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return FPC_ERR_CUDA;                                   \
        }                                                          \
    } while (0)

static inline int fpc_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute() is per DEVICE: a launcher remembers which devices it has configured (one process may drive several,
// from several host threads).  Usage:  if (once.need()) { FPC_CUDA(cudaFuncSetAttribute(...)); once.done(); }
// The bit is set only after the attribute call succeeded; two threads racing here both make the (idempotent) call.
struct FpcPerDeviceOnce {
    std::atomic<unsigned long long> mask[2];
    FpcPerDeviceOnce() { mask[0].store(0ull); mask[1].store(0ull); }
    static int device()
    {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) return -1;
        return d & 127;
    }
    bool need() const
    {
        const int d = device();
        if (d < 0) return true;
        return !((mask[d >> 6].load(std::memory_order_acquire) >> (d & 63)) & 1ull);
    }
    void done()
    {
        const int d = device();
        if (d >= 0) mask[d >> 6].fetch_or(1ull << (d & 63), std::memory_order_release);
    }
};

// ---- exact fp32 ops: never contracted to FMA, so coverage / depth decisions are bit-identical to the
//      CPU golden model compiled with -ffp-contract=off (DESIGN.md "Rasterizer semantics") ----------------
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
// correctly rounded 1/x: the same bits as xdiv(1, x) (and as 1.0f / x on the CPU), in fewer instructions
__device__ __forceinline__ float xrcp(float x) { return __frcp_rn(x); }

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// rast.w holds float(tri_id + 1); ids < 2^24 are exact in fp32
__device__ __forceinline__ int rast_tri(float w) { return (int)w - 1; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Shading formula of SURVEY App. A.1 (perspective-correct barycentrics from clip-space vertices).
struct Shade { float u, v, zw, iw; };
struct ShadeLazy { float u, v, iw, zn, wn; };      // z/w = xdiv(zn, wn), left to the caller (fused kernel: only for rast_out)

__device__ __forceinline__ Shade shade_pixel(const float4& p0, const float4& p1, const float4& p2,
                                             float fx, float fy)
{
    float p0x = xsub(p0.x, xmul(fx, p0.w)), p0y = xsub(p0.y, xmul(fy, p0.w));
    float p1x = xsub(p1.x, xmul(fx, p1.w)), p1y = xsub(p1.y, xmul(fy, p1.w));
    float p2x = xsub(p2.x, xmul(fx, p2.w)), p2y = xsub(p2.y, xmul(fy, p2.w));
    float a0 = xsub(xmul(p1x, p2y), xmul(p1y, p2x));
    float a1 = xsub(xmul(p2x, p0y), xmul(p2y, p0x));
    float a2 = xsub(xmul(p0x, p1y), xmul(p0y, p1x));
    float at = xadd(xadd(a0, a1), a2);
    float iw = xrcp(at);
    Shade s;
    s.iw = iw;
    s.u = xmul(a0, iw);
    s.v = xmul(a1, iw);
    float z = xadd(xadd(xmul(p0.z, a0), xmul(p1.z, a1)), xmul(p2.z, a2));
    float w = xadd(xadd(xmul(p0.w, a0), xmul(p1.w, a1)), xmul(p2.w, a2));
    s.zw = xdiv(z, w);
    return s;
}

__device__ __forceinline__ ShadeLazy shade_pixel_lazy(const float4& p0, const float4& p1, const float4& p2, float fx, float fy)
{
    float p0x = xsub(p0.x, xmul(fx, p0.w)), p0y = xsub(p0.y, xmul(fy, p0.w));
    float p1x = xsub(p1.x, xmul(fx, p1.w)), p1y = xsub(p1.y, xmul(fy, p1.w));
    float p2x = xsub(p2.x, xmul(fx, p2.w)), p2y = xsub(p2.y, xmul(fy, p2.w));
    float a0 = xsub(xmul(p1x, p2y), xmul(p1y, p2x));
    float a1 = xsub(xmul(p2x, p0y), xmul(p2y, p0x));
    float a2 = xsub(xmul(p0x, p1y), xmul(p0y, p1x));
    float at = xadd(xadd(a0, a1), a2);
    ShadeLazy s;
    s.iw = xrcp(at);
    s.u = xmul(a0, s.iw);
    s.v = xmul(a1, s.iw);
    s.zn = xadd(xadd(xmul(p0.z, a0), xmul(p1.z, a1)), xmul(p2.z, a2));
    s.wn = xadd(xadd(xmul(p0.w, a0), xmul(p1.w, a1)), xmul(p2.w, a2));
    return s;
}

// The same barycentrics for the BACKWARD pass: the cross products are formed with one fused multiply-add each (the first
// product enters unrounded), which halves the cancellation error of a_k on sliver triangles.  The forward outputs must keep
// the op order shared with oracle/golden.c (above); the gradient only has to be accurate — the op-level backward kernels get
// the same contraction from the compiler.  Measured at 20k vertices / 1024^2: worst-vertex gradient error vs float64 3e-3 of
// the largest gradient with the forward's values, 1.5e-4 with these.
struct ShadeGrad { float u, v, iw; };

__device__ __forceinline__ ShadeGrad shade_pixel_grad(const float4& p0, const float4& p1, const float4& p2, float fx, float fy)
{
    float p0x = __fmaf_rn(-fx, p0.w, p0.x), p0y = __fmaf_rn(-fy, p0.w, p0.y);
    float p1x = __fmaf_rn(-fx, p1.w, p1.x), p1y = __fmaf_rn(-fy, p1.w, p1.y);
    float p2x = __fmaf_rn(-fx, p2.w, p2.x), p2y = __fmaf_rn(-fy, p2.w, p2.y);
    float a0 = __fmaf_rn(p1x, p2y, -(p1y * p2x));
    float a1 = __fmaf_rn(p2x, p0y, -(p2y * p0x));
    float a2 = __fmaf_rn(p0x, p1y, -(p0y * p1x));
    ShadeGrad s;
    s.iw = 1.f / (a0 + a1 + a2);
    s.u = a0 * s.iw;
    s.v = a1 * s.iw;
    return s;
}

// image loss per channel value e = ref - 255 colour (fit.py:579 is the L2 form; the north-star also asks for L1):
//   L2: term e^2, d term / d colour = -510 e;   L1: term |e|, d term / d colour = -255 sign(e)  (sign(0) = 0, as torch.abs)
__device__ __forceinline__ float loss_term(float e, int l1) { return l1 ? fabsf(e) : e * e; }
__device__ __forceinline__ float loss_dcolour(float e, float k, int l1)
{
    return l1 ? (-255.f * k) * (float)((e > 0.f) - (e < 0.f)) : (-510.f * k) * e;
}

// pixel centre in NDC: fx = (2/W) * px + (1/W - 1), evaluated as separate mul/add
__device__ __forceinline__ float pixel_ndc(int p, float scale, float offset) { return xadd(xmul(scale, (float)p), offset); }
