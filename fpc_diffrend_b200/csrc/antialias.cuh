// Pair analysis and silhouette position gradient of the antialias op, shared by the op-level
// kernels (antialias.cu) and the fused render + antialias + loss kernel (fused_aa.cu).
// Every arithmetic op of aa_analyze mirrors oracle/golden.c:aa_analyze in order and rounding.
#pragma once
#include "common.cuh"

namespace fpc {

struct AAParams {
    const float* rast;
    const float* pos;
    const int32_t* tri;
    const int32_t* tri_opp;
    int N, V, T, H, W, C;
    float xh, yh;
};

struct AAPair { bool valid; int di; int tri; float alpha; int px, py; };

__device__ __forceinline__ bool same_sign(float a, float b) { return (__float_as_int(a) ^ __float_as_int(b)) >= 0; }

__device__ __forceinline__ bool rational_gt(float n0, float n1, float d0, float d1)
{
    float p0 = xmul(n0, d1), p1 = xmul(n1, d0);
    return same_sign(d0, d1) ? (p0 > p1) : (p0 < p1);
}

__device__ __forceinline__ int max_idx3(float n0, float n1, float n2, float d0, float d1, float d2)
{
    bool g10 = rational_gt(n1, n0, d1, d0);
    bool g20 = rational_gt(n2, n0, d2, d0);
    bool g21 = rational_gt(n2, n1, d2, d1);
    if (g20 && g21) return 2;
    if (g10) return 1;
    return 0;
}

__device__ __forceinline__ float cross2(float ax, float ay, float bx, float by) { return xsub(xmul(ax, by), xmul(bx, ay)); }

// Analysis of the pair (px,py) -> (px+1,py) [d=0] or (px,py+1) [d=1].  zt0 / zt1 = (z/w, id) of the two pixels.
// Every arithmetic op mirrors oracle/golden.c:aa_analyze in order and rounding.
__device__ __forceinline__ AAPair aa_analyze(const AAParams& ap, int n, int px, int py, int d, float2 zt0, float2 zt1)
{
    AAPair r; r.valid = false; r.di = 0; r.tri = -1; r.alpha = 0.f; r.px = px; r.py = py;
    int tri0 = rast_tri(zt0.y), tri1 = rast_tri(zt1.y);
    if (tri0 == tri1) return r;
    int t = (tri0 >= 0) ? tri0 : tri1;
    if (tri0 >= 0 && tri1 >= 0) t = (zt0.x < zt1.x) ? tri0 : tri1;
    if (t == tri1) { px += 1 - d; py += d; }
    if (t < 0 || t >= ap.T) return r;
    int vi0 = __ldg(ap.tri + 3 * t), vi1 = __ldg(ap.tri + 3 * t + 1), vi2 = __ldg(ap.tri + 3 * t + 2);
    if ((unsigned)vi0 >= (unsigned)ap.V || (unsigned)vi1 >= (unsigned)ap.V || (unsigned)vi2 >= (unsigned)ap.V) return r;
    int op0 = __ldg(ap.tri_opp + 3 * t), op1 = __ldg(ap.tri_opp + 3 * t + 1), op2 = __ldg(ap.tri_opp + 3 * t + 2);
    const float* P = ap.pos + (size_t)n * ap.V * 4;
    float4 p0 = ldg4(P + 4 * (size_t)vi0), p1 = ldg4(P + 4 * (size_t)vi1), p2 = ldg4(P + 4 * (size_t)vi2);
    float4 o0 = (op0 < 0) ? p0 : ldg4(P + 4 * (size_t)op0);
    float4 o1 = (op1 < 0) ? p1 : ldg4(P + 4 * (size_t)op1);
    float4 o2 = (op2 < 0) ? p2 : ldg4(P + 4 * (size_t)op2);
    float xh = ap.xh, yh = ap.yh;
    float w0 = xrcp(p0.w), w1 = xrcp(p1.w), w2 = xrcp(p2.w);
    float ow0 = xrcp(o0.w), ow1 = xrcp(o1.w), ow2 = xrcp(o2.w);
    float fx = xsub(xadd((float)px, 0.5f), xh), fy = xsub(xadd((float)py, 0.5f), yh);
    float x0 = xsub(xmul(xmul(p0.x, w0), xh), fx), y0 = xsub(xmul(xmul(p0.y, w0), yh), fy);
    float x1 = xsub(xmul(xmul(p1.x, w1), xh), fx), y1 = xsub(xmul(xmul(p1.y, w1), yh), fy);
    float x2 = xsub(xmul(xmul(p2.x, w2), xh), fx), y2 = xsub(xmul(xmul(p2.y, w2), yh), fy);
    float ox0 = xsub(xmul(xmul(o0.x, ow0), xh), fx), oy0 = xsub(xmul(xmul(o0.y, ow0), yh), fy);
    float ox1 = xsub(xmul(xmul(o1.x, ow1), xh), fx), oy1 = xsub(xmul(xmul(o1.y, ow1), yh), fy);
    float ox2 = xsub(xmul(xmul(o2.x, ow2), xh), fx), oy2 = xsub(xmul(xmul(o2.y, ow2), yh), fy);
    float bb = cross2(xsub(x1, x0), xsub(y1, y0), xsub(x2, x0), xsub(y2, y0));
    float a0 = cross2(xsub(x1, ox0), xsub(y1, oy0), xsub(x2, ox0), xsub(y2, oy0));
    float a1 = cross2(xsub(x2, ox1), xsub(y2, oy1), xsub(x0, ox1), xsub(y0, oy1));
    float a2 = cross2(xsub(x0, ox2), xsub(y0, oy2), xsub(x1, ox2), xsub(y1, oy2));
    bool s0 = same_sign(a0, bb), s1 = same_sign(a1, bb), s2 = same_sign(a2, bb);
    if (!(s0 || s1 || s2)) return r;
    if (d) { float s; s = x0; x0 = y0; y0 = s; s = x1; x1 = y1; y1 = s; s = x2; x2 = y2; y2 = s; }
    float dx0 = xsub(x2, x1), dx1 = xsub(x0, x2), dx2 = xsub(x1, x0);
    float dy0 = xsub(y2, y1), dy1 = xsub(y0, y2), dy2 = xsub(y1, y0);
    const float FMAXV = 3.402823466e38f;
    float dc = -FMAXV;
    float ds = (t == tri0) ? 1.f : -1.f;
    float d0 = xmul(ds, xsub(xmul(x1, dy0), xmul(y1, dx0)));
    float d1 = xmul(ds, xsub(xmul(x2, dy1), xmul(y2, dx1)));
    float d2 = xmul(ds, xsub(xmul(x0, dy2), xmul(y0, dx2)));
    if (same_sign(y1, y2)) { d0 = -FMAXV; dy0 = 1.f; }
    if (same_sign(y2, y0)) { d1 = -FMAXV; dy1 = 1.f; }
    if (same_sign(y0, y1)) { d2 = -FMAXV; dy2 = 1.f; }
    int di = max_idx3(d0, d1, d2, dy0, dy1, dy2);
    if (di == 0 && s0 && fabsf(dy0) >= fabsf(dx0)) dc = xdiv(d0, dy0);
    if (di == 1 && s1 && fabsf(dy1) >= fabsf(dx1)) dc = xdiv(d1, dy1);
    if (di == 2 && s2 && fabsf(dy2) >= fabsf(dx2)) dc = xdiv(d2, dy2);
    const float eps = 0.0625f;
    if (dc > -eps && dc < 1.f + eps) {
        dc = clamp01(dc);
        r.valid = true; r.di = di; r.tri = t; r.alpha = xmul(ds, xsub(0.5f, dc)); r.px = px; r.py = py;
    }
    return r;
}

// Silhouette position gradient of one accepted pair: gradient (x, y, w) of the two end points of the crossing edge —
// corners (di + 1) % 3 and (di + 2) % 3 of the pair's triangle, clip-space positions p1, p2 — for the pair analysed at pixel
// (apx, apy) in direction d, given dd = sum_c d loss / d out_c (colour_1 - colour_0) of the pair.
__device__ __forceinline__ void aa_pair_corner_grads(float xh, float yh, const float4& p1, const float4& p2, int apx, int apy, int d, float dd,
                                                     float (&gp1)[3], float (&gp2)[3])
{
    float w1 = xrcp(p1.w), w2 = xrcp(p2.w);
    float fx = xsub(xadd((float)apx, 0.5f), xh), fy = xsub(xadd((float)apy, 0.5f), yh);
    float x1 = xsub(xmul(xmul(p1.x, w1), xh), fx), y1 = xsub(xmul(xmul(p1.y, w1), yh), fy);
    float x2 = xsub(xmul(xmul(p2.x, w2), xh), fx), y2 = xsub(xmul(xmul(p2.y, w2), yh), fy);
    if (d) { float s; s = x1; x1 = y1; y1 = s; s = x2; x2 = y2; y2 = s; }
    float dxx = x2 - x1, dyy = y2 - y1;
    float db = x1 * dyy - y1 * dxx;
    float iy = 1.f / (dyy + copysignf(1e-3f, dyy));
    float dby = db * iy;
    float iw1 = -w1 * iy * dd, iw2 = w2 * iy * dd;
    float s1 = d ? yh : xh, s2 = d ? xh : yh;
    float gp1x = iw1 * s1 * y2, gp2x = iw2 * s1 * y1;
    float gp1y = iw1 * s2 * (dby - x2), gp2y = iw2 * s2 * (dby - x1);
    if (d) { float s; s = gp1x; gp1x = gp1y; gp1y = s; s = gp2x; gp2x = gp2y; gp2y = s; }
    gp1[0] = gp1x; gp1[1] = gp1y; gp1[2] = -(p1.x * gp1x + p1.y * gp1y) * w1;
    gp2[0] = gp2x; gp2[1] = gp2y; gp2[2] = -(p2.x * gp2x + p2.y * gp2y) * w2;
}

// op-level antialias backward: the pair's gradient is scattered with float REDs (as upstream); handled by the thread that
// owns pix0
__device__ __forceinline__ void aa_pos_grad(const AAParams& ap, int n, const AAPair& a, int d, float dd, float* __restrict__ g_pos)
{
    int t = a.tri;
    int i1 = __ldg(ap.tri + 3 * t + (a.di + 1) % 3), i2 = __ldg(ap.tri + 3 * t + (a.di + 2) % 3);
    const float* P = ap.pos + (size_t)n * ap.V * 4;
    float4 p1 = ldg4(P + 4 * (size_t)i1), p2 = ldg4(P + 4 * (size_t)i2);
    float gp1[3], gp2[3];
    aa_pair_corner_grads(ap.xh, ap.yh, p1, p2, a.px, a.py, d, dd, gp1, gp2);
    float* G = g_pos + (size_t)n * ap.V * 4;
    atomicAdd(G + 4 * (size_t)i1 + 0, gp1[0]); atomicAdd(G + 4 * (size_t)i1 + 1, gp1[1]); atomicAdd(G + 4 * (size_t)i1 + 3, gp1[2]);
    atomicAdd(G + 4 * (size_t)i2 + 0, gp2[0]); atomicAdd(G + 4 * (size_t)i2 + 1, gp2[1]); atomicAdd(G + 4 * (size_t)i2 + 3, gp2[2]);
}

}  // namespace fpc
