// Mip-mapped texturing path of the drop-in (reference fit.py:153-155 with enable_mip):
//   dr.rasterize(...)                                   -> rast_db, differentiable            (fpc_rasterize_bwd_db)
//   dr.interpolate(uv, rast, uv_idx, rast_db=rast_db, diff_attrs='all') -> (texc, texd)       (fpc_interpolate_da_fwd/bwd)
//   dr.texture(tex, texc, texd, filter_mode='linear-mipmap-linear', max_mip_level=k)          (fpc_texture_mip_*)
// Semantics: SURVEY App. A.1-A.3 and oracle/torch_ref.py (barycentric_diffs, interpolate_da, texture_mip); one thread per
// pixel, float REDs for the scatters (as the non-mip op-level kernels), the mip chain itself is built and back-propagated
// level by level without atomics (every fine texel has exactly one parent).
#include "common.cuh"

namespace {

constexpr int MAX_DIFF = 32;
constexpr int MAX_LEVELS = 16;

struct DiffSel { int k; int idx[MAX_DIFF]; };

struct MipDesc {
    int L;                          // number of mip levels above level 0
    int w[MAX_LEVELS + 1], h[MAX_LEVELS + 1];
    long long off[MAX_LEVELS + 1];  // float offset of level l (l >= 1) inside the mip buffer; off[0] unused
};

__host__ MipDesc mip_desc(int Nt, int Ht, int Wt, int C, int L)
{
    MipDesc d;
    d.L = L; d.w[0] = Wt; d.h[0] = Ht; d.off[0] = 0;
    long long o = 0;
    for (int l = 1; l <= MAX_LEVELS; l++) {
        d.w[l] = d.w[l - 1] >> 1; d.h[l] = d.h[l - 1] >> 1;
        d.off[l] = o;
        if (l <= L) o += (long long)Nt * d.h[l] * d.w[l] * C;
    }
    return d;
}

// ---- interpolate with pixel differentials -----------------------------------------------------------------

__global__ void __launch_bounds__(256) k_interp_da_fwd(const float* __restrict__ attr, int attr_stride, int Vt, int A,
                                                       const float* __restrict__ rast, const float* __restrict__ rast_db,
                                                       const int32_t* __restrict__ tri, DiffSel sel, long long npx_total, long long npx_inst,
                                                       int T, float* __restrict__ out, float* __restrict__ out_da)
{
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= npx_total) return;
    float4 r = ldg4(rast + 4 * pi);
    int t = rast_tri(r.w);
    float* o = out + pi * A;
    float* oda = out_da + pi * 2 * sel.k;
    bool ok = t >= 0 && t < T;
    int i0 = 0, i1 = 0, i2 = 0;
    if (ok) {
        i0 = __ldg(tri + 3 * t); i1 = __ldg(tri + 3 * t + 1); i2 = __ldg(tri + 3 * t + 2);
        ok = (unsigned)i0 < (unsigned)Vt && (unsigned)i1 < (unsigned)Vt && (unsigned)i2 < (unsigned)Vt;
    }
    if (!ok) {
        for (int c = 0; c < A; c++) o[c] = 0.f;
        for (int j = 0; j < 2 * sel.k; j++) oda[j] = 0.f;
        return;
    }
    const float* at = attr + (size_t)(pi / npx_inst) * attr_stride;
    const float* a0 = at + (size_t)i0 * A;
    const float* a1 = at + (size_t)i1 * A;
    const float* a2 = at + (size_t)i2 * A;
    float b0 = r.x, b1 = r.y, b2 = 1.f - r.x - r.y;
    for (int c = 0; c < A; c++) o[c] = b0 * __ldg(a0 + c) + b1 * __ldg(a1 + c) + b2 * __ldg(a2 + c);
    float4 db = ldg4(rast_db + 4 * pi);
    for (int j = 0; j < sel.k; j++) {
        int c = sel.idx[j];
        float v2 = __ldg(a2 + c), e0 = __ldg(a0 + c) - v2, e1 = __ldg(a1 + c) - v2;
        oda[2 * j] = db.x * e0 + db.z * e1;
        oda[2 * j + 1] = db.y * e0 + db.w * e1;
    }
}

__global__ void __launch_bounds__(256) k_interp_da_bwd(const float* __restrict__ attr, int attr_stride, int Vt, int A,
                                                       const float* __restrict__ rast, const float* __restrict__ rast_db,
                                                       const int32_t* __restrict__ tri, DiffSel sel, const float* __restrict__ dy,
                                                       const float* __restrict__ dda, long long npx_total, long long npx_inst, int T,
                                                       float* __restrict__ g_attr, float* __restrict__ g_rast, float* __restrict__ g_db)
{
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= npx_total) return;
    float4 r = ldg4(rast + 4 * pi);
    int t = rast_tri(r.w);
    float4 gr = make_float4(0.f, 0.f, 0.f, 0.f), gdb = make_float4(0.f, 0.f, 0.f, 0.f);
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
        float b0 = r.x, b1 = r.y, b2 = 1.f - r.x - r.y;
        if (dy) {
            const float* d = dy + pi * A;
            for (int c = 0; c < A; c++) {
                float g = __ldg(d + c);
                if (g != 0.f) {
                    atomicAdd(ga + (size_t)i0 * A + c, b0 * g);
                    atomicAdd(ga + (size_t)i1 * A + c, b1 * g);
                    atomicAdd(ga + (size_t)i2 * A + c, b2 * g);
                }
                float v2 = __ldg(a2 + c);
                gr.x += g * (__ldg(a0 + c) - v2);
                gr.y += g * (__ldg(a1 + c) - v2);
            }
        }
        if (dda) {
            float4 db = ldg4(rast_db + 4 * pi);
            const float* d = dda + pi * 2 * sel.k;
            for (int j = 0; j < sel.k; j++) {
                int c = sel.idx[j];
                float gx = __ldg(d + 2 * j), gy = __ldg(d + 2 * j + 1);
                if (gx == 0.f && gy == 0.f) continue;
                float v2 = __ldg(a2 + c), e0 = __ldg(a0 + c) - v2, e1 = __ldg(a1 + c) - v2;
                gdb.x += gx * e0; gdb.y += gy * e0; gdb.z += gx * e1; gdb.w += gy * e1;
                float w0 = gx * db.x + gy * db.y, w1 = gx * db.z + gy * db.w;
                atomicAdd(ga + (size_t)i0 * A + c, w0);
                atomicAdd(ga + (size_t)i1 * A + c, w1);
                atomicAdd(ga + (size_t)i2 * A + c, -(w0 + w1));
            }
        }
    }
    reinterpret_cast<float4*>(g_rast)[pi] = gr;
    if (g_db) reinterpret_cast<float4*>(g_db)[pi] = gdb;
}

// ---- rasterize backward including the gradient of rast_db ---------------------------------------------------
// a_k = C_k + A_k fx + B_k fy (barycentric numerators, affine in the pixel's NDC position), at = sum a_k, iw = 1/at,
// u = a0 iw, v = a1 iw, du/dX = xs iw (A0 - u At), du/dY = ys iw (B0 - u Bt), dv/dX = xs iw (A1 - v At), dv/dY = ys iw (B1 - v Bt).
__global__ void __launch_bounds__(256) k_raster_bwd_db(const float* __restrict__ pos, const int32_t* __restrict__ tri,
                                                       const float* __restrict__ rast, const float* __restrict__ dy, const float* __restrict__ ddb,
                                                       int N, int V, int T, int H, int W, float xs, float xo, float ys, float yo,
                                                       float* __restrict__ grad_pos)
{
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= (long long)N * H * W) return;
    float4 r = ldg4(rast + 4 * pi);
    int t = rast_tri(r.w);
    if (t < 0 || t >= T) return;
    float4 g = ldg4(dy + 4 * pi);
    float4 gd = ldg4(ddb + 4 * pi);
    if (g.x == 0.f && g.y == 0.f && gd.x == 0.f && gd.y == 0.f && gd.z == 0.f && gd.w == 0.f) return;
    int px = (int)(pi % W), py = (int)((pi / W) % H), n = (int)(pi / ((long long)W * H));
    int i[3] = {__ldg(tri + 3 * t), __ldg(tri + 3 * t + 1), __ldg(tri + 3 * t + 2)};
    const float* P = pos + (size_t)n * V * 4;
    // coordinates relative to the pixel, p_k = (x_k - fx w_k, y_k - fy w_k): the numerators become a_k = p_k1 x p_k2 (the
    // forward formula, no cancellation against the pixel position) and A_k, B_k keep their values
    float fx = pixel_ndc(px, xs, xo), fy = pixel_ndc(py, ys, yo);
    float x[3], y[3], w[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { float4 q = ldg4(P + 4 * (size_t)i[k]); w[k] = q.w; x[k] = q.x - fx * q.w; y[k] = q.y - fy * q.w; }
    float A[3], B[3], a[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
        A[k] = y[k1] * w[k2] - w[k1] * y[k2];
        B[k] = w[k1] * x[k2] - x[k1] * w[k2];
        a[k] = x[k1] * y[k2] - y[k1] * x[k2];
    }
    const float at = a[0] + a[1] + a[2], iw = 1.f / at;
    const float u = a[0] * iw, v = a[1] * iw;
    const float At = A[0] + A[1] + A[2], Bt = B[0] + B[1] + B[2];
    const float s1 = xs * gd.x, s2 = ys * gd.y, s3 = xs * gd.z, s4 = ys * gd.w;
    // through the differentials
    const float Giw_db = s1 * (A[0] - u * At) + s2 * (B[0] - u * Bt) + s3 * (A[1] - v * At) + s4 * (B[1] - v * Bt);
    const float Gu = g.x - iw * (s1 * At + s2 * Bt), Gv = g.y - iw * (s3 * At + s4 * Bt);
    const float GAt = -iw * (s1 * u + s3 * v), GBt = -iw * (s2 * u + s4 * v);
    const float GA[3] = {iw * s1 + GAt, iw * s3 + GAt, GAt};
    const float GB[3] = {iw * s2 + GBt, iw * s4 + GBt, GBt};
    // through u, v, iw
    const float Giw = Giw_db + Gu * a[0] + Gv * a[1];
    const float Gat = -iw * iw * Giw;
    const float GC[3] = {Gu * iw + Gat, Gv * iw + Gat, Gat};
    // A, B, a are bilinear in the (relative) vertex coordinates: vertex m appears in the two numerators that do not carry its index
    float* G = grad_pos + (size_t)n * V * 4;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const int p = (m + 1) % 3, q = (m + 2) % 3;          // numerators p and q contain vertex m
        // numerator p = (p+1, p+2) = (q, m): A_p = y_q w_m - w_q y_m, B_p = w_q x_m - x_q w_m, a_p = x_q y_m - y_q x_m
        // numerator q = (q+1, q+2) = (m, p): A_q = y_m w_p - w_m y_p, B_q = w_m x_p - x_m w_p, a_q = x_m y_p - y_m x_p
        float gx = w[q] * GB[p] - y[q] * GC[p] - w[p] * GB[q] + y[p] * GC[q];
        float gy = -w[q] * GA[p] + x[q] * GC[p] + w[p] * GA[q] - x[p] * GC[q];
        float gw = y[q] * GA[p] - x[q] * GB[p] - y[p] * GA[q] + x[p] * GB[q];
        gw -= fx * gx + fy * gy;                             // the relative coordinates depend on w_m
        if (gx != 0.f) atomicAdd(G + 4 * (size_t)i[m] + 0, gx);
        if (gy != 0.f) atomicAdd(G + 4 * (size_t)i[m] + 1, gy);
        if (gw != 0.f) atomicAdd(G + 4 * (size_t)i[m] + 3, gw);
    }
}

// ---- mip chain ----------------------------------------------------------------------------------------------

// dst [Nt,h,w,C] = 2x2 box filter of src [Nt,2h,2w,C]
__global__ void __launch_bounds__(256) k_mip_down(const float* __restrict__ src, float* __restrict__ dst, int Nt, int h, int w, int C)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Nt * h * w * C) return;
    int c = (int)(i % C);
    long long p = i / C;
    int x = (int)(p % w), y = (int)((p / w) % h), n = (int)(p / ((long long)w * h));
    const float* s = src + (((size_t)n * 2 * h + 2 * y) * 2 * w + 2 * x) * C + c;
    const size_t row = (size_t)2 * w * C;
    dst[i] = 0.25f * (((__ldg(s) + __ldg(s + C)) + __ldg(s + row)) + __ldg(s + row + C));
}

// g_fine [Nt,2h,2w,C] += 0.25 * g_coarse [Nt,h,w,C] (each fine texel has exactly one parent: plain read-modify-write)
__global__ void __launch_bounds__(256) k_mip_up(const float* __restrict__ g_coarse, float* __restrict__ g_fine, int Nt, int h, int w, int C)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Nt * 4 * h * w * C) return;
    int c = (int)(i % C);
    long long p = i / C;
    int x = (int)(p % (2 * w)), y = (int)((p / (2 * w)) % (2 * h)), n = (int)(p / ((long long)4 * w * h));
    g_fine[i] += 0.25f * __ldg(g_coarse + (((size_t)n * h + (y >> 1)) * w + (x >> 1)) * C + c);
}

struct Bilin { int i00, i10, i01, i11; float fx, fy; };

__device__ __forceinline__ Bilin bilin_index(float u, float v, int Wt, int Ht)
{
    Bilin f;
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

struct LevelSel { int l0, l1; float f; bool grad; };      // grad: the level is strictly inside (0, L) -> d out / d level exists

// level = 0.5 log2(major axis^2 of the pixel footprint in texels) + bias, clamped to [0, L]
__device__ __forceinline__ float mip_level_raw(float4 da, int Wt, int Ht, float bias, bool has_da)
{
    float lev = bias;
    if (has_da) {
        float dsdx = da.x * (float)Wt, dsdy = da.y * (float)Wt, dtdx = da.z * (float)Ht, dtdy = da.w * (float)Ht;
        float A = dsdx * dsdx + dtdx * dtdx, B = dsdy * dsdy + dtdy * dtdy, Cc = dsdx * dsdy + dtdx * dtdy;
        float l2b = 0.5f * (A + B), l2n = 0.25f * (A - B) * (A - B) + Cc * Cc;
        lev += 0.5f * log2f(l2b + sqrtf(l2n));
    }
    return lev;
}

__device__ __forceinline__ LevelSel mip_select(float lev, int L, int nearest)
{
    LevelSel s;
    s.grad = false; s.f = 0.f;
    if (!(lev > 0.f)) { s.l0 = s.l1 = 0; return s; }          // also NaN / -inf (degenerate footprint)
    if (lev >= (float)L) { s.l0 = s.l1 = L; return s; }
    if (nearest) { s.l0 = s.l1 = min((int)floorf(lev + 0.5f), L); return s; }
    s.l0 = (int)floorf(lev); s.l1 = s.l0 + 1; s.f = lev - (float)s.l0; s.grad = true;
    return s;
}

__device__ __forceinline__ const float* level_ptr(const float* tex, const float* mip, const MipDesc& md, int l, int n, int C)
{
    return (l == 0 ? tex : mip + md.off[l]) + (size_t)n * md.h[l] * md.w[l] * C;
}

__global__ void __launch_bounds__(256) k_tex_mip_fwd(const float* __restrict__ tex, const float* __restrict__ mip, MipDesc md, int Nt, int C,
                                                     const float* __restrict__ uv, const float* __restrict__ uv_da, const float* __restrict__ bias,
                                                     int nearest, long long npx_total, long long npx_inst, float* __restrict__ out)
{
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= npx_total) return;
    float2 q = __ldg(reinterpret_cast<const float2*>(uv) + pi);
    float4 da = uv_da ? ldg4(uv_da + 4 * pi) : make_float4(0.f, 0.f, 0.f, 0.f);
    LevelSel s = mip_select(mip_level_raw(da, md.w[0], md.h[0], bias ? __ldg(bias + pi) : 0.f, uv_da != nullptr), md.L, nearest);
    const int n = Nt > 1 ? (int)(pi / npx_inst) : 0;
    float* o = out + pi * C;
    Bilin f0 = bilin_index(q.x, q.y, md.w[s.l0], md.h[s.l0]);
    const float* t0 = level_ptr(tex, mip, md, s.l0, n, C);
    for (int c = 0; c < C; c++) {
        float t00 = __ldg(t0 + (size_t)f0.i00 * C + c), t10 = __ldg(t0 + (size_t)f0.i10 * C + c);
        float t01 = __ldg(t0 + (size_t)f0.i01 * C + c), t11 = __ldg(t0 + (size_t)f0.i11 * C + c);
        float a = t00 + (t10 - t00) * f0.fx, b = t01 + (t11 - t01) * f0.fx;
        o[c] = a + (b - a) * f0.fy;
    }
    if (s.l1 != s.l0) {
        Bilin f1 = bilin_index(q.x, q.y, md.w[s.l1], md.h[s.l1]);
        const float* t1 = level_ptr(tex, mip, md, s.l1, n, C);
        for (int c = 0; c < C; c++) {
            float t00 = __ldg(t1 + (size_t)f1.i00 * C + c), t10 = __ldg(t1 + (size_t)f1.i10 * C + c);
            float t01 = __ldg(t1 + (size_t)f1.i01 * C + c), t11 = __ldg(t1 + (size_t)f1.i11 * C + c);
            float a = t00 + (t10 - t00) * f1.fx, b = t01 + (t11 - t01) * f1.fx;
            float c1 = a + (b - a) * f1.fy;
            o[c] = o[c] + (c1 - o[c]) * s.f;
        }
    }
}

// one level's share of the backward pass: scatters weight * dy into the level's gradient, returns sum_c dy_c colour_c and
// accumulates d / d(u, v)
__device__ __forceinline__ float tex_level_bwd(const float* __restrict__ t, float* __restrict__ gt, const Bilin& f, int Wl, int Hl, int C,
                                               const float* __restrict__ d, float weight, float& gu, float& gv)
{
    const float w00 = (1.f - f.fx) * (1.f - f.fy), w10 = f.fx * (1.f - f.fy), w01 = (1.f - f.fx) * f.fy, w11 = f.fx * f.fy;
    float dot = 0.f, su = 0.f, sv = 0.f;
    for (int c = 0; c < C; c++) {
        float g = __ldg(d + c);
        float t00 = __ldg(t + (size_t)f.i00 * C + c), t10 = __ldg(t + (size_t)f.i10 * C + c);
        float t01 = __ldg(t + (size_t)f.i01 * C + c), t11 = __ldg(t + (size_t)f.i11 * C + c);
        float a = t00 + (t10 - t00) * f.fx, b = t01 + (t11 - t01) * f.fx;
        dot += g * (a + (b - a) * f.fy);
        su += g * ((t10 - t00) * (1.f - f.fy) + (t11 - t01) * f.fy);
        sv += g * ((t01 - t00) * (1.f - f.fx) + (t11 - t10) * f.fx);
        const float gw = g * weight;
        if (gt && gw != 0.f) {
            atomicAdd(gt + (size_t)f.i00 * C + c, gw * w00);
            atomicAdd(gt + (size_t)f.i10 * C + c, gw * w10);
            atomicAdd(gt + (size_t)f.i01 * C + c, gw * w01);
            atomicAdd(gt + (size_t)f.i11 * C + c, gw * w11);
        }
    }
    gu += weight * su * (float)Wl;
    gv += weight * sv * (float)Hl;
    return dot;
}

__global__ void __launch_bounds__(256) k_tex_mip_bwd(const float* __restrict__ tex, const float* __restrict__ mip, MipDesc md, int Nt, int C,
                                                     const float* __restrict__ uv, const float* __restrict__ uv_da, const float* __restrict__ bias,
                                                     int nearest, const float* __restrict__ dy, long long npx_total, long long npx_inst,
                                                     float* __restrict__ g_tex, float* __restrict__ g_mip, float* __restrict__ g_uv,
                                                     float* __restrict__ g_da, float* __restrict__ g_bias)
{
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= npx_total) return;
    float2 q = __ldg(reinterpret_cast<const float2*>(uv) + pi);
    float4 da = uv_da ? ldg4(uv_da + 4 * pi) : make_float4(0.f, 0.f, 0.f, 0.f);
    LevelSel s = mip_select(mip_level_raw(da, md.w[0], md.h[0], bias ? __ldg(bias + pi) : 0.f, uv_da != nullptr), md.L, nearest);
    const int n = Nt > 1 ? (int)(pi / npx_inst) : 0;
    const float* d = dy + pi * C;
    float gu = 0.f, gv = 0.f;
    auto grad_ptr = [&](int l) -> float* {
        if (!g_tex) return nullptr;
        return (l == 0 ? g_tex : g_mip + md.off[l]) + (size_t)n * md.h[l] * md.w[l] * C;
    };
    Bilin f0 = bilin_index(q.x, q.y, md.w[s.l0], md.h[s.l0]);
    float dot0 = tex_level_bwd(level_ptr(tex, mip, md, s.l0, n, C), grad_ptr(s.l0), f0, md.w[s.l0], md.h[s.l0], C, d, 1.f - s.f, gu, gv);
    float glev = 0.f;
    if (s.l1 != s.l0) {
        Bilin f1 = bilin_index(q.x, q.y, md.w[s.l1], md.h[s.l1]);
        float dot1 = tex_level_bwd(level_ptr(tex, mip, md, s.l1, n, C), grad_ptr(s.l1), f1, md.w[s.l1], md.h[s.l1], C, d, s.f, gu, gv);
        glev = dot1 - dot0;                                   // d out / d level = c1 - c0
    }
    reinterpret_cast<float2*>(g_uv)[pi] = make_float2(gu, gv);
    if (g_bias) g_bias[pi] = s.grad ? glev : 0.f;
    if (g_da) {
        float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s.grad && glev != 0.f) {
            const float Wf = (float)md.w[0], Hf = (float)md.h[0];
            float dsdx = da.x * Wf, dsdy = da.y * Wf, dtdx = da.z * Hf, dtdy = da.w * Hf;
            float A = dsdx * dsdx + dtdx * dtdx, B = dsdy * dsdy + dtdy * dtdy, Cc = dsdx * dsdy + dtdx * dtdy;
            float l2a = sqrtf(0.25f * (A - B) * (A - B) + Cc * Cc), major = 0.5f * (A + B) + l2a;
            if (l2a > 0.f && major > 0.f) {
                float gm = glev * 0.5f / (0.6931471805599453f * major);       // d level / d major
                float gA = gm * (0.5f + 0.25f * (A - B) / l2a), gB = gm * (0.5f - 0.25f * (A - B) / l2a), gC = gm * Cc / l2a;
                gd.x = (2.f * dsdx * gA + dsdy * gC) * Wf;
                gd.y = (2.f * dsdy * gB + dsdx * gC) * Wf;
                gd.z = (2.f * dtdx * gA + dtdy * gC) * Hf;
                gd.w = (2.f * dtdy * gB + dtdx * gC) * Hf;
            }
        }
        reinterpret_cast<float4*>(g_da)[pi] = gd;
    }
}

int check_sel(const char* who, const int32_t* diff_attrs, int K, int A, DiffSel& sel)
{
    if (!diff_attrs) K = A;
    FPC_CHECK_ARG(K >= 0 && K <= MAX_DIFF, "%s: at most %d differentiated attributes (got %d)", who, MAX_DIFF, K);
    sel.k = K;
    for (int j = 0; j < K; j++) {
        sel.idx[j] = diff_attrs ? diff_attrs[j] : j;
        FPC_CHECK_ARG(sel.idx[j] >= 0 && sel.idx[j] < A, "%s: diff_attrs[%d] = %d is outside [0, %d)", who, j, sel.idx[j], A);
    }
    return FPC_OK;
}

}  // namespace

extern "C" int fpc_interpolate_da_fwd(const float* attr, int Na, int Vt, int A, const float* rast, const float* rast_db, const int32_t* tri,
                                      const int32_t* diff_attrs, int K, int N, int T, int H, int W, float* out, float* out_da,
                                      fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(attr && rast && rast_db && tri && out && out_da, "interpolate_da_fwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && T > 0 && H > 0 && W > 0 && Vt > 0 && A > 0, "interpolate_da_fwd: sizes must be positive");
    FPC_CHECK_ARG(Na == 1 || Na == N, "interpolate_da_fwd: attr batch must be 1 or N (got %d, N=%d)", Na, N);
    DiffSel sel;
    int st = check_sel("interpolate_da_fwd", diff_attrs, K, A, sel);
    if (st != FPC_OK) return st;
    long long npx_inst = (long long)H * W, npx = npx_inst * N;
    k_interp_da_fwd<<<fpc_div_up(npx, 256), 256, 0, stream>>>(attr, Na == 1 ? 0 : Vt * A, Vt, A, rast, rast_db, tri, sel, npx, npx_inst, T, out, out_da);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_interpolate_da_bwd(const float* attr, int Na, int Vt, int A, const float* rast, const float* rast_db, const int32_t* tri,
                                      const int32_t* diff_attrs, int K, const float* dy, const float* dda, int N, int T, int H, int W,
                                      float* grad_attr, float* grad_rast, float* grad_rast_db, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(attr && rast && rast_db && tri && grad_attr && grad_rast, "interpolate_da_bwd: null pointer argument");
    FPC_CHECK_ARG(dy || dda, "interpolate_da_bwd: at least one of dy, dda must be given");
    FPC_CHECK_ARG(N > 0 && T > 0 && H > 0 && W > 0 && Vt > 0 && A > 0, "interpolate_da_bwd: sizes must be positive");
    FPC_CHECK_ARG(Na == 1 || Na == N, "interpolate_da_bwd: attr batch must be 1 or N (got %d, N=%d)", Na, N);
    DiffSel sel;
    int st = check_sel("interpolate_da_bwd", diff_attrs, K, A, sel);
    if (st != FPC_OK) return st;
    long long npx_inst = (long long)H * W, npx = npx_inst * N;
    FPC_CUDA(cudaMemsetAsync(grad_attr, 0, (size_t)Na * Vt * A * sizeof(float), stream));
    k_interp_da_bwd<<<fpc_div_up(npx, 256), 256, 0, stream>>>(attr, Na == 1 ? 0 : Vt * A, Vt, A, rast, rast_db, tri, sel, dy, dda, npx, npx_inst, T,
                                                               grad_attr, grad_rast, grad_rast_db);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_rasterize_bwd_db(const float* pos, const int32_t* tri, const float* rast, const float* d_rast, const float* d_rast_db,
                                    int N, int V, int T, int H, int W, float* grad_pos, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(pos && tri && rast && d_rast && d_rast_db && grad_pos, "rasterize_bwd_db: null pointer argument");
    FPC_CHECK_ARG(N > 0 && V > 0 && T > 0 && H > 0 && W > 0, "rasterize_bwd_db: sizes must be positive");
    FPC_CUDA(cudaMemsetAsync(grad_pos, 0, (size_t)N * V * 4 * sizeof(float), stream));
    long long npx = (long long)N * H * W;
    k_raster_bwd_db<<<fpc_div_up(npx, 256), 256, 0, stream>>>(pos, tri, rast, d_rast, d_rast_db, N, V, T, H, W, 2.f / (float)W,
                                                              1.f / (float)W - 1.f, 2.f / (float)H, 1.f / (float)H - 1.f, grad_pos);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_texture_mip_levels(int Ht, int Wt, int max_mip_level)
{
    int L = 0;
    while (L < MAX_LEVELS && (max_mip_level < 0 || L < max_mip_level) && Ht >= 2 && Wt >= 2 && Ht % 2 == 0 && Wt % 2 == 0) {
        Ht >>= 1; Wt >>= 1; L++;
    }
    return L;
}

extern "C" size_t fpc_texture_mip_floats(int Nt, int Ht, int Wt, int C, int L)
{
    size_t total = 0;
    for (int l = 1; l <= L && l <= MAX_LEVELS; l++) total += (size_t)Nt * (Ht >> l) * (Wt >> l) * C;
    return total > 0 ? total : 1;
}

static int check_mip_args(const char* who, int Nt, int Ht, int Wt, int C, int L)
{
    FPC_CHECK_ARG(Nt > 0 && Ht > 0 && Wt > 0 && C > 0, "%s: texture sizes must be positive", who);
    FPC_CHECK_ARG(L >= 0 && L <= fpc_texture_mip_levels(Ht, Wt, -1),
                  "%s: %d mip levels are not available for a %dx%d texture (extents must stay even down to the last level)", who, L, Wt, Ht);
    return FPC_OK;
}

extern "C" int fpc_texture_mip_build(const float* tex, int Nt, int Ht, int Wt, int C, int L, float* mip, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(tex && (mip || L == 0), "texture_mip_build: null pointer argument");
    int st = check_mip_args("texture_mip_build", Nt, Ht, Wt, C, L);
    if (st != FPC_OK) return st;
    MipDesc md = mip_desc(Nt, Ht, Wt, C, L);
    for (int l = 1; l <= L; l++) {
        const float* src = l == 1 ? tex : mip + md.off[l - 1];
        long long n = (long long)Nt * md.h[l] * md.w[l] * C;
        k_mip_down<<<fpc_div_up(n, 256), 256, 0, stream>>>(src, mip + md.off[l], Nt, md.h[l], md.w[l], C);
        FPC_LAUNCH_CHECK();
    }
    return FPC_OK;
}

extern "C" int fpc_texture_mip_fwd(const float* tex, const float* mip, int Nt, int Ht, int Wt, int C, int L, const float* uv,
                                   const float* uv_da, const float* mip_level_bias, int nearest_level, int N, int H, int W, float* out,
                                   fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(tex && uv && out && (mip || L == 0), "texture_mip_fwd: null pointer argument");
    FPC_CHECK_ARG(uv_da || mip_level_bias, "texture_mip_fwd: uv_da or mip_level_bias must be given");
    FPC_CHECK_ARG(N > 0 && H > 0 && W > 0 && (Nt == 1 || Nt == N), "texture_mip_fwd: bad sizes (tex batch must be 1 or N)");
    int st = check_mip_args("texture_mip_fwd", Nt, Ht, Wt, C, L);
    if (st != FPC_OK) return st;
    long long npx_inst = (long long)H * W, npx = npx_inst * N;
    k_tex_mip_fwd<<<fpc_div_up(npx, 256), 256, 0, stream>>>(tex, mip, mip_desc(Nt, Ht, Wt, C, L), Nt, C, uv, uv_da, mip_level_bias, nearest_level,
                                                            npx, npx_inst, out);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_texture_mip_bwd(const float* tex, const float* mip, int Nt, int Ht, int Wt, int C, int L, const float* uv,
                                   const float* uv_da, const float* mip_level_bias, int nearest_level, const float* dy, int N, int H, int W,
                                   float* grad_tex, float* grad_mip, int mip_is_constant, float* grad_uv, float* grad_uv_da, float* grad_bias,
                                   fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(tex && uv && dy && grad_uv && (mip || L == 0), "texture_mip_bwd: null pointer argument");
    FPC_CHECK_ARG(uv_da || mip_level_bias, "texture_mip_bwd: uv_da or mip_level_bias must be given");
    FPC_CHECK_ARG(!grad_tex || grad_mip || L == 0, "texture_mip_bwd: grad_tex needs grad_mip (scratch of fpc_texture_mip_floats floats)");
    FPC_CHECK_ARG(!grad_uv_da || uv_da, "texture_mip_bwd: grad_uv_da without uv_da");
    FPC_CHECK_ARG(!grad_bias || mip_level_bias, "texture_mip_bwd: grad_bias without mip_level_bias");
    FPC_CHECK_ARG(N > 0 && H > 0 && W > 0 && (Nt == 1 || Nt == N), "texture_mip_bwd: bad sizes (tex batch must be 1 or N)");
    int st = check_mip_args("texture_mip_bwd", Nt, Ht, Wt, C, L);
    if (st != FPC_OK) return st;
    MipDesc md = mip_desc(Nt, Ht, Wt, C, L);
    if (grad_tex) {
        FPC_CUDA(cudaMemsetAsync(grad_tex, 0, (size_t)Nt * Ht * Wt * C * sizeof(float), stream));
        if (L > 0) FPC_CUDA(cudaMemsetAsync(grad_mip, 0, fpc_texture_mip_floats(Nt, Ht, Wt, C, L) * sizeof(float), stream));
    }
    long long npx_inst = (long long)H * W, npx = npx_inst * N;
    k_tex_mip_bwd<<<fpc_div_up(npx, 256), 256, 0, stream>>>(tex, mip, md, Nt, C, uv, uv_da, mip_level_bias, nearest_level, dy, npx, npx_inst,
                                                            grad_tex, grad_mip, grad_uv, grad_uv_da, grad_bias);
    FPC_LAUNCH_CHECK();
    if (grad_tex && !mip_is_constant) {
        // fold the gradients of the coarse levels down the chain: level L -> L-1 -> ... -> the texture itself
        for (int l = L; l >= 1; l--) {
            float* fine = l == 1 ? grad_tex : grad_mip + md.off[l - 1];
            long long n = (long long)Nt * 4 * md.h[l] * md.w[l] * C;
            k_mip_up<<<fpc_div_up(n, 256), 256, 0, stream>>>(grad_mip + md.off[l], fine, Nt, md.h[l], md.w[l], C);
            FPC_LAUNCH_CHECK();
        }
    }
    return FPC_OK;
}
