/*
 * oracle/golden.c — scalar CPU golden model of the four differentiable-rendering ops on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fpc_diffrend_b200/ may call into this file; it is linked
 * and executed only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the arithmetic of these ops lives in NVlabs/nvdiffrast (un-vendored, un-pinned
 * dependency of /root/reference: src/torch/fit.py:13,151-160), which is not present in this environment
 * and for which the reference holds no tests or golden vectors.  This file restates
 * the published algorithm (Laine et al. 2020, "Modular Primitives for High-Performance Differentiable
 * Rendering", §3) and the op semantics of SURVEY.md Appendix A, anchored on the reference's call sites:
 *   rasterize   fit.py:151   -> gold_rasterize_fwd / gold_rasterize_bwd      (App. A.1)
 *   interpolate fit.py:157   -> gold_interpolate_fwd / gold_interpolate_bwd  (App. A.2)
 *   texture     fit.py:158   -> gold_texture_linear_fwd / _bwd               (App. A.3, 'linear', wrap)
 *   antialias   fit.py:160   -> gold_topology_build / gold_antialias_fwd / _bwd (App. A.4)
 * Forward and backward passes run the views of a batch in parallel (OpenMP); shared-tensor gradients are summed in view order.
 * Its analytic backward passes are themselves pinned against a float64 torch-autograd restatement
 * (oracle/torch_ref.py) in tests/test_oracle.py.
 *
 * Determinism contract shared with the CUDA path (DESIGN.md "Rasterizer semantics"):
 *   - window coords are snapped to 1/16 px with round-half-even of  (x * (1/w)) * (8*W) + (8*W)
 *     evaluated as separate IEEE fp32 mul / add (no FMA contraction: compile with -ffp-contract=off);
 *   - coverage: pixel centre (16*px+8, 16*py+8) against integer edge functions, shared edges owned by
 *     exactly one side (rule in edge_bias());  no back-face culling;  zero-area triangles dropped;
 *   - visibility: smallest depth wins (LESS), ties keep the lower triangle index;  the depth of a fragment
 *     is the triangle's screen-space z/w PLANE (z/w is affine in window space; upstream's CUDA rasterizer
 *     also depth-tests on a per-triangle plane equation) evaluated as in depth_plane() / plane_eval():
 *     per-vertex z/w in fp32, plane set-up in fp64 on the snapped vertices, per-fragment evaluation with two
 *     fp32 FMAs;  fragments whose plane depth is outside [-1,1] are discarded.  The z/w written to
 *     rast[...,2] is the fp32 value of the shading formula (what upstream's shader pass writes);
 *   - triangles with a vertex at w <= 0 are clipped against the near plane z + w >= 0 (clip_near()); the pieces keep
 *     the parent's id and are shaded with the parent's vertices.  A (sub-)triangle with a window coordinate beyond
 *     +-2^20 px is dropped (no guard-band clipper).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SUBPIX 16
#define SNAP_LIMIT 16777216.0f /* 2^24 sub-pixel units = 2^20 px */

typedef struct { float x, y, z, w; } f4;

/* ------------------------------------------------------------------------------------------------ */
/* rasterize                                                                                        */
/* ------------------------------------------------------------------------------------------------ */

static int snap_vertex(f4 p, int W, int H, int32_t* sx, int32_t* sy)
{
    if (!(p.w > 0.0f)) return 0;
    float rw = 1.0f / p.w;
    float sxs = 8.0f * (float)W, sys = 8.0f * (float)H;
    float xf = (p.x * rw) * sxs + sxs;
    float yf = (p.y * rw) * sys + sys;
    if (!(fabsf(xf) < SNAP_LIMIT) || !(fabsf(yf) < SNAP_LIMIT)) return 0;
    *sx = (int32_t)lrintf(xf);
    *sy = (int32_t)lrintf(yf);
    return 1;
}

/* A sample exactly on an edge belongs to the triangle iff the edge direction (dx,dy) (triangle oriented
 * to positive area) satisfies dy > 0 || (dy == 0 && dx < 0).  The reversed edge gives the complement,
 * so a shared edge is drawn exactly once. */
static inline int64_t edge_bias(int64_t dx, int64_t dy)
{
    return (dy > 0 || (dy == 0 && dx < 0)) ? 0 : -1;
}

/* Depth plane of a triangle.  (x_i,y_i): snapped vertices in ORIGINAL order, (pxa,pya): first candidate pixel
 * of the triangle's image-clamped bbox (the plane's reference point).  All fp64 products of the snapped
 * integer coordinates are exact; every fp64 op is separate (no contraction). */
typedef struct { float zref, dzdx, dzdy; } plane_t;

static inline plane_t depth_plane(f4 p0, f4 p1, f4 p2, int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                                  int32_t x2, int32_t y2, int pxa, int pya)
{
    float zv0 = p0.z / p0.w, zv1 = p1.z / p1.w, zv2 = p2.z / p2.w;
    double X1 = (double)(x1 - x0), Y1 = (double)(y1 - y0), X2 = (double)(x2 - x0), Y2 = (double)(y2 - y0);
    double A = X1 * Y2 - X2 * Y1;
    double dz1 = (double)zv1 - (double)zv0, dz2 = (double)zv2 - (double)zv0;
    double gx = (dz1 * Y2 - dz2 * Y1) / A;                 /* d(z/w) per sub-pixel unit in x */
    double gy = (dz2 * X1 - dz1 * X2) / A;
    double rx = (double)(16 * pxa + 8 - x0), ry = (double)(16 * pya + 8 - y0);
    plane_t pl;
    pl.zref = (float)(((double)zv0 + gx * rx) + gy * ry);
    pl.dzdx = (float)(gx * 16.0);
    pl.dzdy = (float)(gy * 16.0);
    return pl;
}

static inline float plane_eval(plane_t pl, int dx, int dy)
{
    return fmaf(pl.dzdx, (float)dx, fmaf(pl.dzdy, (float)dy, pl.zref));
}

typedef struct { float u, v, zw; float db[4]; } shade_t;

/* The shading formula (App. A.1).  Every operation is a separate IEEE fp32 op in this order. */
static inline int shade_pixel(f4 p0, f4 p1, f4 p2, int px, int py, int W, int H, shade_t* s, int want_db)
{
    float xs = 2.0f / (float)W, xo = 1.0f / (float)W - 1.0f;
    float ys = 2.0f / (float)H, yo = 1.0f / (float)H - 1.0f;
    float fx = xs * (float)px + xo;
    float fy = ys * (float)py + yo;
    float p0x = p0.x - fx * p0.w, p0y = p0.y - fy * p0.w;
    float p1x = p1.x - fx * p1.w, p1y = p1.y - fy * p1.w;
    float p2x = p2.x - fx * p2.w, p2y = p2.y - fy * p2.w;
    float a0 = p1x * p2y - p1y * p2x;
    float a1 = p2x * p0y - p2y * p0x;
    float a2 = p0x * p1y - p0y * p1x;
    float at = (a0 + a1) + a2;
    float iw = 1.0f / at;
    float b0 = a0 * iw, b1 = a1 * iw;
    float z = (p0.z * a0 + p1.z * a1) + p2.z * a2;
    float w = (p0.w * a0 + p1.w * a1) + p2.w * a2;
    float zw = z / w;
    s->zw = zw;
    s->u = fminf(fmaxf(b0, 0.0f), 1.0f);
    s->v = fminf(fmaxf(b1, 0.0f), 1.0f);
    if (want_db) {
        float dfxdx = xs * iw, dfydy = ys * iw;
        float da0dx = p2.y * p1.w - p1.y * p2.w, da0dy = p1.x * p2.w - p2.x * p1.w;
        float da1dx = p0.y * p2.w - p2.y * p0.w, da1dy = p2.x * p0.w - p0.x * p2.w;
        float da2dx = p1.y * p0.w - p0.y * p1.w, da2dy = p0.x * p1.w - p1.x * p0.w;
        float datdx = (da0dx + da1dx) + da2dx, datdy = (da0dy + da1dy) + da2dy;
        s->db[0] = dfxdx * (b0 * datdx - da0dx);
        s->db[1] = dfydy * (b0 * datdy - da0dy);
        s->db[2] = dfxdx * (b1 * datdx - da1dx);
        s->db[3] = dfydy * (b1 * datdy - da1dy);
    }
    return zw >= -1.0f && zw <= 1.0f; /* false for NaN too */
}

/* rast [N,H,W,4] = (u, v, z/w, tri_id+1), rast_db [N,H,W,4] or NULL,
 * second_zw [N,H,W,2] or NULL: plane depth of the winning and of the runner-up fragment (2.0 if none) so tests
 * can mask depth near-ties. */
/* Coverage + depth test of one (sub-)triangle with vertices p0,p1,p2 on behalf of triangle id t. */
static void raster_one(f4 p0, f4 p1, f4 p2, int t, int H, int W, float* depth, float* depth2, int32_t* id)
{
    int32_t x0, y0, x1, y1, x2, y2;
    if (!snap_vertex(p0, W, H, &x0, &y0) || !snap_vertex(p1, W, H, &x1, &y1) ||
        !snap_vertex(p2, W, H, &x2, &y2)) return;
    int64_t area = (int64_t)(x1 - x0) * (y2 - y0) - (int64_t)(x2 - x0) * (y1 - y0);
    if (area == 0) return;
    const int32_t qx1 = x1, qy1 = y1, qx2 = x2, qy2 = y2;   /* original order, for the depth plane */
    if (area < 0) { int32_t tx = x1, ty = y1; x1 = x2; y1 = y2; x2 = tx; y2 = ty; }
    int32_t minx = x0 < x1 ? (x0 < x2 ? x0 : x2) : (x1 < x2 ? x1 : x2);
    int32_t maxx = x0 > x1 ? (x0 > x2 ? x0 : x2) : (x1 > x2 ? x1 : x2);
    int32_t miny = y0 < y1 ? (y0 < y2 ? y0 : y2) : (y1 < y2 ? y1 : y2);
    int32_t maxy = y0 > y1 ? (y0 > y2 ? y0 : y2) : (y1 > y2 ? y1 : y2);
    /* pixel px is a candidate iff minx <= 16*px+8 <= maxx */
    int pxa = (minx - 8 + 15) >> 4, pxb = (maxx - 8) >> 4;
    int pya = (miny - 8 + 15) >> 4, pyb = (maxy - 8) >> 4;
    if (pxa < 0) pxa = 0;
    if (pya < 0) pya = 0;
    if (pxb > W - 1) pxb = W - 1;
    if (pyb > H - 1) pyb = H - 1;
    if (pxa > pxb || pya > pyb) return;
    int64_t ex[3] = { x1 - x0, x2 - x1, x0 - x2 };
    int64_t ey[3] = { y1 - y0, y2 - y1, y0 - y2 };
    int32_t ox[3] = { x0, x1, x2 }, oy[3] = { y0, y1, y2 };
    int64_t bias[3];
    for (int e = 0; e < 3; e++) bias[e] = edge_bias(ex[e], ey[e]);
    plane_t pl = depth_plane(p0, p1, p2, x0, y0, qx1, qy1, qx2, qy2, pxa, pya);
    for (int py = pya; py <= pyb; py++) {
        for (int px = pxa; px <= pxb; px++) {
            int64_t sx = 16 * px + 8, sy = 16 * py + 8;
            int inside = 1;
            for (int e = 0; e < 3; e++) {
                int64_t E = ex[e] * (sy - oy[e]) - ey[e] * (sx - ox[e]) + bias[e];
                if (E < 0) { inside = 0; break; }
            }
            if (!inside) continue;
            float zd = plane_eval(pl, px - pxa, py - pya);
            if (!(zd >= -1.0f && zd <= 1.0f)) continue;
            size_t pi = (size_t)py * W + px;
            if (zd < depth[pi]) { depth2[pi] = depth[pi]; depth[pi] = zd; id[pi] = t; }
            else if (zd < depth2[pi]) depth2[pi] = zd;
        }
    }
}

/* Near-plane clipper (triangles with a vertex at w <= 0 are clipped, the pieces keep the parent id).
 * Sutherland-Hodgman against z + w >= 0; a crossing is always interpolated from the INSIDE vertex towards the outside
 * one, t = d_in / (d_in - d_out), so both triangles sharing the edge produce the same point (watertight).  out[4]. */
static inline f4 clip_lerp(f4 a, float da, f4 b, float db)
{
    float t = da / (da - db);
    f4 r;
    r.x = a.x + t * (b.x - a.x);
    r.y = a.y + t * (b.y - a.y);
    r.z = a.z + t * (b.z - a.z);
    r.w = a.w + t * (b.w - a.w);
    return r;
}

static int clip_near(const f4* v, f4* out)
{
    float d[3];
    int in[3], n = 0;
    for (int i = 0; i < 3; i++) { d[i] = v[i].z + v[i].w; in[i] = d[i] >= 0.0f; }
    for (int i = 0; i < 3; i++) {
        int j = (i + 1) % 3;
        if (in[i]) out[n++] = v[i];
        if (in[i] != in[j]) out[n++] = in[i] ? clip_lerp(v[i], d[i], v[j], d[j]) : clip_lerp(v[j], d[j], v[i], d[i]);
    }
    return n;
}

void gold_rasterize_fwd(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                        float* rast, float* rast_db, float* second_zw)
{
    size_t npx = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; n++) {
        const f4* P = (const f4*)pos + (size_t)n * V;
        float* depth = (float*)malloc(npx * sizeof(float));
        float* depth2 = (float*)malloc(npx * sizeof(float));
        int32_t* id = (int32_t*)malloc(npx * sizeof(int32_t));
        for (size_t i = 0; i < npx; i++) { depth[i] = 2.0f; depth2[i] = 2.0f; id[i] = -1; }
        for (int t = 0; t < T; t++) {
            int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
            if (i0 < 0 || i0 >= V || i1 < 0 || i1 >= V || i2 < 0 || i2 >= V) continue;
            f4 v[3] = { P[i0], P[i1], P[i2] };
            if (v[0].w > 0.0f && v[1].w > 0.0f && v[2].w > 0.0f) {
                raster_one(v[0], v[1], v[2], t, H, W, depth, depth2, id);
            } else {
                f4 c[4];
                int nc = clip_near(v, c);
                if (nc >= 3) raster_one(c[0], c[1], c[2], t, H, W, depth, depth2, id);
                if (nc == 4) raster_one(c[0], c[2], c[3], t, H, W, depth, depth2, id);
            }
        }
        for (int py = 0; py < H; py++) {
            for (int px = 0; px < W; px++) {
                size_t pi = (size_t)py * W + px;
                float* o = rast + ((size_t)n * npx + pi) * 4;
                float* odb = rast_db ? rast_db + ((size_t)n * npx + pi) * 4 : NULL;
                if (second_zw) { second_zw[((size_t)n * npx + pi) * 2] = depth[pi]; second_zw[((size_t)n * npx + pi) * 2 + 1] = depth2[pi]; }
                int t = id[pi];
                if (t < 0) {
                    o[0] = o[1] = o[2] = o[3] = 0.0f;
                    if (odb) odb[0] = odb[1] = odb[2] = odb[3] = 0.0f;
                    continue;
                }
                shade_t s;
                shade_pixel(P[tri[3 * t]], P[tri[3 * t + 1]], P[tri[3 * t + 2]], px, py, W, H, &s, odb != NULL);
                o[0] = s.u; o[1] = s.v; o[2] = fminf(fmaxf(s.zw, -1.0f), 1.0f); o[3] = (float)(t + 1);
                if (odb) { odb[0] = s.db[0]; odb[1] = s.db[1]; odb[2] = s.db[2]; odb[3] = s.db[3]; }
            }
        }
        free(depth); free(depth2); free(id);
    }
}

/* d pos[x,y,w] from (d u, d v) = dy[...,0:2]; formula differentiated without the clamps; d(z/w), d(id) and
 * d pos.z are zero (App. A.1 "Backward").  grad_pos [N,V,4] is overwritten.  Accumulates in double. */
void gold_rasterize_bwd(const float* pos, const int32_t* tri, const float* rast, const float* dy,
                        int N, int V, int T, int H, int W, float* grad_pos)
{
    size_t npx = (size_t)H * W;
    double* acc = (double*)calloc((size_t)N * V * 4, sizeof(double));
    float xs = 2.0f / (float)W, xo = 1.0f / (float)W - 1.0f;
    float ys = 2.0f / (float)H, yo = 1.0f / (float)H - 1.0f;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; n++) {
        const f4* P = (const f4*)pos + (size_t)n * V;
        double* A = acc + (size_t)n * V * 4;
        for (int py = 0; py < H; py++) for (int px = 0; px < W; px++) {
            size_t pi = (size_t)n * npx + (size_t)py * W + px;
            int t = (int)rast[pi * 4 + 3] - 1;
            if (t < 0 || t >= T) continue;
            double gu = dy[pi * 4 + 0], gv = dy[pi * 4 + 1];
            if (gu == 0.0 && gv == 0.0) continue;
            int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
            f4 q0 = P[i0], q1 = P[i1], q2 = P[i2];
            double fx = (double)(xs * (float)px + xo), fy = (double)(ys * (float)py + yo);
            double p0x = q0.x - fx * q0.w, p0y = q0.y - fy * q0.w;
            double p1x = q1.x - fx * q1.w, p1y = q1.y - fy * q1.w;
            double p2x = q2.x - fx * q2.w, p2y = q2.y - fy * q2.w;
            double a0 = p1x * p2y - p1y * p2x, a1 = p2x * p0y - p2y * p0x, a2 = p0x * p1y - p0y * p1x;
            double iw = 1.0 / (a0 + a1 + a2);
            double u = a0 * iw, v = a1 * iw;
            double gbb = gu * u + gv * v;
            double g0 = iw * (gu - gbb), g1 = iw * (gv - gbb), g2 = -iw * gbb;
            double g0x = -g1 * p2y + g2 * p1y, g0y = g1 * p2x - g2 * p1x;
            double g1x = g0 * p2y - g2 * p0y, g1y = -g0 * p2x + g2 * p0x;
            double g2x = -g0 * p1y + g1 * p0y, g2y = g0 * p1x - g1 * p0x;
            A[i0 * 4 + 0] += g0x; A[i0 * 4 + 1] += g0y; A[i0 * 4 + 3] += -fx * g0x - fy * g0y;
            A[i1 * 4 + 0] += g1x; A[i1 * 4 + 1] += g1y; A[i1 * 4 + 3] += -fx * g1x - fy * g1y;
            A[i2 * 4 + 0] += g2x; A[i2 * 4 + 1] += g2y; A[i2 * 4 + 3] += -fx * g2x - fy * g2y;
        }
    }
    for (size_t i = 0; i < (size_t)N * V * 4; i++) grad_pos[i] = (float)acc[i];
    free(acc);
}

/* ------------------------------------------------------------------------------------------------ */
/* interpolate                                                                                      */
/* ------------------------------------------------------------------------------------------------ */

/* attr [Na,Vt,A] with Na in {1,N};  out [N,H,W,A] */
void gold_interpolate_fwd(const float* attr, int Na, int Vt, int A, const float* rast, const int32_t* tri,
                          int N, int T, int H, int W, float* out)
{
    size_t npx = (size_t)H * W;
#pragma omp parallel for
    for (int n = 0; n < N; n++) {
        const float* at = attr + (Na > 1 ? (size_t)n * Vt * A : 0);
        for (size_t p = 0; p < npx; p++) {
            const float* r = rast + ((size_t)n * npx + p) * 4;
            float* o = out + ((size_t)n * npx + p) * A;
            int t = (int)r[3] - 1;
            if (t < 0 || t >= T) { for (int c = 0; c < A; c++) o[c] = 0.0f; continue; }
            int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
            if (i0 < 0 || i0 >= Vt || i1 < 0 || i1 >= Vt || i2 < 0 || i2 >= Vt) {
                for (int c = 0; c < A; c++) o[c] = 0.0f;
                continue;
            }
            float b0 = r[0], b1 = r[1], b2 = 1.0f - r[0] - r[1];
            for (int c = 0; c < A; c++)
                o[c] = b0 * at[i0 * A + c] + b1 * at[i1 * A + c] + b2 * at[i2 * A + c];
        }
    }
}

/* g_attr [Na,Vt,A] (summed over N when Na==1), g_rast [N,H,W,4] = (du, dv, 0, 0) */
void gold_interpolate_bwd(const float* attr, int Na, int Vt, int A, const float* rast, const int32_t* tri,
                          const float* dy, int N, int T, int H, int W, float* g_attr, float* g_rast)
{
    size_t npx = (size_t)H * W;
    size_t na = (size_t)(Na > 1 ? N : 1) * Vt * A;
    size_t per = (size_t)Vt * A;
    /* one accumulator per view (views run in parallel); a shared attribute tensor sums them in view order below */
    double* acc = (double*)calloc((size_t)N * per, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; n++) {
        size_t ao = (Na > 1 ? (size_t)n * Vt * A : 0);
        const float* at = attr + ao;
        double* ga = acc + (size_t)n * per;
        for (size_t p = 0; p < npx; p++) {
            const float* r = rast + ((size_t)n * npx + p) * 4;
            const float* d = dy + ((size_t)n * npx + p) * A;
            float* gr = g_rast + ((size_t)n * npx + p) * 4;
            gr[0] = gr[1] = gr[2] = gr[3] = 0.0f;
            int t = (int)r[3] - 1;
            if (t < 0 || t >= T) continue;
            int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
            if (i0 < 0 || i0 >= Vt || i1 < 0 || i1 >= Vt || i2 < 0 || i2 >= Vt) continue;
            double b0 = r[0], b1 = r[1], b2 = 1.0 - b0 - b1, gu = 0.0, gv = 0.0;
            for (int c = 0; c < A; c++) {
                double g = d[c];
                ga[i0 * A + c] += b0 * g; ga[i1 * A + c] += b1 * g; ga[i2 * A + c] += b2 * g;
                gu += g * ((double)at[i0 * A + c] - at[i2 * A + c]);
                gv += g * ((double)at[i1 * A + c] - at[i2 * A + c]);
            }
            gr[0] = (float)gu; gr[1] = (float)gv;
        }
    }
    if (Na > 1) {
        for (size_t i = 0; i < na; i++) g_attr[i] = (float)acc[i];
    } else {
#pragma omp parallel for
        for (size_t i = 0; i < per; i++) {
            double t = 0.0;
            for (int n = 0; n < N; n++) t += acc[(size_t)n * per + i];
            g_attr[i] = (float)t;
        }
    }
    free(acc);
}

/* ------------------------------------------------------------------------------------------------ */
/* texture, filter_mode='linear', boundary_mode='wrap'                                              */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { int i00, i10, i01, i11; float fx, fy; } texfetch_t;

static inline texfetch_t tex_index(float u, float v, int Wt, int Ht)
{
    texfetch_t f;
    u = u - floorf(u);
    v = v - floorf(v);
    float x = u * (float)Wt - 0.5f, y = v * (float)Ht - 0.5f;
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

/* tex [Nt,Ht,Wt,C], Nt in {1,N}; uv [N,H,W,2]; out [N,H,W,C] */
void gold_texture_linear_fwd(const float* tex, int Nt, int Ht, int Wt, int C, const float* uv,
                             int N, int H, int W, float* out)
{
    size_t npx = (size_t)H * W;
#pragma omp parallel for
    for (int n = 0; n < N; n++) {
        const float* tx = tex + (Nt > 1 ? (size_t)n * Ht * Wt * C : 0);
        for (size_t p = 0; p < npx; p++) {
            const float* q = uv + ((size_t)n * npx + p) * 2;
            texfetch_t f = tex_index(q[0], q[1], Wt, Ht);
            float* o = out + ((size_t)n * npx + p) * C;
            for (int c = 0; c < C; c++) {
                float t00 = tx[(size_t)f.i00 * C + c], t10 = tx[(size_t)f.i10 * C + c];
                float t01 = tx[(size_t)f.i01 * C + c], t11 = tx[(size_t)f.i11 * C + c];
                float a = t00 + (t10 - t00) * f.fx, b = t01 + (t11 - t01) * f.fx;
                o[c] = a + (b - a) * f.fy;
            }
        }
    }
}

/* g_tex [Nt,Ht,Wt,C] overwritten, g_uv [N,H,W,2] */
void gold_texture_linear_bwd(const float* tex, int Nt, int Ht, int Wt, int C, const float* uv, const float* dy,
                             int N, int H, int W, float* g_tex, float* g_uv)
{
    size_t npx = (size_t)H * W;
    size_t nt = (size_t)(Nt > 1 ? N : 1) * Ht * Wt * C;
    size_t per = (size_t)Ht * Wt * C;
    double* acc = (double*)calloc((size_t)N * per, sizeof(double));     /* per view, summed in view order below */
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; n++) {
        size_t to = (Nt > 1 ? (size_t)n * Ht * Wt * C : 0);
        const float* tx = tex + to;
        double* gt = acc + (size_t)n * per;
        for (size_t p = 0; p < npx; p++) {
            const float* q = uv + ((size_t)n * npx + p) * 2;
            const float* d = dy + ((size_t)n * npx + p) * C;
            texfetch_t f = tex_index(q[0], q[1], Wt, Ht);
            double fx = f.fx, fy = f.fy, gu = 0.0, gv = 0.0;
            for (int c = 0; c < C; c++) {
                double g = d[c];
                double t00 = tx[(size_t)f.i00 * C + c], t10 = tx[(size_t)f.i10 * C + c];
                double t01 = tx[(size_t)f.i01 * C + c], t11 = tx[(size_t)f.i11 * C + c];
                gt[(size_t)f.i00 * C + c] += g * (1 - fx) * (1 - fy);
                gt[(size_t)f.i10 * C + c] += g * fx * (1 - fy);
                gt[(size_t)f.i01 * C + c] += g * (1 - fx) * fy;
                gt[(size_t)f.i11 * C + c] += g * fx * fy;
                gu += g * ((t10 - t00) * (1 - fy) + (t11 - t01) * fy);
                gv += g * ((t01 - t00) * (1 - fx) + (t11 - t10) * fx);
            }
            g_uv[((size_t)n * npx + p) * 2 + 0] = (float)(gu * Wt);
            g_uv[((size_t)n * npx + p) * 2 + 1] = (float)(gv * Ht);
        }
    }
    if (Nt > 1) {
        for (size_t i = 0; i < nt; i++) g_tex[i] = (float)acc[i];
    } else {
#pragma omp parallel for
        for (size_t i = 0; i < per; i++) {
            double t = 0.0;
            for (int n = 0; n < N; n++) t += acc[(size_t)n * per + i];
            g_tex[i] = (float)t;
        }
    }
    free(acc);
}

/* ------------------------------------------------------------------------------------------------ */
/* antialias                                                                                        */
/* ------------------------------------------------------------------------------------------------ */

/* tri_opp [T,3] (built by oracle/golden.py:topology_build with a numpy edge sort): for triangle t and edge e
 * (edge e is opposite corner e: e0=(v1,v2), e1=(v2,v0), e2=(v0,v1)) the third vertex of the other triangle
 * sharing that edge, or -1 for a boundary edge.  With more than two triangles on an edge the other triangle
 * with the lowest (index, corner) is used. */

static inline int same_sign(float a, float b)
{
    int32_t ia, ib;
    memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
    return (ia ^ ib) >= 0;
}

/* n0/d0 > n1/d1 without dividing */
static inline int rational_gt(float n0, float n1, float d0, float d1)
{
    float p0 = n0 * d1, p1 = n1 * d0;
    return same_sign(d0, d1) ? (p0 > p1) : (p0 < p1);
}

static inline int max_idx3(float n0, float n1, float n2, float d0, float d1, float d2)
{
    int g10 = rational_gt(n1, n0, d1, d0);
    int g20 = rational_gt(n2, n0, d2, d0);
    int g21 = rational_gt(n2, n1, d2, d1);
    if (g20 && g21) return 2;
    if (g10) return 1;
    return 0;
}

typedef struct { int valid; int di; int tri; float ds; float alpha; int px, py; } aa_pair_t;

/* Analysis of the pixel pair (px,py) -> (px+1,py) [d=0] or (px,py+1) [d=1]   (App. A.4 step 3) */
static aa_pair_t aa_analyze(const float* rast, const f4* P, const int32_t* tri, const int32_t* tri_opp,
                            int T, int V, int H, int W, int px, int py, int d)
{
    aa_pair_t r; r.valid = 0; r.di = 0; r.tri = -1; r.ds = 1.0f; r.alpha = 0.0f; r.px = px; r.py = py;
    size_t pix0 = (size_t)py * W + px, pix1 = pix0 + (d ? W : 1);
    int tri0 = (int)rast[pix0 * 4 + 3] - 1, tri1 = (int)rast[pix1 * 4 + 3] - 1;
    if (tri0 == tri1) return r;
    float z0 = rast[pix0 * 4 + 2], z1 = rast[pix1 * 4 + 2];
    int t = (tri0 >= 0) ? tri0 : tri1;
    if (tri0 >= 0 && tri1 >= 0) t = (z0 < z1) ? tri0 : tri1;
    if (t == tri1) { px += 1 - d; py += d; }
    if (t < 0 || t >= T) return r;
    int vi0 = tri[3 * t], vi1 = tri[3 * t + 1], vi2 = tri[3 * t + 2];
    if (vi0 < 0 || vi0 >= V || vi1 < 0 || vi1 >= V || vi2 < 0 || vi2 >= V) return r;
    int op0 = tri_opp[3 * t], op1 = tri_opp[3 * t + 1], op2 = tri_opp[3 * t + 2];
    f4 p0 = P[vi0], p1 = P[vi1], p2 = P[vi2];
    f4 o0 = (op0 < 0) ? p0 : P[op0], o1 = (op1 < 0) ? p1 : P[op1], o2 = (op2 < 0) ? p2 : P[op2];
    float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    float w0 = 1.0f / p0.w, w1 = 1.0f / p1.w, w2 = 1.0f / p2.w;
    float ow0 = 1.0f / o0.w, ow1 = 1.0f / o1.w, ow2 = 1.0f / o2.w;
    float fx = (float)px + 0.5f - xh, fy = (float)py + 0.5f - yh;
    float x0 = p0.x * w0 * xh - fx, y0 = p0.y * w0 * yh - fy;
    float x1 = p1.x * w1 * xh - fx, y1 = p1.y * w1 * yh - fy;
    float x2 = p2.x * w2 * xh - fx, y2 = p2.y * w2 * yh - fy;
    float ox0 = o0.x * ow0 * xh - fx, oy0 = o0.y * ow0 * yh - fy;
    float ox1 = o1.x * ow1 * xh - fx, oy1 = o1.y * ow1 * yh - fy;
    float ox2 = o2.x * ow2 * xh - fx, oy2 = o2.y * ow2 * yh - fy;
    float bb = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    float a0 = (x1 - ox0) * (y2 - oy0) - (x2 - ox0) * (y1 - oy0);
    float a1 = (x2 - ox1) * (y0 - oy1) - (x0 - ox1) * (y2 - oy1);
    float a2 = (x0 - ox2) * (y1 - oy2) - (x1 - ox2) * (y0 - oy2);
    if (!(same_sign(a0, bb) || same_sign(a1, bb) || same_sign(a2, bb))) return r;
    if (d) { float s; s = x0; x0 = y0; y0 = s; s = x1; x1 = y1; y1 = s; s = x2; x2 = y2; y2 = s; }
    float dx0 = x2 - x1, dx1 = x0 - x2, dx2 = x1 - x0;
    float dy0 = y2 - y1, dy1 = y0 - y2, dy2 = y1 - y0;
    float dc = -3.402823466e38f;
    float ds = (t == tri0) ? 1.0f : -1.0f;
    float d0 = ds * (x1 * dy0 - y1 * dx0);
    float d1 = ds * (x2 * dy1 - y2 * dx1);
    float d2 = ds * (x0 * dy2 - y0 * dx2);
    if (same_sign(y1, y2)) { d0 = -3.402823466e38f; dy0 = 1.0f; }
    if (same_sign(y2, y0)) { d1 = -3.402823466e38f; dy1 = 1.0f; }
    if (same_sign(y0, y1)) { d2 = -3.402823466e38f; dy2 = 1.0f; }
    int di = max_idx3(d0, d1, d2, dy0, dy1, dy2);
    if (di == 0 && same_sign(a0, bb) && fabsf(dy0) >= fabsf(dx0)) dc = d0 / dy0;
    if (di == 1 && same_sign(a1, bb) && fabsf(dy1) >= fabsf(dx1)) dc = d1 / dy1;
    if (di == 2 && same_sign(a2, bb) && fabsf(dy2) >= fabsf(dx2)) dc = d2 / dy2;
    const float eps = 0.0625f;
    if (dc > -eps && dc < 1.0f + eps) {
        dc = fminf(fmaxf(dc, 0.0f), 1.0f);
        r.valid = 1; r.di = di; r.tri = t; r.ds = ds; r.alpha = ds * (0.5f - dc); r.px = px; r.py = py;
    }
    return r;
}

/* color [N,H,W,C], rast [N,H,W,4], pos [N,V,4], tri [T,3], tri_opp [T,3] -> out [N,H,W,C] */
void gold_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                        const int32_t* tri_opp, int N, int V, int T, int H, int W, int C, float* out)
{
    size_t npx = (size_t)H * W;
    memcpy(out, color, (size_t)N * npx * C * sizeof(float));
#pragma omp parallel for
    for (int n = 0; n < N; n++) {
        const float* r = rast + (size_t)n * npx * 4;
        const float* col = color + (size_t)n * npx * C;
        float* o = out + (size_t)n * npx * C;
        const f4* P = (const f4*)pos + (size_t)n * V;
        for (int py = 0; py < H; py++) for (int px = 0; px < W; px++) for (int d = 0; d < 2; d++) {
            if ((d == 0 && px == W - 1) || (d == 1 && py == H - 1)) continue;
            aa_pair_t a = aa_analyze(r, P, tri, tri_opp, T, V, H, W, px, py, d);
            if (!a.valid) continue;
            size_t pix0 = (size_t)py * W + px, pix1 = pix0 + (d ? W : 1);
            size_t dst = (a.alpha > 0.0f) ? pix0 : pix1;
            for (int c = 0; c < C; c++) o[dst * C + c] += a.alpha * (col[pix1 * C + c] - col[pix0 * C + c]);
        }
    }
}

/* g_color [N,H,W,C], g_pos [N,V,4] (both overwritten).  Position gradient follows the published edge-crossing
 * derivative with the 1e-3 px regulariser on 1/dy (App. A.4 step 5). */
void gold_antialias_bwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                        const int32_t* tri_opp, const float* dy, int N, int V, int T, int H, int W, int C,
                        float* g_color, float* g_pos)
{
    size_t npx = (size_t)H * W;
    double* gc = (double*)calloc((size_t)N * npx * C, sizeof(double));
    double* gp = (double*)calloc((size_t)N * V * 4, sizeof(double));
    for (size_t i = 0; i < (size_t)N * npx * C; i++) gc[i] = dy[i];
    float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; n++) {
        const float* r = rast + (size_t)n * npx * 4;
        const float* col = color + (size_t)n * npx * C;
        const float* g = dy + (size_t)n * npx * C;
        double* oc = gc + (size_t)n * npx * C;
        double* op = gp + (size_t)n * V * 4;
        const f4* P = (const f4*)pos + (size_t)n * V;
        for (int py = 0; py < H; py++) for (int px = 0; px < W; px++) for (int d = 0; d < 2; d++) {
            if ((d == 0 && px == W - 1) || (d == 1 && py == H - 1)) continue;
            aa_pair_t a = aa_analyze(r, P, tri, tri_opp, T, V, H, W, px, py, d);
            if (!a.valid) continue;
            size_t pix0 = (size_t)py * W + px, pix1 = pix0 + (d ? W : 1);
            size_t dst = (a.alpha > 0.0f) ? pix0 : pix1;
            double dd = 0.0;
            for (int c = 0; c < C; c++) {
                double gy = g[dst * C + c];
                dd += gy * ((double)col[pix1 * C + c] - col[pix0 * C + c]);
                oc[pix0 * C + c] -= a.alpha * gy;
                oc[pix1 * C + c] += a.alpha * gy;
            }
            if (dd == 0.0) continue;
            int t = a.tri;
            /* endpoints of the crossing edge di: e0=(v1,v2), e1=(v2,v0), e2=(v0,v1) */
            int i1 = tri[3 * t + (a.di + 1) % 3], i2 = tri[3 * t + (a.di + 2) % 3];
            f4 p1 = P[i1], p2 = P[i2];
            float w1 = 1.0f / p1.w, w2 = 1.0f / p2.w;
            float fx = (float)a.px + 0.5f - xh, fy = (float)a.py + 0.5f - yh;
            float x1 = p1.x * w1 * xh - fx, y1 = p1.y * w1 * yh - fy;
            float x2 = p2.x * w2 * xh - fx, y2 = p2.y * w2 * yh - fy;
            if (d) { float s; s = x1; x1 = y1; y1 = s; s = x2; x2 = y2; y2 = s; }
            /* alpha = ds*0.5 - db/dyy with db = x1*dyy - y1*dxx (ds*ds = 1) */
            double dxx = (double)x2 - x1, dyy = (double)y2 - y1;
            double db = x1 * dyy - y1 * dxx;
            double ep = copysign(1e-3, dyy);
            double iy = 1.0 / (dyy + ep);
            double dby = db * iy;
            double iw1 = -w1 * iy * dd, iw2 = w2 * iy * dd;
            double s1 = d ? yh : xh, s2 = d ? xh : yh;   /* scale of the (possibly swapped) x / y axes */
            double gp1x = iw1 * s1 * y2, gp2x = iw2 * s1 * y1;
            double gp1y = iw1 * s2 * (dby - x2), gp2y = iw2 * s2 * (dby - x1);
            if (d) { double s; s = gp1x; gp1x = gp1y; gp1y = s; s = gp2x; gp2x = gp2y; gp2y = s; }
            double gp1w = -(p1.x * gp1x + p1.y * gp1y) * w1;
            double gp2w = -(p2.x * gp2x + p2.y * gp2y) * w2;
            op[i1 * 4 + 0] += gp1x; op[i1 * 4 + 1] += gp1y; op[i1 * 4 + 3] += gp1w;
            op[i2 * 4 + 0] += gp2x; op[i2 * 4 + 1] += gp2y; op[i2 * 4 + 3] += gp2w;
        }
    }
    for (size_t i = 0; i < (size_t)N * npx * C; i++) g_color[i] = (float)gc[i];
    for (size_t i = 0; i < (size_t)N * V * 4; i++) g_pos[i] = (float)gp[i];
    free(gc); free(gp);
}

/* threads the OpenMP loops above run on (1 when built without OpenMP): reported as cpu_baseline.cores by bench.py */
#ifdef _OPENMP
#include <omp.h>
int gold_omp_threads(void) { return omp_get_max_threads(); }
#else
int gold_omp_threads(void) { return 1; }
#endif
