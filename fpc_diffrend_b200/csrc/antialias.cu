// Silhouette-edge antialiasing fwd/bwd (replaces dr.antialias, reference fit.py:160).
//
// Differences from the upstream structure, by design:
//   * topology is a per-mesh adjacency table tri_opp [T,3] built once (fpc_topology_build) instead of an
//     edge hash rebuilt on every call (the reference never passes topology_hash, fit.py:160);
//   * the forward pass and the colour gradient are pixel-parallel GATHERS: every pixel analyses the four
//     pixel pairs it belongs to and sums what it receives in a fixed order — no work queue, no atomics,
//     run-to-run deterministic.  Only the (sparse) silhouette position gradient is scattered.
#include "common.cuh"

namespace {

constexpr unsigned long long HKEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned CODE_NONE = 0xFFFFFFFFu;

struct TopoTable {
    unsigned long long* keys;
    unsigned* c0;
    unsigned* c1;
    unsigned mask;
};

__device__ __forceinline__ unsigned hash_key(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

__device__ __forceinline__ unsigned long long edge_key(const int32_t* tri, int t, int e)
{
    unsigned va = (unsigned)tri[3 * t + (e + 1) % 3], vb = (unsigned)tri[3 * t + (e + 2) % 3];
    unsigned lo = min(va, vb), hi = max(va, vb);
    return ((unsigned long long)lo << 32) | hi;
}

__global__ void k_topo_init(TopoTable tb)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= tb.mask) { tb.keys[i] = HKEY_EMPTY; tb.c0[i] = CODE_NONE; tb.c1[i] = CODE_NONE; }
}

// pass 0: claim a slot and record the lowest (triangle,corner) code on the edge
// pass 1: record the second lowest code
// pass 2: emit the opposite vertex
__global__ void k_topo_pass(TopoTable tb, const int32_t* __restrict__ tri, int T, int pass, int32_t* __restrict__ tri_opp)
{
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= 3 * T) return;
    int t = gid / 3, e = gid - 3 * t;
    unsigned long long key = edge_key(tri, t, e);
    unsigned code = (unsigned)t * 4u + (unsigned)e;
    unsigned slot = hash_key(key) & tb.mask;
    if (pass == 0) {
        while (true) {
            unsigned long long prev = atomicCAS(tb.keys + slot, HKEY_EMPTY, key);
            if (prev == HKEY_EMPTY || prev == key) break;
            slot = (slot + 1) & tb.mask;
        }
        atomicMin(tb.c0 + slot, code);
        return;
    }
    while (tb.keys[slot] != key) slot = (slot + 1) & tb.mask;
    if (pass == 1) {
        if (tb.c0[slot] != code) atomicMin(tb.c1 + slot, code);
        return;
    }
    unsigned other = (tb.c0[slot] == code) ? tb.c1[slot] : tb.c0[slot];
    tri_opp[gid] = (other == CODE_NONE) ? -1 : tri[3 * (other >> 2) + (other & 3u)];
}

// ---------------------------------------------------------------------------------------------------------

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
    float w0 = xdiv(1.f, p0.w), w1 = xdiv(1.f, p1.w), w2 = xdiv(1.f, p2.w);
    float ow0 = xdiv(1.f, o0.w), ow1 = xdiv(1.f, o1.w), ow2 = xdiv(1.f, o2.w);
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

__device__ __forceinline__ float2 load_zt(const float* rast, size_t pix) { return __ldg(reinterpret_cast<const float2*>(rast) + 2 * pix + 1); }

// out[p] = color[p] + sum over the 4 pairs containing p of what the pair deposits on p
__global__ void __launch_bounds__(256) k_aa_fwd(AAParams ap, const float* __restrict__ color, float* __restrict__ out)
{
    long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long npx_inst = (long long)ap.H * ap.W;
    if (gi >= npx_inst * ap.N) return;
    int n = (int)(gi / npx_inst);
    int px = (int)(gi % ap.W), py = (int)((gi / ap.W) % ap.H);
    size_t pix = (size_t)gi;
    const int C = ap.C;
    float2 zc = load_zt(ap.rast, pix);
    // neighbour pairs: 0 = (p, right) as pix0, 1 = (left, p) as pix1, 2 = (p, up) as pix0, 3 = (down, p) as pix1
    float alpha[4] = {0.f, 0.f, 0.f, 0.f};
    size_t other[4] = {pix, pix, pix, pix};
    bool any = false;
    if (px + 1 < ap.W) {
        float2 zn = load_zt(ap.rast, pix + 1);
        if (zn.y != zc.y) { AAPair a = aa_analyze(ap, n, px, py, 0, zc, zn); if (a.valid && a.alpha > 0.f) { alpha[0] = a.alpha; other[0] = pix + 1; any = true; } }
    }
    if (px > 0) {
        float2 zn = load_zt(ap.rast, pix - 1);
        if (zn.y != zc.y) { AAPair a = aa_analyze(ap, n, px - 1, py, 0, zn, zc); if (a.valid && !(a.alpha > 0.f)) { alpha[1] = a.alpha; other[1] = pix - 1; any = true; } }
    }
    if (py + 1 < ap.H) {
        float2 zn = load_zt(ap.rast, pix + ap.W);
        if (zn.y != zc.y) { AAPair a = aa_analyze(ap, n, px, py, 1, zc, zn); if (a.valid && a.alpha > 0.f) { alpha[2] = a.alpha; other[2] = pix + ap.W; any = true; } }
    }
    if (py > 0) {
        float2 zn = load_zt(ap.rast, pix - ap.W);
        if (zn.y != zc.y) { AAPair a = aa_analyze(ap, n, px, py - 1, 1, zn, zc); if (a.valid && !(a.alpha > 0.f)) { alpha[3] = a.alpha; other[3] = pix - ap.W; any = true; } }
    }
    for (int c = 0; c < C; c++) {
        float cc = __ldg(color + pix * C + c);
        float o = cc;
        if (any) {
            // pair k as pix0: += alpha (c1 - c0) = alpha (other - self);  as pix1: += alpha (c1 - c0) = alpha (self - other)
            if (alpha[0] != 0.f) o += alpha[0] * (__ldg(color + other[0] * C + c) - cc);
            if (alpha[1] != 0.f) o += alpha[1] * (cc - __ldg(color + other[1] * C + c));
            if (alpha[2] != 0.f) o += alpha[2] * (__ldg(color + other[2] * C + c) - cc);
            if (alpha[3] != 0.f) o += alpha[3] * (cc - __ldg(color + other[3] * C + c));
        }
        out[pix * C + c] = o;
    }
}

// silhouette position gradient of one accepted pair (handled by the thread that owns pix0)
__device__ __forceinline__ void aa_pos_grad(const AAParams& ap, int n, const AAPair& a, int d, float dd, float* __restrict__ g_pos)
{
    int t = a.tri;
    int i1 = __ldg(ap.tri + 3 * t + (a.di + 1) % 3), i2 = __ldg(ap.tri + 3 * t + (a.di + 2) % 3);
    const float* P = ap.pos + (size_t)n * ap.V * 4;
    float4 p1 = ldg4(P + 4 * (size_t)i1), p2 = ldg4(P + 4 * (size_t)i2);
    float xh = ap.xh, yh = ap.yh;
    float w1 = xdiv(1.f, p1.w), w2 = xdiv(1.f, p2.w);
    float fx = xsub(xadd((float)a.px, 0.5f), xh), fy = xsub(xadd((float)a.py, 0.5f), yh);
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
    float gp1w = -(p1.x * gp1x + p1.y * gp1y) * w1;
    float gp2w = -(p2.x * gp2x + p2.y * gp2y) * w2;
    float* G = g_pos + (size_t)n * ap.V * 4;
    atomicAdd(G + 4 * (size_t)i1 + 0, gp1x); atomicAdd(G + 4 * (size_t)i1 + 1, gp1y); atomicAdd(G + 4 * (size_t)i1 + 3, gp1w);
    atomicAdd(G + 4 * (size_t)i2 + 0, gp2x); atomicAdd(G + 4 * (size_t)i2 + 1, gp2y); atomicAdd(G + 4 * (size_t)i2 + 3, gp2w);
}

__global__ void __launch_bounds__(256) k_aa_bwd(AAParams ap, const float* __restrict__ color, const float* __restrict__ dy,
                                                float* __restrict__ g_color, float* __restrict__ g_pos)
{
    long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long npx_inst = (long long)ap.H * ap.W;
    if (gi >= npx_inst * ap.N) return;
    int n = (int)(gi / npx_inst);
    int px = (int)(gi % ap.W), py = (int)((gi / ap.W) % ap.H);
    size_t pix = (size_t)gi;
    const int C = ap.C;
    float2 zc = load_zt(ap.rast, pix);
    // For pair k: alpha, the other pixel, the destination pixel whose dy weights the term, and the sign with
    // which this pixel's colour enters  (pix0: -alpha dy[dst],  pix1: +alpha dy[dst]).
    float alpha[4] = {0.f, 0.f, 0.f, 0.f};
    size_t dst[4] = {pix, pix, pix, pix};
    bool any = false;
    AAPair own[2];
    own[0].valid = false; own[1].valid = false;
    size_t own_other[2] = {pix, pix};
    if (px + 1 < ap.W) {
        float2 zn = load_zt(ap.rast, pix + 1);
        if (zn.y != zc.y) { AAPair a = aa_analyze(ap, n, px, py, 0, zc, zn); if (a.valid) { alpha[0] = a.alpha; dst[0] = a.alpha > 0.f ? pix : pix + 1; own[0] = a; own_other[0] = pix + 1; any = true; } }
    }
    if (px > 0) {
        float2 zn = load_zt(ap.rast, pix - 1);
        if (zn.y != zc.y) { AAPair a = aa_analyze(ap, n, px - 1, py, 0, zn, zc); if (a.valid) { alpha[1] = a.alpha; dst[1] = a.alpha > 0.f ? pix - 1 : pix; any = true; } }
    }
    if (py + 1 < ap.H) {
        float2 zn = load_zt(ap.rast, pix + ap.W);
        if (zn.y != zc.y) { AAPair a = aa_analyze(ap, n, px, py, 1, zc, zn); if (a.valid) { alpha[2] = a.alpha; dst[2] = a.alpha > 0.f ? pix : pix + ap.W; own[1] = a; own_other[1] = pix + ap.W; any = true; } }
    }
    if (py > 0) {
        float2 zn = load_zt(ap.rast, pix - ap.W);
        if (zn.y != zc.y) { AAPair a = aa_analyze(ap, n, px, py - 1, 1, zn, zc); if (a.valid) { alpha[3] = a.alpha; dst[3] = a.alpha > 0.f ? pix - ap.W : pix; any = true; } }
    }
    float dd0 = 0.f, dd1 = 0.f;
    for (int c = 0; c < C; c++) {
        float g = __ldg(dy + pix * C + c);
        if (any) {
            if (alpha[0] != 0.f) g -= alpha[0] * __ldg(dy + dst[0] * C + c);
            if (alpha[1] != 0.f) g += alpha[1] * __ldg(dy + dst[1] * C + c);
            if (alpha[2] != 0.f) g -= alpha[2] * __ldg(dy + dst[2] * C + c);
            if (alpha[3] != 0.f) g += alpha[3] * __ldg(dy + dst[3] * C + c);
            float cc = __ldg(color + pix * C + c);
            if (own[0].valid) dd0 += __ldg(dy + dst[0] * C + c) * (__ldg(color + own_other[0] * C + c) - cc);
            if (own[1].valid) dd1 += __ldg(dy + dst[2] * C + c) * (__ldg(color + own_other[1] * C + c) - cc);
        }
        g_color[pix * C + c] = g;
    }
    if (own[0].valid && dd0 != 0.f) aa_pos_grad(ap, n, own[0], 0, dd0, g_pos);
    if (own[1].valid && dd1 != 0.f) aa_pos_grad(ap, n, own[1], 1, dd1, g_pos);
}

unsigned next_pow2(unsigned x) { unsigned p = 1; while (p < x) p <<= 1; return p; }

}  // namespace

extern "C" size_t fpc_topology_scratch_bytes(int T)
{
    if (T <= 0) return 256;
    size_t M = next_pow2((unsigned)(6u * (unsigned)T));
    return M * 16 + 256;
}

extern "C" int fpc_topology_build(const int32_t* tri, int T, int V, int32_t* tri_opp, void* scratch, size_t scratch_bytes,
                                  fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)V;
    FPC_CHECK_ARG(tri && tri_opp, "topology_build: null pointer argument");
    FPC_CHECK_ARG(T > 0 && T < (1 << 28), "topology_build: 0 < T < 2^28 required (got %d)", T);
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_topology_scratch_bytes(T), "topology_build: scratch too small");
    unsigned M = next_pow2(6u * (unsigned)T);
    TopoTable tb;
    tb.keys = (unsigned long long*)scratch;
    tb.c0 = (unsigned*)((char*)scratch + (size_t)M * 8);
    tb.c1 = tb.c0 + M;
    tb.mask = M - 1;
    k_topo_init<<<fpc_div_up(M, 256), 256, 0, stream>>>(tb);
    FPC_LAUNCH_CHECK();
    for (int pass = 0; pass < 3; pass++) {
        k_topo_pass<<<fpc_div_up(3LL * T, 256), 256, 0, stream>>>(tb, tri, T, pass, tri_opp);
        FPC_LAUNCH_CHECK();
    }
    return FPC_OK;
}

static int aa_fill(AAParams& ap, const float* rast, const float* pos, const int32_t* tri, const int32_t* tri_opp,
                   int N, int V, int T, int H, int W, int C)
{
    ap.rast = rast; ap.pos = pos; ap.tri = tri; ap.tri_opp = tri_opp;
    ap.N = N; ap.V = V; ap.T = T; ap.H = H; ap.W = W; ap.C = C;
    ap.xh = 0.5f * (float)W; ap.yh = 0.5f * (float)H;
    return 0;
}

extern "C" int fpc_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                                 const int32_t* tri_opp, int N, int V, int T, int H, int W, int C, float* out,
                                 fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(color && rast && pos && tri && tri_opp && out, "antialias_fwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && V > 0 && T > 0 && H > 0 && W > 0 && C > 0, "antialias_fwd: sizes must be positive");
    AAParams ap;
    aa_fill(ap, rast, pos, tri, tri_opp, N, V, T, H, W, C);
    long long npx = (long long)N * H * W;
    k_aa_fwd<<<fpc_div_up(npx, 256), 256, 0, stream>>>(ap, color, out);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_antialias_bwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                                 const int32_t* tri_opp, const float* dy, int N, int V, int T, int H, int W, int C,
                                 float* grad_color, float* grad_pos, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(color && rast && pos && tri && tri_opp && dy && grad_color && grad_pos, "antialias_bwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && V > 0 && T > 0 && H > 0 && W > 0 && C > 0, "antialias_bwd: sizes must be positive");
    AAParams ap;
    aa_fill(ap, rast, pos, tri, tri_opp, N, V, T, H, W, C);
    FPC_CUDA(cudaMemsetAsync(grad_pos, 0, (size_t)N * V * 4 * sizeof(float), stream));
    long long npx = (long long)N * H * W;
    k_aa_bwd<<<fpc_div_up(npx, 256), 256, 0, stream>>>(ap, color, dy, grad_color, grad_pos);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
