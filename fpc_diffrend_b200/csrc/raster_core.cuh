// Rasterizer core shared by the stand-alone rasterize op (raster.cu) and the fused fit kernel (fused.cu).
//
// Binning (k_setup / k_scan / k_fill, raster.cu) produces, per (instance, 64x64-px bin), a list of SMALL
// triangles (snapped bbox <= 128 px in both axes and <= 2x2 bins: int32 edge math is exact) and, per instance,
// a list of LARGE triangles (everything else, int64 edge math, whole CTA cooperates on each).
//
// raster_bin(): one CTA resolves visibility of its bin into a 64x64 shared-memory array of 64-bit keys
//   key = order_preserving(depth) << 32 | triangle_id         (atomicMin == LESS test, lower id wins ties)
// Small triangles are processed warp by warp, 32 at a time: each lane sets up one triangle and stages it in
// shared memory, then the warp walks the (triangle, bbox-row) work items of the batch with a balanced
// lane <-> item mapping (prefix sum + binary search), so lanes stay busy whatever the triangle sizes are.
//
// Semantics are bit-identical to oracle/golden.c (DESIGN.md "Rasterizer semantics").
#pragma once
#include "common.cuh"

namespace fpc {

// Tuning knobs (scripts/exp_variants.py builds variants; measured on B200 at config 2, see profiles/):
// 32x32-px bins with 128-thread CTAs and 64 registers/thread put 8 independent CTAs on an SM, which hides the
// raster -> shade barrier and the gather latency of the shading phase better than 64x64 bins / 256 threads.
#ifndef FPC_BIN_LOG2
#define FPC_BIN_LOG2 5
#endif
#ifndef FPC_FINE_THREADS
#define FPC_FINE_THREADS 128
#endif
#ifndef FPC_DYN_BATCH
#define FPC_DYN_BATCH 1
#endif
#ifndef FPC_WALK_MASK
#define FPC_WALK_MASK 1      // row coverage as a bit mask, then only the covered span is depth-tested
#endif
#ifndef FPC_KEY32
#define FPC_KEY32 1          // 32-bit packed (depth - bin base, id) keys + native ATOMS.MIN where a bin's depth range allows it
#endif
#ifndef FPC_KEY32_TEST_OVERFLOW
#define FPC_KEY32_TEST_OVERFLOW 0
#endif
#ifndef FPC_CAS_MANUAL
#define FPC_CAS_MANUAL 0     // 1: CAS loop seeded by the early-out read — measured 40 % SLOWER than atomicMin on B200
#endif
constexpr int BIN_LOG2 = FPC_BIN_LOG2;
constexpr int BIN = 1 << BIN_LOG2;   // bin edge in pixels
constexpr int FINE_THREADS = FPC_FINE_THREADS;
constexpr int FINE_WARPS = FINE_THREADS / 32;
constexpr float SNAP_LIMIT = 16777216.0f;
constexpr int SMALL_EXTENT = 2048;   // 128 px in 1/16-px units
constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned KEY32_MARGIN = 256u;

struct RasterParams {
    const float* pos;
    const int32_t* tri;
    int N, V, T, H, W;
    int BW, BH, NB;
    int halo;                        // bins are widened by `halo` px on every side when triangles are binned (fused_aa.cu)
    float xs, xo, ys, yo;            // pixel -> NDC
    float sxs, sys;                  // NDC -> 1/16 px:  8*W, 8*H
    int* bin_count;                  // [N*NB]
    int* bin_cursor;                 // [N*NB]
    int* large_count;                // [N]
    int* bin_offset;                 // [N*NB]
    int* tri_info;                   // [N*T]
    int* pairs;                      // [N*4T]
    int* large_list;                 // [N*2T]  triangle ids, or T + slot of a clipped piece (clip_* below)
    float4* clip_verts;              // [clip_cap*3] vertices of the pieces of near-clipped triangles
    int* clip_parent;                // [clip_cap]   id of the triangle a piece belongs to
    int* clip_count;                 // [1]          pieces allocated so far (all instances share the pool)
    int clip_cap;
    int* tri_anchor;                 // [N*T]  px | py << 16: pixel of the triangle's centroid, clamped to the image (moment origin)
    int4* tri4;                      // [T] (i0, i1, i2, 0): 16-byte copy of tri written by k_setup, one load per triangle
    const int32_t* pad_i_src; int4* pad_i_dst; int pad_i_n;          // nullable job for k_setup: int [n,3] -> int4 [n]
    ushort4* tri_bbox;               // [N*T] (pxa, pya, pxb, pyb): candidate pixel range of a SMALL triangle, clamped to the image
    float* slot_grad;                // nullable: [N*T*4*9] per-(view, triangle, bin k) gradient slots of fused.cu; k_setup zeroes
                                     // the accumulator slots (1 and 2) of LARGE triangles, every other slot is written at most once
    unsigned* slot_valid;            // [N*T] (with slot_grad): bits 6-7 of every byte = class of the triangle (0 none, 1 small, 2 large: k_setup),
                                     //       bit 0 of byte k = slot k was written (the fused kernels)
    unsigned* bin_nzlo;              // [N*NB] ~min and
    unsigned* bin_zhi;               // [N*NB]  max of depth_key(z/w) over the vertices of the SMALL triangles listed in the bin (k_setup;
                                     //         zero-initialised with the counters: an empty bin reads (0xFFFFFFFF, 0))
    int idbits;                      // bits of a triangle id: ceil(log2(T))
    // launch order of the fused kernels' CTAs (k_fill): within groups of order_gv views, bins sorted by list-length class, longest
    // first, empty bins last — the short background CTAs fill the gaps the long ones leave at the end of the launch
    int* order_count;                // [ngroups * ORDER_CLASSES] bins per class (zeroed with the counters)
    int* bin_order;                  // [ngroups * ORDER_CLASSES * ORDER_GROUP_VIEWS * NB] (view in group << 16 | bin) per class; null: identity order
};

constexpr int ORDER_CLASSES = 8;
constexpr int ORDER_GROUP_VIEWS = 16;

// 0: empty bin (background), 1: 1-7 triangles, 2: 8-15, 3: 16-31, ... 7: >= 256
__device__ __forceinline__ int order_class(int count)
{
    if (count <= 0) return 0;
    const int c = 32 - __clz(count) - 2;
    return c < 1 ? 1 : (c > 7 ? 7 : c);
}

// (blockIdx.x, blockIdx.y) of a (NB, N) grid -> the (view, bin) this CTA works on.  MODE 0: identity; 1: the group's bins by class,
// longest lists first, background bins last; 2: the same order for the non-empty bins with the background bins spread evenly
// between them (a background CTA only streams its reference tile: it mixes well with the compute-bound ones)
template <int MODE>
__device__ __forceinline__ int ordered_bin(const RasterParams& rp, int& n, int& bin)      // returns the bin's class, -1 = unknown (identity order)
{
    n = blockIdx.y; bin = blockIdx.x;
    if (MODE == 0 || !rp.bin_order) return -1;
    const int g = n / ORDER_GROUP_VIEWS;
    int r = (n - g * ORDER_GROUP_VIEWS) * rp.NB + bin;           // rank of this CTA inside its group of views
    const int4 hi = __ldg(reinterpret_cast<const int4*>(rp.order_count + ORDER_CLASSES * g) + 1);
    const int4 lo = __ldg(reinterpret_cast<const int4*>(rp.order_count + ORDER_CLASSES * g));
    const int cnt[ORDER_CLASSES] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    int cls = ORDER_CLASSES - 1;
    if (MODE == 2 && cnt[0] > 0) {
        const int total = min(ORDER_GROUP_VIEWS, rp.N - g * ORDER_GROUP_VIEWS) * rp.NB;
        const int k = total / cnt[0];                            // every k-th CTA of the group is a background bin
        const int q = r / k;
        if (r - q * k == k - 1 && q < cnt[0]) { cls = 0; r = q; }
        else r -= min(cnt[0], (r + 1) / k);
    }
    if (cls != 0) {
#pragma unroll
        for (int c = ORDER_CLASSES - 1; c > 0; c--)
            if (cls == c && r >= cnt[c]) { r -= cnt[c]; cls = c - 1; }
    }
    const int id = __ldg(rp.bin_order + ((size_t)g * ORDER_CLASSES + cls) * ((size_t)ORDER_GROUP_VIEWS * rp.NB) + r);    // view in group << 16 | bin
    n = g * ORDER_GROUP_VIEWS + (id >> 16);
    bin = id & 0xffff;
    return cls;
}

struct SnappedTri {
    int x0, y0, x1, y1, x2, y2;      // 1/16 px, ORIGINAL vertex order
    bool flip;                       // true when the original order has negative area
    int minx, maxx, miny, maxy;      // unclamped bbox, 1/16 px
    int pxa, pxb, pya, pyb;          // candidate pixel range clamped to the image
};

__device__ __forceinline__ bool snap_vertex(const float4& p, float sxs, float sys, int& sx, int& sy)
{
    if (!(p.w > 0.f)) return false;
    float rw = xrcp(p.w);
    float xf = xadd(xmul(xmul(p.x, rw), sxs), sxs);
    float yf = xadd(xmul(xmul(p.y, rw), sys), sys);
    if (!(fabsf(xf) < SNAP_LIMIT) || !(fabsf(yf) < SNAP_LIMIT)) return false;
    sx = __float2int_rn(xf);
    sy = __float2int_rn(yf);
    return true;
}

// Returns false when the triangle produces no fragments at all.
__device__ __forceinline__ bool setup_triangle(const float4& p0, const float4& p1, const float4& p2,
                                               const RasterParams& rp, SnappedTri& s)
{
    if (!snap_vertex(p0, rp.sxs, rp.sys, s.x0, s.y0) || !snap_vertex(p1, rp.sxs, rp.sys, s.x1, s.y1) ||
        !snap_vertex(p2, rp.sxs, rp.sys, s.x2, s.y2))
        return false;
    long long area = (long long)(s.x1 - s.x0) * (s.y2 - s.y0) - (long long)(s.x2 - s.x0) * (s.y1 - s.y0);
    if (area == 0) return false;
    s.flip = area < 0;
    s.minx = min(s.x0, min(s.x1, s.x2)); s.maxx = max(s.x0, max(s.x1, s.x2));
    s.miny = min(s.y0, min(s.y1, s.y2)); s.maxy = max(s.y0, max(s.y1, s.y2));
    s.pxa = max((s.minx - 8 + 15) >> 4, 0);
    s.pxb = min((s.maxx - 8) >> 4, rp.W - 1);
    s.pya = max((s.miny - 8 + 15) >> 4, 0);
    s.pyb = min((s.maxy - 8) >> 4, rp.H - 1);
    return s.pxa <= s.pxb && s.pya <= s.pyb;
}

// Bins whose (halo-widened) pixel window overlaps the triangle's candidate pixel range.
struct BinRange { int bx0, bx1, by0, by1; };

__device__ __forceinline__ BinRange bin_range(const SnappedTri& s, const RasterParams& rp)
{
    BinRange r;
    r.bx0 = max(s.pxa - rp.halo, 0) >> BIN_LOG2;
    r.bx1 = min((s.pxb + rp.halo) >> BIN_LOG2, rp.BW - 1);
    r.by0 = max(s.pya - rp.halo, 0) >> BIN_LOG2;
    r.by1 = min((s.pyb + rp.halo) >> BIN_LOG2, rp.BH - 1);
    return r;
}

__device__ __forceinline__ bool is_small(const SnappedTri& s, const BinRange& r)
{
    return (s.maxx - s.minx) <= SMALL_EXTENT && (s.maxy - s.miny) <= SMALL_EXTENT && (r.bx1 - r.bx0) <= 1 && (r.by1 - r.by0) <= 1;
}

// vertex indices of triangle t from the 16-byte copy (valid in every kernel that runs after k_setup)
__device__ __forceinline__ int4 tri_indices(const RasterParams& rp, int t) { return __ldg(rp.tri4 + t); }

// Near-plane clipper, the exact mirror of oracle/golden.c: clip_near() / clip_lerp().
__device__ __forceinline__ float4 clip_lerp(const float4& a, float da, const float4& b, float db)
{
    const float t = xdiv(da, xsub(da, db));
    return make_float4(xadd(a.x, xmul(t, xsub(b.x, a.x))), xadd(a.y, xmul(t, xsub(b.y, a.y))),
                       xadd(a.z, xmul(t, xsub(b.z, a.z))), xadd(a.w, xmul(t, xsub(b.w, a.w))));
}

__device__ __forceinline__ int clip_near(const float4 (&v)[3], float4 (&out)[4])
{
    float d[3];
    bool in[3];
    int n = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) { d[i] = xadd(v[i].z, v[i].w); in[i] = d[i] >= 0.f; }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int j = (i + 1) % 3;
        if (in[i]) out[n++] = v[i];
        if (in[i] != in[j]) out[n++] = in[i] ? clip_lerp(v[i], d[i], v[j], d[j]) : clip_lerp(v[j], d[j], v[i], d[i]);
    }
    return n;
}

// id under which the fragments of list entry t compete: the triangle itself, or the parent of a clipped piece
__device__ __forceinline__ int entry_triangle_id(const RasterParams& rp, int t) { return t < rp.T ? t : rp.clip_parent[t - rp.T]; }

template <bool PADDED = true>
__device__ __forceinline__ bool load_triangle(const RasterParams& rp, int n, int t, float4& p0, float4& p1, float4& p2)
{
    if (PADDED && t >= rp.T) {        // a piece of a near-clipped triangle: explicit vertices
        const float4* c = rp.clip_verts + 3 * (size_t)(t - rp.T);
        p0 = c[0]; p1 = c[1]; p2 = c[2];
        return true;
    }
    int i0, i1, i2;
    if (PADDED) { const int4 q = tri_indices(rp, t); i0 = q.x; i1 = q.y; i2 = q.z; }
    else { i0 = __ldg(rp.tri + 3 * t); i1 = __ldg(rp.tri + 3 * t + 1); i2 = __ldg(rp.tri + 3 * t + 2); }
    if ((unsigned)i0 >= (unsigned)rp.V || (unsigned)i1 >= (unsigned)rp.V || (unsigned)i2 >= (unsigned)rp.V) return false;
    const float* P = rp.pos + (size_t)n * rp.V * 4;
    p0 = ldg4(P + 4 * (size_t)i0);
    p1 = ldg4(P + 4 * (size_t)i1);
    p2 = ldg4(P + 4 * (size_t)i2);
    return true;
}

// ---- depth plane (oracle/golden.c: depth_plane / plane_eval) ---------------------------------------------
struct Plane { float zref, dzdx, dzdy; };

__device__ __forceinline__ Plane depth_plane(const float4& p0, const float4& p1, const float4& p2, const SnappedTri& s)
{
    float zv0 = xdiv(p0.z, p0.w), zv1 = xdiv(p1.z, p1.w), zv2 = xdiv(p2.z, p2.w);
    double X1 = (double)(s.x1 - s.x0), Y1 = (double)(s.y1 - s.y0), X2 = (double)(s.x2 - s.x0), Y2 = (double)(s.y2 - s.y0);
    double A = __dsub_rn(__dmul_rn(X1, Y2), __dmul_rn(X2, Y1));
    double dz1 = __dsub_rn((double)zv1, (double)zv0), dz2 = __dsub_rn((double)zv2, (double)zv0);
    double gx = __ddiv_rn(__dsub_rn(__dmul_rn(dz1, Y2), __dmul_rn(dz2, Y1)), A);
    double gy = __ddiv_rn(__dsub_rn(__dmul_rn(dz2, X1), __dmul_rn(dz1, X2)), A);
    double rx = (double)(16 * s.pxa + 8 - s.x0), ry = (double)(16 * s.pya + 8 - s.y0);
    Plane pl;
    pl.zref = __double2float_rn(__dadd_rn(__dadd_rn((double)zv0, __dmul_rn(gx, rx)), __dmul_rn(gy, ry)));
    pl.dzdx = __double2float_rn(__dmul_rn(gx, 16.0));
    pl.dzdy = __double2float_rn(__dmul_rn(gy, 16.0));
    return pl;
}

__device__ __forceinline__ float plane_eval(float zref, float dzdx, float dzdy, int dx, int dy)
{
    return __fmaf_rn(dzdx, (float)dx, __fmaf_rn(dzdy, (float)dy, zref));
}

__device__ __forceinline__ unsigned depth_key(float zw)
{
    // order-preserving map float -> unsigned: negative values are complemented, non-negative ones get the top bit
    const unsigned b = __float_as_uint(zw);
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
constexpr unsigned DEPTH_KEY_MINUS1 = 0x407FFFFFu;     // depth_key(-1.0f)
constexpr unsigned DEPTH_KEY_PLUS1 = 0xBF800000u;      // depth_key(+1.0f)

// depth test + visibility update of one fragment at key slot kp; the CAS loop starts from the value the early-out read
__device__ __forceinline__ void emit_fragment_at(unsigned long long* kp, float zd, int t)
{
    if (!(zd >= -1.f && zd <= 1.f)) return;
    const unsigned long long key = ((unsigned long long)depth_key(zd) << 32) | (unsigned)t;
#if FPC_CAS_MANUAL
    unsigned long long old = *kp;
    while (key < old) {
        const unsigned long long prev = atomicCAS(kp, old, key);
        if (prev == old) break;
        old = prev;
    }
#else
    if (key < *kp) atomicMin(kp, key);
#endif
}

// 32-bit mode (FPC_KEY32): key = (depth_key - base) << idbits | id, one native shared-memory atomicMin per fragment.  A depth
// outside the window the bin promised (never observed; the window carries a margin) raises `overflow` and the CTA redoes the
// bin with 64-bit keys, so the result is the 64-bit result in every case.  The window [base, base + limit) always lies inside
// [depth_key(-1), depth_key(+1)] (raster_tile clamps it), so the one unsigned compare below is also the depth-range test: a
// fragment outside [-1, 1] (or a NaN) lands outside the window and sends the bin to the 64-bit path, which discards it.
struct Key32Mode { unsigned base, limit; int idbits; int* overflow; };

__device__ __forceinline__ void emit_fragment32(unsigned* kp, float zd, int t, const Key32Mode& km)
{
    const unsigned d = depth_key(zd) - km.base;
#if FPC_KEY32_TEST_OVERFLOW
    if ((t % 5) == 0) { *km.overflow = 1; return; }           // test build only: exercises the 64-bit redo of the bin
#endif
    if (d >= km.limit) { *km.overflow = 1; return; }
    atomicMin(kp, (d << km.idbits) | (unsigned)t);
}

template <int TW>
__device__ __forceinline__ void emit_fragment(unsigned long long* keys, float zd, int t, int lx, int ly)
{
    if (!(zd >= -1.f && zd <= 1.f)) return;
    unsigned long long key = ((unsigned long long)depth_key(zd) << 32) | (unsigned)t;
    // 64-bit shared atomicMin is a CAS loop: skip it for fragments that already lose against the stored key
    if (key < keys[ly * TW + lx]) atomicMin(keys + ly * TW + lx, key);
}

// ---- fill rule -------------------------------------------------------------------------------------------
__device__ __forceinline__ int edge_bias(int dx, int dy) { return (dy > 0 || (dy == 0 && dx < 0)) ? 0 : -1; }
__device__ __forceinline__ long long edge_bias64(long long dx, long long dy) { return (dy > 0 || (dy == 0 && dx < 0)) ? 0 : -1; }

// per-warp staging area for a batch of 32 small triangles (structure of arrays: lane-contiguous)
struct WarpStage {
    int e[3][32];        // edge functions (+bias) at the first pixel (xa, ya) of the bin-clipped bbox
    int ax[3][32];       // step per +1 px in x
    int ay[3][32];       // step per +1 px in y
    int xy[32];          // xa | ya << 16          (absolute pixel coordinates)
    int wn[32];          // row width | rows << 16
    int off[32];         // (xa - pxa) | (ya - pya) << 16   offsets from the plane's reference pixel
    float zref[32], dzdx[32], dzdy[32];
    int tri[32];
    int prefix[32];      // inclusive prefix sum of rows
};

__device__ __forceinline__ void emit_any(unsigned long long* kp, float zd, int t, const Key32Mode&) { emit_fragment_at(kp, zd, t); }
__device__ __forceinline__ void emit_any(unsigned* kp, float zd, int t, const Key32Mode& km) { emit_fragment32(kp, zd, t, km); }

// The small triangles of a bin: warps claim batches of 32, every lane sets one triangle up and stages it, the warp then walks
// the batch's (triangle, row) items.  KeyT = unsigned long long (depth_key << 32 | id, CAS-loop atomicMin) or unsigned
// (Key32Mode, native atomicMin).
// `ewin` (nullable, shared memory, EWIN_CAP words): the gather phase of the fused kernels wants, for list entry i, the pixel
// window of the triangle inside this tile and the index k of this bin among the triangle's (at most 2 x 2) bins — both fall
// out of the set-up below:  ewin[i] = (xa-ox) | (xb-ox) << 6 | (ya-oy) << 12 | (yb-oy) << 18 | k << 24, or k << 24 | EWIN_NONE
// when the triangle has no candidate pixel in the tile.
constexpr int EWIN_CAP = 1024;
constexpr unsigned EWIN_NONE = 1u << 26;

template <int TW, int NT, typename KeyT>
__device__ __forceinline__ void walk_small(const RasterParams& rp, int n, const int* __restrict__ list, int count, int ox, int oy,
                                           int min_x, int min_y, int lim_x, int lim_y, KeyT* keys, WarpStage& st, int* next_batch,
                                           const Key32Mode& km, unsigned* ewin, int bin_x, int bin_y)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    (void)warp;
#if FPC_DYN_BATCH
    // (smaller claims for sparsely filled bins were measured: no gain — the barrier stalls after this phase are not a
    // batch-granularity effect, gpurun_out/exp_variants.jsonl round 1)
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(next_batch, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
#else
    for (int base = warp * 32; base < count; base += NT) {
#endif
        int i = base + lane;
        int rows = 0;
        if (i < count) {
            int t = list[i];
            float4 p0, p1, p2;
            SnappedTri s;
            if (load_triangle(rp, n, t, p0, p1, p2) && setup_triangle(p0, p1, p2, rp, s)) {
                int xa = max(s.pxa, min_x), xb = min(s.pxb, lim_x), ya = max(s.pya, min_y), yb = min(s.pyb, lim_y);
                if (ewin && i < EWIN_CAP) {
                    const BinRange br = bin_range(s, rp);
                    const unsigned k = (unsigned)(((bin_y - br.by0) << 1) | (bin_x - br.bx0)) & 3u;
                    ewin[i] = (xa <= xb && ya <= yb) ? ((unsigned)(xa - ox) | ((unsigned)(xb - ox) << 6) | ((unsigned)(ya - oy) << 12) |
                                                        ((unsigned)(yb - oy) << 18) | (k << 24))
                                                     : ((k << 24) | EWIN_NONE);
                }
                if (xa <= xb && ya <= yb) {
                    rows = yb - ya + 1;
                    // oriented vertex order (positive area): swap 1 <-> 2 when flipped
                    int ax1 = s.flip ? s.x2 : s.x1, ay1 = s.flip ? s.y2 : s.y1;
                    int ax2 = s.flip ? s.x1 : s.x2, ay2 = s.flip ? s.y1 : s.y2;
                    int sx = 16 * xa + 8, sy = 16 * ya + 8;
                    int ex0 = ax1 - s.x0, ey0 = ay1 - s.y0;
                    int ex1 = ax2 - ax1, ey1 = ay2 - ay1;
                    int ex2 = s.x0 - ax2, ey2 = s.y0 - ay2;
                    st.e[0][lane] = ex0 * (sy - s.y0) - ey0 * (sx - s.x0) + edge_bias(ex0, ey0);
                    st.e[1][lane] = ex1 * (sy - ay1) - ey1 * (sx - ax1) + edge_bias(ex1, ey1);
                    st.e[2][lane] = ex2 * (sy - ay2) - ey2 * (sx - ax2) + edge_bias(ex2, ey2);
                    st.ax[0][lane] = -16 * ey0; st.ax[1][lane] = -16 * ey1; st.ax[2][lane] = -16 * ey2;
                    st.ay[0][lane] = 16 * ex0;  st.ay[1][lane] = 16 * ex1;  st.ay[2][lane] = 16 * ex2;
                    st.xy[lane] = xa | (ya << 16);
                    st.wn[lane] = (xb - xa + 1) | (rows << 16);
                    st.off[lane] = (xa - s.pxa) | ((ya - s.pya) << 16);
                    Plane pl = depth_plane(p0, p1, p2, s);
                    st.zref[lane] = pl.zref; st.dzdx[lane] = pl.dzdx; st.dzdy[lane] = pl.dzdy;
                    st.tri[lane] = t;
                }
            }
        }
        int incl = rows;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        st.prefix[lane] = incl;
        int total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        for (int k = lane; k < total; k += 32) {
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (st.prefix[j + step - 1] <= k) j += step;
            int wn = st.wn[j];
            int r = k - (st.prefix[j] - (wn >> 16));
            int wd = wn & 0xffff;
            int e0 = st.e[0][j] + r * st.ay[0][j], e1 = st.e[1][j] + r * st.ay[1][j], e2 = st.e[2][j] + r * st.ay[2][j];
            int a0 = st.ax[0][j], a1 = st.ax[1][j], a2 = st.ax[2][j];
            int xy = st.xy[j], off = st.off[j];
            int lx = (xy & 0xffff) - ox, ly = (int)((unsigned)xy >> 16) + r - oy;
            int dx = off & 0xffff, dy = (off >> 16) + r;
            float zref = st.zref[j], dzdx = st.dzdx[j], dzdy = st.dzdy[j];
            int t = st.tri[j];
            float zrow = __fmaf_rn(dzdy, (float)dy, zref);
#if FPC_WALK_MASK
            // coverage of the row as a 32-bit mask (branch-free body), then only the covered span — contiguous, the row cuts
            // a convex region — is depth-tested: lanes do not idle through the misses.  Pixels beyond 32 (halo tiles only)
            // take the plain per-pixel loop.
            const int wm = min(wd, 32);
            unsigned m = 0;
            for (int x = 0; x < wm; x++) {
                m |= ((unsigned)~(e0 | e1 | e2) >> 31) << x;
                e0 += a0; e1 += a1; e2 += a2;
            }
            if (m) {
                const int first = __ffs(m) - 1, cnt = __popc(m);
                KeyT* kp = keys + ly * TW + lx + first;
                float xf = (float)(dx + first);
                for (int c = 0; c < cnt; c++, kp++, xf += 1.f) emit_any(kp, __fmaf_rn(dzdx, xf, zrow), t, km);
            }
            if (TW > 32) {
                for (int x = 32; x < wd; x++) {
                    if ((e0 | e1 | e2) >= 0) emit_any(keys + ly * TW + lx + x, __fmaf_rn(dzdx, (float)(dx + x), zrow), t, km);
                    e0 += a0; e1 += a1; e2 += a2;
                }
            }
#else
            for (int x = 0; x < wd; x++) {
                if ((e0 | e1 | e2) >= 0) emit_any(keys + ly * TW + lx + x, __fmaf_rn(dzdx, (float)(dx + x), zrow), t, km);
                e0 += a0; e1 += a1; e2 += a2;
            }
#endif
        }
        __syncwarp();
    }

}

// Resolve the visibility of the TW x TW pixel tile with origin (ox, oy) — bin `bin` of instance n, widened by
// rp.halo px on every side when TW == BIN + 2 halo (the origin may then be negative) — into keys[TW*TW] (shared
// memory, initialised here).  `stage` is NT/32 WarpStage records in shared memory.  All NT threads of the CTA
// must call this.
// Returns true when the keys were left in the packed 32-bit layout (only if `widen` is false): read them with tile_key().
template <int TW, int NT = FINE_THREADS>
__device__ __forceinline__ bool raster_tile(const RasterParams& rp, int n, int bin, int ox, int oy, unsigned long long* keys, WarpStage* stage,
                                            bool widen = true, unsigned* ewin = nullptr)
{
    const int bin_x = bin % rp.BW, bin_y = bin / rp.BW;
    const int min_x = max(ox, 0), min_y = max(oy, 0);
    const int lim_x = min(ox + TW, rp.W) - 1, lim_y = min(oy + TW, rp.H) - 1;
    const int count = rp.bin_count[(size_t)n * rp.NB + bin];
    const int nlarge = rp.large_count[n];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    __shared__ int next_batch;       // warps claim batches of 32 triangles dynamically (balances uneven batches)
    __shared__ int s_overflow;
    if (threadIdx.x == 0) { next_batch = 0; s_overflow = 0; }
    __syncthreads();

    // ---- small triangles ----
    const int* list = rp.pairs + (size_t)n * 4 * rp.T + rp.bin_offset[(size_t)n * rp.NB + bin];
    WarpStage& st = stage[warp];
    Key32Mode km;
    km.base = 0u; km.limit = 0u; km.idbits = rp.idbits; km.overflow = &s_overflow;
    bool mode32 = false;
#if FPC_KEY32
    // depth window of the bin from the per-triangle vertex depth ranges (k_setup); a fragment of a covered pixel lies inside its
    // triangle, so its plane depth stays within that range up to rounding: KEY32_MARGIN keys of slack on either side
    if (nlarge == 0 && count > 0 && rp.idbits <= 24) {
        // (k_setup accumulated the window per bin while it counted the lists: no pass over the list, no barrier here)
        const unsigned zlo = ~rp.bin_nzlo[(size_t)n * rp.NB + bin], zhi = rp.bin_zhi[(size_t)n * rp.NB + bin];
        const unsigned span = (1u << (32 - rp.idbits)) - 1u;          // the all-ones key stays free for "empty"
        if (span > 2u * KEY32_MARGIN && zlo >= KEY32_MARGIN && zhi >= zlo && (zhi - zlo) < span - 2u * KEY32_MARGIN) {
            // clamp the window to the valid depth range [-1, 1] (see emit_fragment32)
            const unsigned lo = max(zlo - KEY32_MARGIN, DEPTH_KEY_MINUS1);
            const unsigned long long hi = min((unsigned long long)(zlo - KEY32_MARGIN) + span, (unsigned long long)DEPTH_KEY_PLUS1 + 1ull);
            if (hi > lo) {
                mode32 = true;
                km.base = lo; km.limit = (unsigned)(hi - lo);
            }
        }
    }
    if (mode32) {
        unsigned* keys32 = reinterpret_cast<unsigned*>(keys);
        for (int i = threadIdx.x; i < TW * TW; i += NT) keys32[i] = 0xFFFFFFFFu;
        __syncthreads();
        walk_small<TW, NT, unsigned>(rp, n, list, count, ox, oy, min_x, min_y, lim_x, lim_y, keys32, st, &next_batch, km, ewin, bin_x, bin_y);
        __syncthreads();
        if (!s_overflow && !widen) return true;               // the caller decodes the packed keys itself (tile_key)
        if (!s_overflow) {
            // widen in place to the 64-bit layout the shading phases read (only the id part is used downstream): every
            // thread reads its keys, then all write
            constexpr int PER = (TW * TW + NT - 1) / NT;
            unsigned k32[PER];
#pragma unroll
            for (int j = 0; j < PER; j++) { const int i = threadIdx.x + j * NT; k32[j] = (i < TW * TW) ? keys32[i] : 0xFFFFFFFFu; }
            __syncthreads();
            const unsigned idmask = (1u << rp.idbits) - 1u;
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const int i = threadIdx.x + j * NT;
                if (i < TW * TW) keys[i] = (k32[j] == 0xFFFFFFFFu) ? KEY_EMPTY : (unsigned long long)(k32[j] & idmask);
            }
            __syncthreads();
            return false;
        }
        // (never observed) a depth left the window: redo the bin with 64-bit keys
        if (threadIdx.x == 0) next_batch = 0;
        __syncthreads();
    }
#endif
    for (int i = threadIdx.x; i < TW * TW; i += NT) keys[i] = KEY_EMPTY;
    __syncthreads();
    walk_small<TW, NT, unsigned long long>(rp, n, list, count, ox, oy, min_x, min_y, lim_x, lim_y, keys, st, &next_batch, km, ewin, bin_x, bin_y);

    // ---- large triangles: the whole CTA cooperates on each one (16 pixels per thread), int64 edge math ----
    const int* llist = rp.large_list + (size_t)n * 2 * rp.T;
    for (int i = 0; i < nlarge; i++) {
        int t = llist[i];
        float4 p0, p1, p2;
        SnappedTri s;
        if (!load_triangle(rp, n, t, p0, p1, p2) || !setup_triangle(p0, p1, p2, rp, s)) continue;
        int xa = max(s.pxa, min_x), xb = min(s.pxb, lim_x), ya = max(s.pya, min_y), yb = min(s.pyb, lim_y);
        if (xa > xb || ya > yb) continue;
        long long ax1 = s.flip ? s.x2 : s.x1, ay1 = s.flip ? s.y2 : s.y1;
        long long ax2 = s.flip ? s.x1 : s.x2, ay2 = s.flip ? s.y1 : s.y2;
        long long sx = 16 * ox + 8, sy = 16 * oy + 8;
        long long ex0 = ax1 - s.x0, ey0 = ay1 - s.y0, ex1 = ax2 - ax1, ey1 = ay2 - ay1, ex2 = s.x0 - ax2, ey2 = s.y0 - ay2;
        long long b0 = ex0 * (sy - s.y0) - ey0 * (sx - s.x0) + edge_bias64(ex0, ey0);
        long long b1 = ex1 * (sy - ay1) - ey1 * (sx - ax1) + edge_bias64(ex1, ey1);
        long long b2 = ex2 * (sy - ay2) - ey2 * (sx - ax2) + edge_bias64(ex2, ey2);
        Plane pl = depth_plane(p0, p1, p2, s);
        const int tid = entry_triangle_id(rp, t);
        for (int idx = threadIdx.x; idx < TW * TW; idx += NT) {
            int lx = idx % TW, ly = idx / TW;
            int px = ox + lx, py = oy + ly;
            if (px < xa || px > xb || py < ya || py > yb) continue;
            long long r0 = b0 - 16 * ey0 * lx + 16 * ex0 * ly;
            long long r1 = b1 - 16 * ey1 * lx + 16 * ex1 * ly;
            long long r2 = b2 - 16 * ey2 * lx + 16 * ex2 * ly;
            if ((r0 | r1 | r2) >= 0) emit_fragment<TW>(keys, plane_eval(pl.zref, pl.dzdx, pl.dzdy, px - s.pxa, py - s.pya), tid, lx, ly);
        }
    }
    __syncthreads();
    return false;
}

// the plain bin (no halo); keys may come back packed: read them with tile_key(keys, idx, packed, rp.idbits)
__device__ __forceinline__ bool raster_bin(const RasterParams& rp, int n, int bin, unsigned long long* keys, WarpStage* stage, unsigned* ewin = nullptr)
{
    return raster_tile<BIN>(rp, n, bin, (bin % rp.BW) * BIN, (bin / rp.BW) * BIN, keys, stage, false, ewin);
}

// key of tile pixel idx in the 64-bit convention the shading phases use (KEY_EMPTY, or the triangle id in the low 32 bits)
__device__ __forceinline__ unsigned long long tile_key(const unsigned long long* keys, int idx, bool packed, int idbits)
{
    if (!packed) return keys[idx];
    const unsigned k = reinterpret_cast<const unsigned*>(keys)[idx];
    return (k == 0xFFFFFFFFu) ? KEY_EMPTY : (unsigned long long)(k & ((1u << idbits) - 1u));
}

// Host side: scratch layout + the three binning launches.
struct ScratchLayout {
    size_t zero_bytes;               // leading region that must be zeroed each call
    size_t off_count, off_cursor, off_large_count, off_nzlo, off_zhi, off_offset, off_info, off_pairs, off_large, off_anchor, off_tri4, off_bbox, off_valid, off_clip_count, off_clip_verts, off_clip_parent, off_order_count, off_order, total;
    int clip_cap;
};

ScratchLayout raster_layout(int N, int T, int NB);
// Validates, fills rp and enqueues memset + k_setup + k_scan + k_fill.  Returns an fpc_status.
// slot_grad [N*T*4*9] (nullable): gradient slots of the fused kernels (see RasterParams).
int raster_bin_triangles(const char* who, const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                         void* scratch, size_t scratch_bytes, cudaStream_t stream, RasterParams& rp,
                         float* slot_grad = nullptr, int halo = 0,
                         const int32_t* pad_i_src = nullptr, int4* pad_i_dst = nullptr, int pad_i_n = 0, bool launch_order = false);

// slot of (view n, triangle t) in bin (bx, by): k = which of the (at most 2 x 2) bins the small triangle was listed in
__device__ __forceinline__ int slot_index_k(int info, int bx, int by) { return ((by - ((info >> 10) & 1023)) << 1) | (bx - (info & 1023)); }
constexpr int SLOT_FLOATS = 12;      // three corners x float4 (d x, d y, d w, 0): one 16-byte load per (vertex, triangle, bin) in the gather
constexpr unsigned SLOT_CLASS_SMALL = 0x40404040u, SLOT_CLASS_LARGE = 0x80808080u;   // slot_valid as k_setup leaves it (class in bits 6-7 of every byte)
constexpr unsigned char SLOT_WRITTEN = 0x41;                                           // byte k once slot k of a small triangle holds data
constexpr int SLOTS_PER_TRI = 4;

}  // namespace fpc
