// Tile-binned fixed-point software rasterizer for sm_100a  (replaces dr.rasterize, reference fit.py:151).
//
// Pipeline per call (no host read-back, all buffers caller-owned):
//   k_setup : one thread per (instance, triangle): snap to 1/16 px, bbox -> 64x64-px bins touched.
//             small triangles (<= 2x2 bins) bump per-bin counters, large ones go to a per-instance list.
//   k_scan  : per instance exclusive scan of the bin counters (bin list offsets; capacity 4T per instance).
//   k_fill  : one thread per (instance, triangle): append the triangle to its bins' lists.
//   k_fine  : one CTA per (bin, instance): a 64x64 array of 64-bit (depth,id) keys lives in shared memory;
//             threads stage one triangle each from the bin list, walk its bbox and atomicMin the key of every
//             covered pixel; after a barrier the CTA shades its 4096 pixels straight from the keys and
//             writes rast / rast_db — the key buffer never touches HBM.
// Semantics (bit-identical to oracle/golden.c): DESIGN.md "Rasterizer semantics".
#include "raster_core.cuh"

using namespace fpc;

namespace {

// tri_info: bits 0..9 bx0, 10..19 by0, 20 (nbx-1), 21 (nby-1), 22..23 class (0 none, 1 small, 2 large)
//
// k_setup / k_fill run one CTA per chunk of BIN_TPB triangles of ONE instance (grid.y = instance).  When the bin
// grid fits in shared memory (HIST) the CTA first histograms its triangles there and then touches each global
// counter once, so the hot counters (a head covers only ~1/4 of the bins) see ~10x fewer global atomics.
#ifndef FPC_BIN_TPB
#define FPC_BIN_TPB 1024
#endif
constexpr int BIN_TPB = FPC_BIN_TPB;
constexpr int HIST_MAX_BINS = 4096;        // three shared-memory tables of NB words each in k_setup (48 KB at 2048 x 2048)

// Origin of a triangle's gradient moments (fused.cu): the pixel that holds the centroid, clamped to the image.  The position
// gradient is assembled from sums of g_k times (pixel - origin); with the origin ON the triangle the vertex offsets stay as
// small as in a per-pixel evaluation.  (A corner of the bounding box is far from a diagonal sliver in its thin direction:
// measured 20x the rounding error of the op-level kernel on such triangles.)
__device__ __forceinline__ int moment_origin(const SnappedTri& s, const RasterParams& rp)
{
    const int cx = (s.x0 + s.x1 + s.x2) / 3, cy = (s.y0 + s.y1 + s.y2) / 3;       // 1/16 px
    const int px = min(max(cx >> 4, 0), rp.W - 1), py = min(max(cy >> 4, 0), rp.H - 1);
    return px | (py << 16);
}

template <bool HIST>
__global__ void __launch_bounds__(BIN_TPB) k_setup(RasterParams rp)
{
    extern __shared__ int hist[];            // HIST: count [NB] | ~min depth key [NB] | max depth key [NB] of the CTA's triangles per bin
    const int n = blockIdx.y;
    const int t = blockIdx.x * BIN_TPB + threadIdx.x;
    unsigned* s_nzlo = reinterpret_cast<unsigned*>(hist) + rp.NB;
    unsigned* s_zhi = s_nzlo + rp.NB;
    if (HIST) {
        for (int b = threadIdx.x; b < 3 * rp.NB; b += BIN_TPB) hist[b] = 0;
        __syncthreads();
    }
    if (t < rp.T) {
        const size_t gid = (size_t)n * rp.T + t;
        float4 p0, p1, p2;
        SnappedTri s;
        int info = 0;
        bool large = false;
        if (load_triangle<false>(rp, n, t, p0, p1, p2)) {
            if (p0.w > 0.f && p1.w > 0.f && p2.w > 0.f) {
                if (setup_triangle(p0, p1, p2, rp, s)) {
                    rp.tri_anchor[gid] = moment_origin(s, rp);
                    const BinRange br = bin_range(s, rp);
                    const int bx0 = br.bx0, bx1 = br.bx1, by0 = br.by0, by1 = br.by1;
                    if (!is_small(s, br)) {
                        int slot = atomicAdd(rp.large_count + n, 1);
                        rp.large_list[(size_t)n * 2 * rp.T + slot] = t;
                        info = 2 << 22;
                        large = true;
                    } else {
                        rp.tri_bbox[gid] = make_ushort4((unsigned short)s.pxa, (unsigned short)s.pya, (unsigned short)s.pxb, (unsigned short)s.pyb);
                        info = bx0 | (by0 << 10) | ((bx1 - bx0) << 20) | ((by1 - by0) << 21) | (1 << 22);
                        // depth window of the triangle (32-bit key mode of the fine rasterizer): the per-vertex z/w of depth_plane()
                        const unsigned k0 = depth_key(xdiv(p0.z, p0.w)), k1 = depth_key(xdiv(p1.z, p1.w)), k2 = depth_key(xdiv(p2.z, p2.w));
                        const unsigned nzl = ~min(k0, min(k1, k2)), zh = max(k0, max(k1, k2));
                        for (int by = by0; by <= by1; by++)
                            for (int bx = bx0; bx <= bx1; bx++) {
                                const int b = by * rp.BW + bx;
                                if (HIST) { atomicAdd(hist + b, 1); atomicMax(s_nzlo + b, nzl); atomicMax(s_zhi + b, zh); }
                                else {
                                    atomicAdd(rp.bin_count + (size_t)n * rp.NB + b, 1);
                                    atomicMax(rp.bin_nzlo + (size_t)n * rp.NB + b, nzl); atomicMax(rp.bin_zhi + (size_t)n * rp.NB + b, zh);
                                }
                            }
                    }
                }
            } else {
                // a vertex at w <= 0: clip against the near plane; the (at most two) pieces go to the instance's large list
                // with explicit vertices and compete under the parent's id (rare path: the whole CTA walks them)
                const float4 v[3] = {p0, p1, p2};
                float4 c[4];
                const int nc = clip_near(v, c);
                bool anchored = false;
                for (int k = 0; k + 2 < nc; k++) {
                    if (!setup_triangle(c[0], c[k + 1], c[k + 2], rp, s)) continue;
                    const int slot = atomicAdd(rp.clip_count, 1);
                    if (slot >= rp.clip_cap) break;                    // pool exhausted: piece dropped (sized N*T/32 + 1024)
                    rp.clip_verts[3 * (size_t)slot] = c[0]; rp.clip_verts[3 * (size_t)slot + 1] = c[k + 1]; rp.clip_verts[3 * (size_t)slot + 2] = c[k + 2];
                    rp.clip_parent[slot] = t;
                    rp.large_list[(size_t)n * 2 * rp.T + atomicAdd(rp.large_count + n, 1)] = rp.T + slot;
                    if (!anchored) { rp.tri_anchor[gid] = moment_origin(s, rp); anchored = true; }
                    info = 2 << 22;
                    large = true;
                }
            }
        }
        rp.tri_info[gid] = info;
        // class of the triangle for the vertex gather; no gradient slot of this (view, triangle) written yet
        if (rp.slot_grad) rp.slot_valid[gid] = large ? SLOT_CLASS_LARGE : ((info >> 22) == 1 ? SLOT_CLASS_SMALL : 0u);
        // large (and near-clipped) triangles are resolved by whole CTAs in every bin they touch: their gradient is accumulated
        // (float REDs) in slots 1 (moments) and 2 (antialias corner terms), which start from zero
        if (large && rp.slot_grad) {
            float* z = rp.slot_grad + (gid * SLOTS_PER_TRI + 1) * SLOT_FLOATS;
#pragma unroll
            for (int i = 0; i < 2 * SLOT_FLOATS; i++) z[i] = 0.f;
        }
    }
    if (n == 0) {
        // 16-byte copies of the index / attribute arrays the per-pixel phases gather from (one load instead of three)
        if (t < rp.T) rp.tri4[t] = make_int4(__ldg(rp.tri + 3 * t), __ldg(rp.tri + 3 * t + 1), __ldg(rp.tri + 3 * t + 2), 0);
        for (int i = t; i < rp.pad_i_n; i += gridDim.x * BIN_TPB)
            rp.pad_i_dst[i] = make_int4(__ldg(rp.pad_i_src + 3 * i), __ldg(rp.pad_i_src + 3 * i + 1), __ldg(rp.pad_i_src + 3 * i + 2), 0);
    }
    if (HIST) {
        __syncthreads();
        for (int b = threadIdx.x; b < rp.NB; b += BIN_TPB) {
            int c = hist[b];
            if (c) {
                atomicAdd(rp.bin_count + (size_t)n * rp.NB + b, c);
                atomicMax(rp.bin_nzlo + (size_t)n * rp.NB + b, s_nzlo[b]);
                atomicMax(rp.bin_zhi + (size_t)n * rp.NB + b, s_zhi[b]);
            }
        }
    }
}

// one CTA per instance; NB is small (256 at 1024^2, 1024 at 2048^2)
__global__ void __launch_bounds__(256) k_scan(RasterParams rp)
{
    __shared__ int warp_sums[8];
    __shared__ int carry;
    int n = blockIdx.x;
    const int* cnt = rp.bin_count + (size_t)n * rp.NB;
    int* off = rp.bin_offset + (size_t)n * rp.NB;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < rp.NB; base += 256) {
        int i = base + threadIdx.x;
        int v = (i < rp.NB) ? cnt[i] : 0;
        int x = v;
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= d) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (threadIdx.x >> 5); w++) wbase += warp_sums[w];
        int c = carry;
        if (i < rp.NB) off[i] = c + wbase + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + wbase + x;
        __syncthreads();
    }
}

// HIST variant: no k_scan launch — every CTA recomputes the exclusive scan of its instance's bin counters in shared
// memory (NB <= 8192 ints: a few hundred cycles) and the first CTA of the instance publishes it for the fine kernels.
template <bool HIST>
__global__ void __launch_bounds__(BIN_TPB) k_fill(RasterParams rp)
{
    extern __shared__ int hist[];            // HIST: hist [NB] | soff [NB]
    __shared__ int warp_tot[32];
    const int n = blockIdx.y;
    const int t = blockIdx.x * BIN_TPB + threadIdx.x;
    int* soff = hist + rp.NB;
    if (HIST) {
        const int per = (rp.NB + BIN_TPB - 1) / BIN_TPB;                 // consecutive bins per thread
        const int b0 = threadIdx.x * per;
        const int* cnt = rp.bin_count + (size_t)n * rp.NB;
        int local = 0;
        for (int k = 0; k < per; k++) {
            int b = b0 + k;
            int c = (b < rp.NB) ? cnt[b] : 0;
            if (b < rp.NB) { hist[b] = 0; soff[b] = local; }
            local += c;
        }
        int x = local;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = (lane < BIN_TPB / 32) ? warp_tot[lane] : 0, xs = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, xs, d);
                if (lane >= d) xs += y;
            }
            warp_tot[lane] = xs - w;                                     // exclusive over warps
        }
        __syncthreads();
        const int base = warp_tot[warp] + x - local;                     // exclusive prefix of this thread's first bin
        for (int k = 0; k < per; k++) {
            int b = b0 + k;
            if (b < rp.NB) {
                int o = soff[b] + base;
                soff[b] = o;
                if (blockIdx.x == 0) rp.bin_offset[(size_t)n * rp.NB + b] = o;
            }
        }
        __syncthreads();
        // the first CTA of the view also files the view's bins by list-length class (launch order of the fused kernels)
        if (blockIdx.x == 0 && rp.bin_order) {
            __shared__ int s_cls[ORDER_CLASSES], s_base[ORDER_CLASSES];
            if (threadIdx.x < ORDER_CLASSES) s_cls[threadIdx.x] = 0;
            __syncthreads();
            const int any_large = rp.large_count[n] != 0;               // large triangles are walked in every bin of the view
            const int g = n / ORDER_GROUP_VIEWS, nin = n - g * ORDER_GROUP_VIEWS;
            int cls[HIST_MAX_BINS / BIN_TPB], rk[HIST_MAX_BINS / BIN_TPB];
#pragma unroll
            for (int k = 0; k < HIST_MAX_BINS / BIN_TPB; k++) {
                const int b = threadIdx.x + k * BIN_TPB;
                if (b < rp.NB) { cls[k] = order_class(cnt[b] + any_large); rk[k] = atomicAdd(s_cls + cls[k], 1); }
            }
            __syncthreads();
            if (threadIdx.x < ORDER_CLASSES) s_base[threadIdx.x] = atomicAdd(rp.order_count + ORDER_CLASSES * g + threadIdx.x, s_cls[threadIdx.x]);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < HIST_MAX_BINS / BIN_TPB; k++) {
                const int b = threadIdx.x + k * BIN_TPB;
                if (b < rp.NB) rp.bin_order[((size_t)g * ORDER_CLASSES + cls[k]) * ((size_t)ORDER_GROUP_VIEWS * rp.NB) + s_base[cls[k]] + rk[k]] = (nin << 16) | b;
            }
        }
    }
    int info = (t < rp.T) ? rp.tri_info[(size_t)n * rp.T + t] : 0;
    const bool small = (info >> 22) == 1;
    const int bx0 = info & 1023, by0 = (info >> 10) & 1023, nbx = (info >> 20) & 1, nby = (info >> 21) & 1;
    int* pairs = rp.pairs + (size_t)n * 4 * rp.T;
    const int* offset = HIST ? soff : rp.bin_offset + (size_t)n * rp.NB;
    int* cursor = rp.bin_cursor + (size_t)n * rp.NB;
    if (!HIST) {
        if (small)
            for (int by = by0; by <= by0 + nby; by++)
                for (int bx = bx0; bx <= bx0 + nbx; bx++) {
                    int b = by * rp.BW + bx;
                    pairs[offset[b] + atomicAdd(cursor + b, 1)] = t;
                }
        return;
    }
    int rank[4] = {0, 0, 0, 0};
    if (small) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((k & 1) <= nbx && (k >> 1) <= nby) rank[k] = atomicAdd(hist + (by0 + (k >> 1)) * rp.BW + bx0 + (k & 1), 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < rp.NB; b += BIN_TPB) {
        int c = hist[b];
        if (c) hist[b] = offset[b] + atomicAdd(cursor + b, c);      // count -> first slot of this CTA's entries
    }
    __syncthreads();
    if (small) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((k & 1) <= nbx && (k >> 1) <= nby) pairs[hist[(by0 + (k >> 1)) * rp.BW + bx0 + (k & 1)] + rank[k]] = t;
    }
}

__global__ void __launch_bounds__(FINE_THREADS) k_fine(RasterParams rp, float* __restrict__ rast, float* __restrict__ rast_db)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    WarpStage* stage = reinterpret_cast<WarpStage*>(smem + sizeof(unsigned long long) * BIN * BIN);
    const int bin = blockIdx.x, n = blockIdx.y;
    const int ox = (bin % rp.BW) * BIN, oy = (bin / rp.BW) * BIN;
    const bool packed = raster_bin(rp, n, bin, keys, stage);

    // ---- shade: (u, v, z/w, id+1) and the barycentric pixel differentials ----
    const float* P = rp.pos + (size_t)n * rp.V * 4;
    for (int idx = threadIdx.x; idx < BIN * BIN; idx += FINE_THREADS) {
        int lx = idx & (BIN - 1), ly = idx >> BIN_LOG2;
        int px = ox + lx, py = oy + ly;
        if (px >= rp.W || py >= rp.H) continue;
        size_t pi = ((size_t)n * rp.H + py) * rp.W + px;
        unsigned long long key = tile_key(keys, idx, packed, rp.idbits);
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f), odb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (key != KEY_EMPTY) {
            int t = (int)(key & 0xFFFFFFFFu);
            const int4 ti = tri_indices(rp, t);
            const int i0 = ti.x, i1 = ti.y, i2 = ti.z;
            float4 p0 = ldg4(P + 4 * (size_t)i0), p1 = ldg4(P + 4 * (size_t)i1), p2 = ldg4(P + 4 * (size_t)i2);
            float fx = pixel_ndc(px, rp.xs, rp.xo), fy = pixel_ndc(py, rp.ys, rp.yo);
            Shade sh = shade_pixel(p0, p1, p2, fx, fy);
            out = make_float4(clamp01(sh.u), clamp01(sh.v), fminf(fmaxf(sh.zw, -1.f), 1.f), (float)(t + 1));
            if (rast_db) {
                float b0 = sh.u, b1 = sh.v;
                float dfxdx = xmul(rp.xs, sh.iw), dfydy = xmul(rp.ys, sh.iw);
                float da0dx = xsub(xmul(p2.y, p1.w), xmul(p1.y, p2.w)), da0dy = xsub(xmul(p1.x, p2.w), xmul(p2.x, p1.w));
                float da1dx = xsub(xmul(p0.y, p2.w), xmul(p2.y, p0.w)), da1dy = xsub(xmul(p2.x, p0.w), xmul(p0.x, p2.w));
                float da2dx = xsub(xmul(p1.y, p0.w), xmul(p0.y, p1.w)), da2dy = xsub(xmul(p0.x, p1.w), xmul(p1.x, p0.w));
                float datdx = xadd(xadd(da0dx, da1dx), da2dx), datdy = xadd(xadd(da0dy, da1dy), da2dy);
                odb.x = xmul(dfxdx, xsub(xmul(b0, datdx), da0dx));
                odb.y = xmul(dfydy, xsub(xmul(b0, datdy), da0dy));
                odb.z = xmul(dfxdx, xsub(xmul(b1, datdx), da1dx));
                odb.w = xmul(dfydy, xsub(xmul(b1, datdy), da1dy));
            }
        }
        reinterpret_cast<float4*>(rast)[pi] = out;
        if (rast_db) reinterpret_cast<float4*>(rast_db)[pi] = odb;
    }
}

// d pos from (d u, d v): one thread per pixel, float atomics into grad_pos (upstream's scheme).
__global__ void __launch_bounds__(256) k_raster_bwd(const float* __restrict__ pos, const int32_t* __restrict__ tri,
                                                    const float* __restrict__ rast, const float* __restrict__ dy,
                                                    int N, int V, int T, int H, int W, float xs, float xo, float ys, float yo,
                                                    float* __restrict__ grad_pos)
{
    long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= (long long)N * H * W) return;
    float4 r = ldg4(rast + 4 * pi);
    int t = rast_tri(r.w);
    if (t < 0 || t >= T) return;
    float4 g = ldg4(dy + 4 * pi);
    if (g.x == 0.f && g.y == 0.f) return;
    int px = (int)(pi % W), py = (int)((pi / W) % H), n = (int)(pi / ((long long)W * H));
    int i0 = __ldg(tri + 3 * t), i1 = __ldg(tri + 3 * t + 1), i2 = __ldg(tri + 3 * t + 2);
    const float* P = pos + (size_t)n * V * 4;
    float4 q0 = ldg4(P + 4 * (size_t)i0), q1 = ldg4(P + 4 * (size_t)i1), q2 = ldg4(P + 4 * (size_t)i2);
    float fx = pixel_ndc(px, xs, xo), fy = pixel_ndc(py, ys, yo);
    float p0x = q0.x - fx * q0.w, p0y = q0.y - fy * q0.w;
    float p1x = q1.x - fx * q1.w, p1y = q1.y - fy * q1.w;
    float p2x = q2.x - fx * q2.w, p2y = q2.y - fy * q2.w;
    float a0 = p1x * p2y - p1y * p2x, a1 = p2x * p0y - p2y * p0x, a2 = p0x * p1y - p0y * p1x;
    float iw = 1.f / (a0 + a1 + a2);
    float u = a0 * iw, v = a1 * iw;
    float gbb = g.x * u + g.y * v;
    float g0 = iw * (g.x - gbb), g1 = iw * (g.y - gbb), g2 = -iw * gbb;
    float g0x = -g1 * p2y + g2 * p1y, g0y = g1 * p2x - g2 * p1x;
    float g1x = g0 * p2y - g2 * p0y, g1y = -g0 * p2x + g2 * p0x;
    float g2x = -g0 * p1y + g1 * p0y, g2y = g0 * p1x - g1 * p0x;
    float* G = grad_pos + (size_t)n * V * 4;
    atomicAdd(G + 4 * (size_t)i0 + 0, g0x); atomicAdd(G + 4 * (size_t)i0 + 1, g0y); atomicAdd(G + 4 * (size_t)i0 + 3, -fx * g0x - fy * g0y);
    atomicAdd(G + 4 * (size_t)i1 + 0, g1x); atomicAdd(G + 4 * (size_t)i1 + 1, g1y); atomicAdd(G + 4 * (size_t)i1 + 3, -fx * g1x - fy * g1y);
    atomicAdd(G + 4 * (size_t)i2 + 0, g2x); atomicAdd(G + 4 * (size_t)i2 + 1, g2y); atomicAdd(G + 4 * (size_t)i2 + 3, -fx * g2x - fy * g2y);
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

namespace fpc {

ScratchLayout raster_layout(int N, int T, int NB)
{
    ScratchLayout L;
    size_t o = 0;
    L.off_count = o;       o += align_up((size_t)N * NB * 4);
    L.off_cursor = o;      o += align_up((size_t)N * NB * 4);
    L.off_large_count = o; o += align_up((size_t)N * 4);
    L.off_nzlo = o;        o += align_up((size_t)N * NB * 4);
    L.off_zhi = o;         o += align_up((size_t)N * NB * 4);
    L.off_clip_count = o;  o += align_up(4);
    const size_t ngroups = ((size_t)N + ORDER_GROUP_VIEWS - 1) / ORDER_GROUP_VIEWS;
    L.off_order_count = o; o += align_up(ngroups * ORDER_CLASSES * 4);
    L.zero_bytes = o;
    L.off_offset = o;      o += align_up((size_t)N * NB * 4);
    L.off_info = o;        o += align_up((size_t)N * T * 4);
    L.off_pairs = o;       o += align_up((size_t)N * T * 16);
    L.off_large = o;       o += align_up((size_t)N * T * 8);
    L.clip_cap = (int)(((size_t)N * T) / 32 + 1024 < 0x3fffffff ? ((size_t)N * T) / 32 + 1024 : 0x3fffffff);
    L.off_clip_verts = o;  o += align_up((size_t)L.clip_cap * 48);
    L.off_clip_parent = o; o += align_up((size_t)L.clip_cap * 4);
    L.off_anchor = o;      o += align_up((size_t)N * T * 4);
    L.off_tri4 = o;        o += align_up((size_t)T * 16);
    L.off_bbox = o;        o += align_up((size_t)N * T * 8);
    L.off_valid = o;       o += align_up((size_t)N * T * 4);
    L.off_order = o;       o += align_up(ngroups * ORDER_CLASSES * ORDER_GROUP_VIEWS * (size_t)NB * 4);
    L.total = o;
    return L;
}

int raster_bin_triangles(const char* who, const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                         void* scratch, size_t scratch_bytes, cudaStream_t stream, RasterParams& rp,
                         float* slot_grad, int halo, const int32_t* pad_i_src, int4* pad_i_dst, int pad_i_n, bool launch_order)
{
    FPC_CHECK_ARG(pos && tri, "%s: pos and tri must be non-null", who);
    FPC_CHECK_ARG(N > 0 && V > 0 && T > 0 && H > 0 && W > 0, "%s: N, V, T, H, W must be positive (got %d %d %d %d %d)", who, N, V, T, H, W);
    FPC_CHECK_ARG(T < (1 << 24), "%s: at most 2^24-1 triangles (got %d)", who, T);
    FPC_CHECK_ARG(H <= 32768 && W <= 32768 && N <= 65535, "%s: resolution <= 32768^2 and N <= 65535 (got %dx%d, N=%d)", who, H, W, N);
    rp.pos = pos; rp.tri = tri; rp.N = N; rp.V = V; rp.T = T; rp.H = H; rp.W = W;
    rp.BW = fpc_div_up(W, BIN); rp.BH = fpc_div_up(H, BIN); rp.NB = rp.BW * rp.BH; rp.halo = halo;
    ScratchLayout L = raster_layout(N, T, rp.NB);
    FPC_CHECK_ARG(scratch && scratch_bytes >= L.total, "%s: scratch too small (%zu < %zu bytes)", who, scratch_bytes, L.total);
    rp.xs = 2.0f / (float)W; rp.xo = 1.0f / (float)W - 1.0f;
    rp.ys = 2.0f / (float)H; rp.yo = 1.0f / (float)H - 1.0f;
    rp.sxs = 8.0f * (float)W; rp.sys = 8.0f * (float)H;
    char* s = (char*)scratch;
    rp.bin_count = (int*)(s + L.off_count);
    rp.bin_cursor = (int*)(s + L.off_cursor);
    rp.large_count = (int*)(s + L.off_large_count);
    rp.bin_offset = (int*)(s + L.off_offset);
    rp.tri_info = (int*)(s + L.off_info);
    rp.pairs = (int*)(s + L.off_pairs);
    rp.large_list = (int*)(s + L.off_large);
    rp.tri_anchor = (int*)(s + L.off_anchor);
    rp.tri4 = (int4*)(s + L.off_tri4);
    rp.bin_nzlo = (unsigned*)(s + L.off_nzlo);
    rp.bin_zhi = (unsigned*)(s + L.off_zhi);
    rp.tri_bbox = (ushort4*)(s + L.off_bbox);
    rp.slot_valid = (unsigned*)(s + L.off_valid);
    rp.idbits = 1;
    while ((1 << rp.idbits) < T) rp.idbits++;
    rp.clip_count = (int*)(s + L.off_clip_count);
    rp.clip_verts = (float4*)(s + L.off_clip_verts);
    rp.clip_parent = (int*)(s + L.off_clip_parent);
    rp.clip_cap = L.clip_cap;
    rp.pad_i_src = pad_i_src; rp.pad_i_dst = pad_i_dst; rp.pad_i_n = pad_i_src ? pad_i_n : 0;
    rp.order_count = (int*)(s + L.off_order_count);
    rp.bin_order = (launch_order && rp.NB <= HIST_MAX_BINS) ? (int*)(s + L.off_order) : nullptr;       // (written by k_fill<true>)
    FPC_CUDA(cudaMemsetAsync(s, 0, L.zero_bytes, stream));
    rp.slot_grad = slot_grad;
    dim3 grid(fpc_div_up(T, BIN_TPB), N);
    if (rp.NB <= HIST_MAX_BINS) {
        size_t hb = (size_t)rp.NB * sizeof(int);
        k_setup<true><<<grid, BIN_TPB, 3 * hb, stream>>>(rp);
        FPC_LAUNCH_CHECK();
        static FpcPerDeviceOnce fill_attr_set;
        if (fill_attr_set.need()) {
            FPC_CUDA(cudaFuncSetAttribute(k_fill<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * HIST_MAX_BINS * (int)sizeof(int)));
            fill_attr_set.done();
        }
        k_fill<true><<<grid, BIN_TPB, 2 * hb, stream>>>(rp);              // scans the bin counters itself
    } else {
        k_setup<false><<<grid, BIN_TPB, 0, stream>>>(rp);
        FPC_LAUNCH_CHECK();
        k_scan<<<N, 256, 0, stream>>>(rp);
        FPC_LAUNCH_CHECK();
        k_fill<false><<<grid, BIN_TPB, 0, stream>>>(rp);
    }
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

}  // namespace fpc

extern "C" int fpc_raster_bin_px(void) { return BIN; }

extern "C" int fpc_rasterize_clip_pieces(const void* scratch, int N, int T, int H, int W, int* requested, int* capacity, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(scratch && requested && capacity, "rasterize_clip_pieces: null pointer argument");
    FPC_CHECK_ARG(N > 0 && T > 0 && H > 0 && W > 0, "rasterize_clip_pieces: N, T, H, W must be positive");
    const ScratchLayout L = raster_layout(N, T, fpc_div_up(W, BIN) * fpc_div_up(H, BIN));
    FPC_CUDA(cudaMemcpyAsync(requested, (const char*)scratch + L.off_clip_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
    FPC_CUDA(cudaStreamSynchronize(stream));
    *capacity = L.clip_cap;
    return FPC_OK;
}

extern "C" size_t fpc_rasterize_scratch_bytes(int N, int T, int H, int W)
{
    if (N <= 0 || T <= 0 || H <= 0 || W <= 0) return 256;
    int NB = fpc_div_up(W, BIN) * fpc_div_up(H, BIN);
    return raster_layout(N, T, NB).total;
}

extern "C" int fpc_rasterize_fwd(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                                 float* rast, float* rast_db, void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(rast, "rasterize_fwd: rast must be non-null");
    RasterParams rp;
    int st = raster_bin_triangles("rasterize_fwd", pos, tri, N, V, T, H, W, scratch, scratch_bytes, stream, rp);
    if (st != FPC_OK) return st;
    const size_t smem = sizeof(unsigned long long) * BIN * BIN + sizeof(WarpStage) * FINE_WARPS;
    static FpcPerDeviceOnce attr_set;
    if (attr_set.need()) {
        FPC_CUDA(cudaFuncSetAttribute(k_fine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.done();
    }
    k_fine<<<dim3(rp.NB, N), FINE_THREADS, smem, stream>>>(rp, rast, rast_db);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_rasterize_bwd(const float* pos, const int32_t* tri, const float* rast, const float* dy,
                                 int N, int V, int T, int H, int W, float* grad_pos, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(pos && tri && rast && dy && grad_pos, "rasterize_bwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && V > 0 && T > 0 && H > 0 && W > 0, "rasterize_bwd: N, V, T, H, W must be positive");
    FPC_CUDA(cudaMemsetAsync(grad_pos, 0, (size_t)N * V * 4 * sizeof(float), stream));
    long long npx = (long long)N * H * W;
    k_raster_bwd<<<fpc_div_up(npx, 256), 256, 0, stream>>>(pos, tri, rast, dy, N, V, T, H, W, 2.0f / (float)W,
                                                            1.0f / (float)W - 1.0f, 2.0f / (float)H,
                                                            1.0f / (float)H - 1.0f, grad_pos);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
