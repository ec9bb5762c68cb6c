// Fused render + loss + gradient kernel of the fit iteration (no antialias):
//   rasterize -> interpolate -> [bilinear texture] -> background composite -> image loss
//   -> d loss / d colour -> [texture bwd] -> interpolate bwd -> rasterize bwd -> d loss / d pos_clip
// i.e. reference fit.py:151-158,161,579 and their part of loss.backward() (fit.py:611) in ONE kernel per
// (64x64-px bin, view).  Nothing per-pixel is written to HBM unless the caller asks for the images: the only
// compulsory HBM traffic is reading the reference frame (4C or C bytes / px) and the geometry.
//
// Per CTA:  (1) raster_bin(): visibility keys in shared memory (raster_core.cuh);
//           (2) pixel-parallel shading, loss and d loss / d (a0, a1, a2) per pixel -> shared memory;
//           (3) triangle-parallel gather: the barycentric numerators are affine in the pixel position, so the position
//               gradient of a triangle needs only nine moments of d loss / d a_k over its visible pixels.  One thread per
//               entry of the bin's triangle list walks the entry's bounding box IN A FIXED ORDER, sums the moments of the
//               pixels the triangle won, converts them to the gradient of its three corners and stores the nine floats
//               in the slot of (view, triangle, bin): written exactly once, no atomics, no zero-fill, bit-reproducible;
//           (4) k_vtx_gather (one thread per (view, vertex)) sums, in the fixed order of a vertex -> triangle adjacency
//               list, the slots of the triangles around the vertex -> d loss / d pos_clip.
// (Triangles too large for the bin lists — more than 2 x 2 bins or 128 px, or clipped by the near plane — are resolved
// by whole CTAs; their moments are the one thing still accumulated with float REDs.)
#include "antialias.cuh"
#include "raster_core.cuh"

using namespace fpc;

namespace {

#ifndef FPC_TRIGRAD_REAL
#define FPC_TRIGRAD_REAL float
#endif

struct FusedParams {
    const float* attr;       // [Va, A]  vertex colours (A == C) or uv (A == 2, textured)
    const int32_t* attr_tri; // [T,3]
    const int4* attr_tri4;   // [T] 16-byte copy (k_setup) or null when attr_tri == tri
    int Va, A;
    const float* tex;        // [Ht,Wt,C] or null
    int Ht, Wt;
    const void* ref;         // [N,H,W,C] f32 or u8
    int ref_u8;
    int C;
    float bg, k;             // k = scale / (H*W*C)
    int l1;                  // 0: squared error (fit.py:579), 1: absolute error
    float* grad_pos;         // [N,V,4] (written by k_vtx_gather)
    float* grad_tex;         // [Ht,Wt,C] or null: d loss / d tex accumulated with REDs (cleared by the host function)
    float* rast_out;         // [N,H,W,4] or null
    float* colour_out;       // [N,H,W,C] or null (composited image)
    int vpf, row_lo, row_hi; // band split: views per frame; the first view of a frame renders bin rows >= row_lo, the last < row_hi
    double* loss_partial;    // [N*NB]
    float* slots;            // [N*T*4*9] gradient slots (RasterParams::slot_grad), or null: forward only
};

// band split of the camera-split mode: bin row `by` of view n belongs to this rank?
__device__ __forceinline__ bool outside_band(const FusedParams& fp, int n, int by)
{
    if (fp.vpf <= 1 && fp.row_lo <= 0 && by < fp.row_hi) return false;          // no band split (the common case): no division
    const int c = n % fp.vpf;
    return (c == 0 && by < fp.row_lo) || (c == fp.vpf - 1 && by >= fp.row_hi);
}

// The barycentric numerators are affine in the pixel position:  a_k(px) = C_k + A_k fx + B_k fy  with
//   C0 = x1 y2 - y1 x2, A0 = y1 w2 - w1 y2, B0 = w1 x2 - x1 w2   (and cyclic),
// so d loss / d pos of a triangle needs only nine sums over its visible pixels, m = (S_k, SX_k, SY_k) with
// g_k = d loss / d a_k at the pixel and (lx, ly) the pixel offset from the triangle's anchor pixel (fx0, fy0).
// out = (d x, d y, d w) of corner 0, 1, 2.  Same result as summing k_raster_bwd's per-pixel formula (oracle:
// gold_rasterize_bwd).
__device__ __forceinline__ void triangle_corner_grads(const float* m, float fx0, float fy0, float xs, float ys,
                                                      const float4& q0, const float4& q1, const float4& q2, float* out)
{
    typedef FPC_TRIGRAD_REAL real_t;
    const real_t S0 = m[0], S1 = m[1], S2 = m[2];
    const real_t X0 = (real_t)xs * m[3], X1 = (real_t)xs * m[4], X2 = (real_t)xs * m[5];     // sum g_k (fx - fx0)
    const real_t Y0 = (real_t)ys * m[6], Y1 = (real_t)ys * m[7], Y2 = (real_t)ys * m[8];     // sum g_k (fy - fy0)
    // vertex positions relative to the anchor: p_m = (x_m - fx0 w_m, y_m - fy0 w_m)
    real_t p0x = q0.x - (real_t)fx0 * q0.w, p0y = q0.y - (real_t)fy0 * q0.w;
    real_t p1x = q1.x - (real_t)fx0 * q1.w, p1y = q1.y - (real_t)fy0 * q1.w;
    real_t p2x = q2.x - (real_t)fx0 * q2.w, p2y = q2.y - (real_t)fy0 * q2.w;
    // sum_px g_k p_my(px) = S_k p_my - w_m Y_k ;  sum_px g_k p_mx(px) = S_k p_mx - w_m X_k
#define SY_(k, pmy, wm) (S##k * (pmy) - (wm) * Y##k)
#define SX_(k, pmx, wm) (S##k * (pmx) - (wm) * X##k)
    out[0] = (float)(-SY_(1, p2y, q2.w) + SY_(2, p1y, q1.w));
    out[1] = (float)(SX_(1, p2x, q2.w) - SX_(2, p1x, q1.w));
    out[3] = (float)(SY_(0, p2y, q2.w) - SY_(2, p0y, q0.w));
    out[4] = (float)(-SX_(0, p2x, q2.w) + SX_(2, p0x, q0.w));
    out[6] = (float)(-SY_(0, p1y, q1.w) + SY_(1, p0y, q0.w));
    out[7] = (float)(SX_(0, p1x, q1.w) - SX_(1, p0x, q0.w));
#undef SY_
#undef SX_
    // d loss / d A_k = sum g_k fx, d loss / d B_k = sum g_k fy
    real_t dA0 = fx0 * S0 + X0, dA1 = fx0 * S1 + X1, dA2 = fx0 * S2 + X2;
    real_t dB0 = fy0 * S0 + Y0, dB1 = fy0 * S1 + Y1, dB2 = fy0 * S2 + Y2;
    out[2] = (float)(q2.y * dA1 - q2.x * dB1 - q1.y * dA2 + q1.x * dB2);
    out[5] = (float)(-q2.y * dA0 + q2.x * dB0 + q0.y * dA2 - q0.x * dB2);
    out[8] = (float)(q1.y * dA0 - q1.x * dB0 - q0.y * dA1 + q0.x * dB1);
}

// ---- layout of the per-pixel gradient terms the gather phase reads (shared memory) ----
// After the visibility phase the key tile holds one 32-bit word per pixel in its first BIN*BIN*4 bytes (packed keys, or ids
// narrowed from the 64-bit keys); the three planes g0 | g1 | g2 (d loss / d a_k per pixel) follow it, over the second half of
// the key tile and the WarpStage records, both free by then.
constexpr int PIX = BIN * BIN;
constexpr int SVIS_N = 1024;          // bytes of the "id won a pixel" filter (indexed by id & 1023)
static_assert(sizeof(unsigned long long) * PIX + sizeof(WarpStage) * FINE_WARPS >= 4 * sizeof(float) * PIX, "g planes must fit behind the ids");

// 64-bit keys -> 32-bit ids in place (0xFFFFFFFF = empty); all NT threads of the CTA; ends with a barrier
template <int NPIX, int NT>
__device__ __forceinline__ void narrow_keys(unsigned long long* keys)
{
    constexpr int PER = (NPIX + NT - 1) / NT;
    unsigned id[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int i = threadIdx.x + j * NT;
        const unsigned long long k = (i < NPIX) ? keys[i] : KEY_EMPTY;
        id[j] = (k == KEY_EMPTY) ? 0xFFFFFFFFu : (unsigned)k;
    }
    __syncthreads();
    unsigned* k32 = reinterpret_cast<unsigned*>(keys);
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int i = threadIdx.x + j * NT;
        if (i < NPIX) k32[i] = id[j];
    }
    __syncthreads();
}

// Moments of one list entry (triangle t) over the pixels it won inside the window [xa,xb] x [ya,yb] of a tile whose ids
// are ids[(y - ty0) * TW + (x - tx0)] and whose gradient planes are g[0..2][(y - gy0) * GW + (x - gx0)], in row-major
// pixel order (fixed -> bit-reproducible).  (anx, any) = the triangle's anchor pixel.
template <int TW, int GW>
__device__ __forceinline__ bool gather_moments(const int* __restrict__ ids, const float* __restrict__ g0p, const float* __restrict__ g1p,
                                               const float* __restrict__ g2p, int t, int xa, int xb, int ya, int yb, int tx0, int ty0,
                                               int gx0, int gy0, int anx, int any, float (&m)[9])
{
#pragma unroll
    for (int c = 0; c < 9; c++) m[c] = 0.f;
    bool seen = false;
    for (int y = ya; y <= yb; y++) {
        const int* idr = ids + (y - ty0) * TW - tx0;
        const int gr = (y - gy0) * GW - gx0;
        const float fly = (float)(y - any);
        for (int x = xa; x <= xb; x++) {
            if (idr[x] != t) continue;
            seen = true;
            const float a = g0p[gr + x], b = g1p[gr + x], c = g2p[gr + x];
            const float flx = (float)(x - anx);
            m[0] += a; m[1] += b; m[2] += c;
            m[3] += a * flx; m[4] += b * flx; m[5] += c * flx;
            m[6] += a * fly; m[7] += b * fly; m[8] += c * fly;
        }
    }
    return seen;
}

// Visible list entries (window not empty, id seen by the filter) of the first `ntab` entries, bucketed by window area class
// (floor(log2(area)), 8 classes, largest first) into vis_list; returns their number.  The order inside a class is arbitrary —
// every entry is summed by ONE thread in a fixed pixel order, so the order in which entries are processed does not reach the
// results.  All NT threads of the CTA; ends with a barrier.
template <int NT>
__device__ __forceinline__ int sort_visible_entries(const unsigned* __restrict__ ewin, const unsigned char* __restrict__ svis,
                                                    const int* __restrict__ list, int ntab, unsigned short* __restrict__ vis_list)
{
    __shared__ int s_cls[8], s_cur[8];
    if (threadIdx.x < 8) { s_cls[threadIdx.x] = 0; }
    __syncthreads();
    auto cls_of = [&](int i) -> int {
        const unsigned win = ewin[i];
        if ((win & EWIN_NONE) || !svis[list[i] & (SVIS_N - 1)]) return -1;
        const int area = (int)(((win >> 6) & 63) - (win & 63) + 1) * (int)(((win >> 18) & 63) - ((win >> 12) & 63) + 1);
        return 7 - min(31 - __clz(area), 7);                      // class 0 = the largest windows
    };
#ifndef FPC_EXP_NOSCAN
    for (int i = threadIdx.x; i < ntab; i += NT) {
        const int c = cls_of(i);
        if (c >= 0) atomicAdd(&s_cls[c], 1);
    }
#endif
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 0; c < 8; c++) { s_cur[c] = acc; acc += s_cls[c]; }
        s_cls[0] = acc;
    }
    __syncthreads();
    const int nvis = s_cls[0];
#ifndef FPC_EXP_NOSCAN
    for (int i = threadIdx.x; i < ntab; i += NT) {
        const int c = cls_of(i);
        if (c >= 0) vis_list[atomicAdd(&s_cur[c], 1)] = (unsigned short)i;
    }
#endif
    __syncthreads();
    return nvis;
}

// nine corner-gradient floats -> the slot's three float4 (x, y, w, 0) + the "written" byte
__device__ __forceinline__ void store_slot(float* slots, unsigned* slot_valid, size_t gid, int k, const float (&out)[9])
{
    float4* o = reinterpret_cast<float4*>(slots + (gid * SLOTS_PER_TRI + k) * SLOT_FLOATS);
    o[0] = make_float4(out[0], out[1], out[2], 0.f);
    o[1] = make_float4(out[3], out[4], out[5], 0.f);
    o[2] = make_float4(out[6], out[7], out[8], 0.f);
    reinterpret_cast<unsigned char*>(slot_valid + gid)[k] = SLOT_WRITTEN;
}

__device__ __forceinline__ float* slot_ptr(float* slots, size_t gid, int k) { return slots + (gid * SLOTS_PER_TRI + k) * SLOT_FLOATS; }

// moments of a LARGE triangle's pixel: float REDs into the triangle's accumulator slot 1 (zeroed by k_setup)
__device__ __forceinline__ void large_pixel_moments(float* slots, size_t gid, float g0, float g1, float g2, int px, int py, int an)
{
    float* M = slot_ptr(slots, gid, 1);
    const float flx = (float)(px - (an & 0xffff)), fly = (float)(py - (int)((unsigned)an >> 16));
    const float v[9] = {g0, g1, g2, g0 * flx, g1 * flx, g2 * flx, g0 * fly, g1 * fly, g2 * fly};
#pragma unroll
    for (int c = 0; c < 9; c++)
        if (v[c] != 0.f) atomicAdd(M + c, v[c]);
}

// d loss / d tex of one pixel: g_c times the four bilinear weights, RED into grad_tex [Ht,Wt,C] (texture.cu: k_tex_bwd).
// ix / iy pack the two (already wrapped) texel columns / rows, 16 bits each (textures up to 65535 texels a side).
template <int C>
__device__ __forceinline__ void tex_grad_scatter(const FusedParams& fp, unsigned ix, unsigned iy, float wx, float wy, const float (&g)[C])
{
    const size_t r0 = (size_t)(iy & 0xffffu) * fp.Wt, r1 = (size_t)(iy >> 16) * fp.Wt;
    const unsigned x0 = ix & 0xffffu, x1 = ix >> 16;
    const float w00 = (1.f - wx) * (1.f - wy), w10 = wx * (1.f - wy), w01 = (1.f - wx) * wy, w11 = wx * wy;
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (g[c] == 0.f) continue;
        atomicAdd(fp.grad_tex + (r0 + x0) * C + c, g[c] * w00);
        atomicAdd(fp.grad_tex + (r0 + x1) * C + c, g[c] * w10);
        atomicAdd(fp.grad_tex + (r1 + x0) * C + c, g[c] * w01);
        atomicAdd(fp.grad_tex + (r1 + x1) * C + c, g[c] * w11);
    }
}

// A bin no triangle touches: every pixel is background.  Its loss term sum (ref - 255 bg)^2 is streamed straight from
// the reference frame (16-byte loads when the layout allows), optional image outputs are filled, nothing else runs.
// Returns the CTA's partial loss in thread 0 via `red` (shared, FINE_WARPS doubles).
template <int C, int NT = FINE_THREADS>
__device__ __forceinline__ void background_bin(const RasterParams& rp, const FusedParams& fp, int n, int bin, int ox, int oy, double* red,
                                               const uint4* pre = nullptr)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int esz = fp.ref_u8 ? 1 : 4;
    const int rows = min(BIN, rp.H - oy), wpx = min(BIN, rp.W - ox);
    const unsigned char* rbase = reinterpret_cast<const unsigned char*>(fp.ref);
    const size_t row_bytes = (size_t)rp.W * C * esz;
    const float b255 = 255.f * fp.bg;
    double acc = 0.0;
    const bool fast = (wpx == BIN) && (row_bytes % 16 == 0) && ((reinterpret_cast<size_t>(rbase) & 15) == 0);
    if (fast) {
        const int cpr = (BIN * C * esz) >> 4;                    // 16-byte chunks per tile row
        for (int i = threadIdx.x; i < rows * cpr; i += NT) {
            int r = i / cpr, ch = i - r * cpr;
            // (the caller may have requested this thread's first chunk before it knew the bin was empty: `pre`)
            const uint4 v = (pre && i == (int)threadIdx.x) ? *pre
                                                           : __ldg(reinterpret_cast<const uint4*>(rbase + ((size_t)n * rp.H + oy + r) * row_bytes + (size_t)ox * C * esz + ch * 16));
            const unsigned w4[4] = {v.x, v.y, v.z, v.w};
            float sacc = 0.f;
            if (fp.ref_u8) {
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        float e = (float)((w4[k] >> (8 * b)) & 0xffu) - b255;
                        sacc += loss_term(e, fp.l1);
                    }
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    float e = __uint_as_float(w4[k]) - b255;
                    sacc += loss_term(e, fp.l1);
                }
            }
            acc += (double)sacc;
        }
    } else {
        for (int i = threadIdx.x; i < rows * wpx * C; i += NT) {
            int r = i / (wpx * C), e0 = i - r * (wpx * C);
            size_t gi = (((size_t)n * rp.H + oy + r) * rp.W + ox) * C + e0;
            float rv = fp.ref_u8 ? (float)__ldg(rbase + gi) : __ldg(reinterpret_cast<const float*>(rbase) + gi);
            float e = rv - b255;
            acc += (double)loss_term(e, fp.l1);
        }
    }
    if (fp.rast_out || fp.colour_out) {
        for (int i = threadIdx.x; i < rows * wpx; i += NT) {
            int r = i / wpx, x = i - r * wpx;
            size_t pi = ((size_t)n * rp.H + oy + r) * rp.W + ox + x;
            if (fp.rast_out) reinterpret_cast<float4*>(fp.rast_out)[pi] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (fp.colour_out) {
#pragma unroll
                for (int c = 0; c < C; c++) fp.colour_out[pi * C + c] = fp.bg;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NT / 32; w++) t += red[w];
        fp.loss_partial[(size_t)n * rp.NB + bin] = t;
    }
}

#ifndef FPC_ORDER_MODE
#define FPC_ORDER_MODE 1
#endif
#ifndef FPC_AA_ORDER_MAX_CTAS
#define FPC_AA_ORDER_MAX_CTAS 20480          // the antialias kernel is launched in by-list-length order only up to this many busy CTAs
#endif
#ifndef FPC_FUSED_MINBLOCKS
#define FPC_FUSED_MINBLOCKS 8
#endif

template <int C, bool TEX>
__global__ void __launch_bounds__(FINE_THREADS, FPC_FUSED_MINBLOCKS) k_fused(RasterParams rp, FusedParams fp)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    WarpStage* stage = reinterpret_cast<WarpStage*>(smem + sizeof(unsigned long long) * BIN * BIN);
    __shared__ double red[FINE_WARPS];

    // which (view, bin) this CTA works on: long triangle lists first (k_fill); looked up by one thread — the look-up is ~150
    // instructions, and for the 2/3 of the CTAs that only stream a background tile it would otherwise be their main cost
    __shared__ int s_ticket[3];
    if (threadIdx.x == 0) {
        int n_, bin_;
        s_ticket[2] = ordered_bin<FPC_ORDER_MODE>(rp, n_, bin_);
        s_ticket[0] = n_; s_ticket[1] = bin_;
    }
    __syncthreads();
    const int n = s_ticket[0], bin = s_ticket[1], cls = s_ticket[2];
    const int ox = (bin % rp.BW) * BIN, oy = (bin / rp.BW) * BIN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (outside_band(fp, n, bin / rp.BW)) {            // another rank renders this bin row (shard.view_band_shard)
        if (threadIdx.x == 0) fp.loss_partial[(size_t)n * rp.NB + bin] = 0.0;
        return;
    }
    {
        // ~2/3 of the bins of a head shot hold no triangle: their only work is streaming the reference tile through the loss.  The
        // first chunk of the tile is requested together with the counters that decide it (one memory round trip instead of two).
        // (class 0 of the launch order = no list entry and no large triangle in the view: no need to wait for the counters)
        const bool empty = (cls >= 0) ? (cls == 0) : (rp.bin_count[(size_t)n * rp.NB + bin] == 0 && rp.large_count[n] == 0);
        const int esz0 = fp.ref_u8 ? 1 : 4;
        const unsigned char* rbase = reinterpret_cast<const unsigned char*>(fp.ref);
        const size_t row_bytes = (size_t)rp.W * C * esz0;
        const int cpr = (BIN * C * esz0) >> 4;
        const bool fast = (ox + BIN <= rp.W) && (row_bytes % 16 == 0) && ((reinterpret_cast<size_t>(rbase) & 15) == 0) &&
                          (int)threadIdx.x < min(BIN, rp.H - oy) * cpr;
        uint4 pre = make_uint4(0u, 0u, 0u, 0u);
        if (fast) {
            const int r = threadIdx.x / cpr, ch = threadIdx.x - r * cpr;
            pre = __ldg(reinterpret_cast<const uint4*>(rbase + ((size_t)n * rp.H + oy + r) * row_bytes + (size_t)ox * C * esz0 + ch * 16));
        }
        if (empty) {
            background_bin<C>(rp, fp, n, bin, ox, oy, red, fast ? &pre : nullptr);
            return;
        }
    }

    // ---- (0) the tile of the reference frame starts its way into shared memory now (cp.async), so its HBM / L2
    //          latency is hidden behind the rasterization phase ----
    unsigned char* sref = smem + sizeof(unsigned long long) * BIN * BIN + sizeof(WarpStage) * FINE_WARPS;
    const int esz = fp.ref_u8 ? 1 : 4;
    const int pitch = BIN * C * esz;                          // bytes per tile row (multiple of 16)
    {
        const unsigned char* rbase = reinterpret_cast<const unsigned char*>(fp.ref);
        const int rows = min(BIN, rp.H - oy);
        const size_t row_bytes = (size_t)rp.W * C * esz;
        const bool fast = (ox + BIN <= rp.W) && (row_bytes % 16 == 0) && ((reinterpret_cast<size_t>(rbase) & 15) == 0);
        if (fast) {
            const int cpr = pitch >> 4;
            for (int i = threadIdx.x; i < rows * cpr; i += FINE_THREADS) {
                int r = i / cpr, ch = i - r * cpr;
                const unsigned char* src = rbase + ((size_t)n * rp.H + oy + r) * row_bytes + (size_t)ox * C * esz + ch * 16;
                unsigned dst = (unsigned)__cvta_generic_to_shared(sref + r * pitch + ch * 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        } else {
            const int wpx = min(BIN, rp.W - ox);
            for (int i = threadIdx.x; i < rows * wpx * C; i += FINE_THREADS) {
                int r = i / (wpx * C), e = i - r * (wpx * C);
                size_t gi = (((size_t)n * rp.H + oy + r) * rp.W + ox) * C + e;
                if (fp.ref_u8) sref[r * pitch + e] = __ldg(rbase + gi);
                else reinterpret_cast<float*>(sref + r * pitch)[e] = __ldg(reinterpret_cast<const float*>(rbase) + gi);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // gather-phase helpers filled on the way: per-entry pixel windows (raster set-up) and a 1024-way "this id won a pixel" filter
    unsigned* ewin = reinterpret_cast<unsigned*>(sref + (size_t)PIX * C * esz);
    unsigned char* svis = reinterpret_cast<unsigned char*>(ewin + EWIN_CAP);
    for (int i = threadIdx.x; i < SVIS_N / 4; i += FINE_THREADS) reinterpret_cast<unsigned*>(svis)[i] = 0u;
    const bool packed = raster_bin(rp, n, bin, keys, stage, fp.slots ? ewin : nullptr);
    if (!packed) narrow_keys<PIX, FINE_THREADS>(keys);           // rare: the bin's depth range did not fit the packed keys
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    int* ids = reinterpret_cast<int*>(keys);                      // packed key or id per pixel; rewritten below as id / -1
    float* sg0 = reinterpret_cast<float*>(smem) + PIX;             // d loss / d a_k per pixel (gather phase)
    float* sg1 = sg0 + PIX;
    float* sg2 = sg1 + PIX;
    const unsigned idmask = packed ? ((1u << rp.idbits) - 1u) : 0xFFFFFFFFu;
    const bool any_large = rp.large_count[n] != 0;

    // ---- (2) shade + loss + (d u, d v) per pixel ----
    const float* P = rp.pos + (size_t)n * rp.V * 4;
    double loss_acc = 0.0;
    for (int idx = threadIdx.x; idx < PIX; idx += FINE_THREADS) {
        int lx = idx & (BIN - 1), ly = idx >> BIN_LOG2;
        int px = ox + lx, py = oy + ly;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        const unsigned kw = (unsigned)ids[idx];
        const bool fg = kw != 0xFFFFFFFFu;
        const int t = fg ? (int)(kw & idmask) : -1;
        if (px < rp.W && py < rp.H) {
            size_t pi = ((size_t)n * rp.H + py) * rp.W + px;
            float refv[C];
#pragma unroll
            for (int c = 0; c < C; c++)
                refv[c] = fp.ref_u8 ? (float)sref[ly * pitch + lx * C + c]
                                    : reinterpret_cast<const float*>(sref + ly * pitch)[lx * C + c];
            float col[C];
            float4 rout = make_float4(0.f, 0.f, 0.f, 0.f);
            float zn = 0.f, wn = 1.f;
            float a0c[TEX ? 2 : C], a1c[TEX ? 2 : C], a2c[TEX ? 2 : C];
            float dudc[C], dvdc[C];      // TEX: d colour_c / d texU, d texV
            float su = 0.f, sv = 0.f, siw = 0.f;
            unsigned tix = 0u, tiy = 0u;     // TEX: packed texel columns / rows (lo 16 bits = index 0, hi = index 1)
            float twx = 0.f, twy = 0.f;
            if (fg) {
                const int4 ti = tri_indices(rp, t);
                const int i0 = ti.x, i1 = ti.y, i2 = ti.z;
                float4 p0 = ldg4(P + 4 * (size_t)i0), p1 = ldg4(P + 4 * (size_t)i1), p2 = ldg4(P + 4 * (size_t)i2);
                float fx = pixel_ndc(px, rp.xs, rp.xo), fy = pixel_ndc(py, rp.ys, rp.yo);
                ShadeLazy sh = shade_pixel_lazy(p0, p1, p2, fx, fy);
                zn = sh.zn; wn = sh.wn;
                float u = clamp01(sh.u), v = clamp01(sh.v);
                if (fp.slots) { const ShadeGrad sg = shade_pixel_grad(p0, p1, p2, fx, fy); su = sg.u; sv = sg.v; siw = sg.iw; }
                rout = make_float4(u, v, 0.f, (float)(t + 1));
                int j0 = i0, j1 = i1, j2 = i2;
                if (fp.attr_tri4) { const int4 tj = __ldg(fp.attr_tri4 + t); j0 = tj.x; j1 = tj.y; j2 = tj.z; }
                bool ok = (unsigned)j0 < (unsigned)fp.Va && (unsigned)j1 < (unsigned)fp.Va && (unsigned)j2 < (unsigned)fp.Va;
                constexpr int AA = TEX ? 2 : C;
                float b2 = 1.f - u - v;
                float at[AA];
                if (TEX) {           // uv pairs: one 8-byte load per corner
                    const float2 z2 = make_float2(0.f, 0.f);
                    const float2 q0 = ok ? __ldg(reinterpret_cast<const float2*>(fp.attr) + j0) : z2;
                    const float2 q1 = ok ? __ldg(reinterpret_cast<const float2*>(fp.attr) + j1) : z2;
                    const float2 q2 = ok ? __ldg(reinterpret_cast<const float2*>(fp.attr) + j2) : z2;
                    a0c[0] = q0.x; a0c[AA - 1] = q0.y; a1c[0] = q1.x; a1c[AA - 1] = q1.y; a2c[0] = q2.x; a2c[AA - 1] = q2.y;
                } else {
#pragma unroll
                    for (int c = 0; c < AA; c++) {
                        a0c[c] = ok ? __ldg(fp.attr + (size_t)j0 * AA + c) : 0.f;
                        a1c[c] = ok ? __ldg(fp.attr + (size_t)j1 * AA + c) : 0.f;
                        a2c[c] = ok ? __ldg(fp.attr + (size_t)j2 * AA + c) : 0.f;
                    }
                }
#pragma unroll
                for (int c = 0; c < AA; c++) at[c] = u * a0c[c] + v * a1c[c] + b2 * a2c[c];
                if (TEX) {
                    // bilinear, wrap (texture.cu: tex_index)
                    float tu = at[0] - floorf(at[0]), tv = at[1] - floorf(at[1]);
                    float x = xsub(xmul(tu, (float)fp.Wt), 0.5f), y = xsub(xmul(tv, (float)fp.Ht), 0.5f);
                    float x0f = floorf(x), y0f = floorf(y);
                    int ix0 = (int)x0f, iy0 = (int)y0f, ix1 = ix0 + 1, iy1 = iy0 + 1;
                    float wx = x - x0f, wy = y - y0f;
                    if (ix0 < 0) ix0 += fp.Wt;
                    if (iy0 < 0) iy0 += fp.Ht;
                    if (ix1 >= fp.Wt) ix1 -= fp.Wt;
                    if (iy1 >= fp.Ht) iy1 -= fp.Ht;
                    size_t i00 = (size_t)iy0 * fp.Wt + ix0, i10 = (size_t)iy0 * fp.Wt + ix1;
                    size_t i01 = (size_t)iy1 * fp.Wt + ix0, i11 = (size_t)iy1 * fp.Wt + ix1;
                    tix = (unsigned)ix0 | ((unsigned)ix1 << 16); tiy = (unsigned)iy0 | ((unsigned)iy1 << 16);
                    twx = wx; twy = wy;
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        float t00 = __ldg(fp.tex + i00 * C + c), t10 = __ldg(fp.tex + i10 * C + c);
                        float t01 = __ldg(fp.tex + i01 * C + c), t11 = __ldg(fp.tex + i11 * C + c);
                        float a = t00 + (t10 - t00) * wx, b = t01 + (t11 - t01) * wx;
                        col[c] = a + (b - a) * wy;
                        dudc[c] = (float)fp.Wt * ((t10 - t00) * (1.f - wy) + (t11 - t01) * wy);
                        dvdc[c] = (float)fp.Ht * ((t01 - t00) * (1.f - wx) + (t11 - t10) * wx);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++) col[c] = at[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < C; c++) col[c] = fp.bg;
            }
            float gc[C];
#pragma unroll
            for (int c = 0; c < C; c++) {
                float e = refv[c] - 255.f * col[c];
                loss_acc += (double)loss_term(e, fp.l1);
                gc[c] = loss_dcolour(e, fp.k, fp.l1);
            }
            if (fg) {
                float gu = 0.f, gv = 0.f;
                if (TEX && fp.grad_tex) tex_grad_scatter<C>(fp, tix, tiy, twx, twy, gc);
                if (TEX) {
                    float gU = 0.f, gV = 0.f;
#pragma unroll
                    for (int c = 0; c < C; c++) { gU += gc[c] * dudc[c]; gV += gc[c] * dvdc[c]; }
                    gu = gU * (a0c[0] - a2c[0]) + gV * (a0c[1] - a2c[1]);
                    gv = gU * (a1c[0] - a2c[0]) + gV * (a1c[1] - a2c[1]);
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++) { gu += gc[c] * (a0c[c] - a2c[c]); gv += gc[c] * (a1c[c] - a2c[c]); }
                }
                // d loss / d a_k  (u = a0/at, v = a1/at, unclamped barycentrics as in the op-level backward)
                float gbb = gu * su + gv * sv;
                g2 = -siw * gbb;
                g0 = siw * gu + g2;
                g1 = siw * gv + g2;
            }
            if (fp.rast_out) {
                if (fg) rout.z = fminf(fmaxf(xdiv(zn, wn), -1.f), 1.f);      // the z/w division only when somebody looks at it
                reinterpret_cast<float4*>(fp.rast_out)[pi] = rout;
            }
            if (fp.colour_out) {
#pragma unroll
                for (int c = 0; c < C; c++) fp.colour_out[pi * C + c] = col[c];
            }
        }
        // ---- per-pixel terms of the gather phase (own slot only: no hazard with the pixels other threads still read) ----
        if (fp.slots) {
            ids[idx] = t;
            sg0[idx] = g0; sg1[idx] = g1; sg2[idx] = g2;
            if (fg) svis[t & (SVIS_N - 1)] = 1;
            if (any_large && fg && (g0 != 0.f || g1 != 0.f || g2 != 0.f)) {
                const size_t gid = (size_t)n * rp.T + t;
                if ((rp.tri_info[gid] >> 22) == 2) large_pixel_moments(fp.slots, gid, g0, g1, g2, px, py, rp.tri_anchor[gid]);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) red[warp] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < FINE_WARPS; w++) s += red[w];
        fp.loss_partial[(size_t)n * rp.NB + bin] = s;
    }
    if (!fp.slots) return;
#ifdef FPC_EXP_NOGATHER
    return;                       // experiment: cost of the gather phase (scripts/exp_variants.py); results are wrong
#endif

    // ---- (3) gather: the moments of every triangle that won pixels, summed over its pixels in a FIXED (row-major) order by
    //          one thread, converted to the gradient of its three corners and stored in the slot of (view, triangle, this
    //          bin).  Entries that won nothing write nothing (their slot stays invalid).  The visible entries are first
    //          bucketed by window area (8 classes) so that the threads of a warp walk windows of similar size. ----
    const int count = rp.bin_count[(size_t)n * rp.NB + bin];
    const int* list = rp.pairs + (size_t)n * 4 * rp.T + rp.bin_offset[(size_t)n * rp.NB + bin];
    unsigned short* vis_list = reinterpret_cast<unsigned short*>(svis + SVIS_N);
    const int nvis = sort_visible_entries<FINE_THREADS>(ewin, svis, list, min(count, EWIN_CAP), vis_list);
    // L = 1, 2 or 4 threads per entry (as many as the CTA has to spare): thread j of an entry's group walks rows j, j + L, ...,
    // the partial sums are combined by a fixed butterfly — fewer, shorter walks at the tail of the CTA
    const int L = (nvis * 4 <= FINE_THREADS) ? 4 : ((nvis * 2 <= FINE_THREADS) ? 2 : 1);
    const int sub = threadIdx.x & (L - 1);
    for (int base = 0; base < nvis; base += FINE_THREADS / L) {
        const int e = base + threadIdx.x / L;
        const bool act = e < nvis;
        const int i = act ? (int)vis_list[e] : 0;
        const int t = act ? list[i] : -2;
        const size_t gid = (size_t)n * rp.T + (act ? t : 0);
        const unsigned win = act ? ewin[i] : (1u << 12);              // inactive: an empty window (wy0 > wy1)
        const int wx0 = win & 63, wx1 = (win >> 6) & 63, wy0 = (win >> 12) & 63, wy1 = (win >> 18) & 63;     // tile-relative window
        // requested now, used after the walk: the triangle's anchor and vertex indices
        int an = 0;
        int4 ti = make_int4(0, 0, 0, 0);
        if (act && sub == 0) { an = rp.tri_anchor[gid]; ti = tri_indices(rp, t); }
        float m[9];
#pragma unroll
        for (int c = 0; c < 9; c++) m[c] = 0.f;
        int seen = 0;
        for (int y = wy0 + sub; y <= wy1; y += L) {
            const int* idr = ids + y * BIN + wx0;
            // which pixels of the row did the triangle win?  (branch-free pass, then only the hits are visited, left to right)
            unsigned hit = 0u;
            for (int x = 0; x <= wx1 - wx0; x++) hit |= (unsigned)(idr[x] == t) << x;
            if (!hit) continue;
            seen = 1;
            const int rowb = y * BIN + wx0;
            float r0 = 0.f, r1 = 0.f, r2 = 0.f;
            do {
                const int x = __ffs(hit) - 1;
                hit &= hit - 1u;
                const float a = sg0[rowb + x], b = sg1[rowb + x], c = sg2[rowb + x];
                const float flx = (float)x;
                r0 += a; r1 += b; r2 += c;
                m[3] += a * flx; m[4] += b * flx; m[5] += c * flx;
            } while (hit);
            const float fly = (float)(y - wy0);
            m[0] += r0; m[1] += r1; m[2] += r2;
            m[6] += r0 * fly; m[7] += r1 * fly; m[8] += r2 * fly;
        }
        for (int o = L >> 1; o > 0; o >>= 1) {                          // all 32 lanes get here (no early exits above)
#pragma unroll
            for (int c = 0; c < 9; c++) m[c] += __shfl_xor_sync(0xffffffffu, m[c], o);
            seen |= __shfl_xor_sync(0xffffffffu, seen, o);
        }
        if (act && sub == 0 && seen) {
            // moments about the window corner -> about the triangle's anchor pixel
            const float dx = (float)(ox + wx0 - (an & 0xffff)), dy = (float)(oy + wy0 - (int)((unsigned)an >> 16));
#pragma unroll
            for (int c = 0; c < 3; c++) { m[3 + c] += dx * m[c]; m[6 + c] += dy * m[c]; }
            const float4 p0 = ldg4(P + 4 * (size_t)ti.x), p1 = ldg4(P + 4 * (size_t)ti.y), p2 = ldg4(P + 4 * (size_t)ti.z);
            const float fx0 = pixel_ndc(an & 0xffff, rp.xs, rp.xo), fy0 = pixel_ndc((int)((unsigned)an >> 16), rp.ys, rp.yo);
            float out[9];
            triangle_corner_grads(m, fx0, fy0, rp.xs, rp.ys, p0, p1, p2, out);
            const int k = (win >> 24) & 3;
            store_slot(fp.slots, rp.slot_valid, gid, k, out);
        }
    }
    // (more entries than the window table holds — > 1024 triangles in a 32 x 32-px bin: one thread per entry, from the binning records)
    const int bx = bin % rp.BW, by = bin / rp.BW;
    for (int i = EWIN_CAP + threadIdx.x; i < count; i += FINE_THREADS) {
        const int t = list[i];
        const size_t gid = (size_t)n * rp.T + t;
        const ushort4 bb = rp.tri_bbox[gid];
        const int an = rp.tri_anchor[gid];
        float m[9];
        if (gather_moments<BIN, BIN>(ids, sg0, sg1, sg2, t, max((int)bb.x, ox), min((int)bb.z, ox + BIN - 1), max((int)bb.y, oy), min((int)bb.w, oy + BIN - 1),
                                     ox, oy, ox, oy, an & 0xffff, (int)((unsigned)an >> 16), m)) {
            const int4 ti = tri_indices(rp, t);
            const float4 p0 = ldg4(P + 4 * (size_t)ti.x), p1 = ldg4(P + 4 * (size_t)ti.y), p2 = ldg4(P + 4 * (size_t)ti.z);
            float out[9];
            triangle_corner_grads(m, pixel_ndc(an & 0xffff, rp.xs, rp.xo), pixel_ndc((int)((unsigned)an >> 16), rp.ys, rp.yo), rp.xs, rp.ys, p0, p1, p2, out);
            const int k = slot_index_k(rp.tri_info[gid], bx, by);
            store_slot(fp.slots, rp.slot_valid, gid, k, out);
        }
    }
}

// fixed-order sum of the per-CTA loss partials (deterministic), by one CTA of 256 threads
__device__ __forceinline__ void loss_reduce_cta(const double* __restrict__ partial, int n, float k, float* __restrict__ loss)
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
    if (threadIdx.x == 0) loss[0] = (float)(red[0] * (double)k);
}

// One thread per (view, vertex): d loss / d pos_clip of the vertex = sum, in the fixed order of its adjacency list, of the
// gradient slots of the triangles around it (each triangle: its at most 2 x 2 bin slots in k order) — a gather: no atomics,
// every element of grad_pos is written exactly once (nothing to zero), bit-reproducible.  vadj_off [V+1], vadj_item [3T]
// (item = triangle * 4 + corner), built once per mesh by fpc_vertex_adjacency_build.  The loss partials of the fused kernel
// are complete by now as well: one extra CTA (the last) sums them, which saves a launch on the critical path.
__global__ void __launch_bounds__(256) k_vtx_gather(RasterParams rp, const float* __restrict__ slots, const int32_t* __restrict__ vadj_off,
                                                    const int32_t* __restrict__ vadj_item, float* __restrict__ grad_pos,
                                                    const double* __restrict__ loss_partial, int n_partial, float k, float* __restrict__ loss)
{
    if (blockIdx.x == gridDim.x - 1) {
        loss_reduce_cta(loss_partial, n_partial, k, loss);
        return;
    }
    const long long gv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gv >= (long long)rp.N * rp.V) return;
    const int n = (int)(gv / rp.V), v = (int)(gv - (long long)n * rp.V);
    float gx = 0.f, gy = 0.f, gw = 0.f;
    const int j0 = __ldg(vadj_off + v), j1 = __ldg(vadj_off + v + 1);
    // four adjacency items at a time: their (independent) index / class loads are in flight together; the sums below still run
    // in list order
    for (int jb = j0; jb < j1; jb += 4) {
        int item[4];
        unsigned valid[4];
#pragma unroll
        for (int u = 0; u < 4; u++) item[u] = (jb + u < j1) ? __ldg(vadj_item + jb + u) : -1;
#pragma unroll
        for (int u = 0; u < 4; u++) valid[u] = (item[u] >= 0) ? __ldg(rp.slot_valid + (size_t)n * rp.T + (item[u] >> 2)) : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = item[u] >> 2, corner = item[u] & 3;
            const size_t gid = (size_t)n * rp.T + t;
            const unsigned vm = valid[u];
            const int cls = (vm >> 6) & 3;
            if (cls == 1) {
                const float4* S = reinterpret_cast<const float4*>(slots + gid * (SLOTS_PER_TRI * SLOT_FLOATS)) + corner;
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                    if ((vm >> (8 * kk)) & 1u) { const float4 g = __ldg(S + 3 * kk); gx += g.x; gy += g.y; gw += g.z; }
            } else if (cls == 2) {
                // large / near-clipped triangle: slot 1 holds its moments (float REDs), slot 2 the antialias corner terms
                const float* M = slots + (gid * SLOTS_PER_TRI + 1) * SLOT_FLOATS;
                float m[9];
                bool any = false;
#pragma unroll
                for (int c = 0; c < 9; c++) { m[c] = M[c]; any = any || (m[c] != 0.f); }
                if (any) {
                    const int4 ti = tri_indices(rp, t);
                    const float* P = rp.pos + (size_t)n * rp.V * 4;
                    const float4 p0 = ldg4(P + 4 * (size_t)ti.x), p1 = ldg4(P + 4 * (size_t)ti.y), p2 = ldg4(P + 4 * (size_t)ti.z);
                    const int an = rp.tri_anchor[gid];
                    float out[9];
                    triangle_corner_grads(m, pixel_ndc(an & 0xffff, rp.xs, rp.xo), pixel_ndc((int)((unsigned)an >> 16), rp.ys, rp.yo), rp.xs, rp.ys, p0, p1, p2, out);
                    gx += out[3 * corner]; gy += out[3 * corner + 1]; gw += out[3 * corner + 2];
                }
                const float* Aa = M + SLOT_FLOATS + 4 * corner;
                gx += Aa[0]; gy += Aa[1]; gw += Aa[2];
            }
        }
    }
    reinterpret_cast<float4*>(grad_pos)[gv] = make_float4(gx, gy, 0.f, gw);
}

// ---- vertex -> (triangle, corner) adjacency (CSR), built once per mesh ------------------------------------------------
__global__ void __launch_bounds__(256) k_vadj_count(const int32_t* __restrict__ tri, int T, int V, int* __restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * T) return;
    const int v = tri[i];
    if ((unsigned)v < (unsigned)V) atomicAdd(cnt + v, 1);
}

// single CTA exclusive scan (V is at most a few 10^5; runs once per mesh)
__global__ void __launch_bounds__(1024) k_vadj_scan(const int* __restrict__ cnt, int V, int32_t* __restrict__ off, int* __restrict__ cursor)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < V; base += 1024) {
        const int i = base + threadIdx.x;
        const int c = (i < V) ? cnt[i] : 0;
        int x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], xs = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, xs, d);
                if (lane >= d) xs += y;
            }
            wsum[lane] = xs - w;
        }
        __syncthreads();
        const int excl = carry + wsum[warp] + x - c;
        if (i < V) { off[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[V] = carry;
}

__global__ void __launch_bounds__(256) k_vadj_fill(const int32_t* __restrict__ tri, int T, int V, int* __restrict__ cursor, int32_t* __restrict__ item)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * T) return;
    const int v = tri[i];
    if ((unsigned)v < (unsigned)V) item[atomicAdd(cursor + v, 1)] = (i / 3) * 4 + (i % 3);
}

// the fill order above depends on the atomics: sort every vertex's (short) list so that the gather order is canonical
__global__ void __launch_bounds__(256) k_vadj_sort(const int32_t* __restrict__ off, int V, int32_t* __restrict__ item)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int a = off[v], b = off[v + 1];
    for (int i = a + 1; i < b; i++) {
        const int x = item[i];
        int j = i - 1;
        while (j >= a && item[j] > x) { item[j + 1] = item[j]; j--; }
        item[j + 1] = x;
    }
}

__global__ void __launch_bounds__(256) k_fused_loss_reduce(const double* __restrict__ partial, int n, float k, float* __restrict__ loss)
{
    loss_reduce_cta(partial, n, k, loss);
}

// keys + per-warp triangle staging + the reference-frame tile (u8 or f32, C channels)
// + the gather phase's window table, visibility filter and list of visible entries
__host__ size_t fused_smem(int C, int ref_u8)
{
    return sizeof(unsigned long long) * BIN * BIN + sizeof(WarpStage) * FINE_WARPS + (size_t)BIN * BIN * C * (ref_u8 ? 1 : 4) + sizeof(unsigned) * EWIN_CAP + SVIS_N +
           sizeof(unsigned short) * EWIN_CAP;
}

template <int C, bool TEX>
int launch_fused(const RasterParams& rp, const FusedParams& fp, cudaStream_t stream)
{
    static FpcPerDeviceOnce attr_set;
    if (attr_set.need()) {
        FPC_CUDA(cudaFuncSetAttribute(k_fused<C, TEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(C, 0)));
        attr_set.done();
    }
    k_fused<C, TEX><<<dim3(rp.NB, rp.N), FINE_THREADS, fused_smem(C, fp.ref_u8), stream>>>(rp, fp);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

#include "fused_aa.cuh"

namespace {

template <int C, bool TEX>
int launch_fused_aa(const RasterParams& rp, const FusedParams& fp, const int32_t* tri_opp, cudaStream_t stream)
{
    static FpcPerDeviceOnce attr_set;
    if (attr_set.need()) {
        FPC_CUDA(cudaFuncSetAttribute(k_fused_aa<C, TEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)aa_smem_layout(C, 4, TEX).total));
        attr_set.done();
    }
    k_fused_aa<C, TEX><<<dim3(rp.NB, rp.N), AA_THREADS, aa_smem_layout(C, fp.ref_u8 ? 1 : 4, TEX && fp.grad_tex).total, stream>>>(rp, fp, tri_opp);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

}  // namespace

extern "C" size_t fpc_render_loss_fused_scratch_bytes(int N, int T, int H, int W)
{
    if (N <= 0 || T <= 0 || H <= 0 || W <= 0) return 256;
    int NB = fpc_div_up(W, BIN) * fpc_div_up(H, BIN);
    return align256(raster_layout(N, T, NB).total) + align256((size_t)N * NB * sizeof(double)) +
           align256((size_t)N * T * SLOTS_PER_TRI * SLOT_FLOATS * sizeof(float)) + align256((size_t)T * sizeof(int4));
}

static int render_loss_fused_impl(const char* who, const float* pos, const int32_t* tri, const int32_t* tri_opp, const float* attr,
                                  const int32_t* attr_tri, int Va, int A, const float* tex, int Ht, int Wt, const void* ref, int ref_is_u8,
                                  int N, int V, int T, int H, int W, int C, float bg, float scale, int loss_kind,
                                  float* loss, float* grad_pos, float* grad_tex, float* rast_out, float* colour_out,
                                  const int32_t* vadj_off, const int32_t* vadj_item, int views_per_frame, int row_lo, int row_hi,
                                  void* scratch, size_t scratch_bytes, cudaStream_t stream)
{
    FPC_CHECK_ARG(!grad_pos || (vadj_off && vadj_item), "%s: grad_pos needs the vertex adjacency of the mesh (fpc_vertex_adjacency_build)", who);
    FPC_CHECK_ARG(attr && attr_tri && ref && loss, "%s: attr, attr_tri, ref and loss must be non-null", who);
    FPC_CHECK_ARG(C == 1 || C == 3, "%s: C must be 1 or 3 (got %d)", who, C);
    FPC_CHECK_ARG(loss_kind == 0 || loss_kind == 1, "%s: loss_kind must be 0 (L2) or 1 (L1), got %d", who, loss_kind);
    FPC_CHECK_ARG(Va > 0, "%s: Va must be positive", who);
    if (tex) FPC_CHECK_ARG(A == 2 && Ht > 0 && Wt > 0, "%s: textured shading needs A == 2 (uv) and a non-empty texture", who);
    else FPC_CHECK_ARG(A == C, "%s: vertex-colour shading needs A == C (got A=%d, C=%d)", who, A, C);
    FPC_CHECK_ARG(scratch_bytes >= fpc_render_loss_fused_scratch_bytes(N, T, H, W), "%s: scratch too small", who);
    if (grad_tex) {
        FPC_CHECK_ARG(tex && grad_pos, "%s: grad_tex needs textured shading and the backward pass (grad_pos)", who);
        FPC_CHECK_ARG(Ht <= 65535 && Wt <= 65535, "%s: grad_tex supports textures up to 65535 texels a side", who);
        FPC_CUDA(cudaMemsetAsync(grad_tex, 0, (size_t)Ht * Wt * C * sizeof(float), stream));
    }
    RasterParams rp;
    const int NB0 = fpc_div_up(W, BIN) * fpc_div_up(H, BIN);
    double* loss_partial = (double*)((char*)scratch + align256(raster_layout(N, T, NB0).total));
    float* slots_mem = (float*)((char*)loss_partial + align256((size_t)N * NB0 * sizeof(double)));
    float* slots = grad_pos ? slots_mem : nullptr;
    int4* attr_tri4 = (attr_tri != tri) ? (int4*)((char*)slots_mem + align256((size_t)N * T * SLOTS_PER_TRI * SLOT_FLOATS * sizeof(float))) : nullptr;
    // no accumulator is zero-filled: every gradient slot and every element of grad_pos is written exactly once (k_setup only
    // zeroes the accumulator slots of large triangles); with antialias the bins are widened by the 2-px halo the fused kernel
    // resolves around its bin
    // Launch order of the fused kernel's CTAs (raster_core.cuh: ordered_bin): longest triangle lists first, background bins last.
    // k_fused gains 10 % from it at config 2.  The antialias kernel LOSES 2.5 % when a rank renders all views (config 3 / 5 on one
    // GPU, 37k - 590k CTAs: neighbouring bins share texels, triangles and vertices, and its background CTAs are slow on their own)
    // but GAINS 6 % when a rank renders a band of ~1 view under the 8-way camera split (4.6k CTAs in the band, 10 waves of 3 CTAs
    // per SM: there the tail of the launch is what counts) and 1.3 % at 2 ranks (18.4k CTAs): it is ordered only up to
    // FPC_AA_ORDER_MAX_CTAS busy CTAs.
    const int bh = fpc_div_up(H, BIN), bw = fpc_div_up(W, BIN);
    long long ctas = (long long)N * bh * bw;
    if (views_per_frame > 0 && row_lo >= 0 && row_hi > 0)
        ctas = (long long)(N / views_per_frame) * ((long long)views_per_frame * bh - row_lo - (bh - row_hi)) * bw;
    const bool launch_order = !tri_opp || ctas <= FPC_AA_ORDER_MAX_CTAS;
    int st = raster_bin_triangles(who, pos, tri, N, V, T, H, W, scratch, scratch_bytes, stream, rp, slots, tri_opp ? AA_HALO : 0,
                                  attr_tri4 ? attr_tri : nullptr, attr_tri4, T, launch_order);
    if (st != FPC_OK) return st;
    FusedParams fp;
    fp.attr = attr; fp.attr_tri = attr_tri; fp.attr_tri4 = attr_tri4; fp.Va = Va; fp.A = A; fp.tex = tex; fp.Ht = Ht; fp.Wt = Wt;
    fp.ref = ref; fp.ref_u8 = ref_is_u8; fp.C = C; fp.bg = bg; fp.k = scale / ((float)H * (float)W * (float)C); fp.l1 = loss_kind;
    fp.grad_pos = grad_pos; fp.grad_tex = grad_tex; fp.rast_out = rast_out; fp.colour_out = colour_out;
    const int rows = fpc_div_up(H, BIN);
    if (views_per_frame <= 0) { views_per_frame = 1; row_lo = 0; row_hi = rows; }         // no band split
    FPC_CHECK_ARG(N % views_per_frame == 0 && row_lo >= 0 && row_lo < rows && row_hi > 0 && row_hi <= rows,
                  "%s: band split needs N %% views_per_frame == 0 and 0 <= row_lo < %d, 0 < row_hi <= %d (got %d, %d, %d)", who, rows, rows,
                  views_per_frame, row_lo, row_hi);
    FPC_CHECK_ARG(views_per_frame > 1 || row_lo < row_hi, "%s: empty band [%d, %d)", who, row_lo, row_hi);
    fp.vpf = views_per_frame; fp.row_lo = row_lo; fp.row_hi = row_hi;
    fp.loss_partial = loss_partial;
    fp.slots = slots;
    if (tri_opp) {
        if (tex) st = (C == 1) ? launch_fused_aa<1, true>(rp, fp, tri_opp, stream) : launch_fused_aa<3, true>(rp, fp, tri_opp, stream);
        else st = (C == 1) ? launch_fused_aa<1, false>(rp, fp, tri_opp, stream) : launch_fused_aa<3, false>(rp, fp, tri_opp, stream);
    } else {
        if (tex) st = (C == 1) ? launch_fused<1, true>(rp, fp, stream) : launch_fused<3, true>(rp, fp, stream);
        else st = (C == 1) ? launch_fused<1, false>(rp, fp, stream) : launch_fused<3, false>(rp, fp, stream);
    }
    if (st != FPC_OK) return st;
    if (grad_pos) {
        // + 1 CTA: the loss reduction rides along
        k_vtx_gather<<<fpc_div_up((long long)N * V, 256) + 1, 256, 0, stream>>>(rp, fp.slots, vadj_off, vadj_item, grad_pos, fp.loss_partial, N * rp.NB, fp.k, loss);
        FPC_LAUNCH_CHECK();
    } else {
        k_fused_loss_reduce<<<1, 256, 0, stream>>>(fp.loss_partial, N * rp.NB, fp.k, loss);
        FPC_LAUNCH_CHECK();
    }
    return FPC_OK;
}

extern "C" int fpc_render_loss_fused(const float* pos, const int32_t* tri, const float* attr, const int32_t* attr_tri, int Va, int A,
                                     const float* tex, int Ht, int Wt, const void* ref, int ref_is_u8,
                                     int N, int V, int T, int H, int W, int C, float bg, float scale, int loss_kind,
                                     float* loss, float* grad_pos, float* grad_tex, float* rast_out, float* colour_out,
                                     const int32_t* vadj_off, const int32_t* vadj_item,
                                     void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    return render_loss_fused_impl("render_loss_fused", pos, tri, nullptr, attr, attr_tri, Va, A, tex, Ht, Wt, ref, ref_is_u8, N, V, T, H, W, C,
                                  bg, scale, loss_kind, loss, grad_pos, grad_tex, rast_out, colour_out, vadj_off, vadj_item, 0, 0, 0, scratch, scratch_bytes,
                                  (cudaStream_t)stream_);
}

extern "C" int fpc_render_loss_fused_aa(const float* pos, const int32_t* tri, const int32_t* tri_opp, const float* attr,
                                        const int32_t* attr_tri, int Va, int A, const float* tex, int Ht, int Wt,
                                        const void* ref, int ref_is_u8, int N, int V, int T, int H, int W, int C, float bg, float scale, int loss_kind,
                                        float* loss, float* grad_pos, float* grad_tex, float* rast_out, float* colour_out,
                                        const int32_t* vadj_off, const int32_t* vadj_item,
                                        void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    FPC_CHECK_ARG(tri_opp, "render_loss_fused_aa: tri_opp must be non-null (fpc_topology_build)");
    return render_loss_fused_impl("render_loss_fused_aa", pos, tri, tri_opp, attr, attr_tri, Va, A, tex, Ht, Wt, ref, ref_is_u8, N, V, T, H, W, C,
                                  bg, scale, loss_kind, loss, grad_pos, grad_tex, rast_out, colour_out, vadj_off, vadj_item, 0, 0, 0, scratch, scratch_bytes,
                                  (cudaStream_t)stream_);
}

extern "C" int fpc_render_loss_fused_band(const float* pos, const int32_t* tri, const int32_t* tri_opp, const float* attr,
                                          const int32_t* attr_tri, int Va, int A, const float* tex, int Ht, int Wt,
                                          const void* ref, int ref_is_u8, int N, int V, int T, int H, int W, int C, float bg, float scale, int loss_kind,
                                          int views_per_frame, int row_lo, int row_hi,
                                          float* loss, float* grad_pos, float* grad_tex, float* rast_out, float* colour_out,
                                          const int32_t* vadj_off, const int32_t* vadj_item,
                                          void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    FPC_CHECK_ARG(views_per_frame > 0, "render_loss_fused_band: views_per_frame must be positive");
    return render_loss_fused_impl("render_loss_fused_band", pos, tri, tri_opp, attr, attr_tri, Va, A, tex, Ht, Wt, ref, ref_is_u8, N, V, T, H, W, C,
                                  bg, scale, loss_kind, loss, grad_pos, grad_tex, rast_out, colour_out, vadj_off, vadj_item, views_per_frame, row_lo, row_hi,
                                  scratch, scratch_bytes, (cudaStream_t)stream_);
}

extern "C" size_t fpc_vertex_adjacency_scratch_bytes(int V)
{
    return V > 0 ? align256((size_t)V * sizeof(int)) * 2 : 256;
}

extern "C" int fpc_vertex_adjacency_build(const int32_t* tri, int T, int V, int32_t* vadj_off, int32_t* vadj_item,
                                          void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(tri && vadj_off && vadj_item, "vertex_adjacency_build: null pointer argument");
    FPC_CHECK_ARG(T > 0 && V > 0 && T < (1 << 24), "vertex_adjacency_build: T, V must be positive, T < 2^24 (got %d, %d)", T, V);
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_vertex_adjacency_scratch_bytes(V), "vertex_adjacency_build: scratch too small");
    int* cnt = (int*)scratch;
    int* cursor = (int*)((char*)scratch + align256((size_t)V * sizeof(int)));
    FPC_CUDA(cudaMemsetAsync(cnt, 0, (size_t)V * sizeof(int), stream));
    k_vadj_count<<<fpc_div_up(3LL * T, 256), 256, 0, stream>>>(tri, T, V, cnt);
    FPC_LAUNCH_CHECK();
    k_vadj_scan<<<1, 1024, 0, stream>>>(cnt, V, vadj_off, cursor);
    FPC_LAUNCH_CHECK();
    k_vadj_fill<<<fpc_div_up(3LL * T, 256), 256, 0, stream>>>(tri, T, V, cursor, vadj_item);
    FPC_LAUNCH_CHECK();
    k_vadj_sort<<<fpc_div_up(V, 256), 256, 0, stream>>>(vadj_off, V, vadj_item);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
