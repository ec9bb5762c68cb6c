// Fused render + ANTIALIAS + loss + gradient kernel (included by fused.cu):
//   rasterize -> interpolate -> [bilinear texture] -> antialias -> background composite -> image loss
//   -> d loss / d colour -> antialias bwd -> [texture bwd] -> interpolate bwd -> rasterize bwd -> d loss / d pos_clip
// i.e. reference fit.py:151-161,579 and their part of loss.backward() (fit.py:611) in ONE kernel per (32x32-px bin, view).
//
// Antialiasing couples a pixel to its 4-neighbourhood (forward) and to the 4-neighbourhoods of those (backward), so a
// CTA resolves visibility and colour for its bin widened by a 2-px halo (36x36 tile; triangles are binned against the
// widened bins, RasterParams::halo) and keeps everything per-pixel in shared memory:
//   (1) raster_tile<36>: (depth,id) keys;
//   (2) shade every tile pixel: colour and z/w (kept in the key slot); for the bin's own pixels also the 3C
//       coefficients of the linear map  d loss/d colour -> d loss/d (a0,a1,a2)  (barycentric numerators);
//   (3) pixel pairs (p,right) / (p,up) with different triangle ids are compacted into a work list and analysed
//       densely (aa_analyze, the op-level code, bit-identical decisions) -> per-pixel alpha of its two own pairs;
//   (4) ring-1 region (34x34): antialiased colour, loss (own pixels only) and d loss / d out;
//   (5) own pixels: d loss / d colour gathered from the 4 pairs (no atomics) -> d loss / d (a0,a1,a2) of the pixel, and
//       dd = sum_c d loss/d out_c (colour_1 - colour_0) of the pixel's two own pairs, all kept in shared memory;
//   (6) triangle-parallel gather (as k_fused): one thread per entry of the bin's list walks the entry's bounding box inside
//       the tile in a fixed order, sums the moments of its own pixels AND the silhouette terms of the pairs this bin owns
//       whose crossing edge belongs to the triangle, and stores the gradient of the three corners in the slot of
//       (view, triangle, bin): no atomics anywhere, bit-reproducible.
// Each pair is owned by its lower/left pixel, each pixel by exactly one bin: nothing is counted twice.
#pragma once

namespace {

// CTA size of the antialias kernel: 256 threads (8 warps) share one tile's ~60 KB of shared memory, 3 CTAs per SM = 24
// resident warps (128-thread CTAs: 4 x 4 = 16 warps at the same footprint per tile; measured slower, see profiles/)
#ifndef FPC_AA_THREADS
#define FPC_AA_THREADS 256
#endif
// launch order of the CTAs (raster_core.cuh: ordered_bin): by list length when the host asks for it (RasterParams::bin_order is
// set only for launches with few busy CTAs, see render_loss_fused_impl), identity otherwise
#ifndef FPC_ORDER_MODE_AA
#define FPC_ORDER_MODE_AA 1
#endif
#ifndef FPC_AA_MINBLOCKS
#define FPC_AA_MINBLOCKS 3
#endif
constexpr int AA_THREADS = FPC_AA_THREADS;
constexpr int AA_WARPS = AA_THREADS / 32;
constexpr int AA_HALO = 2;
constexpr int AA_TW = BIN + 2 * AA_HALO;       // tile edge
constexpr int AA_NT = AA_TW * AA_TW;
constexpr int AA_R1 = BIN + 2;                 // ring-1 region edge (pixels whose d loss / d out is needed)
constexpr int AA_REF_MARGIN = 4;               // reference tile starts 4 px left of the bin: 4-byte aligned for every C / dtype
constexpr int AA_REF_W = BIN + 2 * AA_REF_MARGIN;

struct AASmem {
    size_t keys, region0, col, alpha_r, alpha_u, info, coef, ref, ewin, svis, vis, texc, total;
};

__host__ __device__ inline size_t aa_align16(size_t x) { return (x + 15) & ~(size_t)15; }

// region0 is time-shared: WarpStage records (phase 1); then pair work list (phase 3) / d loss / d out (phases 4-5) in its first
// part and, behind them, the 3C coefficients per own pixel (written in phase 2, once the records are dead)
// texgrad: also keep the texture coordinates of the bin's own pixels (phase 5 scatters d loss / d tex from them)
__host__ __device__ inline AASmem aa_smem_layout(int C, int esz, bool texgrad = false)
{
    AASmem L;
    size_t o = 0;
    L.keys = o; o += sizeof(unsigned long long) * AA_NT;
    size_t r0 = sizeof(WarpStage) * AA_WARPS;
    size_t lst = sizeof(unsigned short) * 2 * AA_NT;
    size_t gc = sizeof(float) * AA_R1 * AA_R1 * C;
    const size_t head = aa_align16(lst > gc ? lst : gc);
    const size_t coef = aa_align16(sizeof(float) * BIN * BIN * 3 * C);
    if (head + coef > r0) r0 = head + coef;
    L.region0 = o; o += aa_align16(r0);
    L.coef = L.region0 + head;
    L.col = o; o += aa_align16(sizeof(float) * AA_NT * C);
    L.alpha_r = o; o += aa_align16(sizeof(float) * AA_NT);
    L.alpha_u = o; o += aa_align16(sizeof(float) * AA_NT);
    L.info = o; o += aa_align16(AA_NT);
    L.ref = o; o += aa_align16((size_t)AA_R1 * AA_REF_W * C * esz);
    L.ewin = o; o += sizeof(unsigned) * EWIN_CAP;
    L.svis = o; o += SVIS_N;
    L.vis = o; o += sizeof(unsigned short) * EWIN_CAP;
    L.texc = o; if (texgrad) o += sizeof(float2) * BIN * BIN;
    L.total = o;
    return L;
}

// texel columns / rows (wrapped, packed 16 + 16 bits) and weights of a bilinear lookup; same op order as tex_bilinear
__device__ __forceinline__ void tex_coords(const FusedParams& fp, float au, float av, unsigned& ix, unsigned& iy, float& wx, float& wy)
{
    float tu = au - floorf(au), tv = av - floorf(av);
    float x = xsub(xmul(tu, (float)fp.Wt), 0.5f), y = xsub(xmul(tv, (float)fp.Ht), 0.5f);
    float x0f = floorf(x), y0f = floorf(y);
    int ix0 = (int)x0f, iy0 = (int)y0f, ix1 = ix0 + 1, iy1 = iy0 + 1;
    wx = x - x0f; wy = y - y0f;
    if (ix0 < 0) ix0 += fp.Wt;
    if (iy0 < 0) iy0 += fp.Ht;
    if (ix1 >= fp.Wt) ix1 -= fp.Wt;
    if (iy1 >= fp.Ht) iy1 -= fp.Ht;
    ix = (unsigned)ix0 | ((unsigned)ix1 << 16);
    iy = (unsigned)iy0 | ((unsigned)iy1 << 16);
}

// bilinear, wrap (texture.cu: tex_index); same op order as k_fused
template <int C>
__device__ __forceinline__ void tex_bilinear(const FusedParams& fp, float au, float av, float (&col)[C], float (&dudc)[C], float (&dvdc)[C])
{
    float tu = au - floorf(au), tv = av - floorf(av);
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
#pragma unroll
    for (int c = 0; c < C; c++) {
        float t00 = __ldg(fp.tex + i00 * C + c), t10 = __ldg(fp.tex + i10 * C + c);
        float t01 = __ldg(fp.tex + i01 * C + c), t11 = __ldg(fp.tex + i11 * C + c);
        float a = t00 + (t10 - t00) * wx, b = t01 + (t11 - t01) * wx;
        col[c] = a + (b - a) * wy;
        dudc[c] = (float)fp.Wt * ((t10 - t00) * (1.f - wy) + (t11 - t01) * wy);
        dvdc[c] = (float)fp.Ht * ((t01 - t00) * (1.f - wx) + (t11 - t10) * wx);
    }
}

template <int C, bool TEX>
__global__ void __launch_bounds__(AA_THREADS, FPC_AA_MINBLOCKS) k_fused_aa(RasterParams rp, FusedParams fp, const int32_t* __restrict__ tri_opp)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int esz = fp.ref_u8 ? 1 : 4;
    const bool texgrad = TEX && fp.grad_tex != nullptr;
    const AASmem L = aa_smem_layout(C, esz, texgrad);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + L.keys);
    WarpStage* stage = reinterpret_cast<WarpStage*>(smem + L.region0);
    unsigned short* s_list = reinterpret_cast<unsigned short*>(smem + L.region0);
    float* s_gc = reinterpret_cast<float*>(smem + L.region0);
    float* s_col = reinterpret_cast<float*>(smem + L.col);
    float* s_ar = reinterpret_cast<float*>(smem + L.alpha_r);
    float* s_au = reinterpret_cast<float*>(smem + L.alpha_u);
    unsigned char* s_info = smem + L.info;
    float* s_coef = reinterpret_cast<float*>(smem + L.coef);
    unsigned char* s_ref = smem + L.ref;
    float2* s_texc = reinterpret_cast<float2*>(smem + L.texc);
    unsigned* ewin = reinterpret_cast<unsigned*>(smem + L.ewin);          // gather phase: per-entry pixel windows (raster set-up)
    unsigned char* svis = smem + L.svis;                                  // gather phase: "this id won a pixel" filter
    __shared__ double red[AA_WARPS];
    __shared__ int s_nlist;
    __shared__ unsigned char s_rowpair[AA_TW + 4];     // tile row -> some pixel of the row is the triangle side of an accepted pair with
                                                       // a non-zero gradient (the gather phase looks for pair terms only in such rows)

    int bin, n;
    ordered_bin<FPC_ORDER_MODE_AA>(rp, n, bin);        // long triangle lists first (k_fill)
    const int ox = (bin % rp.BW) * BIN, oy = (bin / rp.BW) * BIN;
    const int tx0 = ox - AA_HALO, ty0 = oy - AA_HALO;           // tile origin (may be negative)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ref_pitch = AA_REF_W * C * esz;                   // bytes per reference tile row (multiple of 4)

    if (outside_band(fp, n, bin / rp.BW)) {            // another rank renders this bin row (shard.view_band_shard)
        if (threadIdx.x == 0) fp.loss_partial[(size_t)n * rp.NB + bin] = 0.0;
        return;
    }
    // bins are widened by the halo when triangles are binned: an empty list means an all-background tile
    if (rp.bin_count[(size_t)n * rp.NB + bin] == 0 && rp.large_count[n] == 0) {
        background_bin<C, AA_THREADS>(rp, fp, n, bin, ox, oy, red);
        return;
    }

    // ---- (0) reference tile (rows oy-1 .. oy+32, columns ox-4 .. ox+35) on its way into shared memory ----
    {
        const unsigned char* rbase = reinterpret_cast<const unsigned char*>(fp.ref);
        const long long row_bytes = (long long)rp.W * C * esz;
        const bool fast = (row_bytes % 4 == 0) && ((reinterpret_cast<size_t>(rbase) & 3) == 0);
        if (fast) {
            const int gpr = ref_pitch >> 2;                     // 4-byte granules per tile row
            const long long x_off = (long long)(ox - AA_REF_MARGIN) * C * esz;
            for (int i = threadIdx.x; i < AA_R1 * gpr; i += AA_THREADS) {
                int r = i / gpr, g = i - r * gpr;
                int py = oy - 1 + r;
                long long b0 = x_off + 4 * g;
                if (py < 0 || py >= rp.H || b0 < 0 || b0 + 4 > row_bytes) continue;
                const unsigned char* src = rbase + ((size_t)n * rp.H + py) * (size_t)row_bytes + b0;
                unsigned dst = (unsigned)__cvta_generic_to_shared(s_ref + r * ref_pitch + 4 * g);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            }
        } else {
            for (int i = threadIdx.x; i < AA_R1 * AA_REF_W * C; i += AA_THREADS) {
                int r = i / (AA_REF_W * C), e = i - r * (AA_REF_W * C);
                int py = oy - 1 + r, px = ox - AA_REF_MARGIN + e / C;
                if (py < 0 || py >= rp.H || px < 0 || px >= rp.W) continue;
                size_t gi = (((size_t)n * rp.H + py) * rp.W + px) * C + (e % C);
                if (fp.ref_u8) s_ref[r * ref_pitch + e] = __ldg(rbase + gi);
                else reinterpret_cast<float*>(s_ref + r * ref_pitch)[e] = __ldg(reinterpret_cast<const float*>(rbase) + gi);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- (1) visibility of the widened tile ----
    for (int i = threadIdx.x; i < SVIS_N / 4; i += AA_THREADS) reinterpret_cast<unsigned*>(svis)[i] = 0u;
    raster_tile<AA_TW, AA_THREADS>(rp, n, bin, tx0, ty0, keys, stage, true, fp.slots ? ewin : nullptr);

    // colour a background pixel hands to the antialias op: texture at uv = (0,0) or 0 (interpolate)
    float bgcol[C];
    if (TEX) {
        float du[C], dv[C];
        tex_bilinear<C>(fp, 0.f, 0.f, bgcol, du, dv);
    } else {
#pragma unroll
        for (int c = 0; c < C; c++) bgcol[c] = 0.f;
    }

    // ---- (2) shade every tile pixel; the key slot becomes (z/w bits << 32 | id + 1) ----
    const float* P = rp.pos + (size_t)n * rp.V * 4;
    for (int idx = threadIdx.x; idx < AA_NT; idx += AA_THREADS) {
        const int tx = idx % AA_TW, ty = idx / AA_TW;
        const int px = tx0 + tx, py = ty0 + ty;
        const unsigned long long key = keys[idx];
        const bool inner = tx >= AA_HALO && tx < AA_HALO + BIN && ty >= AA_HALO && ty < AA_HALO + BIN;
        float col[C];
#pragma unroll
        for (int c = 0; c < C; c++) col[c] = bgcol[c];
        unsigned long long slot = 0ull;
        float K[3 * C];
#pragma unroll
        for (int c = 0; c < 3 * C; c++) K[c] = 0.f;
        float4 rout = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 tcoord = make_float2(0.f, 0.f);      // background pixels sample the texture at uv = (0,0)
        if (key != KEY_EMPTY) {       // only in-image pixels receive fragments
            int t = (int)(key & 0xFFFFFFFFu);
            const int4 ti = tri_indices(rp, t);
            const int i0 = ti.x, i1 = ti.y, i2 = ti.z;
            float4 p0 = ldg4(P + 4 * (size_t)i0), p1 = ldg4(P + 4 * (size_t)i1), p2 = ldg4(P + 4 * (size_t)i2);
            float fx = pixel_ndc(px, rp.xs, rp.xo), fy = pixel_ndc(py, rp.ys, rp.yo);
            Shade sh = shade_pixel(p0, p1, p2, fx, fy);
            float u = clamp01(sh.u), v = clamp01(sh.v);
            float zw = fminf(fmaxf(sh.zw, -1.f), 1.f);
            rout = make_float4(u, v, zw, (float)(t + 1));
            slot = ((unsigned long long)__float_as_uint(zw) << 32) | (unsigned)(t + 1);
            if (inner) svis[t & (SVIS_N - 1)] = 1;            // triangles seen only in the halo are added by their pairs (phase 5)
            int j0 = i0, j1 = i1, j2 = i2;
            if (fp.attr_tri4) { const int4 tj = __ldg(fp.attr_tri4 + t); j0 = tj.x; j1 = tj.y; j2 = tj.z; }
            bool ok = (unsigned)j0 < (unsigned)fp.Va && (unsigned)j1 < (unsigned)fp.Va && (unsigned)j2 < (unsigned)fp.Va;
            constexpr int AA = TEX ? 2 : C;
            float b2 = 1.f - u - v;
            float a0c[AA], a1c[AA], a2c[AA], at[AA];
            if (TEX) {               // uv pairs: one 8-byte load per corner
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
            float ku[C], kv[C];        // d colour_c / d u, d colour_c / d v
            if (TEX) {
                float dudc[C], dvdc[C];
                tex_bilinear<C>(fp, at[0], at[1], col, dudc, dvdc);
                tcoord = make_float2(at[0], at[AA - 1]);
#pragma unroll
                for (int c = 0; c < C; c++) {
                    ku[c] = dudc[c] * (a0c[0] - a2c[0]) + dvdc[c] * (a0c[1] - a2c[1]);
                    kv[c] = dudc[c] * (a1c[0] - a2c[0]) + dvdc[c] * (a1c[1] - a2c[1]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    col[c] = at[c];
                    ku[c] = a0c[c] - a2c[c];
                    kv[c] = a1c[c] - a2c[c];
                }
            }
            // d loss / d a_k = sum_c g_c K[k][c]   (u = a0/at, v = a1/at, unclamped barycentrics as in the op-level backward)
            if (inner && fp.slots) {
                const ShadeGrad sg = shade_pixel_grad(p0, p1, p2, fx, fy);      // backward-only barycentrics (common.cuh)
#pragma unroll
                for (int c = 0; c < C; c++) {
                    float k2 = -sg.iw * (ku[c] * sg.u + kv[c] * sg.v);
                    K[2 * C + c] = k2;
                    K[0 * C + c] = sg.iw * ku[c] + k2;
                    K[1 * C + c] = sg.iw * kv[c] + k2;
                }
            }
        }
        keys[idx] = slot;
#pragma unroll
        for (int c = 0; c < C; c++) s_col[idx * C + c] = col[c];
        s_ar[idx] = 0.f; s_au[idx] = 0.f; s_info[idx] = 0;
        if (inner) {
            const int ii = (ty - AA_HALO) * BIN + (tx - AA_HALO);
#pragma unroll
            for (int c = 0; c < 3 * C; c++) s_coef[ii * 3 * C + c] = K[c];
            if (texgrad) s_texc[ii] = tcoord;
            if (fp.rast_out && px < rp.W && py < rp.H) reinterpret_cast<float4*>(fp.rast_out)[((size_t)n * rp.H + py) * rp.W + px] = rout;
        }
    }
    if (threadIdx.x == 0) s_nlist = 0;
    if (threadIdx.x < AA_TW + 4) s_rowpair[threadIdx.x] = 0;
    __syncthreads();

    // ---- (3a) work list of pixel pairs with different triangle ids (both pixels inside the tile and the image) ----
    for (int base = 0; base < AA_NT; base += AA_THREADS) {
        const int idx = base + threadIdx.x;
        bool cr = false, cu = false;
        if (idx < AA_NT) {
            const int tx = idx % AA_TW, ty = idx / AA_TW;
            const int px = tx0 + tx, py = ty0 + ty;
            if (px >= 0 && py >= 0 && px < rp.W && py < rp.H) {
                unsigned id = (unsigned)keys[idx];
                cr = (tx + 1 < AA_TW) && (px + 1 < rp.W) && ((unsigned)keys[idx + 1] != id);
                cu = (ty + 1 < AA_TW) && (py + 1 < rp.H) && ((unsigned)keys[idx + AA_TW] != id);
            }
        }
        const unsigned mr = __ballot_sync(0xffffffffu, cr), mu = __ballot_sync(0xffffffffu, cu);
        const int cnt = __popc(mr) + __popc(mu);
        int wbase = 0;
        if (lane == 0 && cnt) wbase = atomicAdd(&s_nlist, cnt);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        const unsigned below = (1u << lane) - 1u;
        if (cr) s_list[wbase + __popc(mr & below)] = (unsigned short)idx;
        if (cu) s_list[wbase + __popc(mr) + __popc(mu & below)] = (unsigned short)(idx | 0x8000);
    }
    __syncthreads();

    // ---- (3b) analyse the pairs densely ----
    AAParams ap;
    ap.rast = nullptr; ap.pos = rp.pos; ap.tri = rp.tri; ap.tri_opp = tri_opp;
    ap.N = rp.N; ap.V = rp.V; ap.T = rp.T; ap.H = rp.H; ap.W = rp.W; ap.C = C;
    ap.xh = 0.5f * (float)rp.W; ap.yh = 0.5f * (float)rp.H;
    const int nlist = s_nlist;
    for (int i = threadIdx.x; i < nlist; i += AA_THREADS) {
        const unsigned e = s_list[i];
        const int idx = e & 0x7fff, d = e >> 15;
        const int tx = idx % AA_TW, ty = idx / AA_TW;
        const int px = tx0 + tx, py = ty0 + ty;
        const unsigned long long k0 = keys[idx], k1 = keys[idx + (d ? AA_TW : 1)];
        float2 z0 = make_float2(__uint_as_float((unsigned)(k0 >> 32)), (float)(unsigned)k0);
        float2 z1 = make_float2(__uint_as_float((unsigned)(k1 >> 32)), (float)(unsigned)k1);
        AAPair a = aa_analyze(ap, n, px, py, d, z0, z1);
        if (a.valid) {
            const int front1 = (a.px != px || a.py != py) ? 1 : 0;
            const unsigned bits = 8u | (unsigned)a.di | ((unsigned)front1 << 2);
            if (d) { s_au[idx] = a.alpha; atomicOr(reinterpret_cast<unsigned*>(s_info + (idx & ~3)), (bits << 4) << (8 * (idx & 3))); }
            else { s_ar[idx] = a.alpha; atomicOr(reinterpret_cast<unsigned*>(s_info + (idx & ~3)), bits << (8 * (idx & 3))); }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    // ---- (4) ring-1 region: antialiased colour, loss (own pixels), d loss / d out ----
    double loss_acc = 0.0;
    for (int r = threadIdx.x; r < AA_R1 * AA_R1; r += AA_THREADS) {
        const int rx = r % AA_R1, ry = r / AA_R1;
        const int tx = rx + 1, ty = ry + 1, idx = ty * AA_TW + tx;
        const int px = tx0 + tx, py = ty0 + ty;
        const bool in_img = px >= 0 && py >= 0 && px < rp.W && py < rp.H;
        const bool inner = rx >= 1 && rx <= BIN && ry >= 1 && ry <= BIN;
        const bool fg = in_img && ((unsigned)keys[idx] != 0u);
        float gcv[C];
#pragma unroll
        for (int c = 0; c < C; c++) gcv[c] = 0.f;
        if (in_img) {
            const float a0 = s_ar[idx], a1 = s_ar[idx - 1], a2 = s_au[idx], a3 = s_au[idx - AA_TW];
#pragma unroll
            for (int c = 0; c < C; c++) {
                float comp = fp.bg;
                if (fg) {
                    const float cc = s_col[idx * C + c];
                    float o = cc;
                    if (a0 > 0.f) o += a0 * (s_col[(idx + 1) * C + c] - cc);
                    if (a1 < 0.f) o += a1 * (cc - s_col[(idx - 1) * C + c]);
                    if (a2 > 0.f) o += a2 * (s_col[(idx + AA_TW) * C + c] - cc);
                    if (a3 < 0.f) o += a3 * (cc - s_col[(idx - AA_TW) * C + c]);
                    comp = o;
                }
                const int rb = ry * ref_pitch + ((rx + AA_REF_MARGIN - 1) * C + c) * esz;
                const float refv = fp.ref_u8 ? (float)s_ref[rb] : *reinterpret_cast<const float*>(s_ref + rb);
                const float e = refv - 255.f * comp;
                if (inner) {
                    loss_acc += (double)loss_term(e, fp.l1);
                    if (fp.colour_out) fp.colour_out[(((size_t)n * rp.H + py) * rp.W + px) * C + c] = comp;
                }
                if (fg) gcv[c] = loss_dcolour(e, fp.k, fp.l1);
            }
        }
        // region0 still holds the work list for other threads of phase 3b? no: a barrier separates the phases
#pragma unroll
        for (int c = 0; c < C; c++) s_gc[r * C + c] = gcv[c];
    }
    __syncthreads();

    // ---- (5) own pixels: d loss / d colour (gather over the 4 pairs) -> d loss / d (a0, a1, a2); dd of the two own pairs ----
    static_assert(PIX % AA_THREADS == 0, "every thread owns PIX / AA_THREADS pixels");
    constexpr int PER = PIX / AA_THREADS;
    const bool any_large = rp.large_count[n] != 0;
    float ddr_[PER], ddu_[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int ii = threadIdx.x + j * AA_THREADS;
        const int ix = ii & (BIN - 1), iy = ii >> BIN_LOG2;
        const int tx = ix + AA_HALO, ty = iy + AA_HALO, idx = ty * AA_TW + tx;
        const int r = (iy + 1) * AA_R1 + (ix + 1);
        const int px = ox + ix, py = oy + iy;
        const bool in_img = px < rp.W && py < rp.H;
        const unsigned idp1 = (unsigned)keys[idx];
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        float ddr = 0.f, ddu = 0.f;
        if (in_img) {
            const float a0 = s_ar[idx], a1 = s_ar[idx - 1], a2 = s_au[idx], a3 = s_au[idx - AA_TW];
            // destination pixel (ring-1 index) of every pair: pix0 when alpha > 0, else pix1
            const int d0 = (a0 > 0.f) ? r : r + 1, d1 = (a1 > 0.f) ? r - 1 : r;
            const int d2 = (a2 > 0.f) ? r : r + AA_R1, d3 = (a3 > 0.f) ? r - AA_R1 : r;
            const unsigned info = s_info[idx];
            float gpre[C];             // d loss / d colour before antialias
#pragma unroll
            for (int c = 0; c < C; c++) {
                float g = s_gc[r * C + c];
                if (a0 != 0.f) g -= a0 * s_gc[d0 * C + c];
                if (a1 != 0.f) g += a1 * s_gc[d1 * C + c];
                if (a2 != 0.f) g -= a2 * s_gc[d2 * C + c];
                if (a3 != 0.f) g += a3 * s_gc[d3 * C + c];
                gpre[c] = g;
                if (idp1) {
                    g0 += g * s_coef[ii * 3 * C + 0 * C + c];
                    g1 += g * s_coef[ii * 3 * C + 1 * C + c];
                    g2 += g * s_coef[ii * 3 * C + 2 * C + c];
                }
                const float cc = s_col[idx * C + c];
                if (info & 8u) ddr += s_gc[d0 * C + c] * (s_col[(idx + 1) * C + c] - cc);
                if (info & 0x80u) ddu += s_gc[d2 * C + c] * (s_col[(idx + AA_TW) * C + c] - cc);
            }
            if (texgrad) {
                // background pixels too: their colour (texture at uv = 0) can blend into a neighbour
                const float2 tc = s_texc[ii];
                unsigned tix, tiy;
                float twx, twy;
                tex_coords(fp, tc.x, tc.y, tix, tiy, twx, twy);
                tex_grad_scatter<C>(fp, tix, tiy, twx, twy, gpre);
            }
            if (fp.slots && any_large) {
                // large / near-clipped triangles are in no bin list: their terms are accumulated with float REDs (slots 1, 2)
                if (idp1 && (g0 != 0.f || g1 != 0.f || g2 != 0.f)) {
                    const size_t gid = (size_t)n * rp.T + (idp1 - 1u);
                    if ((rp.tri_info[gid] >> 22) == 2) large_pixel_moments(fp.slots, gid, g0, g1, g2, px, py, rp.tri_anchor[gid]);
                }
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    const unsigned inf = d ? (info >> 4) : info;
                    const float dd = d ? ddu : ddr;
                    if (!(inf & 8u) || dd == 0.f) continue;
                    const int f1 = (inf >> 2) & 1;
                    const int tt = (int)(unsigned)keys[idx + f1 * (d ? AA_TW : 1)] - 1;
                    const size_t gid = (size_t)n * rp.T + tt;
                    if (tt < 0 || (rp.tri_info[gid] >> 22) != 2) continue;
                    const int di = inf & 3;
                    const int4 ti = tri_indices(rp, tt);
                    const int vi[3] = {ti.x, ti.y, ti.z};
                    const int c1 = (di + 1) % 3, c2 = (di + 2) % 3;
                    float gp1[3], gp2[3];
                    aa_pair_corner_grads(ap.xh, ap.yh, ldg4(P + 4 * (size_t)vi[c1]), ldg4(P + 4 * (size_t)vi[c2]), px + (d ? 0 : f1), py + (d ? f1 : 0), d, dd, gp1, gp2);
                    float* Aa = slot_ptr(fp.slots, gid, 2);
#pragma unroll
                    for (int c = 0; c < 3; c++) { atomicAdd(Aa + 4 * c1 + c, gp1[c]); atomicAdd(Aa + 4 * c2 + c, gp2[c]); }
                }
            }
        }
        // the pixel's own coefficient slots become its gradient terms (read only by this thread above)
        if (fp.slots) { s_coef[ii * 3 * C + 0] = g0; s_coef[ii * 3 * C + 1] = g1; s_coef[ii * 3 * C + 2] = g2; }
        // tile rows that hold the triangle side of this pixel's pairs: its own row, and the row above for an up pair whose
        // crossing edge belongs to the upper pixel's triangle (racing stores of the same value)
        if (ddr != 0.f) s_rowpair[ty] = 1;
        if (ddu != 0.f) { s_rowpair[ty] = 1; s_rowpair[ty + 1] = 1; }
        // ... and the triangle on the far side of such a pair may have been seen in the halo only: let the gather visit it
        if (fp.slots && (ddr != 0.f || ddu != 0.f)) {
            const unsigned ir = (unsigned)keys[idx + 1], iu = (unsigned)keys[idx + AA_TW];
            if (ddr != 0.f && ir) svis[(ir - 1u) & (SVIS_N - 1)] = 1;
            if (ddu != 0.f && iu) svis[(iu - 1u) & (SVIS_N - 1)] = 1;
        }
        ddr_[j] = ddr; ddu_[j] = ddu;
    }
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) red[warp] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < AA_WARPS; w++) s += red[w];
        fp.loss_partial[(size_t)n * rp.NB + bin] = s;
    }
    if (!fp.slots) return;
    // the alphas have been consumed (barrier above): the alpha planes of the own pixels now carry the pairs' dd
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int ii = threadIdx.x + j * AA_THREADS;
        const int idx = ((ii >> BIN_LOG2) + AA_HALO) * AA_TW + (ii & (BIN - 1)) + AA_HALO;
        s_ar[idx] = ddr_[j]; s_au[idx] = ddu_[j];
    }
    __syncthreads();

    // ---- (6) gather per list entry: moments of the triangle's own pixels + silhouette terms of the pairs this bin owns ----
    const int count = rp.bin_count[(size_t)n * rp.NB + bin];
    const int* list = rp.pairs + (size_t)n * 4 * rp.T + rp.bin_offset[(size_t)n * rp.NB + bin];
    const int bx = bin % rp.BW, by = bin / rp.BW;
    const int* idw = reinterpret_cast<const int*>(keys);            // low word of the 64-bit slot = id + 1
    unsigned short* vis_list = reinterpret_cast<unsigned short*>(smem + L.vis);
    const int ntab = min(count, EWIN_CAP);
    const int nvis = sort_visible_entries<AA_THREADS>(ewin, svis, list, ntab, vis_list);
    // visible entries of the window table (bucketed by window area), then whatever the table could not hold.
    // L = 1, 2 or 4 threads per entry (as many as the CTA has to spare): thread j of an entry's group walks rows j, j + L, ...,
    // the partial sums are combined by a fixed butterfly — fewer, shorter walks at the tail of the CTA
    const int ntot = nvis + max(count - EWIN_CAP, 0);
    const int LG = (ntot * 4 <= AA_THREADS) ? 4 : ((ntot * 2 <= AA_THREADS) ? 2 : 1);
    const int sub = threadIdx.x & (LG - 1);
    for (int base = 0; base < ntot; base += AA_THREADS / LG) {
        const int e = base + threadIdx.x / LG;
        const bool act = e < ntot;
        int t = -2, xa = 1, xb = 0, ya = 1, yb = 0, kslot = 0;
        size_t gid = 0;
        if (act) {
            const int i = e < nvis ? (int)vis_list[e] : EWIN_CAP + (e - nvis);
            t = list[i];
            gid = (size_t)n * rp.T + t;
            if (i < EWIN_CAP) {
                const unsigned win = ewin[i];
                kslot = (win >> 24) & 3;
                xa = tx0 + (win & 63); xb = tx0 + ((win >> 6) & 63); ya = ty0 + ((win >> 12) & 63); yb = ty0 + ((win >> 18) & 63);
            } else {                  // more entries than the window table holds: the same from the binning records
                const ushort4 bb = rp.tri_bbox[gid];
                kslot = slot_index_k(rp.tri_info[gid], bx, by);
                xa = max((int)bb.x, max(tx0, 0)); xb = min((int)bb.z, tx0 + AA_TW - 1);
                ya = max((int)bb.y, max(ty0, 0)); yb = min((int)bb.w, ty0 + AA_TW - 1);
                if (xa > xb) { ya = 1; yb = 0; }
            }
        }
        // requested now, used by the pair terms and after the walk: the triangle's anchor and vertices
        int an = 0;
        float4 pc[3];
        pc[0] = pc[1] = pc[2] = make_float4(0.f, 0.f, 0.f, 1.f);
        if (act && ya <= yb) {
            an = rp.tri_anchor[gid];
            const int4 ti = tri_indices(rp, t);
            pc[0] = ldg4(P + 4 * (size_t)ti.x); pc[1] = ldg4(P + 4 * (size_t)ti.y); pc[2] = ldg4(P + 4 * (size_t)ti.z);
        }
        const int anx = an & 0xffff, any = (int)((unsigned)an >> 16);
        float m[9], cg[9];
#pragma unroll
        for (int c = 0; c < 9; c++) { m[c] = 0.f; cg[c] = 0.f; }
        int seen = 0;                               // the triangle has an own pixel or a pair term in this tile
        auto pair_term = [&](int apx, int apy, int d, unsigned inf, float dd) {
            if (dd == 0.f) return;
            seen = 1;
            const int di = inf & 3, c1 = (di + 1) % 3, c2 = (di + 2) % 3;
            float gp1[3], gp2[3];
            // (dynamic corner indices would put pc / cg in local memory: select with compares instead)
            const float4 q1 = c1 == 0 ? pc[0] : (c1 == 1 ? pc[1] : pc[2]);
            const float4 q2 = c2 == 0 ? pc[0] : (c2 == 1 ? pc[1] : pc[2]);
            aa_pair_corner_grads(ap.xh, ap.yh, q1, q2, apx, apy, d, dd, gp1, gp2);
#pragma unroll
            for (int cnr = 0; cnr < 3; cnr++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (cnr == c1) cg[3 * cnr + c] += gp1[c];
                    if (cnr == c2) cg[3 * cnr + c] += gp2[c];
                }
        };
        for (int y = ya + sub; y <= yb; y += LG) {
            // which pixels of the row did the triangle win?  (branch-free pass, then only the hits are visited, left to right)
            const int rowi = (y - ty0) * AA_TW - tx0;
            unsigned long long hit = 0ull;
            for (int x = xa; x <= xb; x++) hit |= (unsigned long long)(idw[2 * (rowi + x)] == t + 1) << (x - xa);
            if (!hit) continue;
            const bool own_y = y >= oy && y < oy + BIN;
            const float fly = (float)(y - any);
            const bool row_has_pairs = s_rowpair[y - ty0] != 0;
            while (hit) {
                const int x = xa + __ffsll((long long)hit) - 1;
                hit &= hit - 1ull;
                const int idx = rowi + x;
                const bool own_x = x >= ox && x < ox + BIN;
                if (own_x && own_y) {
                    seen = 1;
                    const int ii = (y - oy) * BIN + (x - ox);
                    const float a = s_coef[ii * 3 * C + 0], b = s_coef[ii * 3 * C + 1], c = s_coef[ii * 3 * C + 2];
                    const float flx = (float)(x - anx);
                    m[0] += a; m[1] += b; m[2] += c;
                    m[3] += a * flx; m[4] += b * flx; m[5] += c * flx;
                    m[6] += a * fly; m[7] += b * fly; m[8] += c * fly;
                    // this pixel's own pairs whose crossing edge belongs to its triangle (front1 == 0)
                    if (row_has_pairs) {
                        const unsigned inf = s_info[idx];
                        if ((inf & 8u) && !((inf >> 2) & 1u)) pair_term(x, y, 0, inf, s_ar[idx]);
                        if ((inf & 0x80u) && !((inf >> 6) & 1u)) pair_term(x, y, 1, inf >> 4, s_au[idx]);
                    }
                }
                if (!row_has_pairs) continue;
                // pairs owned by the left / lower neighbour (a pixel of this bin) whose crossing edge belongs to THIS pixel's triangle
                if (own_y && x - 1 >= ox && x - 1 < ox + BIN) {
                    const unsigned inf = s_info[idx - 1];
                    if ((inf & 8u) && ((inf >> 2) & 1u)) pair_term(x, y, 0, inf, s_ar[idx - 1]);
                }
                if (own_x && y - 1 >= oy && y - 1 < oy + BIN) {
                    const unsigned inf = s_info[idx - AA_TW];
                    if ((inf & 0x80u) && ((inf >> 6) & 1u)) pair_term(x, y, 1, inf >> 4, s_au[idx - AA_TW]);
                }
            }
        }
        for (int o = LG >> 1; o > 0; o >>= 1) {                         // all 32 lanes get here (no early exits above)
#pragma unroll
            for (int c = 0; c < 9; c++) { m[c] += __shfl_xor_sync(0xffffffffu, m[c], o); cg[c] += __shfl_xor_sync(0xffffffffu, cg[c], o); }
            seen |= __shfl_xor_sync(0xffffffffu, seen, o);
        }
        if (act && sub == 0 && seen) {
            float out[9];
            triangle_corner_grads(m, pixel_ndc(anx, rp.xs, rp.xo), pixel_ndc(any, rp.ys, rp.yo), rp.xs, rp.ys, pc[0], pc[1], pc[2], out);
#pragma unroll
            for (int c = 0; c < 9; c++) out[c] += cg[c];
            store_slot(fp.slots, rp.slot_valid, gid, kslot, out);
        }
    }
}

}  // namespace
