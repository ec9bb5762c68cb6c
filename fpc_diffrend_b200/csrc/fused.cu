// Fused render + loss + gradient kernel of the fit iteration (no antialias):
//   rasterize -> interpolate -> [bilinear texture] -> background composite -> image loss
//   -> d loss / d colour -> [texture bwd] -> interpolate bwd -> rasterize bwd -> d loss / d pos_clip
// i.e. reference fit.py:151-158,161,579 and their part of loss.backward() (fit.py:611) in ONE kernel per
// (64x64-px bin, view).  Nothing per-pixel is written to HBM unless the caller asks for the images: the only
// compulsory HBM traffic is reading the reference frame (4C or C bytes / px) and the geometry.
//
// Per CTA:  (1) raster_bin(): visibility keys in shared memory (raster_core.cuh);
//           (2) pixel-parallel shading, loss and (d u, d v) per pixel -> shared memory;
//           (3) triangle-parallel gradient pass over the same balanced (triangle,row) work items: every lane sums
//               the position gradient of its row in registers, rows of one triangle are combined with a segmented
//               warp-shuffle reduction, and each visible (triangle, bin) pair issues 9 float REDs to grad_pos.
#include "raster_core.cuh"

using namespace fpc;

namespace {

struct FusedParams {
    const float* attr;       // [Va, A]  vertex colours (A == C) or uv (A == 2, textured)
    const int32_t* attr_tri; // [T,3]
    int Va, A;
    const float* tex;        // [Ht,Wt,C] or null
    int Ht, Wt;
    const void* ref;         // [N,H,W,C] f32 or u8
    int ref_u8;
    int C;
    float bg, k;             // k = scale / (H*W*C)
    float* grad_pos;         // [N,V,4]
    float* rast_out;         // [N,H,W,4] or null
    float* colour_out;       // [N,H,W,C] or null (composited image)
    double* loss_partial;    // [N*NB]
};

// second use of the per-warp staging memory (gradient pass)
struct GradStage {
    float px[3][32], py[3][32], pw[3][32];
    int xy[32], wn[32], tri[32], vidx[3][32], prefix[32];
};
static_assert(sizeof(GradStage) <= sizeof(WarpStage), "GradStage must fit in the WarpStage memory");

struct PosGrad { float v[9]; };   // (x,y,w) of vertex 0, 1, 2

// d pos of one pixel from (gu, gv) = d loss / d (u, v)   (same formula as k_raster_bwd / gold_rasterize_bwd)
__device__ __forceinline__ void pixel_pos_grad(float gu, float gv, float fx, float fy, float q0x, float q0y, float q0w,
                                               float q1x, float q1y, float q1w, float q2x, float q2y, float q2w, PosGrad& a)
{
    float p0x = q0x - fx * q0w, p0y = q0y - fy * q0w;
    float p1x = q1x - fx * q1w, p1y = q1y - fy * q1w;
    float p2x = q2x - fx * q2w, p2y = q2y - fy * q2w;
    float a0 = p1x * p2y - p1y * p2x, a1 = p2x * p0y - p2y * p0x, a2 = p0x * p1y - p0y * p1x;
    float iw = 1.f / (a0 + a1 + a2);
    float u = a0 * iw, v = a1 * iw;
    float gbb = gu * u + gv * v;
    float g0 = iw * (gu - gbb), g1 = iw * (gv - gbb), g2 = -iw * gbb;
    float g0x = -g1 * p2y + g2 * p1y, g0y = g1 * p2x - g2 * p1x;
    float g1x = g0 * p2y - g2 * p0y, g1y = -g0 * p2x + g2 * p0x;
    float g2x = -g0 * p1y + g1 * p0y, g2y = g0 * p1x - g1 * p0x;
    a.v[0] += g0x; a.v[1] += g0y; a.v[2] += -fx * g0x - fy * g0y;
    a.v[3] += g1x; a.v[4] += g1y; a.v[5] += -fx * g1x - fy * g1y;
    a.v[6] += g2x; a.v[7] += g2y; a.v[8] += -fx * g2x - fy * g2y;
}

__device__ __forceinline__ void red_vertex(float* G, int vi, float gx, float gy, float gw)
{
    if (gx != 0.f) atomicAdd(G + 4 * (size_t)vi + 0, gx);
    if (gy != 0.f) atomicAdd(G + 4 * (size_t)vi + 1, gy);
    if (gw != 0.f) atomicAdd(G + 4 * (size_t)vi + 3, gw);
}

template <int C, bool TEX>
__global__ void __launch_bounds__(FINE_THREADS) k_fused(RasterParams rp, FusedParams fp)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    WarpStage* stage = reinterpret_cast<WarpStage*>(smem + sizeof(unsigned long long) * BIN * BIN);
    float2* guv = reinterpret_cast<float2*>(smem + sizeof(unsigned long long) * BIN * BIN + sizeof(WarpStage) * FINE_WARPS);
    float* acc_all = reinterpret_cast<float*>(guv + BIN * BIN);          // [FINE_WARPS][9][32]
    __shared__ double red[FINE_WARPS];

    const int bin = blockIdx.x, n = blockIdx.y;
    const int ox = (bin % rp.BW) * BIN, oy = (bin / rp.BW) * BIN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    raster_bin(rp, n, bin, keys, stage);

    // ---- (2) shade + loss + (d u, d v) per pixel ----
    const float* P = rp.pos + (size_t)n * rp.V * 4;
    double loss_acc = 0.0;
    for (int idx = threadIdx.x; idx < BIN * BIN; idx += FINE_THREADS) {
        int lx = idx & (BIN - 1), ly = idx >> BIN_LOG2;
        int px = ox + lx, py = oy + ly;
        float2 g = make_float2(0.f, 0.f);
        if (px < rp.W && py < rp.H) {
            size_t pi = ((size_t)n * rp.H + py) * rp.W + px;
            unsigned long long key = keys[idx];
            float col[C];
            float4 rout = make_float4(0.f, 0.f, 0.f, 0.f);
            bool fg = key != KEY_EMPTY;
            float a0c[TEX ? 2 : C], a1c[TEX ? 2 : C], a2c[TEX ? 2 : C];
            float dudc[C], dvdc[C];      // TEX: d colour_c / d texU, d texV
            if (fg) {
                int t = (int)(key & 0xFFFFFFFFu);
                int i0 = __ldg(rp.tri + 3 * t), i1 = __ldg(rp.tri + 3 * t + 1), i2 = __ldg(rp.tri + 3 * t + 2);
                float4 p0 = ldg4(P + 4 * (size_t)i0), p1 = ldg4(P + 4 * (size_t)i1), p2 = ldg4(P + 4 * (size_t)i2);
                float fx = pixel_ndc(px, rp.xs, rp.xo), fy = pixel_ndc(py, rp.ys, rp.yo);
                Shade sh = shade_pixel(p0, p1, p2, fx, fy);
                float u = clamp01(sh.u), v = clamp01(sh.v);
                rout = make_float4(u, v, fminf(fmaxf(sh.zw, -1.f), 1.f), (float)(t + 1));
                int j0 = __ldg(fp.attr_tri + 3 * t), j1 = __ldg(fp.attr_tri + 3 * t + 1), j2 = __ldg(fp.attr_tri + 3 * t + 2);
                bool ok = (unsigned)j0 < (unsigned)fp.Va && (unsigned)j1 < (unsigned)fp.Va && (unsigned)j2 < (unsigned)fp.Va;
                constexpr int AA = TEX ? 2 : C;
                float b2 = 1.f - u - v;
                float at[AA];
#pragma unroll
                for (int c = 0; c < AA; c++) {
                    a0c[c] = ok ? __ldg(fp.attr + (size_t)j0 * AA + c) : 0.f;
                    a1c[c] = ok ? __ldg(fp.attr + (size_t)j1 * AA + c) : 0.f;
                    a2c[c] = ok ? __ldg(fp.attr + (size_t)j2 * AA + c) : 0.f;
                    at[c] = u * a0c[c] + v * a1c[c] + b2 * a2c[c];
                }
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
                float r = fp.ref_u8 ? (float)__ldg(reinterpret_cast<const unsigned char*>(fp.ref) + pi * C + c)
                                    : __ldg(reinterpret_cast<const float*>(fp.ref) + pi * C + c);
                float e = r - 255.f * col[c];
                loss_acc += (double)(e * e);
                gc[c] = (-510.f * fp.k) * e;
            }
            if (fg) {
                if (TEX) {
                    float gU = 0.f, gV = 0.f;
#pragma unroll
                    for (int c = 0; c < C; c++) { gU += gc[c] * dudc[c]; gV += gc[c] * dvdc[c]; }
                    g.x = gU * (a0c[0] - a2c[0]) + gV * (a0c[1] - a2c[1]);
                    g.y = gU * (a1c[0] - a2c[0]) + gV * (a1c[1] - a2c[1]);
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++) { g.x += gc[c] * (a0c[c] - a2c[c]); g.y += gc[c] * (a1c[c] - a2c[c]); }
                }
            }
            if (fp.rast_out) reinterpret_cast<float4*>(fp.rast_out)[pi] = rout;
            if (fp.colour_out) {
#pragma unroll
                for (int c = 0; c < C; c++) fp.colour_out[pi * C + c] = col[c];
            }
        }
        guv[idx] = g;
    }
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) red[warp] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < FINE_WARPS; w++) s += red[w];
        fp.loss_partial[(size_t)n * rp.NB + bin] = s;
    }
    if (!fp.grad_pos) return;

    // ---- (3) gradient pass: small triangles, warp-balanced (triangle,row) items ----
    const int lim_x = min(ox + BIN, rp.W) - 1, lim_y = min(oy + BIN, rp.H) - 1;
    const int count = rp.bin_count[(size_t)n * rp.NB + bin];
    const int* list = rp.pairs + (size_t)n * 4 * rp.T + rp.bin_offset[(size_t)n * rp.NB + bin];
    GradStage& st = *reinterpret_cast<GradStage*>(&stage[warp]);
    float* acc = acc_all + warp * 9 * 32;
    float* G = fp.grad_pos + (size_t)n * rp.V * 4;
    for (int base = warp * 32; base < count; base += FINE_THREADS) {
        int i = base + lane;
        int rows = 0;
        if (i < count) {
            int t = list[i];
            float4 p0, p1, p2;
            SnappedTri s;
            if (load_triangle(rp, n, t, p0, p1, p2) && setup_triangle(p0, p1, p2, rp, s)) {
                int xa = max(s.pxa, ox), xb = min(s.pxb, lim_x), ya = max(s.pya, oy), yb = min(s.pyb, lim_y);
                if (xa <= xb && ya <= yb) {
                    rows = yb - ya + 1;
                    st.px[0][lane] = p0.x; st.py[0][lane] = p0.y; st.pw[0][lane] = p0.w;
                    st.px[1][lane] = p1.x; st.py[1][lane] = p1.y; st.pw[1][lane] = p1.w;
                    st.px[2][lane] = p2.x; st.py[2][lane] = p2.y; st.pw[2][lane] = p2.w;
                    st.xy[lane] = xa | (ya << 16);
                    st.wn[lane] = (xb - xa + 1) | (rows << 16);
                    st.tri[lane] = t;
                    st.vidx[0][lane] = __ldg(rp.tri + 3 * t); st.vidx[1][lane] = __ldg(rp.tri + 3 * t + 1); st.vidx[2][lane] = __ldg(rp.tri + 3 * t + 2);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 9; c++) acc[c * 32 + lane] = 0.f;
        int incl = rows;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        st.prefix[lane] = incl;
        int total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        for (int k0 = 0; k0 < total; k0 += 32) {
            int k = k0 + lane;
            int j = -1;
            PosGrad a;
#pragma unroll
            for (int c = 0; c < 9; c++) a.v[c] = 0.f;
            if (k < total) {
                j = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (st.prefix[j + step - 1] <= k) j += step;
                int wn = st.wn[j];
                int r = k - (st.prefix[j] - (wn >> 16));
                int wd = wn & 0xffff;
                int xy = st.xy[j];
                int x0 = xy & 0xffff, yy = (int)((unsigned)xy >> 16) + r;
                int t = st.tri[j];
                int kidx = (yy - oy) * BIN + (x0 - ox);
                float fy = pixel_ndc(yy, rp.ys, rp.yo);
                bool loaded = false;
                float q0x = 0, q0y = 0, q0w = 0, q1x = 0, q1y = 0, q1w = 0, q2x = 0, q2y = 0, q2w = 0;
                for (int x = 0; x < wd; x++) {
                    if ((int)(keys[kidx + x] & 0xFFFFFFFFu) != t || keys[kidx + x] == KEY_EMPTY) continue;
                    float2 g = guv[kidx + x];
                    if (g.x == 0.f && g.y == 0.f) continue;
                    if (!loaded) {
                        q0x = st.px[0][j]; q0y = st.py[0][j]; q0w = st.pw[0][j];
                        q1x = st.px[1][j]; q1y = st.py[1][j]; q1w = st.pw[1][j];
                        q2x = st.px[2][j]; q2y = st.py[2][j]; q2w = st.pw[2][j];
                        loaded = true;
                    }
                    pixel_pos_grad(g.x, g.y, pixel_ndc(x0 + x, rp.xs, rp.xo), fy, q0x, q0y, q0w, q1x, q1y, q1w, q2x, q2y, q2w, a);
                }
            }
            // segmented reduction over runs of equal j (items of one triangle are consecutive lanes)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int jo = __shfl_down_sync(0xffffffffu, j, d);
                bool take = (lane + d < 32) && (jo == j) && (j >= 0);
#pragma unroll
                for (int c = 0; c < 9; c++) {
                    float o = __shfl_down_sync(0xffffffffu, a.v[c], d);
                    if (take) a.v[c] += o;
                }
            }
            int jp = __shfl_up_sync(0xffffffffu, j, 1);
            if (j >= 0 && (lane == 0 || jp != j)) {
#pragma unroll
                for (int c = 0; c < 9; c++) acc[c * 32 + j] += a.v[c];
            }
            __syncwarp();
        }
        if (rows > 0) {
            red_vertex(G, st.vidx[0][lane], acc[0 * 32 + lane], acc[1 * 32 + lane], acc[2 * 32 + lane]);
            red_vertex(G, st.vidx[1][lane], acc[3 * 32 + lane], acc[4 * 32 + lane], acc[5 * 32 + lane]);
            red_vertex(G, st.vidx[2][lane], acc[6 * 32 + lane], acc[7 * 32 + lane], acc[8 * 32 + lane]);
        }
        __syncwarp();
    }

    // ---- gradient pass: large triangles (rare), whole CTA per triangle ----
    const int nlarge = rp.large_count[n];
    const int* llist = rp.large_list + (size_t)n * rp.T;
    for (int i = 0; i < nlarge; i++) {
        int t = llist[i];
        float4 p0, p1, p2;
        SnappedTri s;
        if (!load_triangle(rp, n, t, p0, p1, p2) || !setup_triangle(p0, p1, p2, rp, s)) continue;
        int xa = max(s.pxa, ox), xb = min(s.pxb, lim_x), ya = max(s.pya, oy), yb = min(s.pyb, lim_y);
        if (xa > xb || ya > yb) continue;
        PosGrad a;
#pragma unroll
        for (int c = 0; c < 9; c++) a.v[c] = 0.f;
        for (int idx = threadIdx.x; idx < BIN * BIN; idx += FINE_THREADS) {
            unsigned long long key = keys[idx];
            if (key == KEY_EMPTY || (int)(key & 0xFFFFFFFFu) != t) continue;
            float2 g = guv[idx];
            if (g.x == 0.f && g.y == 0.f) continue;
            int px = ox + (idx & (BIN - 1)), py = oy + (idx >> BIN_LOG2);
            pixel_pos_grad(g.x, g.y, pixel_ndc(px, rp.xs, rp.xo), pixel_ndc(py, rp.ys, rp.yo), p0.x, p0.y, p0.w, p1.x, p1.y, p1.w,
                           p2.x, p2.y, p2.w, a);
        }
#pragma unroll
        for (int c = 0; c < 9; c++)
            for (int o = 16; o > 0; o >>= 1) a.v[c] += __shfl_xor_sync(0xffffffffu, a.v[c], o);
        if (lane == 0) {
            red_vertex(G, __ldg(rp.tri + 3 * t), a.v[0], a.v[1], a.v[2]);
            red_vertex(G, __ldg(rp.tri + 3 * t + 1), a.v[3], a.v[4], a.v[5]);
            red_vertex(G, __ldg(rp.tri + 3 * t + 2), a.v[6], a.v[7], a.v[8]);
        }
    }
}

__global__ void __launch_bounds__(256) k_fused_loss_reduce(const double* __restrict__ partial, int n, float k, float* __restrict__ loss)
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

constexpr size_t FUSED_SMEM = sizeof(unsigned long long) * BIN * BIN + sizeof(WarpStage) * FINE_WARPS + sizeof(float2) * BIN * BIN +
                              sizeof(float) * FINE_WARPS * 9 * 32;

template <int C, bool TEX>
int launch_fused(const RasterParams& rp, const FusedParams& fp, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        FPC_CUDA(cudaFuncSetAttribute(k_fused<C, TEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_SMEM));
        attr_set = true;
    }
    k_fused<C, TEX><<<dim3(rp.NB, rp.N), FINE_THREADS, FUSED_SMEM, stream>>>(rp, fp);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t fpc_render_loss_fused_scratch_bytes(int N, int T, int H, int W)
{
    if (N <= 0 || T <= 0 || H <= 0 || W <= 0) return 256;
    int NB = fpc_div_up(W, BIN) * fpc_div_up(H, BIN);
    return align256(raster_layout(N, T, NB).total) + align256((size_t)N * NB * sizeof(double));
}

extern "C" int fpc_render_loss_fused(const float* pos, const int32_t* tri, const float* attr, const int32_t* attr_tri, int Va, int A,
                                     const float* tex, int Ht, int Wt, const void* ref, int ref_is_u8,
                                     int N, int V, int T, int H, int W, int C, float bg, float scale,
                                     float* loss, float* grad_pos, float* rast_out, float* colour_out,
                                     void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(attr && attr_tri && ref && loss, "render_loss_fused: attr, attr_tri, ref and loss must be non-null");
    FPC_CHECK_ARG(C == 1 || C == 3, "render_loss_fused: C must be 1 or 3 (got %d)", C);
    FPC_CHECK_ARG(Va > 0, "render_loss_fused: Va must be positive");
    if (tex) FPC_CHECK_ARG(A == 2 && Ht > 0 && Wt > 0, "render_loss_fused: textured shading needs A == 2 (uv) and a non-empty texture");
    else FPC_CHECK_ARG(A == C, "render_loss_fused: vertex-colour shading needs A == C (got A=%d, C=%d)", A, C);
    FPC_CHECK_ARG(scratch_bytes >= fpc_render_loss_fused_scratch_bytes(N, T, H, W), "render_loss_fused: scratch too small");
    RasterParams rp;
    int st = raster_bin_triangles("render_loss_fused", pos, tri, N, V, T, H, W, scratch, scratch_bytes, stream, rp);
    if (st != FPC_OK) return st;
    FusedParams fp;
    fp.attr = attr; fp.attr_tri = attr_tri; fp.Va = Va; fp.A = A; fp.tex = tex; fp.Ht = Ht; fp.Wt = Wt;
    fp.ref = ref; fp.ref_u8 = ref_is_u8; fp.C = C; fp.bg = bg; fp.k = scale / ((float)H * (float)W * (float)C);
    fp.grad_pos = grad_pos; fp.rast_out = rast_out; fp.colour_out = colour_out;
    fp.loss_partial = (double*)((char*)scratch + align256(raster_layout(N, T, rp.NB).total));
    if (grad_pos) FPC_CUDA(cudaMemsetAsync(grad_pos, 0, (size_t)N * V * 4 * sizeof(float), stream));
    if (tex) st = (C == 1) ? launch_fused<1, true>(rp, fp, stream) : launch_fused<3, true>(rp, fp, stream);
    else st = (C == 1) ? launch_fused<1, false>(rp, fp, stream) : launch_fused<3, false>(rp, fp, stream);
    if (st != FPC_OK) return st;
    k_fused_loss_reduce<<<1, 256, 0, stream>>>(fp.loss_partial, N * rp.NB, fp.k, loss);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
