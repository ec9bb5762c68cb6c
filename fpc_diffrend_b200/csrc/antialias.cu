// Silhouette-edge antialiasing fwd/bwd (replaces dr.antialias, reference fit.py:160).
//
// Differences from the upstream structure, by design:
//   * topology is a per-mesh adjacency table tri_opp [T,3] built once (fpc_topology_build) instead of an
//     edge hash rebuilt on every call (the reference never passes topology_hash, fit.py:160);
//   * the forward pass and the colour gradient are pixel-parallel GATHERS: every pixel analyses the four
//     pixel pairs it belongs to and sums what it receives in a fixed order — no work queue, no atomics,
//     run-to-run deterministic.  Only the (sparse) silhouette position gradient is scattered.
#include "antialias.cuh"

using namespace fpc;

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
