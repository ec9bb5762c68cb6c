// Mesh regularisers of the shipped loss (reference fit.py:578-582): pytorch3d's mesh_laplacian_smoothing(method='uniform')
// — SQUARED, as the reference does (fit.py:581) —, mesh_edge_loss(target) and mesh_normal_consistency, for a batch of
// frames, with their gradient w.r.t. the blended vertices (added to d V before the D^T contraction).
//
// The reference rebuilds a pytorch3d Meshes object (edge list, sparse Laplacian) on every iteration; here the topology
// (vertex neighbour CSR, face pairs across manifold edges) is precomputed once (fpc_diffrend_b200/topology.py) and
//   k_reg_vertex1 : per vertex  lv = mean(neighbours) - v, r = |lv|, g = lv / r  + this vertex's share of the edge term
//   k_reg_nc_fwd  : per face pair 1 - cos(n0, -n1)                                   (only when w_nc != 0)
//   k_reg_final   : fixed-order sums -> per-frame terms, loss += sum_f (w_lap lap_f^2 + w_edge edge_f + w_nc nc_f)
//   k_reg_vertex2 : GATHER over the neighbour CSR (no atomics):  d v_k = 2 w_lap lap_f (-g_k + sum_{i in N(k)} g_i / deg_i)
//                   + (2 w_edge / E) sum_{j in N(k)} (|v_k - v_j| - target) (v_k - v_j) / |v_k - v_j|
//   k_reg_nc_bwd  : per face pair, 12 float REDs                                     (only when w_nc != 0)
// pytorch3d definitions (v0.7, restated; the package is not installed here -> parity unpinned, see oracle/golden.py):
//   laplacian: L v = (1/deg_i) sum_j v_j - v_i over unique edges; loss = (1/V) sum_i |L v|_i
//   edge     : (1/E) sum over unique edges (|v0 - v1| - target)^2
//   normal   : faces sharing edge (v0,v1) with opposite vertices a, b: n0 = (v1-v0) x (a-v0), n1 = (v1-v0) x (b-v0),
//              loss = mean(1 - cos(n0, -n1))
#include "common.cuh"

namespace {

constexpr int REG_THREADS = 256;

struct RegParams {
    const float* verts;        // [F,V,3]
    int F, V, E, E2;
    const int32_t* nbr_off;    // [V+1]
    const int32_t* nbr_idx;    // [2E]
    const int32_t* quads;      // [E2,4]  v0, v1, a, b
    float w_lap, w_edge, edge_target, w_nc;
    float* g;                  // [F,V,3] scratch: normalised Laplacian vectors
    double* part;              // [F][nblk][2] scratch: sum r, sum edge term (each edge seen from both ends)
    double* part_nc;           // [F][nblk2] scratch
    float* coef;               // [F] scratch: 2 w_lap lap_f
    int nblk, nblk2;
};

__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }

__device__ __forceinline__ double block_sum(double v, double* red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < REG_THREADS / 32; w++) s += red[w];
    __syncthreads();
    return s;           // valid in thread 0
}

__global__ void __launch_bounds__(REG_THREADS) k_reg_vertex1(RegParams rp)
{
    __shared__ double red[REG_THREADS / 32];
    const int f = blockIdx.y, i = blockIdx.x * REG_THREADS + threadIdx.x;
    const float* Vf = rp.verts + (size_t)f * rp.V * 3;
    double r_acc = 0.0, e_acc = 0.0;
    if (i < rp.V) {
        const float3 v = ld3(Vf + 3 * (size_t)i);
        const int a = __ldg(rp.nbr_off + i), b = __ldg(rp.nbr_off + i + 1);
        float sx = 0.f, sy = 0.f, sz = 0.f, es = 0.f;
        for (int k = a; k < b; k++) {
            const float3 q = ld3(Vf + 3 * (size_t)__ldg(rp.nbr_idx + k));
            sx += q.x; sy += q.y; sz += q.z;
            if (rp.w_edge != 0.f) {
                float dx = v.x - q.x, dy = v.y - q.y, dz = v.z - q.z;
                float d = sqrtf(dx * dx + dy * dy + dz * dz) - rp.edge_target;
                es += d * d;
            }
        }
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (b > a) {
            const float inv = 1.f / (float)(b - a);
            const float lx = sx * inv - v.x, ly = sy * inv - v.y, lz = sz * inv - v.z;
            const float r = sqrtf(lx * lx + ly * ly + lz * lz);
            r_acc = (double)r;
            if (r > 0.f) { const float ir = 1.f / r; gx = lx * ir; gy = ly * ir; gz = lz * ir; }
        }
        float* G = rp.g + ((size_t)f * rp.V + i) * 3;
        G[0] = gx; G[1] = gy; G[2] = gz;
        e_acc = (double)es;
    }
    double rs = block_sum(r_acc, red);
    double es = block_sum(e_acc, red);
    if (threadIdx.x == 0) {
        double* P = rp.part + ((size_t)f * rp.nblk + blockIdx.x) * 2;
        P[0] = rs; P[1] = es;
    }
}

struct Quad { float3 v0, v1, a, b; };

__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__global__ void __launch_bounds__(REG_THREADS) k_reg_nc_fwd(RegParams rp)
{
    __shared__ double red[REG_THREADS / 32];
    const int f = blockIdx.y, e = blockIdx.x * REG_THREADS + threadIdx.x;
    const float* Vf = rp.verts + (size_t)f * rp.V * 3;
    double acc = 0.0;
    if (e < rp.E2) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(rp.quads) + e);
        const float3 v0 = ld3(Vf + 3 * (size_t)q.x), v1 = ld3(Vf + 3 * (size_t)q.y), a = ld3(Vf + 3 * (size_t)q.z), b = ld3(Vf + 3 * (size_t)q.w);
        const float3 ed = sub3(v1, v0);
        const float3 n0 = cross3(ed, sub3(a, v0)), n1 = cross3(ed, sub3(b, v0));
        const float l0 = sqrtf(dot3(n0, n0)), l1 = sqrtf(dot3(n1, n1));
        const float c = -dot3(n0, n1) / (fmaxf(l0, 1e-8f) * fmaxf(l1, 1e-8f));          // cos(n0, -n1), torch's eps clamp
        acc = (double)(1.f - c);
    }
    double s = block_sum(acc, red);
    if (threadIdx.x == 0) rp.part_nc[(size_t)f * rp.nblk2 + blockIdx.x] = s;
}

// one CTA: per frame fixed-order sums; loss += total; terms [F,3] (optional) = raw lap, edge, nc values
__global__ void __launch_bounds__(REG_THREADS) k_reg_final(RegParams rp, float* __restrict__ loss, float* __restrict__ terms)
{
    __shared__ double red[REG_THREADS / 32];
    __shared__ double total;
    if (threadIdx.x == 0) total = 0.0;
    __syncthreads();
    for (int f = 0; f < rp.F; f++) {
        double r = 0.0, e = 0.0, c = 0.0;
        for (int k = threadIdx.x; k < rp.nblk; k += REG_THREADS) {
            const double* P = rp.part + ((size_t)f * rp.nblk + k) * 2;
            r += P[0]; e += P[1];
        }
        if (rp.w_nc != 0.f)
            for (int k = threadIdx.x; k < rp.nblk2; k += REG_THREADS) c += rp.part_nc[(size_t)f * rp.nblk2 + k];
        r = block_sum(r, red);
        e = block_sum(e, red);
        c = block_sum(c, red);
        if (threadIdx.x == 0) {
            const float lap = (float)(r / (double)rp.V);
            const float edge = rp.E > 0 ? (float)(e / (2.0 * (double)rp.E)) : 0.f;      // every edge was seen from both ends
            const float nc = (rp.w_nc != 0.f && rp.E2 > 0) ? (float)(c / (double)rp.E2) : 0.f;
            rp.coef[f] = 2.f * rp.w_lap * lap;
            if (terms) { terms[3 * f] = lap; terms[3 * f + 1] = edge; terms[3 * f + 2] = nc; }
            total += (double)(rp.w_lap * lap * lap + rp.w_edge * edge + rp.w_nc * nc);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && loss) loss[0] += (float)total;
}

__global__ void __launch_bounds__(REG_THREADS) k_reg_vertex2(RegParams rp, float* __restrict__ d_verts, int accumulate)
{
    const int f = blockIdx.y, k = blockIdx.x * REG_THREADS + threadIdx.x;
    if (k >= rp.V) return;
    const float* Vf = rp.verts + (size_t)f * rp.V * 3;
    const float* Gf = rp.g + (size_t)f * rp.V * 3;
    const float coef = rp.coef[f];
    const float ce = rp.E > 0 ? 2.f * rp.w_edge / (float)rp.E : 0.f;
    const float3 v = ld3(Vf + 3 * (size_t)k);
    const int a = __ldg(rp.nbr_off + k), b = __ldg(rp.nbr_off + k + 1);
    float lx = 0.f, ly = 0.f, lz = 0.f, ex = 0.f, ey = 0.f, ez = 0.f;
    for (int j = a; j < b; j++) {
        const int i = __ldg(rp.nbr_idx + j);
        const int deg = __ldg(rp.nbr_off + i + 1) - __ldg(rp.nbr_off + i);
        const float inv = 1.f / (float)deg;
        const float* gi = Gf + 3 * (size_t)i;
        lx += gi[0] * inv; ly += gi[1] * inv; lz += gi[2] * inv;
        if (ce != 0.f) {
            const float3 q = ld3(Vf + 3 * (size_t)i);
            const float dx = v.x - q.x, dy = v.y - q.y, dz = v.z - q.z;
            const float len = sqrtf(dx * dx + dy * dy + dz * dz);
            if (len > 0.f) {
                const float s = (len - rp.edge_target) / len;
                ex += s * dx; ey += s * dy; ez += s * dz;
            }
        }
    }
    const float* gk = Gf + 3 * (size_t)k;
    const float ox = coef * ((1.f / (float)rp.V) * (lx - gk[0])) + ce * ex;
    const float oy = coef * ((1.f / (float)rp.V) * (ly - gk[1])) + ce * ey;
    const float oz = coef * ((1.f / (float)rp.V) * (lz - gk[2])) + ce * ez;
    float* o = d_verts + ((size_t)f * rp.V + k) * 3;
    if (accumulate) { o[0] += ox; o[1] += oy; o[2] += oz; }
    else { o[0] = ox; o[1] = oy; o[2] = oz; }
}

__global__ void __launch_bounds__(REG_THREADS) k_reg_nc_bwd(RegParams rp, float* __restrict__ d_verts)
{
    const int f = blockIdx.y, e = blockIdx.x * REG_THREADS + threadIdx.x;
    if (e >= rp.E2) return;
    const float* Vf = rp.verts + (size_t)f * rp.V * 3;
    const int4 q = __ldg(reinterpret_cast<const int4*>(rp.quads) + e);
    const float3 v0 = ld3(Vf + 3 * (size_t)q.x), v1 = ld3(Vf + 3 * (size_t)q.y), a = ld3(Vf + 3 * (size_t)q.z), b = ld3(Vf + 3 * (size_t)q.w);
    const float3 ed = sub3(v1, v0), p = sub3(a, v0), qq = sub3(b, v0);
    const float3 n0 = cross3(ed, p), n1 = cross3(ed, qq);
    const float l0 = sqrtf(dot3(n0, n0)), l1 = sqrtf(dot3(n1, n1));
    if (!(l0 > 1e-8f) || !(l1 > 1e-8f)) return;               // clamped norms: zero gradient through the clamp
    const float il = 1.f / (l0 * l1);
    const float c = dot3(n0, n1) * il;                          // cos(n0, n1); loss = 1 + c
    const float s = rp.w_nc / (float)rp.E2;                     // d total / d c
    // d c / d n0 = n1 / (l0 l1) - c n0 / l0^2
    const float3 g0 = make_float3(s * (n1.x * il - c * n0.x / (l0 * l0)), s * (n1.y * il - c * n0.y / (l0 * l0)), s * (n1.z * il - c * n0.z / (l0 * l0)));
    const float3 g1 = make_float3(s * (n0.x * il - c * n1.x / (l1 * l1)), s * (n0.y * il - c * n1.y / (l1 * l1)), s * (n0.z * il - c * n1.z / (l1 * l1)));
    // n0 = ed x p, n1 = ed x qq:  d ed = p x g0 + qq x g1,  d p = g0 x ed,  d qq = g1 x ed
    const float3 de0 = cross3(p, g0), de1 = cross3(qq, g1);
    const float3 de = make_float3(de0.x + de1.x, de0.y + de1.y, de0.z + de1.z);
    const float3 dp = cross3(g0, ed), dq = cross3(g1, ed);
    float* D = d_verts + (size_t)f * rp.V * 3;
    atomicAdd(D + 3 * (size_t)q.y + 0, de.x); atomicAdd(D + 3 * (size_t)q.y + 1, de.y); atomicAdd(D + 3 * (size_t)q.y + 2, de.z);
    atomicAdd(D + 3 * (size_t)q.z + 0, dp.x); atomicAdd(D + 3 * (size_t)q.z + 1, dp.y); atomicAdd(D + 3 * (size_t)q.z + 2, dp.z);
    atomicAdd(D + 3 * (size_t)q.w + 0, dq.x); atomicAdd(D + 3 * (size_t)q.w + 1, dq.y); atomicAdd(D + 3 * (size_t)q.w + 2, dq.z);
    atomicAdd(D + 3 * (size_t)q.x + 0, -(de.x + dp.x + dq.x)); atomicAdd(D + 3 * (size_t)q.x + 1, -(de.y + dp.y + dq.y));
    atomicAdd(D + 3 * (size_t)q.x + 2, -(de.z + dp.z + dq.z));
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t fpc_mesh_reg_scratch_bytes(int F, int V, int E2)
{
    if (F <= 0 || V <= 0) return 256;
    const int nblk = fpc_div_up(V, REG_THREADS), nblk2 = fpc_div_up(E2 > 0 ? E2 : 1, REG_THREADS);
    return align256((size_t)F * V * 3 * sizeof(float)) + align256((size_t)F * nblk * 2 * sizeof(double)) +
           align256((size_t)F * nblk2 * sizeof(double)) + align256((size_t)F * sizeof(float));
}

extern "C" int fpc_mesh_reg_fwd_bwd(const float* verts, int F, int V, const int32_t* nbr_off, const int32_t* nbr_idx, int E,
                                    const int32_t* edge_quads, int E2, float w_lap, float w_edge, float edge_target, float w_nc,
                                    float* loss_accum, float* terms, float* d_verts, int accumulate,
                                    void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(verts && nbr_off && nbr_idx && d_verts, "mesh_reg: verts, nbr_off, nbr_idx and d_verts must be non-null");
    FPC_CHECK_ARG(F > 0 && V > 0 && E >= 0 && E2 >= 0 && F <= 65535, "mesh_reg: invalid sizes (F=%d V=%d E=%d E2=%d)", F, V, E, E2);
    FPC_CHECK_ARG(w_nc == 0.f || (edge_quads && (reinterpret_cast<size_t>(edge_quads) & 15) == 0), "mesh_reg: w_nc != 0 needs 16-byte aligned edge_quads [E2,4]");
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_mesh_reg_scratch_bytes(F, V, E2), "mesh_reg: scratch too small");
    RegParams rp;
    rp.verts = verts; rp.F = F; rp.V = V; rp.E = E; rp.E2 = E2; rp.nbr_off = nbr_off; rp.nbr_idx = nbr_idx; rp.quads = edge_quads;
    rp.w_lap = w_lap; rp.w_edge = w_edge; rp.edge_target = edge_target; rp.w_nc = w_nc;
    rp.nblk = fpc_div_up(V, REG_THREADS); rp.nblk2 = fpc_div_up(E2 > 0 ? E2 : 1, REG_THREADS);
    char* s = (char*)scratch;
    rp.g = (float*)s;                    s += align256((size_t)F * V * 3 * sizeof(float));
    rp.part = (double*)s;                s += align256((size_t)F * rp.nblk * 2 * sizeof(double));
    rp.part_nc = (double*)s;             s += align256((size_t)F * rp.nblk2 * sizeof(double));
    rp.coef = (float*)s;
    const bool nc = (w_nc != 0.f) && E2 > 0;
    if (!nc) rp.w_nc = 0.f;
    k_reg_vertex1<<<dim3(rp.nblk, F), REG_THREADS, 0, stream>>>(rp);
    FPC_LAUNCH_CHECK();
    if (nc) {
        k_reg_nc_fwd<<<dim3(rp.nblk2, F), REG_THREADS, 0, stream>>>(rp);
        FPC_LAUNCH_CHECK();
    }
    k_reg_final<<<1, REG_THREADS, 0, stream>>>(rp, loss_accum, terms);
    FPC_LAUNCH_CHECK();
    k_reg_vertex2<<<dim3(rp.nblk, F), REG_THREADS, 0, stream>>>(rp, d_verts, accumulate);
    FPC_LAUNCH_CHECK();
    if (nc) {
        k_reg_nc_bwd<<<dim3(rp.nblk2, F), REG_THREADS, 0, stream>>>(rp, d_verts);
        FPC_LAUNCH_CHECK();
    }
    return FPC_OK;
}
