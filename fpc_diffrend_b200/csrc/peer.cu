// Exchange steps of the camera-split mode over NVLink peer memory (multi-GPU part of the north-star — the
// reference is single-process, fit.py has no counterpart).  The ranks of one box map each other's buffers (symmetric memory,
// set up by the host: fpc_diffrend_b200/fit.py) and the kernels below read / write them directly, so the three exchanges of an
// iteration — blended vertices out to every rank, vertex gradients summed by rows, packed parameter gradient summed — cost one
// device-side barrier each instead of one NCCL collective each:
//   fpc_blend_fwd_bcast : V = base + D w for this rank's ROWS of D, the GEMV epilogue stores every result into the vertex
//                         buffer of every rank (lane p of the warp writes to peer p): compute + all-gather in one kernel;
//   fpc_peer_store_rows : the same distribution for rows computed elsewhere (frame batches);
//   fpc_peer_sum_rows   : this rank's rows of sum_p x_p (fixed rank order: bit-identical on every rank, run to run) = the
//                         reduce-scatter of the vertex gradients, feeding D^T;
//   fpc_peer_sum        : sum_p x_p of a small vector on every rank = the all-reduce of the packed gradient.
// Ordering between ranks is the host's job (a symmetric-memory barrier between producer and consumer kernels).
#include "common.cuh"

namespace {

constexpr int MAX_PEERS = 16;
struct PeerTable { float* p[MAX_PEERS]; };

int fill_table(const char* who, void* const* peers, int world, PeerTable& tab)
{
    FPC_CHECK_ARG(peers && world >= 1 && world <= MAX_PEERS, "%s: needs 1 <= world <= %d peer pointers (got %d)", who, MAX_PEERS, world);
    for (int i = 0; i < MAX_PEERS; i++) tab.p[i] = i < world ? (float*)peers[i] : nullptr;
    for (int i = 0; i < world; i++) FPC_CHECK_ARG(tab.p[i], "%s: peer pointer %d is null", who, i);
    return FPC_OK;
}

// one warp per row (blend.cu: k_blend_gemv); every lane ends up with the row's dot product, lane p stores it to peer p
__global__ void __launch_bounds__(256) k_blend_gemv_bcast(const float* __restrict__ D, const float* __restrict__ base,
                                                          const float* __restrict__ w, int R, int B, long long row0, PeerTable tab, int world)
{
    extern __shared__ float sw[];
    for (int i = threadIdx.x; i < B; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool vec = (B & 3) == 0;
    for (int r = warp; r < R; r += nwarps) {
        const float* row = D + (size_t)r * B;
        float acc = 0.f;
        if (vec) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            for (int i = lane; i < (B >> 2); i += 32) {
                const float4 d = __ldg(row4 + i);
                acc += d.x * sw[4 * i] + d.y * sw[4 * i + 1] + d.z * sw[4 * i + 2] + d.w * sw[4 * i + 3];
            }
        } else {
            for (int i = lane; i < B; i += 32) acc += __ldg(row + i) * sw[i];
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const float v = __ldg(base + r) + acc;
        for (int p = lane; p < world; p += 32) tab.p[p][row0 + r] = v;
    }
}

__global__ void __launch_bounds__(256) k_peer_store_rows(const float* __restrict__ src, PeerTable tab, int world, int F, long long rl, long long rt, long long row0)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)F * rl) return;
    const long long f = i / rl, r = i - f * rl;
    const float v = src[i];
    for (int p = 0; p < world; p++) tab.p[p][f * rt + row0 + r] = v;
}

__global__ void __launch_bounds__(256) k_peer_sum_rows(PeerTable tab, int world, int F, long long rl, long long rt, long long row0, float* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)F * rl) return;
    const long long f = i / rl, r = i - f * rl;
    float s = 0.f;
    for (int p = 0; p < world; p++) s += tab.p[p][f * rt + row0 + r];          // rank order: the same bits on every rank
    out[i] = s;
}

__global__ void __launch_bounds__(256) k_peer_sum(PeerTable tab, int world, long long n, float* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int p = 0; p < world; p++) s += tab.p[p][i];
    out[i] = s;
}

}  // namespace

extern "C" int fpc_blend_fwd_bcast(const float* D, const float* base, const float* w, int R_local, int B, long long row0,
                                   void* const* peer_verts, int world, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(D && base && w, "blend_fwd_bcast: null pointer argument");
    FPC_CHECK_ARG(R_local > 0 && B > 0 && row0 >= 0 && (size_t)B * 4 <= 48 * 1024, "blend_fwd_bcast: R_local, B must be positive, B <= 12288 (got %d %d)", R_local, B);
    PeerTable tab;
    int st = fill_table("blend_fwd_bcast", peer_verts, world, tab);
    if (st != FPC_OK) return st;
    const int grid = fpc_div_up(R_local, 8) < 148 * 8 ? fpc_div_up(R_local, 8) : 148 * 8;
    k_blend_gemv_bcast<<<grid, 256, (size_t)B * 4, stream>>>(D, base, w, R_local, B, row0, tab, world);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_peer_store_rows(const float* src, void* const* peers, int world, int F, long long rows_local, long long rows_total,
                                   long long row0, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(src && F > 0 && rows_local > 0 && row0 >= 0 && row0 + rows_local <= rows_total, "peer_store_rows: bad row range");
    PeerTable tab;
    int st = fill_table("peer_store_rows", peers, world, tab);
    if (st != FPC_OK) return st;
    k_peer_store_rows<<<fpc_div_up((long long)F * rows_local, 256), 256, 0, stream>>>(src, tab, world, F, rows_local, rows_total, row0);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_peer_sum_rows(void* const* peers, int world, int F, long long rows_local, long long rows_total, long long row0,
                                 float* out, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(out && F > 0 && rows_local > 0 && row0 >= 0 && row0 + rows_local <= rows_total, "peer_sum_rows: bad row range");
    PeerTable tab;
    int st = fill_table("peer_sum_rows", peers, world, tab);
    if (st != FPC_OK) return st;
    k_peer_sum_rows<<<fpc_div_up((long long)F * rows_local, 256), 256, 0, stream>>>(tab, world, F, rows_local, rows_total, row0, out);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_peer_sum(void* const* peers, int world, long long n, float* out, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(out && n > 0, "peer_sum: out must be non-null and n positive");
    PeerTable tab;
    int st = fill_table("peer_sum", peers, world, tab);
    if (st != FPC_OK) return st;
    k_peer_sum<<<fpc_div_up(n, 256), 256, 0, stream>>>(tab, world, n, out);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
