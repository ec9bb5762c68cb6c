// Batched blendshape combination on the 5th-generation tensor cores (north-star item 1; reference fit.py:103-129 for a
// frame batch):
//     forward   verts [F,R] = base + w [F,B] D^T          -> C[m=r, n=f] = sum_b D[r,b]  w[f,b]        (K = B)
//     backward  d_w   [F,B] = d_verts [F,R] D             -> C[m=b, n=f] = sum_r DT[b,r] d_verts[f,r]  (K = R, split-K)
// Both are C[M,N] = A[M,K] B[N,K]^T with K-major operands, computed by ONE kernel:
//   * TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) streams 128x32 / 64x32 fp32 tiles into a 4-stage shared-memory ring;
//   * the 1e-5 absolute tolerance of the north-star rules out plain TF32 (10-bit mantissa), so the kernel runs the
//     3xTF32 split: four "splitter" warps compute lo = x - (x & 0xFFFFE000) (exact in fp32) of every landed tile into a twin
//     buffer with the same swizzled layout; the landed tile itself is the `hi` operand — the tensor core reads only the upper
//     19 bits of a kind::tf32 operand, i.e. exactly x & 0xFFFFE000 (measured: bit-identical results with and without masking
//     the tile in place, round 2);
//   * one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=64, K=8) three times per k-step
//     (hi*hi + hi*lo + lo*hi) into a 128x64 fp32 accumulator in TMEM; tcgen05.commit releases the stage;
//   * four epilogue warps read the accumulator back with tcgen05.ld (one TMEM lane = one row per thread) and store
//     it transposed, out[(z*N + n)*ldo + m], so that the lanes of a warp write consecutive addresses;
//   * the CTAs are persistent (one per SM) and the accumulator is double-buffered in TMEM: the epilogue of one tile overlaps
//     the loads, splits and MMAs of the next.
// The GEMM is far left of the tensor ridge (AI ~ 24 flop/B at F = 64): the roofline that bounds it is HBM (D is read
// once: 48 MB at config 3); the tensor pipe only has to keep up with the stream.
#include <cuda.h>

#include "common.cuh"
#include "tma.cuh"

using namespace fpc;

namespace {

constexpr int TC_BM = 128, TC_BN = 64, TC_BK = 32;          // tile; BK fp32 = 128 bytes = one swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;               // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;               //  8 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // A hi | A lo | B hi | B lo
constexpr int TC_THREADS = 320;                             // warp 0: TMA, warp 1: MMA + TMEM, warps 2-5: split, warps 6-9: epilogue
constexpr int TC_SPLIT_THREADS = 128;
constexpr int TC_EPI_THREADS = 128;
constexpr unsigned TC_TMEM_COLS = 128;                      // two 128 x 64 fp32 accumulators
constexpr size_t TC_SMEM = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

// kind::tf32 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9 and 10-12), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28   (cute::UMMA::InstrDescriptor)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4, LBO = 1 (ignored for
// swizzled K-major), SBO = 1024 B between 8-row groups, version 1 (Blackwell), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// C[M,N] = A[M,K] B[N,K]^T (3xTF32); out[(z*N + n)*ldo + m] = C[m,n] (+ bias[m]).  PERSISTENT: one CTA per SM walks the tiles
// (m tile, n tile, K split) blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA ring and the splitters run straight on into the next
// tile, the MMA warp alternates between TWO accumulators in TMEM and a separate group of four epilogue warps drains accumulator
// j while the MMAs of tile j + 1 are already running (round 1 had one tile per CTA: prologue, 7 k-steps and epilogue in series).
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_3xtf32(const __grid_constant__ CUtensorMap map_a,
                                                               const __grid_constant__ CUtensorMap map_b, int M, int N,
                                                               int num_kb, int kb_per_split, int tiles_m, int tiles_n, int total_tiles,
                                                               const float* __restrict__ bias, float* __restrict__ out, int ldo)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<size_t>(smem_raw) + 1023) & ~(size_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)TC_STAGES * TC_STAGE_BYTES);
    // bars[0..S) full (TMA landed), [S..2S) ready (split done), [2S..3S) empty (MMA done), [3S..3S+2) accumulator full, [3S+2..3S+4) accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 4);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full0 = bar0, ready0 = bar0 + 8 * TC_STAGES, empty0 = bar0 + 16 * TC_STAGES;
    const uint32_t accf0 = bar0 + 24 * TC_STAGES, acce0 = accf0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 32) {
        for (int s = 0; s < TC_STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(ready0 + 8 * s, TC_SPLIT_THREADS);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(accf0 + 8 * a, 1);
            mbar_init(acce0 + 8 * a, TC_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);

    // tile -> (m block, n block, first k block, k blocks)
    auto decode = [&](int tile, int& m_blk, int& n_blk, int& z, int& kb0, int& nkb) {
        m_blk = tile % tiles_m;
        const int rest = tile / tiles_m;
        n_blk = rest % tiles_n;
        z = rest / tiles_n;
        kb0 = z * kb_per_split;
        nkb = min(kb_per_split, num_kb - kb0);                  // >= 1 (host)
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int m_blk, n_blk, z, kb0, nkb;
                decode(tile, m_blk, n_blk, z, kb0, nkb);
                for (int i = 0; i < nkb; i++, it++) {
                    const int s = it % TC_STAGES;
                    const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    const uint32_t sa = smem0 + (uint32_t)s * TC_STAGE_BYTES;
                    mbar_expect_tx(full0 + 8 * s, TC_A_BYTES + TC_B_BYTES);
                    tma_load_2d(sa, &map_a, full0 + 8 * s, (kb0 + i) * TC_BK, m_blk * TC_BM);
                    tma_load_2d(sa + 2 * TC_A_BYTES, &map_b, full0 + 8 * s, (kb0 + i) * TC_BK, n_blk * TC_BN);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int it = 0, j = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, j++) {
                int m_blk, n_blk, z, kb0, nkb;
                decode(tile, m_blk, n_blk, z, kb0, nkb);
                const int acc = j & 1;
                mbar_wait(acce0 + 8 * acc, ((uint32_t)(j >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * TC_BN);
                for (int i = 0; i < nkb; i++, it++) {
                    const int s = it % TC_STAGES;
                    const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                    mbar_wait(ready0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem0 + (uint32_t)s * TC_STAGE_BYTES;
                    const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + TC_A_BYTES);
                    const uint64_t b_hi = umma_desc(sa + 2 * TC_A_BYTES), b_lo = umma_desc(sa + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; k++) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);          // 32 bytes per K = 8 step inside the swizzle row
                        umma_tf32(tmem_acc, a_hi + adv, b_hi + adv, (i > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(tmem_acc, a_hi + adv, b_lo + adv, 1u);
                        umma_tf32(tmem_acc, a_lo + adv, b_hi + adv, 1u);
                    }
                    umma_commit(empty0 + 8 * s);                                    // stage free once these MMAs have read it
                }
                umma_commit(accf0 + 8 * acc);                                       // accumulator complete
            }
        }
    } else if (warp < 2 + TC_SPLIT_THREADS / 32) {
        // ===== splitters (hi / lo) =====
        const int tid = threadIdx.x - 64;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int m_blk, n_blk, z, kb0, nkb;
            decode(tile, m_blk, n_blk, z, kb0, nkb);
            for (int i = 0; i < nkb; i++, it++) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                unsigned char* st = smem + (size_t)s * TC_STAGE_BYTES;
                uint4* a_hi = reinterpret_cast<uint4*>(st);
                uint4* a_lo = reinterpret_cast<uint4*>(st + TC_A_BYTES);
                uint4* b_hi = reinterpret_cast<uint4*>(st + 2 * TC_A_BYTES);
                uint4* b_lo = reinterpret_cast<uint4*>(st + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
                for (int jj = 0; jj < TC_A_BYTES / 16 / TC_SPLIT_THREADS; jj++) {
                    const int o = tid + jj * TC_SPLIT_THREADS;
                    uint4 v = a_hi[o], h, l;
                    h.x = v.x & 0xFFFFE000u; h.y = v.y & 0xFFFFE000u; h.z = v.z & 0xFFFFE000u; h.w = v.w & 0xFFFFE000u;
                    l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
                    l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
                    l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
                    l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
                    a_lo[o] = l;            // (the raw tile serves as `hi`: kind::tf32 ignores the low 13 mantissa bits)
                }
#pragma unroll
                for (int jj = 0; jj < TC_B_BYTES / 16 / TC_SPLIT_THREADS; jj++) {
                    const int o = tid + jj * TC_SPLIT_THREADS;
                    uint4 v = b_hi[o], h, l;
                    h.x = v.x & 0xFFFFE000u; h.y = v.y & 0xFFFFE000u; h.z = v.z & 0xFFFFE000u; h.w = v.w & 0xFFFFE000u;
                    l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
                    l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
                    l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
                    l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
                    b_lo[o] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the MMA (async proxy)
                mbar_arrive(ready0 + 8 * s);
            }
        }
    } else {
        // ===== epilogue: TMEM lane quarter of this warp is (warp index % 4) =====
        int j = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, j++) {
            int m_blk, n_blk, z, kb0, nkb;
            decode(tile, m_blk, n_blk, z, kb0, nkb);
            const int acc = j & 1;
            mbar_wait(accf0 + 8 * acc, (uint32_t)(j >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = 32 * (warp & 3) + lane;
            const int m = m_blk * TC_BM + row;
            const float bv = (bias && m < M) ? __ldg(bias + m) : 0.f;
            float v[TC_BN];
#pragma unroll
            for (int half = 0; half < TC_BN / 32; half++)
                tmem_ld32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(acc * TC_BN + 32 * half), v + 32 * half);
            // the accumulator is in registers: hand it back to the MMA warp before the (slow) global stores
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(acce0 + 8 * acc);
            if (m < M) {
#pragma unroll
                for (int jn = 0; jn < TC_BN; jn++) {
                    const int n = n_blk * TC_BN + jn;
                    if (n < N) out[((size_t)z * N + n) * ldo + m] = v[jn] + bv;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// d_w [F*B] = sum over the K splits of part [S][F*B], in split order (deterministic)
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ part, int S, long long n, float* __restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < S; k++) s += part[(size_t)k * n + i];
    out[i] = s;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode()
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
    }
    return fn;
}

// row-major fp32 matrix [rows, cols] (cols contiguous) -> tiles of box_rows x 32 columns, 128-byte swizzle, zero OOB fill
int make_map(CUtensorMap* map, const float* ptr, long long rows, long long cols, int box_rows)
{
    encode_tiled_fn enc = get_encode();
    if (!enc) { fpc_set_error("blend_tc: cuTensorMapEncodeTiled is not available from this driver"); return FPC_ERR_UNSUPPORTED; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fpc_set_error("blend_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return FPC_ERR_CUDA; }
    return FPC_OK;
}

int launch_gemm(const float* A, const float* Bm, long long M, long long N, long long K, int splits, const float* bias, float* out, int ldo,
                cudaStream_t stream, int* splits_used)
{
    alignas(64) CUtensorMap map_a, map_b;
    int st = make_map(&map_a, A, M, K, TC_BM);
    if (st != FPC_OK) return st;
    st = make_map(&map_b, Bm, N, K, TC_BN);
    if (st != FPC_OK) return st;
    static FpcPerDeviceOnce attr_set;
    if (attr_set.need()) {
        FPC_CUDA(cudaFuncSetAttribute(k_gemm_3xtf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        attr_set.done();
    }
    const int num_kb = fpc_div_up(K, TC_BK);
    const int per = fpc_div_up(num_kb, splits);
    const int zs = fpc_div_up(num_kb, per);                 // every split owns >= 1 k-block
    const int tiles_m = fpc_div_up(M, TC_BM), tiles_n = fpc_div_up(N, TC_BN);
    const long long total = (long long)tiles_m * tiles_n * zs;
    FPC_CHECK_ARG(total < (1LL << 30), "blend_tc: too many tiles (%lld)", total);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(total < sms ? total : sms);         // persistent: one CTA per SM (192 KB of shared memory each)
    k_gemm_3xtf32<<<grid, TC_THREADS, TC_SMEM, stream>>>(map_a, map_b, (int)M, (int)N, num_kb, per, tiles_m, tiles_n, (int)total, bias, out, ldo);
    FPC_LAUNCH_CHECK();
    *splits_used = zs;
    return FPC_OK;
}

int bwd_splits(int R, int B)
{
    const int mt = fpc_div_up(B, TC_BM);
    int s = 148 / mt;                                        // one wave of CTAs on 148 SMs
    const int num_kb = fpc_div_up(R, TC_BK);
    if (s > num_kb) s = num_kb;
    return s < 1 ? 1 : s;
}

}  // namespace

extern "C" int fpc_blend_tc_supported(int R, int B, int F)
{
    // TMA needs 16-byte row pitches for D [R,B], w [F,B], DT [B,R] and d_verts [F,R]
    return R > 0 && B > 0 && F > 0 && (B % 4) == 0 && (R % 4) == 0;
}

extern "C" int fpc_blend_fwd_tc(const float* D, const float* base, const float* w, int R, int B, int F, float* verts, fpc_stream_t stream_)
{
    FPC_CHECK_ARG(D && base && w && verts, "blend_fwd_tc: null pointer argument");
    FPC_CHECK_ARG(fpc_blend_tc_supported(R, B, F), "blend_fwd_tc: needs B %% 4 == 0 and R %% 4 == 0 (got R=%d B=%d); use fpc_blend_fwd", R, B);
    int zs = 0;
    return launch_gemm(D, w, R, F, B, 1, base, verts, R, (cudaStream_t)stream_, &zs);
}

extern "C" size_t fpc_blend_bwd_tc_scratch_bytes(int R, int B, int F)
{
    if (R <= 0 || B <= 0 || F <= 0) return 256;
    return (size_t)bwd_splits(R, B) * F * B * sizeof(float) + 256;
}

extern "C" int fpc_blend_bwd_tc(const float* DT, const float* d_verts, int R, int B, int F, float* d_w,
                                void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(DT && d_verts && d_w, "blend_bwd_tc: null pointer argument");
    FPC_CHECK_ARG(fpc_blend_tc_supported(R, B, F), "blend_bwd_tc: needs B %% 4 == 0 and R %% 4 == 0 (got R=%d B=%d); use fpc_blend_bwd", R, B);
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_blend_bwd_tc_scratch_bytes(R, B, F), "blend_bwd_tc: scratch too small");
    float* part = (float*)scratch;
    int zs = 0;
    int st = launch_gemm(DT, d_verts, B, F, R, bwd_splits(R, B), nullptr, part, B, stream, &zs);
    if (st != FPC_OK) return st;
    const long long n = (long long)F * B;
    k_splitk_reduce<<<fpc_div_up(n, 256), 256, 0, stream>>>(part, zs, n, d_w);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
