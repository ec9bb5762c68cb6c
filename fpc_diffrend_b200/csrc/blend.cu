// Blendshape combination V = base + D w and its transpose gradient d_w = D^T d_V
// (replaces fit.blend, reference fit.py:103-129, in the north-star form).
//
// D [R,B] row-major (R = 3V), w [F,B], verts [F,R].
//   F == 1  : HBM-bound GEMV, one warp per row, float4 loads of the row, shuffle reduction.
//   F  > 1  : skinny GEMM  verts[f,r] = base[r] + sum_b D[r,b] w[f,b]  on CUDA cores in full fp32
//             (64x64x16 shared-memory tiles, 4x4 register blocking).
// Backward: split over row chunks, each CTA produces a partial [F,B]; a second kernel sums the partials in
// a fixed order (deterministic, no atomics).
#include "common.cuh"

namespace {

// init(r) of the *_ex entry: the value the product is added to -- verts itself (accumulate), base, or 0
__device__ __forceinline__ float blend_init(const float* base, const float* verts, size_t vi, int r, int accumulate)
{
    return accumulate ? verts[vi] : (base ? __ldg(base + r) : 0.f);
}

__global__ void __launch_bounds__(256) k_blend_gemv(const float* __restrict__ D, const float* __restrict__ base,
                                                    const float* __restrict__ w, int R, int B, float coef, int accumulate, float* verts)
{
    extern __shared__ float sw[];
    for (int i = threadIdx.x; i < B; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool vec = (B & 3) == 0;
    for (int r = warp; r < R; r += nwarps) {
        const float* row = D + (size_t)r * B;
        float acc = 0.f;
        if (vec) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            for (int i = lane; i < (B >> 2); i += 32) {
                float4 d = __ldg(row4 + i);
                acc += d.x * sw[4 * i] + d.y * sw[4 * i + 1] + d.z * sw[4 * i + 2] + d.w * sw[4 * i + 3];
            }
        } else {
            for (int i = lane; i < B; i += 32) acc += __ldg(row + i) * sw[i];
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) verts[r] = blend_init(base, verts, r, r, accumulate) + coef * acc;
    }
}

constexpr int TM = 64, TN = 64, TK = 16;

// C[f, r] = base[r] + sum_k D[r,k] * w[f,k];  tile: 64 rows (r) x 64 frames (f)
__global__ void __launch_bounds__(256) k_blend_gemm(const float* __restrict__ D, const float* __restrict__ base,
                                                    const float* __restrict__ w, int R, int B, int F, float coef, int accumulate, float* verts)
{
    __shared__ float sA[TK][TM + 1];   // D tile, k-major
    __shared__ float sB[TK][TN + 1];   // w tile, k-major
    int r0 = blockIdx.x * TM, f0 = blockIdx.y * TN;
    int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx -> rows (4 each), ty -> frames (4 each)
    float acc[4][4] = {};
    for (int k0 = 0; k0 < B; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            int rr = i / TK, kk = i % TK;
            int r = r0 + rr, k = k0 + kk;
            sA[kk][rr] = (r < R && k < B) ? __ldg(D + (size_t)r * B + k) : 0.f;
        }
        for (int i = threadIdx.x; i < TN * TK; i += 256) {
            int ff = i / TK, kk = i % TK;
            int f = f0 + ff, k = k0 + kk;
            sB[kk][ff] = (f < F && k < B) ? __ldg(w + (size_t)f * B + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = sA[kk][tx * 4 + i]; b[i] = sB[kk][ty * 4 + i]; }
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc[j][i] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int f = f0 + ty * 4 + j;
        if (f >= F) continue;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int r = r0 + tx * 4 + i;
            if (r < R) verts[(size_t)f * R + r] = blend_init(base, verts, (size_t)f * R + r, r, accumulate) + coef * acc[j][i];
        }
    }
}

constexpr int BWD_FCHUNK = 8;   // frames accumulated in registers per pass

// partial[blk, f, b] = sum_{r in chunk(blk)} D[r,b] * dV[f,r]
__global__ void __launch_bounds__(256) k_blend_bwd_partial(const float* __restrict__ D, const float* __restrict__ dV,
                                                           int R, int B, int F, int rows_per_blk, float* __restrict__ partial)
{
    extern __shared__ float sdv[];     // [BWD_FCHUNK][rows_per_blk]
    int r_begin = blockIdx.x * rows_per_blk, r_end = min(R, r_begin + rows_per_blk);
    int nrows = max(0, r_end - r_begin);
    for (int f0 = 0; f0 < F; f0 += BWD_FCHUNK) {
        int nf = min(BWD_FCHUNK, F - f0);
        __syncthreads();
        for (int i = threadIdx.x; i < nf * nrows; i += blockDim.x) {
            int ff = i / nrows, rr = i - ff * nrows;
            sdv[ff * rows_per_blk + rr] = __ldg(dV + (size_t)(f0 + ff) * R + r_begin + rr);
        }
        __syncthreads();
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            float acc[BWD_FCHUNK];
#pragma unroll
            for (int ff = 0; ff < BWD_FCHUNK; ff++) acc[ff] = 0.f;
            for (int rr = 0; rr < nrows; rr++) {
                float d = __ldg(D + (size_t)(r_begin + rr) * B + b);
#pragma unroll
                for (int ff = 0; ff < BWD_FCHUNK; ff++)
                    if (ff < nf) acc[ff] += d * sdv[ff * rows_per_blk + rr];
            }
#pragma unroll
            for (int ff = 0; ff < BWD_FCHUNK; ff++)
                if (ff < nf) partial[((size_t)blockIdx.x * F + f0 + ff) * B + b] = acc[ff];
        }
    }
}

__global__ void __launch_bounds__(256) k_blend_bwd_reduce(const float* __restrict__ partial, int nblk, int FB, float* __restrict__ d_w)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= FB) return;
    float acc = 0.f;
    for (int k = 0; k < nblk; k++) acc += partial[(size_t)k * FB + i];
    d_w[i] = acc;
}

constexpr int BWD_ROWS = 128;

}  // namespace

extern "C" int fpc_blend_fwd_ex(const float* D, const float* base, const float* w, int R, int B, int F, float coef, int accumulate,
                                float* verts, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(D && w && verts, "blend_fwd: null pointer argument");
    FPC_CHECK_ARG(R > 0 && B > 0 && F > 0, "blend_fwd: R, B, F must be positive (got %d %d %d)", R, B, F);
    if (F == 1 && (size_t)B * 4 <= 48 * 1024) {
        int grid = min(fpc_div_up(R, 8), 148 * 8);
        k_blend_gemv<<<grid, 256, (size_t)B * 4, stream>>>(D, base, w, R, B, coef, accumulate, verts);
    } else {
        k_blend_gemm<<<dim3(fpc_div_up(R, TM), fpc_div_up(F, TN)), 256, 0, stream>>>(D, base, w, R, B, F, coef, accumulate, verts);
    }
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_blend_fwd(const float* D, const float* base, const float* w, int R, int B, int F, float* verts,
                             fpc_stream_t stream_)
{
    FPC_CHECK_ARG(base, "blend_fwd: null pointer argument");
    return fpc_blend_fwd_ex(D, base, w, R, B, F, 1.f, 0, verts, stream_);
}

extern "C" size_t fpc_blend_bwd_scratch_bytes(int R, int B, int F)
{
    if (R <= 0 || B <= 0 || F <= 0) return 256;
    return (size_t)fpc_div_up(R, BWD_ROWS) * F * B * sizeof(float) + 256;
}

extern "C" int fpc_blend_bwd(const float* D, const float* d_verts, int R, int B, int F, float* d_w,
                             void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(D && d_verts && d_w, "blend_bwd: null pointer argument");
    FPC_CHECK_ARG(R > 0 && B > 0 && F > 0, "blend_bwd: R, B, F must be positive (got %d %d %d)", R, B, F);
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_blend_bwd_scratch_bytes(R, B, F), "blend_bwd: scratch too small");
    int nblk = fpc_div_up(R, BWD_ROWS);
    float* partial = (float*)scratch;
    k_blend_bwd_partial<<<nblk, 256, (size_t)BWD_FCHUNK * BWD_ROWS * sizeof(float), stream>>>(D, d_verts, R, B, F, BWD_ROWS, partial);
    FPC_LAUNCH_CHECK();
    k_blend_bwd_reduce<<<fpc_div_up((long long)F * B, 256), 256, 0, stream>>>(partial, nblk, F * B, d_w);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
