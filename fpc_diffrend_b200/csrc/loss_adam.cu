// Background composite + image loss with its gradient, Adam step, quaternion renormalisation, library plumbing.
// (replaces reference fit.py:161, the first term of fit.py:579, torch.optim.Adam + LambdaLR fit.py:493-505,
//  610-613 and the renorm at fit.py:616-618)
#include <stdarg.h>

#include "common.cuh"

// ---- error plumbing -------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void fpc_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* fpc_last_error(void) { return g_err; }
extern "C" int fpc_abi_version(void) { return FPC_B200_ABI_VERSION; }

extern "C" int fpc_check_device(void)
{
    int dev = 0;
    FPC_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    FPC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    FPC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) {
        fpc_set_error("fpc_b200 is built for sm_100a only; device %d is compute capability %d.%d and there is no fallback path", dev, major, minor);
        return FPC_ERR_UNSUPPORTED;
    }
    return FPC_OK;
}

namespace {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_PER_THREAD = 8;

// comp = id > 0 ? colour : bg ; e = ref - 255 comp ; loss += e^2 ; d_colour = id > 0 ? -2*255*k*e : 0
// k = scale / (H*W*C).  Block partial sums are written out and reduced in a fixed order by k_loss_reduce.
__global__ void __launch_bounds__(LOSS_THREADS) k_image_loss(const float* __restrict__ colour, const float* __restrict__ rast,
                                                             const float* __restrict__ ref, long long npx, int C, float bg, float k, int l1,
                                                             float* __restrict__ d_colour, float* __restrict__ comp,
                                                             double* __restrict__ partial)
{
    __shared__ double red[LOSS_THREADS / 32];
    double acc = 0.0;
    long long base = (long long)blockIdx.x * LOSS_THREADS * LOSS_PER_THREAD;
    for (int it = 0; it < LOSS_PER_THREAD; it++) {
        long long pi = base + (long long)it * LOSS_THREADS + threadIdx.x;
        if (pi >= npx) break;
        bool fg = __ldg(rast + 4 * pi + 3) > 0.f;
        for (int c = 0; c < C; c++) {
            float col = fg ? __ldg(colour + pi * C + c) : bg;
            float e = __ldg(ref + pi * C + c) - 255.f * col;
            acc += (double)loss_term(e, l1);
            d_colour[pi * C + c] = fg ? loss_dcolour(e, k, l1) : 0.f;
            if (comp) comp[pi * C + c] = col;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; w++) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) k_loss_reduce(const double* __restrict__ partial, int n, float k, float* __restrict__ loss)
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

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                              float lr_ramp, float max_iter, const float* __restrict__ step_count, float start_step)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float step0 = __ldg(step_count);              // steps taken so far (LambdaLR epoch)
    if (step0 < start_step) return;               // group still locked (requires_grad False: torch skips it, state untouched)
    float tstep = step0 - start_step + 1.f;       // Adam's t counts the updates of THIS group
    float lr_eff = lr * powf(lr_ramp, step0 / max_iter);
    float bc1 = 1.f - powf(b1, tstep), bc2 = 1.f - powf(b2, tstep);
    float gi = g[i];
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] -= (lr_eff / bc1) * (mi / denom);
}

__global__ void k_adam_advance(float* step_count) { step_count[0] += 1.f; }

__global__ void __launch_bounds__(256) k_quat_renorm(float* __restrict__ q, int n, int mode)
{
    __shared__ float red[256];
    if (mode == 0) {
        int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        // q may sit at any 4-byte offset inside a packed parameter buffer: scalar accesses
        float x = q[4 * i], y = q[4 * i + 1], z = q[4 * i + 2], w = q[4 * i + 3];
        float s = 1.f / sqrtf(x * x + y * y + z * z + w * w);
        q[4 * i] = x * s; q[4 * i + 1] = y * s; q[4 * i + 2] = z * s; q[4 * i + 3] = w * s;
        return;
    }
    // Frobenius: single CTA
    float s = 0.f;
    for (int i = threadIdx.x; i < 4 * n; i += 256) s += q[i] * q[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    float inv = 1.f / sqrtf(red[0]);
    for (int i = threadIdx.x; i < 4 * n; i += 256) q[i] *= inv;
}

}  // namespace

extern "C" size_t fpc_image_loss_scratch_bytes(int N, int H, int W, int C)
{
    (void)C;
    if (N <= 0 || H <= 0 || W <= 0) return 256;
    long long npx = (long long)N * H * W;
    return (size_t)fpc_div_up(npx, LOSS_THREADS * LOSS_PER_THREAD) * sizeof(double) + 256;
}

extern "C" int fpc_image_loss_fwd_bwd(const float* colour, const float* rast, const float* ref, int N, int H, int W, int C,
                                      float bg, float scale, int loss_kind, float* loss, float* d_colour, float* comp,
                                      void* scratch, size_t scratch_bytes, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(colour && rast && ref && loss && d_colour, "image_loss_fwd_bwd: null pointer argument");
    FPC_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0, "image_loss_fwd_bwd: sizes must be positive");
    FPC_CHECK_ARG(scratch && scratch_bytes >= fpc_image_loss_scratch_bytes(N, H, W, C), "image_loss_fwd_bwd: scratch too small");
    long long npx = (long long)N * H * W;
    int nblk = fpc_div_up(npx, LOSS_THREADS * LOSS_PER_THREAD);
    float k = scale / ((float)H * (float)W * (float)C);
    FPC_CHECK_ARG(loss_kind == 0 || loss_kind == 1, "image_loss_fwd_bwd: loss_kind must be 0 (L2) or 1 (L1)");
    k_image_loss<<<nblk, LOSS_THREADS, 0, stream>>>(colour, rast, ref, npx, C, bg, k, loss_kind, d_colour, comp, (double*)scratch);
    FPC_LAUNCH_CHECK();
    k_loss_reduce<<<1, 256, 0, stream>>>((const double*)scratch, nblk, k, loss);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_adam_step_from(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                                  float lr_ramp, float max_iter, const float* step_count, float start_step, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(p && g && m && v && step_count, "adam_step: null pointer argument");
    FPC_CHECK_ARG(n > 0 && max_iter > 0.f && lr_ramp > 0.f && start_step >= 0.f, "adam_step: n, max_iter and lr_ramp must be positive, start_step >= 0");
    k_adam<<<fpc_div_up(n, 256), 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, lr_ramp, max_iter, step_count, start_step);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                             float lr_ramp, float max_iter, const float* step_count, fpc_stream_t stream_)
{
    return fpc_adam_step_from(p, g, m, v, n, lr, b1, b2, eps, lr_ramp, max_iter, step_count, 0.f, stream_);
}

extern "C" int fpc_adam_advance(float* step_count, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(step_count, "adam_advance: null pointer argument");
    k_adam_advance<<<1, 1, 0, stream>>>(step_count);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

extern "C" int fpc_quat_renorm(float* q, int n, int mode, fpc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FPC_CHECK_ARG(q && n > 0, "quat_renorm: q must be non-null and n positive");
    FPC_CHECK_ARG(mode == 0 || mode == 1, "quat_renorm: mode must be 0 (per row) or 1 (Frobenius)");
    if (mode == 0) k_quat_renorm<<<fpc_div_up(n, 256), 256, 0, stream>>>(q, n, 0);
    else k_quat_renorm<<<1, 256, 0, stream>>>(q, n, 1);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
