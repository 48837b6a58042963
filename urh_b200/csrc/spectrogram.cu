// Spectrogram: STFT -> fftshift -> dB (reference: Spectrogram.stft Spectrogram.py:94-116,
// __calculate_spectrogram :156-162, util.arr2decibel util.pyx:38-48).
//
// The reference multiplies complex frames by np.hanning in float64 and runs numpy's complex128 FFT, then casts
// to complex64 and takes 10*log10f(|X|^2) in float32.  To keep the weak bins (down to ~-100 dB below the peak)
// within the stated 1e-3 dB, the FFT here is cuFFT Z2Z (double) — cuFFT is used for the FFT only, as the
// north_star prescribes; windowing, scaling, fftshift, the complex64 cast, the dB map and the left-right flip are
// fused into two hand-written kernels around it.  Frames are processed in batches to bound the working set.
#include "common.cuh"

#include <cufft.h>
#include <math.h>

#define URH_CUFFT(ctx, call)                                                                    \
    do {                                                                                        \
        cufftResult r__ = (call);                                                               \
        if (r__ != CUFFT_SUCCESS) {                                                             \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #call, (int)r__); \
            return URH_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// frames[f][w] = x[(f0+f)*hop + w] * window[w]  (complex128; samples beyond n are zero: Spectrogram.py:102-103)
__global__ void k_stft_window(const float2* __restrict__ x, int64_t n, int W, int hop, const double* __restrict__ window,
                              int64_t f0, int64_t nframes, double2* __restrict__ frames) {
    const int64_t total = nframes * W;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t f = idx / W;
        const int w = (int)(idx - f * W);
        const int64_t i = (f0 + f) * hop + w;
        double2 v = make_double2(0.0, 0.0);
        if (i < n) {
            const float2 s = x[i];
            const double g = window[w];
            v = make_double2((double)s.x * g, (double)s.y * g);
        }
        frames[idx] = v;
    }
}

// out[f][w] = X[f][w] / W   (Spectrogram.stft result, complex128)
__global__ void k_stft_scale(const double2* __restrict__ X, int W, int64_t total, double2* __restrict__ out) {
    const double inv = (double)W;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const double2 v = X[idx];
        out[idx] = make_double2(v.x / inv, v.y / inv);
    }
}

// out[f][j] = dB(shifted[f][W-1-j]), shifted[j] = X[(j + W - W/2... np.fft.fftshift: shifted[j] = X[(j - W/2) mod W] for even/odd W
__global__ void k_stft_db(const double2* __restrict__ X, int W, int64_t nframes, float* __restrict__ out) {
    const int64_t total = nframes * W;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double dW = (double)W;
    const int half = W / 2;  // fftshift moves index n//2.. to the front: shifted[j] = X[(j + (W+1)/2) % W]
    const int shift = (W + 1) / 2;
    (void)half;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t f = idx / W;
        const int j = (int)(idx - f * W);
        const int js = W - 1 - j;                 // fliplr
        const int src = (js + shift) % W;         // fftshift
        const double2 v = X[f * W + src];
        const float re = (float)(v.x / dW), im = (float)(v.y / dW);   // complex128 / W, then astype(complex64)
        out[idx] = __fmul_rn(10.0f, log10f(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im))));
    }
}

static int ensure_plan(urh_ctx* ctx, int W, int64_t batch) {
    if (ctx->fft_valid && ctx->fft_nfft == W && ctx->fft_batch == batch) return URH_OK;
    if (ctx->fft_valid) {
        cufftDestroy((cufftHandle)ctx->fft_plan);
        ctx->fft_valid = false;
    }
    cufftHandle plan;
    int nfft[1] = {W};
    URH_CUFFT(ctx, cufftPlanMany(&plan, 1, nfft, nullptr, 1, W, nullptr, 1, W, CUFFT_Z2Z, (int)batch));
    URH_CUFFT(ctx, cufftSetStream(plan, ctx->stream));
    ctx->fft_plan = (int)plan;
    ctx->fft_nfft = W;
    ctx->fft_batch = batch;
    ctx->fft_valid = true;
    return URH_OK;
}

// mode 0: d_out = complex128 [F][W] stft (Spectrogram.stft);  mode 1: d_out = float32 [F][W] dB map
static int stft_run(urh_ctx* ctx, const float* d_x, int64_t n, int W, int hop, const double* d_window, int64_t num_frames,
                    void* d_out, int mode) {
    if (W <= 0 || hop <= 0 || num_frames <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "stft: bad window/hop/frames");
    urh_arena_reset(ctx);
    const int64_t max_batch = max((int64_t)1, ((int64_t)512 << 20) / ((int64_t)W * 16));
    const int64_t batch = min(num_frames, max_batch);
    double2* buf;
    URH_CHECK(urh_arena(ctx, (size_t)batch * W, &buf));
    const unsigned grid = (unsigned)min(urh_div_up(batch * W, 256), (int64_t)ctx->sm_count * 32);
    for (int64_t f0 = 0; f0 < num_frames; f0 += batch) {
        const int64_t nf = min(batch, num_frames - f0);
        URH_CHECK(ensure_plan(ctx, W, nf));
        URH_LAUNCH(ctx, k_stft_window, grid, 256, 0, (const float2*)d_x, n, W, hop, d_window, f0, nf, buf);
        URH_CUFFT(ctx, cufftExecZ2Z((cufftHandle)ctx->fft_plan, (cufftDoubleComplex*)buf, (cufftDoubleComplex*)buf, CUFFT_FORWARD));
        ctx->launches++;
        if (mode == 0) URH_LAUNCH(ctx, k_stft_scale, grid, 256, 0, buf, W, nf * W, (double2*)d_out + f0 * W);
        else URH_LAUNCH(ctx, k_stft_db, grid, 256, 0, buf, W, nf, (float*)d_out + f0 * W);
    }
    return URH_OK;
}

extern "C" int urh_stft(urh_ctx* ctx, const float* d_x, int64_t n, int window_size, int hop, const double* d_window,
                        int64_t num_frames, double* d_out) {
    return stft_run(ctx, d_x, n, window_size, hop, d_window, num_frames, d_out, 0);
}

extern "C" int urh_spectrogram_db(urh_ctx* ctx, const float* d_x, int64_t n, int window_size, int hop, const double* d_window,
                                  int64_t num_frames, float* d_out) {
    return stft_run(ctx, d_x, n, window_size, hop, d_window, num_frames, d_out, 1);
}
