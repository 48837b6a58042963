// Spectrogram: STFT -> fftshift -> dB (reference: Spectrogram.stft Spectrogram.py:94-116,
// __calculate_spectrogram :156-162, util.arr2decibel util.pyx:38-48).
//
// The reference multiplies complex frames by np.hanning in float64 and runs numpy's complex128 FFT, then casts
// to complex64 and takes 10*log10f(|X|^2) in float32.  To keep the weak bins (down to ~-100 dB below the peak)
// within the stated 1e-3 dB, the FFT is done in double.  Power-of-two windows (URH's: 1024) take ONE fused kernel (window ->
// shared-memory FFT -> scale / fftshift / cast / dB / flip, see k_stft_fused); other sizes use cuFFT Z2Z — for the FFT only, as
// the north_star prescribes — between two hand-written kernels, in batches that bound the working set.
#include "common.cuh"

#include <cufft.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>

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

// ---- fused path (power-of-two windows): window -> FFT in shared memory -> scale / fftshift / dB, one block per frame ----------
// The cuFFT path above moves every frame through HBM three times as complex128 (window kernel -> Z2Z -> dB kernel: 36 GB for
// 2^28 samples at W = 1024, hop = 512).  Here a frame is read once as complex64 (the 50 % overlap with its neighbour comes from
// L2), transformed in double in shared memory (Stockham autosort, radix-4 stages + one radix-2 stage when log2 W is odd) and
// written once as float32 dB (or complex128 for urh_stft): 8 + 8 B/sample at the reference's parameters.
// tw[q] = exp(-2 pi i q / W), q < W, built once per window size with sincospi (double).
__global__ void k_fft_twiddles(int W, double2* __restrict__ tw) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= W) return;
    double s, c;
    sincospi(-2.0 * (double)q / (double)W, &s, &c);
    tw[q] = make_double2(c, s);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// MODE 0: out = complex128 [F][W] = X / W;  MODE 1: out = float32 [F][W] dB map (fftshift + fliplr + complex64 cast + 10 log10f)
// W = 2^LOG2W and the thread count are compile-time: every loop below is fully unrolled (the W / THREADS loads of a thread are
// in flight together — the first version, with run-time W, was bound by the latency of one load after the other — and the index
// arithmetic of the stages is shifts and masks).
template <int LOG2W, int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS) k_stft_fused(const float2* __restrict__ x, int64_t n, int hop,
                                                       const double* __restrict__ window, const double2* __restrict__ tw,
                                                       int64_t nframes, void* __restrict__ out_) {
    constexpr int W = 1 << LOG2W;
    constexpr int PER = W / THREADS;          // elements per thread in the load / store phases
    constexpr int BPT = (W / 4) / THREADS > 0 ? (W / 4) / THREADS : 1;   // radix-4 butterflies per thread and stage
    extern __shared__ double2 s_buf[];        // two W-element buffers
    double2* a = s_buf;
    double2* b = s_buf + W;
    const int64_t f = blockIdx.x;
    const int64_t base = f * hop;
    {
        float2 sm[PER];
        double g[PER];
#pragma unroll
        for (int r = 0; r < PER; r++) {
            const int w = threadIdx.x + r * THREADS;
            const int64_t i = base + w;
            sm[r] = (i < n) ? x[i] : make_float2(0.0f, 0.0f);
            g[r] = window[w];
        }
#pragma unroll
        for (int r = 0; r < PER; r++) a[threadIdx.x + r * THREADS] = make_double2((double)sm[r].x * g[r], (double)sm[r].y * g[r]);
    }
    __syncthreads();
    if (LOG2W & 1) {   // one radix-2 stage first
#pragma unroll
        for (int r = 0; r < (W / 2) / THREADS; r++) {
            const int j = threadIdx.x + r * THREADS;
            const double2 u = a[j], v = a[j + W / 2];
            b[2 * j] = make_double2(u.x + v.x, u.y + v.y);
            b[2 * j + 1] = make_double2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
        double2* t = a; a = b; b = t;
    }
#pragma unroll
    for (int st = 0; st < LOG2W / 2; st++) {
        const int ns = 1 << (2 * st + (LOG2W & 1));   // length of the sub-transforms finished so far (compile-time after unrolling)
        constexpr int quarter = W / 4;
        const int tstep = W / (4 * ns);               // exp(-2 pi i r k / (4 ns)) = tw[r * k * tstep]
#pragma unroll
        for (int r = 0; r < BPT; r++) {
            const int j = threadIdx.x + r * THREADS;
            if (j < quarter) {
                const int k = j & (ns - 1);
                double2 v0 = a[j], v1 = a[j + quarter], v2 = a[j + 2 * quarter], v3 = a[j + 3 * quarter];
                if (ns > 1) {
                    v1 = cmul(v1, tw[k * tstep]);
                    v2 = cmul(v2, tw[2 * k * tstep]);
                    v3 = cmul(v3, tw[3 * k * tstep]);
                }
                // DFT of length 4 (forward: -i rotation)
                const double2 s02 = make_double2(v0.x + v2.x, v0.y + v2.y), d02 = make_double2(v0.x - v2.x, v0.y - v2.y);
                const double2 s13 = make_double2(v1.x + v3.x, v1.y + v3.y), d13 = make_double2(v1.x - v3.x, v1.y - v3.y);
                const int j0 = ((j - k) << 2) + k;   // (j / ns) * 4 ns + k
                b[j0] = make_double2(s02.x + s13.x, s02.y + s13.y);
                b[j0 + ns] = make_double2(d02.x + d13.y, d02.y - d13.x);      // d02 - i d13
                b[j0 + 2 * ns] = make_double2(s02.x - s13.x, s02.y - s13.y);
                b[j0 + 3 * ns] = make_double2(d02.x - d13.y, d02.y + d13.x);  // d02 + i d13
            }
        }
        __syncthreads();
        double2* t = a; a = b; b = t;
    }
    const double inv = 1.0 / (double)W;   // W is a power of two: multiplying by 1/W IS the division by W, bit for bit
    if (MODE == 0) {
        double2* out = (double2*)out_ + f * W;
#pragma unroll
        for (int r = 0; r < PER; r++) {
            const int w = threadIdx.x + r * THREADS;
            out[w] = make_double2(a[w].x * inv, a[w].y * inv);
        }
    } else {
        float* out = (float*)out_ + f * W;
        constexpr int shift = (W + 1) / 2;
#pragma unroll
        for (int r = 0; r < PER; r++) {
            const int j = threadIdx.x + r * THREADS;
            const int src = ((W - 1 - j) + shift) & (W - 1);   // fliplr, then fftshift
            const double2 v = a[src];
            const float re = (float)(v.x * inv), im = (float)(v.y * inv);   // complex128 / W, then astype(complex64)
            out[j] = __fmul_rn(10.0f, log10f(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im))));
        }
    }
}

// ---- radix-16 variant (W = 256, 1024, 4096): two radix-4 levels per pass held in registers, so a frame crosses shared memory
// three times (1024 = 16 * 16 * 4) instead of five: the radix-4 kernel above is bound by shared-memory bandwidth.
// One thread owns 16 points of a pass; W / 16 threads per frame.  Both buffers are padded by one element every 16 (P(i)) so that
// the stride-16 stores of the first pass do not pile onto the same banks.
__device__ __forceinline__ int stft_pad(int i) { return i + (i >> 4); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)
// forward DFT of length 4, in place
__device__ __forceinline__ void dft4(double2& x0, double2& x1, double2& x2, double2& x3) {
    const double2 s02 = cadd(x0, x2), d02 = csub(x0, x2), s13 = cadd(x1, x3), d13 = cmul_mi(csub(x1, x3));
    x0 = cadd(s02, s13); x1 = cadd(d02, d13); x2 = csub(s02, s13); x3 = csub(d02, d13);
}
// forward DFT of length 16, in place: v[c + 4 r'] -> columns, twiddle w16^(c r), rows; output y[r + 4 s] (natural order)
__device__ __forceinline__ void dft16(double2 (&v)[16]) {
    const double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173, H = 0.70710678118654752440;
#pragma unroll
    for (int c = 0; c < 4; c++) dft4(v[c], v[c + 4], v[c + 8], v[c + 12]);   // now v[c + 4 r] = u_c[r]
    // u_c[r] *= w16^(c r)
    v[1 + 4] = cmul(v[1 + 4], make_double2(C1, -S1));    // c=1, r=1: w^1
    v[1 + 8] = cmul(v[1 + 8], make_double2(H, -H));      // w^2
    v[1 + 12] = cmul(v[1 + 12], make_double2(S1, -C1));  // w^3
    v[2 + 4] = cmul(v[2 + 4], make_double2(H, -H));      // c=2, r=1: w^2
    v[2 + 8] = cmul_mi(v[2 + 8]);                        // w^4 = -i
    v[2 + 12] = cmul(v[2 + 12], make_double2(-H, -H));   // w^6
    v[3 + 4] = cmul(v[3 + 4], make_double2(S1, -C1));    // c=3, r=1: w^3
    v[3 + 8] = cmul(v[3 + 8], make_double2(-H, -H));     // w^6
    v[3 + 12] = cmul(v[3 + 12], make_double2(-C1, S1));  // w^9
    // rows: y[r + 4 s] = sum_c u_c[r] w4^(c s): DFT4 over c for each r; u_c[r] sits at v[c + 4 r]
#pragma unroll
    for (int r = 0; r < 4; r++) dft4(v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]);   // v[4 r + s] = y[r + 4 s]
}

template <int LOG2W, int MODE>
__global__ void __launch_bounds__((1 << LOG2W) / 16) k_stft_r16(const float2* __restrict__ x, int64_t n, int hop,
                                                               const double* __restrict__ window, const double2* __restrict__ tw,
                                                               int64_t nframes, void* __restrict__ out_) {
    constexpr int W = 1 << LOG2W;
    constexpr int T = W / 16;                 // threads per frame
    constexpr int PADW = W + W / 16;
    extern __shared__ double2 s_buf[];        // two padded buffers
    double2* a = s_buf;
    double2* b = s_buf + PADW;
    const int tid = threadIdx.x;
    const int64_t f = blockIdx.x;
    const int64_t base = f * hop;
    {
        float2 sm[16];
        double g[16];
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int w = tid + T * m;
            const int64_t i = base + w;
            sm[m] = (i < n) ? x[i] : make_float2(0.0f, 0.0f);
            g[m] = window[w];
        }
#pragma unroll
        for (int m = 0; m < 16; m++) a[stft_pad(tid + T * m)] = make_double2((double)sm[m].x * g[m], (double)sm[m].y * g[m]);
    }
    __syncthreads();
    constexpr int NPASS16 = LOG2W / 4;
#pragma unroll
    for (int st = 0; st < NPASS16; st++) {
        const int ns = 1 << (4 * st);
        const int k = tid & (ns - 1);
        const int tstep = W / (16 * ns);
        double2 v[16];
#pragma unroll
        for (int m = 0; m < 16; m++) v[m] = a[stft_pad(tid + T * m)];
        if (ns > 1) {
#pragma unroll
            for (int m = 1; m < 16; m++) v[m] = cmul(v[m], tw[m * k * tstep]);
        }
        dft16(v);   // v[4 r + s] = y[r + 4 s]
        const int o = ((tid - k) << 4) + k;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) b[stft_pad(o + ns * (r + 4 * q))] = v[4 * r + q];
        __syncthreads();
        double2* t = a; a = b; b = t;
    }
    if ((LOG2W & 3) == 2) {   // one radix-4 pass left (sub-transforms of length W / 4)
        constexpr int ns = W / 4;
        double2 y[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int j = tid + T * r;   // butterfly index < W / 4; k = j (ns = W / 4 > j)
            double2 v0 = a[stft_pad(j)], v1 = a[stft_pad(j + ns)], v2 = a[stft_pad(j + 2 * ns)], v3 = a[stft_pad(j + 3 * ns)];
            v1 = cmul(v1, tw[j]);
            v2 = cmul(v2, tw[2 * j]);
            v3 = cmul(v3, tw[3 * j]);
            dft4(v0, v1, v2, v3);
            y[r][0] = v0; y[r][1] = v1; y[r][2] = v2; y[r][3] = v3;
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) b[stft_pad(tid + T * r + ns * q)] = y[r][q];
        __syncthreads();
        double2* t = a; a = b; b = t;
    }
    const double inv = 1.0 / (double)W;
    if (MODE == 0) {
        double2* out = (double2*)out_ + f * W;
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int w = tid + T * m;
            const double2 v = a[stft_pad(w)];
            out[w] = make_double2(v.x * inv, v.y * inv);
        }
    } else {
        float* out = (float*)out_ + f * W;
        constexpr int shift = (W + 1) / 2;
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int j = tid + T * m;
            const int src = ((W - 1 - j) + shift) & (W - 1);   // fliplr, then fftshift
            const double2 v = a[stft_pad(src)];
            const float re = (float)(v.x * inv), im = (float)(v.y * inv);
            out[j] = __fmul_rn(10.0f, log10f(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im))));
        }
    }
}

template <int LOG2W, int MODE>
static int stft_r16_launch(urh_ctx* ctx, const float* d_x, int64_t n, int hop, const double* d_window, const double2* tw,
                           int64_t num_frames, void* d_out) {
    constexpr int W = 1 << LOG2W;
    const size_t smem = (size_t)2 * (W + W / 16) * sizeof(double2);
    if (smem > 48 * 1024)
        URH_CUDA(ctx, cudaFuncSetAttribute(k_stft_r16<LOG2W, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    URH_LAUNCH(ctx, (k_stft_r16<LOG2W, MODE>), (unsigned)num_frames, W / 16, smem, (const float2*)d_x, n, hop, d_window, tw, num_frames,
               d_out);
    return URH_OK;
}

template <int LOG2W, int MODE>
static int stft_fused_launch(urh_ctx* ctx, const float* d_x, int64_t n, int hop, const double* d_window, const double2* tw,
                             int64_t num_frames, void* d_out) {
    constexpr int W = 1 << LOG2W;
    constexpr int THREADS = (W / 4 >= 256) ? 256 : (W / 4 >= 32 ? W / 4 : 32);
    const size_t smem = (size_t)2 * W * sizeof(double2);
    if (smem > 48 * 1024)
        URH_CUDA(ctx, cudaFuncSetAttribute(k_stft_fused<LOG2W, THREADS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    URH_LAUNCH(ctx, (k_stft_fused<LOG2W, THREADS, MODE>), (unsigned)num_frames, THREADS, smem, (const float2*)d_x, n, hop, d_window, tw,
               num_frames, d_out);
    return URH_OK;
}

static int stft_fused(urh_ctx* ctx, const float* d_x, int64_t n, int W, int hop, const double* d_window, int64_t num_frames,
                      void* d_out, int mode) {
    int log2w = 0;
    while ((1 << log2w) < W) log2w++;
    urh_arena_reset(ctx);
    double2* tw;
    URH_CHECK(urh_arena(ctx, (size_t)W, &tw));
    URH_LAUNCH(ctx, k_fft_twiddles, (unsigned)urh_div_up(W, 256), 256, 0, W, tw);
    // grid.x is limited to 2^31 - 1 frames: far beyond any capture that fits the device
#define STFT_CASE(L)                                                                                                          \
    case L:                                                                                                                   \
        return mode == 0 ? stft_fused_launch<L, 0>(ctx, d_x, n, hop, d_window, (const double2*)tw, num_frames, d_out)         \
                         : stft_fused_launch<L, 1>(ctx, d_x, n, hop, d_window, (const double2*)tw, num_frames, d_out);
    if (!getenv("URH_B200_STFT_RADIX4")) {   // 16 | W: three passes through shared memory instead of five
        if (log2w == 10) return mode == 0 ? stft_r16_launch<10, 0>(ctx, d_x, n, hop, d_window, (const double2*)tw, num_frames, d_out)
                                          : stft_r16_launch<10, 1>(ctx, d_x, n, hop, d_window, (const double2*)tw, num_frames, d_out);
        if (log2w == 12) return mode == 0 ? stft_r16_launch<12, 0>(ctx, d_x, n, hop, d_window, (const double2*)tw, num_frames, d_out)
                                          : stft_r16_launch<12, 1>(ctx, d_x, n, hop, d_window, (const double2*)tw, num_frames, d_out);
        if (log2w == 8) return mode == 0 ? stft_r16_launch<8, 0>(ctx, d_x, n, hop, d_window, (const double2*)tw, num_frames, d_out)
                                         : stft_r16_launch<8, 1>(ctx, d_x, n, hop, d_window, (const double2*)tw, num_frames, d_out);
    }
    switch (log2w) {
        STFT_CASE(7) STFT_CASE(8) STFT_CASE(9) STFT_CASE(10) STFT_CASE(11) STFT_CASE(12)
        default: break;
    }
#undef STFT_CASE
    URH_FAIL(ctx, URH_ERR_INVALID, "stft_fused: unsupported window size");
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
    // power-of-two windows 128 .. 4096 (128 KB of shared memory): the fused kernel; anything else: cuFFT with two kernels around it
    if ((W & (W - 1)) == 0 && W >= 128 && W <= 4096 && num_frames < ((int64_t)1 << 31) && !getenv("URH_B200_STFT_CUFFT"))
        return stft_fused(ctx, d_x, n, W, hop, d_window, num_frames, d_out, mode);
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

// ---- BGRA colormap look-up (Spectrogram.apply_bgra_lookup, Spectrogram.py:192-206) --------------------------------
// out[c][r] = colormap[clip(int((L - 1) * ((data[r][c] - data_min) / (data_max - data_min))))]   (data.T: the image is transposed)
// float32 arithmetic in numpy's order: subtract, divide, multiply, truncate toward zero (astype(int)), np.take(mode="clip").
__global__ void k_bgra_lookup(const float* __restrict__ data, int64_t rows, int64_t cols, const uint32_t* __restrict__ colormap, int L,
                              float data_min, float range, int normalize, uint32_t* __restrict__ out) {
    __shared__ uint32_t s_map[1024];
    const bool in_smem = L <= 1024;
    if (in_smem)
        for (int i = threadIdx.x; i < L; i += blockDim.x) s_map[i] = colormap[i];
    __syncthreads();
    const int64_t total = rows * cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float scale = (float)(L - 1);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t c = idx / rows, r = idx - c * rows;   // output index (c, r) <- data[r][c]
        float v = data[r * cols + c];
        if (normalize) v = __fmul_rn(scale, __fdiv_rn(__fsub_rn(v, data_min), range));
        // astype(int): truncation; NaN and out-of-range values become INT64_MIN in numpy, which mode="clip" maps to entry 0
        long long k = (v == v && fabsf(v) < 9.0e18f) ? (long long)v : LLONG_MIN;
        k = k < 0 ? 0 : (k > L - 1 ? L - 1 : k);
        out[idx] = in_smem ? s_map[k] : colormap[k];
    }
}

extern "C" int urh_bgra_lookup(urh_ctx* ctx, const float* d_data, int64_t rows, int64_t cols, const uint8_t* d_colormap, int entries,
                               float data_min, float data_max, int normalize, uint8_t* d_out) {
    if (rows <= 0 || cols <= 0) return URH_OK;
    if (entries <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "bgra_lookup: empty colormap");
    // the reference forms (data_max - data_min) in Python floats; as a weak scalar it meets the float32 array as a float32
    const float r32 = (float)((double)data_max - (double)data_min);
    const unsigned grid = (unsigned)min(urh_div_up(rows * cols, 256), (int64_t)ctx->sm_count * 16);
    URH_LAUNCH(ctx, k_bgra_lookup, grid, 256, 0, d_data, rows, cols, (const uint32_t*)d_colormap, entries, data_min, r32, normalize,
               (uint32_t*)d_out);
    return URH_OK;
}
