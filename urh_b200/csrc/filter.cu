// FIR filtering and DC correction (reference: signal_functions.fir_filter signal_functions.pyx:513-525,
// Filter.work / apply_fir_filter / apply_bandpass_filter Filter.py:31-46, 84-101).
//
// fir_filter is reproduced in the reference's exact accumulation order: y[k] = sum over i ascending of
// x[i]*taps[k-i], every complex64 product and every add individually rounded (no FMA) — bit-identical output.
// A block stages its input tile (outputs + M-1 halo samples) in shared memory with cp.async.bulk (TMA 1-D bulk
// copy, mbarrier completion); each thread then produces 4 consecutive outputs so that every tap fetched from
// shared memory is used four times.  This kernel is FP32-ALU-bound (8*M unfused flops per sample), not
// HBM-bound: SURVEY §8d reports FP32 utilisation for it.
//
// The band-pass path (complex128 taps, numpy 'same'/FFT convolution in the reference) is evaluated as a
// direct convolution with double accumulation — the reference's own result is a complex128 FFT product, so
// parity is tolerance-based (1e-5 of the signal scale) as stated in BASELINE.md.
#include "common.cuh"

#include <cuda/barrier>
#include <math.h>

#define FIR_THREADS 256
#define FIR_PER_THREAD 4
#define FIR_TILE (FIR_THREADS * FIR_PER_THREAD)

__device__ __forceinline__ void fir_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void fir_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes));
}
__device__ __forceinline__ void fir_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void fir_mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(phase));
}

// Exact-order complex64 FIR.  smem: [taps M float2][tile FIR_TILE + M - 1 float2]
__global__ void __launch_bounds__(FIR_THREADS) k_fir_exact(const float2* __restrict__ x, int64_t n, const float2* __restrict__ taps,
                                                            int m, float2* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    float2* s_taps = (float2*)smem_raw;
    float2* s_x = s_taps + ((m + 1) & ~1);  // keep 16-byte alignment of the tile
    const int64_t tile0 = (int64_t)blockIdx.x * FIR_TILE;
    const int64_t first = tile0 - (m - 1);          // first input sample the tile needs (may be < 0)
    const int64_t lo = first < 0 ? 0 : first;
    const int64_t hi = min(tile0 + FIR_TILE, n);    // one past the last input sample
    const int halo_missing = (int)(lo - first);     // zero initial state: samples before the capture are 0
    for (int j = threadIdx.x; j < m; j += FIR_THREADS) s_taps[j] = taps[j];
    for (int j = threadIdx.x; j < halo_missing; j += FIR_THREADS) s_x[j] = make_float2(0.f, 0.f);
    // bulk-copy [lo, hi) into s_x + halo_missing: needs 16-byte aligned addresses and size
    const int64_t cnt = hi - lo;
    const bool bulk_ok = (((uintptr_t)(x + lo)) % 16 == 0) && ((halo_missing % 2) == 0) && (cnt % 2 == 0) && cnt > 0;
    if (threadIdx.x == 0) {
        fir_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (bulk_ok) {
        if (threadIdx.x == 0) {
            fir_mbar_expect_tx(&bar, (uint32_t)(cnt * sizeof(float2)));
            fir_bulk_g2s(s_x + halo_missing, x + lo, (uint32_t)(cnt * sizeof(float2)), &bar);
        }
        fir_mbar_wait(&bar, 0);
    } else {
        for (int64_t j = threadIdx.x; j < cnt; j += FIR_THREADS) s_x[halo_missing + j] = x[lo + j];
    }
    __syncthreads();
    // thread t -> outputs k = tile0 + 4t .. 4t+3 ; s_x[j] holds x[first + j], so x[k - q] = s_x[k - q - first]
    const int64_t k0 = tile0 + (int64_t)threadIdx.x * FIR_PER_THREAD;
    if (k0 >= n) return;
    float2 acc[FIR_PER_THREAD];
#pragma unroll
    for (int r = 0; r < FIR_PER_THREAD; r++) acc[r] = make_float2(0.f, 0.f);
    // ascending input index i  <=>  descending tap index q = k - i, from q = m-1 down to 0
    const int base = (int)(k0 - first);  // s_x index of x[k0]
    for (int q = m - 1; q >= 0; q--) {
        const float2 h = s_taps[q];
#pragma unroll
        for (int r = 0; r < FIR_PER_THREAD; r++) {
            const float2 v = s_x[base + r - q];
            // complex64 product then += , each operation rounded (GCC's std::complex<float> without fast-math)
            const float pr = __fsub_rn(__fmul_rn(v.x, h.x), __fmul_rn(v.y, h.y));
            const float pi = __fadd_rn(__fmul_rn(v.x, h.y), __fmul_rn(v.y, h.x));
            acc[r].x = __fadd_rn(acc[r].x, pr);
            acc[r].y = __fadd_rn(acc[r].y, pi);
        }
    }
#pragma unroll
    for (int r = 0; r < FIR_PER_THREAD; r++)
        if (k0 + r < n) y[k0 + r] = acc[r];
}

// replaces signal_functions.fir_filter: x, y complex64[n] (device), taps complex64[m] (device)
extern "C" int urh_fir_filter(urh_ctx* ctx, const float* d_x, int64_t n, const float* d_taps, int m, float* d_y) {
    if (n <= 0) return URH_OK;
    if (m <= 0) {
        URH_CUDA(ctx, cudaMemsetAsync(d_y, 0, (size_t)n * 8, ctx->stream));
        return URH_OK;
    }
    const size_t smem = (size_t)(((m + 1) & ~1) + FIR_TILE + m - 1 + 2) * sizeof(float2);
    if (smem > 200 * 1024) URH_FAIL(ctx, URH_ERR_INVALID, "fir_filter: %d taps exceed the shared-memory tile (max ~11000)", m);
    URH_CUDA(ctx, cudaFuncSetAttribute(k_fir_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // NOTE on the reference's zero products: for the first m-1 outputs the reference simply has fewer terms; adding
    // the products of zero-padded samples (0*h = +-0) to a non-zero accumulator changes nothing, and the very first
    // term of every output is x[0]*h (k < m) or a real sample, so the sums are bit-identical except for the sign of
    // an all-zero result, which the reference (np.zeros start) also produces as +0 -> handled by starting at +0.
    URH_LAUNCH(ctx, k_fir_exact, (unsigned)urh_div_up(n, FIR_TILE), FIR_THREADS, smem, (const float2*)d_x, n, (const float2*)d_taps, m,
               (float2*)d_y);
    return URH_OK;
}

// Direct convolution sample c[t + offset], c = full convolution of x (complex64) with h (complex128 taps),
// double accumulation, complex64 result (band-pass path).
__global__ void k_conv_c128(const float2* __restrict__ x, int64_t n, const double2* __restrict__ h, int m, int64_t offset,
                            int64_t out_len, float2* __restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < out_len; k += stride) {
        const int64_t t = k + offset;
        double re = 0.0, im = 0.0;
        const int jlo = (int)max((int64_t)0, t - (n - 1)), jhi = (int)min((int64_t)m - 1, t);
        for (int j = jlo; j <= jhi; j++) {
            const float2 v = x[t - j];
            const double2 c = h[j];
            re += (double)v.x * c.x - (double)v.y * c.y;
            im += (double)v.x * c.y + (double)v.y * c.x;
        }
        y[k] = make_float2((float)re, (float)im);
    }
}

// The same convolution, tiled: the block converts its input span to double ONCE into shared memory (the naive kernel
// converts every sample once per tap -- the float->double converter, not the FMA pipe, was its limit), the taps sit in
// shared memory too, and every thread produces CONV_PER consecutive outputs from a sliding register window: one 16-byte
// shared-memory read feeds 4 * CONV_PER double FMAs.  CONV_PER is odd so that the lanes' 16-byte reads (stride CONV_PER * 16 B)
// fall into distinct banks.  Same accumulation order per output as k_conv_c128 (ascending tap index, fused multiply-adds).
#define CONV_PER 5
#define CONV_THREADS 256
#define CONV_TILE (CONV_PER * CONV_THREADS)
#define CONV_MAX_TAPS 768
__global__ void __launch_bounds__(CONV_THREADS) k_conv_c128_tiled(const float2* __restrict__ x, int64_t n, const double2* __restrict__ h, int m,
                                                                 int64_t offset, int64_t out_len, float2* __restrict__ y) {
    extern __shared__ double2 s_conv[];
    double2* s_h = s_conv;            // [m]
    double2* s_x = s_conv + m;        // [CONV_TILE + m - 1]: s_x[i] = x[first + i], zero outside the array
    for (int j = threadIdx.x; j < m; j += CONV_THREADS) s_h[j] = h[j];
    const int span = CONV_TILE + m - 1;
    for (int64_t k0 = (int64_t)blockIdx.x * CONV_TILE; k0 < out_len; k0 += (int64_t)gridDim.x * CONV_TILE) {
        const int64_t first = k0 + offset - (m - 1);   // input index of s_x[0]
        __syncthreads();                                // the previous tile's readers are done
        for (int i = threadIdx.x; i < span; i += CONV_THREADS) {
            const int64_t g = first + i;
            double2 v = make_double2(0.0, 0.0);
            if (g >= 0 && g < n) { const float2 f = x[g]; v = make_double2((double)f.x, (double)f.y); }
            s_x[i] = v;
        }
        __syncthreads();
        // output o (tile-relative) at tap j reads input index (k0 + o + offset) - j = first + (o + m - 1 - j)
        const int o0 = threadIdx.x * CONV_PER;
        double2 w[CONV_PER];   // w[i] = s_x[o0 + i + m - 1 - j]
#pragma unroll
        for (int i = 0; i < CONV_PER; i++) w[i] = s_x[o0 + i + m - 1];
        double re[CONV_PER], im[CONV_PER];
#pragma unroll
        for (int i = 0; i < CONV_PER; i++) { re[i] = 0.0; im[i] = 0.0; }
        for (int j = 0; j < m; j++) {
            const double2 c = s_h[j];
#pragma unroll
            for (int i = 0; i < CONV_PER; i++) {
                re[i] = fma(w[i].x, c.x, re[i]);
                re[i] = fma(-w[i].y, c.y, re[i]);
                im[i] = fma(w[i].x, c.y, im[i]);
                im[i] = fma(w[i].y, c.x, im[i]);
            }
            // slide the window one input sample down
#pragma unroll
            for (int i = CONV_PER - 1; i > 0; i--) w[i] = w[i - 1];
            if (j + 1 < m) w[0] = s_x[o0 + m - 2 - j];
        }
#pragma unroll
        for (int i = 0; i < CONV_PER; i++) {
            const int64_t k = k0 + o0 + i;
            if (k < out_len) y[k] = make_float2((float)re[i], (float)im[i]);
        }
    }
}

extern "C" int urh_convolve_c128(urh_ctx* ctx, const float* d_x, int64_t n, const double* d_taps, int m, int64_t offset,
                                 int64_t out_len, float* d_y) {
    if (out_len <= 0) return URH_OK;
    if (m >= 1 && m <= CONV_MAX_TAPS) {
        const size_t smem = (size_t)(m + CONV_TILE + m - 1) * sizeof(double2);   // <= 44.5 KB for m <= 768
        const unsigned grid = (unsigned)min(urh_div_up(out_len, CONV_TILE), (int64_t)ctx->sm_count * 8);
        URH_LAUNCH(ctx, k_conv_c128_tiled, grid, CONV_THREADS, smem, (const float2*)d_x, n, (const double2*)d_taps, m, offset, out_len, (float2*)d_y);
        return URH_OK;
    }
    const unsigned grid = (unsigned)min(urh_div_up(out_len, 256), (int64_t)ctx->sm_count * 32);
    URH_LAUNCH(ctx, k_conv_c128, grid, 256, 0, (const float2*)d_x, n, (const double2*)d_taps, m, offset, out_len, (float2*)d_y);
    return URH_OK;
}

// ---- DC correction: x - mean(x, axis=0) (Filter.py:32-33) --------------------------------------------------------------
// numpy's np.mean over axis 0 of a C-contiguous float32 (n,2) array accumulates each column naively in float32 in
// row order.  exact != 0 reproduces that serial chain (one lane per column, the warp streams the data
// through shared memory); exact == 0 uses a double reduction (accurate, NOT what the reference computes for large n).
__global__ void __launch_bounds__(32) k_dc_mean_serial(const float2* __restrict__ x, int64_t n, float* __restrict__ mean) {
    __shared__ float2 buf[2][1024];
    const int lane = threadIdx.x;
    float acc = 0.0f;  // lane 0: I column, lane 1: Q column
    const int64_t nchunks = (n + 1023) / 1024;
    for (int j = lane; j < 1024; j += 32) buf[0][j] = (j < n) ? x[j] : make_float2(0.f, 0.f);
    __syncwarp();
    for (int64_t c = 0; c < nchunks; c++) {
        const int b = (int)(c & 1);
        const int len = (int)min((int64_t)1024, n - c * 1024);
        if (lane < 2) {
            const float* col = (const float*)buf[b] + lane;
            for (int j = 0; j < len; j++) acc = __fadd_rn(acc, col[2 * j]);
        } else if (c + 1 < nchunks) {
            for (int j = lane - 2; j < 1024; j += 30) {
                const int64_t i = (c + 1) * 1024 + j;
                buf[b ^ 1][j] = (i < n) ? x[i] : make_float2(0.f, 0.f);
            }
        }
        __syncwarp();
    }
    if (lane < 2) mean[lane] = __fdiv_rn(acc, (float)n);  // np.mean: sum / count in float32
}

__global__ void k_dc_mean_partial(const float2* __restrict__ x, int64_t n, double* __restrict__ part) {
    double sr = 0.0, si = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 v = x[i];
        sr += v.x; si += v.y;
    }
    __shared__ double s_r[256], s_i[256];
    s_r[threadIdx.x] = sr; s_i[threadIdx.x] = si;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) { s_r[threadIdx.x] += s_r[threadIdx.x + off]; s_i[threadIdx.x] += s_i[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s_r[0]; part[2 * blockIdx.x + 1] = s_i[0]; }
}
__global__ void k_dc_mean_fold(const double* __restrict__ part, int nblocks, int64_t n, float* __restrict__ mean) {
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int b = 0; b < nblocks; b++) s += part[2 * b + threadIdx.x];
        mean[threadIdx.x] = (float)(s / (double)n);
    }
}
__global__ void k_dc_subtract(const float2* __restrict__ x, int64_t n, const float* __restrict__ mean, float2* __restrict__ y) {
    const float mr = mean[0], mi = mean[1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 v = x[i];
        y[i] = make_float2(__fsub_rn(v.x, mr), __fsub_rn(v.y, mi));
    }
}

extern "C" int urh_dc_correction(urh_ctx* ctx, const float* d_iq, int64_t n, float* d_out, int exact_order) {
    if (n <= 0) return URH_OK;
    urh_arena_reset(ctx);
    float* mean;
    URH_CHECK(urh_arena(ctx, 4, &mean));
    if (exact_order) {
        URH_LAUNCH(ctx, k_dc_mean_serial, 1, 32, 0, (const float2*)d_iq, n, mean);
    } else {
        const int nb = ctx->sm_count * 4;
        double* part;
        URH_CHECK(urh_arena(ctx, (size_t)nb * 2, &part));
        URH_LAUNCH(ctx, k_dc_mean_partial, nb, 256, 0, (const float2*)d_iq, n, part);
        URH_LAUNCH(ctx, k_dc_mean_fold, 1, 32, 0, part, nb, n, mean);
    }
    const unsigned grid = (unsigned)min(urh_div_up(n, 256), (int64_t)ctx->sm_count * 16);
    URH_LAUNCH(ctx, k_dc_subtract, grid, 256, 0, (const float2*)d_iq, n, mean, (float2*)d_out);
    return URH_OK;
}

// ---- DC correction of an INTEGER capture: numpy promotes `x - np.mean(x, axis=0)` to float64; the column sums of integers
// are exact (int64 here, float64 pairwise in numpy: both exact below 2^53), mean = sum / n in double, result double[n][2].
template <typename T>
__global__ void __launch_bounds__(256) k_dc_int_partial(const T* __restrict__ x, int64_t n, long long* __restrict__ part) {
    long long sr = 0, si = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        sr += (long long)x[2 * i];
        si += (long long)x[2 * i + 1];
    }
    __shared__ long long s_r[256], s_i[256];
    s_r[threadIdx.x] = sr; s_i[threadIdx.x] = si;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) { s_r[threadIdx.x] += s_r[threadIdx.x + off]; s_i[threadIdx.x] += s_i[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s_r[0]; part[2 * blockIdx.x + 1] = s_i[0]; }
}
__global__ void k_dc_int_fold(const long long* __restrict__ part, int nblocks, int64_t n, double* __restrict__ mean) {
    if (threadIdx.x < 2) {
        long long s = 0;
        for (int b = 0; b < nblocks; b++) s += part[2 * b + threadIdx.x];
        mean[threadIdx.x] = __ddiv_rn((double)s, (double)n);
    }
}
template <typename T>
__global__ void k_dc_int_subtract(const T* __restrict__ x, int64_t n, const double* __restrict__ mean, double* __restrict__ y) {
    const double mr = mean[0], mi = mean[1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        y[2 * i] = __dsub_rn((double)x[2 * i], mr);
        y[2 * i + 1] = __dsub_rn((double)x[2 * i + 1], mi);
    }
}

template <typename T>
static int dc_int(urh_ctx* ctx, const void* d_iq, int64_t n, double* d_out) {
    const int nb = ctx->sm_count * 4;
    long long* part;
    double* mean;
    URH_CHECK(urh_arena(ctx, (size_t)nb * 2, &part));
    URH_CHECK(urh_arena(ctx, 2, &mean));
    URH_LAUNCH(ctx, k_dc_int_partial<T>, nb, 256, 0, (const T*)d_iq, n, part);
    URH_LAUNCH(ctx, k_dc_int_fold, 1, 32, 0, (const long long*)part, nb, n, mean);
    const unsigned grid = (unsigned)min(urh_div_up(n, 256), (int64_t)ctx->sm_count * 16);
    URH_LAUNCH(ctx, k_dc_int_subtract<T>, grid, 256, 0, (const T*)d_iq, n, (const double*)mean, d_out);
    return URH_OK;
}

extern "C" int urh_dc_correction_int(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, double* d_out) {
    if (n <= 0) return URH_OK;
    urh_arena_reset(ctx);
    switch (dtype) {
        case URH_DT_I8: return dc_int<int8_t>(ctx, d_iq, n, d_out);
        case URH_DT_U8: return dc_int<uint8_t>(ctx, d_iq, n, d_out);
        case URH_DT_I16: return dc_int<int16_t>(ctx, d_iq, n, d_out);
        case URH_DT_U16: return dc_int<uint16_t>(ctx, d_iq, n, d_out);
        default: URH_FAIL(ctx, URH_ERR_DTYPE, "urh_dc_correction_int: integer capture expected");
    }
}
