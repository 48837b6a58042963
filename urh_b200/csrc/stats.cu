// Auto-interpretation statistics on the GPU:
//   magnitudes / noise level  — util.get_magnitudes (util.pyx:128-136), AutoInterpretation.detect_noise_level (:60-91)
//   center detection          — AutoInterpretation.detect_center (:226-277): rank trimming, min/max/var, histogram
//   message segmentation      — auto_interpretation.segment_messages_from_magnitudes (auto_interpretation.pyx:55-111)
//   plateau lengths           — auto_interpretation.get_plateau_lengths (:179-208)
//   median filter, dB         — auto_interpretation.median_filter (:211-240), util.arr2decibel (util.pyx:38-48)
// The small, data-dependent decision logic (which chunks are quiet, which histogram bins are local maxima)
// stays on the host in urh_b200/ainterpretation/AutoInterpretation.py, exactly as in the reference; the
// sample-rate reductions run here.
#include "dense_f32.cuh"
#include "scan.cuh"
#include "sparse.cuh"

#include <math.h>

// ---- magnitudes ------------------------------------------------------------------------------------------
// float32 IQ: (double)sqrtf(fl(re*re + im*im)); integer IQ: squares and sum in (wrapping) int32, double sqrt.
template <int DT>
__device__ __forceinline__ double urh_magnitude(const void* iq, int64_t i) {
    typedef typename UrhElem<DT>::type E;
    const E* p = (const E*)iq + 2 * i;
    if (DT == URH_DT_F32) {
        const float re = (float)p[0], im = (float)p[1];
        return (double)__fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
    } else {
        const uint32_t re = (uint32_t)(int32_t)p[0], im = (uint32_t)(int32_t)p[1];
        const int32_t ssum = (int32_t)(re * re + im * im);
        return sqrt((double)ssum);
    }
}

template <int DT>
__global__ void k_magnitudes(const void* __restrict__ iq, int64_t n, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = urh_magnitude<DT>(iq, i);
}

#define URH_DISPATCH_DT(dtype, KERNEL_CALL)                         \
    switch (dtype) {                                                \
        case URH_DT_I8: { constexpr int DT = URH_DT_I8; KERNEL_CALL; } break;   \
        case URH_DT_U8: { constexpr int DT = URH_DT_U8; KERNEL_CALL; } break;   \
        case URH_DT_I16: { constexpr int DT = URH_DT_I16; KERNEL_CALL; } break; \
        case URH_DT_U16: { constexpr int DT = URH_DT_U16; KERNEL_CALL; } break; \
        case URH_DT_F32: { constexpr int DT = URH_DT_F32; KERNEL_CALL; } break; \
        default: URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");  \
    }

extern "C" int urh_get_magnitudes(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, double* d_out) {
    if (n <= 0) return URH_OK;
    const unsigned grid = (unsigned)min((int64_t)ctx->sm_count * 16, urh_div_up(n, 256));
    URH_DISPATCH_DT(dtype, URH_LAUNCH(ctx, k_magnitudes<DT>, grid, 256, 0, d_iq, n, d_out));
    return URH_OK;
}

// ---- chunk statistics for detect_noise_level -----------------------------------------------------------------
// Chunks are counted from the END of the array (AutoInterpretation.py:66-72): chunk j covers
// [n - (j+1)*cs, n - j*cs).  Each block reduces a slice of one chunk to (sum, max) in double; a second kernel
// folds the slices in a fixed order, so the result is deterministic.
#define STAT_BLOCK 256
#define STAT_SLICES 64

template <typename LOADER>
__global__ void __launch_bounds__(STAT_BLOCK) k_chunk_partial(LOADER ld, int64_t n, int64_t cs, int nchunks,
                                                              double* __restrict__ psum, double* __restrict__ pmax) {
    const int chunk = blockIdx.x / STAT_SLICES, slice = blockIdx.x % STAT_SLICES;
    const int64_t c0 = n - (int64_t)(chunk + 1) * cs;
    const int64_t per = urh_div_up(cs, STAT_SLICES);
    const int64_t s0 = c0 + (int64_t)slice * per;
    const int64_t s1 = min(s0 + per, c0 + cs);
    double sum = 0.0, mx = -1.0;
    for (int64_t i = s0 + threadIdx.x; i < s1; i += STAT_BLOCK) {
        const double m = ld(i);
        sum += m;
        mx = fmax(mx, m);   // NaN-ignoring like a sequence of `if e > maximum`
    }
    __shared__ double s_sum[STAT_BLOCK], s_max[STAT_BLOCK];
    s_sum[threadIdx.x] = sum;
    s_max[threadIdx.x] = mx;
    __syncthreads();
    for (int off = STAT_BLOCK / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + off];
            s_max[threadIdx.x] = fmax(s_max[threadIdx.x], s_max[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        psum[blockIdx.x] = s_sum[0];
        pmax[blockIdx.x] = s_max[0];
    }
}

__global__ void k_chunk_final(const double* __restrict__ psum, const double* __restrict__ pmax, int nchunks,
                              double* __restrict__ sum, double* __restrict__ mx) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    double s = 0.0, m = -1.0;
    for (int j = 0; j < STAT_SLICES; j++) {
        s += psum[c * STAT_SLICES + j];
        m = fmax(m, pmax[c * STAT_SLICES + j]);
    }
    sum[c] = s;
    mx[c] = m;
}

template <int DT>
struct LoadMagIQ {
    const void* iq;
    __device__ __forceinline__ double operator()(int64_t i) const { return urh_magnitude<DT>(iq, i); }
};
template <typename T>
struct LoadReal {
    const T* x;
    __device__ __forceinline__ double operator()(int64_t i) const { return (double)x[i]; }
};

template <typename LOADER>
static int chunk_stats(urh_ctx* ctx, LOADER ld, int64_t n, int64_t cs, int nchunks, double* h_sum, double* h_max) {
    urh_arena_reset(ctx);
    double *psum, *pmax, *sum, *mx;
    URH_CHECK(urh_arena(ctx, (size_t)nchunks * STAT_SLICES, &psum));
    URH_CHECK(urh_arena(ctx, (size_t)nchunks * STAT_SLICES, &pmax));
    URH_CHECK(urh_arena(ctx, (size_t)nchunks, &sum));
    URH_CHECK(urh_arena(ctx, (size_t)nchunks, &mx));
    URH_LAUNCH(ctx, (k_chunk_partial<LOADER>), (unsigned)(nchunks * STAT_SLICES), STAT_BLOCK, 0, ld, n, cs, nchunks, psum, pmax);
    URH_LAUNCH(ctx, k_chunk_final, (unsigned)urh_div_up(nchunks, 128), 128, 0, psum, pmax, nchunks, sum, mx);
    URH_CUDA(ctx, cudaMemcpyAsync(h_sum, sum, nchunks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaMemcpyAsync(h_max, mx, nchunks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return URH_OK;
}

// per-chunk (sum, max) of the magnitudes of an IQ capture, never materialising the float64 magnitude array
extern "C" int urh_noise_chunk_stats_iq(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, int64_t chunksize,
                                        int nchunks, double* h_sum, double* h_max) {
    if (nchunks <= 0 || chunksize <= 0 || (int64_t)nchunks * chunksize > n) URH_FAIL(ctx, URH_ERR_INVALID, "bad chunking");
    URH_DISPATCH_DT(dtype, { LoadMagIQ<DT> ld; ld.iq = d_iq; URH_CHECK(chunk_stats(ctx, ld, n, chunksize, nchunks, h_sum, h_max)); });
    return URH_OK;
}
// the same on an existing magnitude array (float32: is_f64 = 0, float64: is_f64 = 1)
extern "C" int urh_noise_chunk_stats(urh_ctx* ctx, const void* d_mags, int is_f64, int64_t n, int64_t chunksize, int nchunks,
                                     double* h_sum, double* h_max) {
    if (nchunks <= 0 || chunksize <= 0 || (int64_t)nchunks * chunksize > n) URH_FAIL(ctx, URH_ERR_INVALID, "bad chunking");
    if (is_f64) {
        LoadReal<double> ld; ld.x = (const double*)d_mags;
        return chunk_stats(ctx, ld, n, chunksize, nchunks, h_sum, h_max);
    }
    LoadReal<float> ld; ld.x = (const float*)d_mags;
    return chunk_stats(ctx, ld, n, chunksize, nchunks, h_sum, h_max);
}

// ---- run tables for the segmenter and the plateau RLE ----------------------------------------------------------------
// mode 0: class = x > thr (segment_messages, tolerance 9 <=> 10 consecutive samples, auto_interpretation.pyx:69)
// mode 1: class = x <= thr ? 0 : 1 (get_plateau_lengths, tolerance 0: every run start)
// Candidates (position, class) are returned to the host (they are few); *h_last = {last_cls, last_len, first_cls}.
template <typename SRC, typename T>
static int run_table(urh_ctx* ctx, const T* d_x, int64_t n, float thr, int tol, int64_t** h_pos, int16_t** h_cls,
                     int64_t* count, int64_t* h_last) {
    urh_arena_reset(ctx);
    UrhClassify cls;
    memset(&cls, 0, sizeof(cls));
    cls.noise_value = 0.0f;
    cls.order = 2;
    cls.thr[0] = thr;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int cap = URH_TILE / (tol + 1) + 2;
    UrhTileSummary* tiles;
    uint32_t* staging;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    const unsigned grid = (unsigned)urh_div_up(ntiles, URH_WARPS_PER_BLOCK);
    const int vec_in = (((uintptr_t)d_x % (2 * sizeof(T))) == 0) ? 1 : 0;
    URH_LAUNCH(ctx, (k_dense_f32<SRC, T>), grid, URH_WARPS_PER_BLOCK * 32, 0, d_x, n, vec_in, cls, tol, tiles, staging, cap,
               (int16_t*)nullptr, 0);
    UrhCandidates cand;
    URH_CHECK(urh_collect_candidates(ctx, n, tol, tiles, staging, cap, &cand));
    *count = cand.count;
    h_last[0] = cand.last_cls;
    h_last[1] = cand.last_len;
    *h_pos = nullptr;
    *h_cls = nullptr;
    if (cand.count > 0) {
        *h_pos = (int64_t*)malloc((size_t)cand.count * sizeof(int64_t));
        *h_cls = (int16_t*)malloc((size_t)cand.count * sizeof(int16_t));
        URH_CUDA(ctx, cudaMemcpyAsync(*h_pos, cand.pos, (size_t)cand.count * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CUDA(ctx, cudaMemcpyAsync(*h_cls, cand.cls, (size_t)cand.count * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return URH_OK;
}

// The two-state machine of segment_messages_from_magnitudes over the run table (host; the runs of >= 10 samples are few):
// h_pos / h_cls = candidates of tolerance 9 (index of the 10th consecutive sample of a run, its class 0 = below / 1 = above the
// noise threshold), first_above = class of sample 0, (last_cls, last_len) = the run that ends the capture.  Pure host code, so the
// sharded path (urh_b200/dist.py: candidates of all shards concatenated) shares it with urh_segment_messages.
extern "C" int urh_segments_from_runs(const int64_t* h_pos, const int16_t* h_cls, int64_t count, int first_above, int last_cls,
                                      int64_t last_len, int64_t n, int64_t* h_segments, int64_t cap, int64_t* k) {
    int state = first_above ? 1 : 0;
    int64_t start = 0, m = 0;
    for (int64_t j = 0; j < count; j++) {
        if (h_cls[j] == state) continue;
        const int64_t p = h_pos[j];  // index of the 10th consecutive sample of the opposite class
        if (state == 1) {
            if (m < cap) { h_segments[2 * m] = start; h_segments[2 * m + 1] = p - 10; }
            m++;
            state = 0;
        } else {
            start = p - 10;
            state = 1;
        }
    }
    if (state == 1) {
        const int64_t conseq_below = (last_cls == 0) ? last_len : 0;
        if (start < n - conseq_below) {
            if (m < cap) { h_segments[2 * m] = start; h_segments[2 * m + 1] = n - conseq_below; }
            m++;
        }
    }
    *k = m;
    return URH_OK;
}

// One shard of a capture whose magnitudes are spread over the ranks: the dense pass (class = above the threshold, tolerance 9)
// leaves the tile table for urh_shard_candidates (carry of the preceding shards, global positions).
// h_summary = {last_cls, last_len, whole, class of the shard's first sample}.
extern "C" int urh_segment_shard_pass(urh_ctx* ctx, const void* d_mags, int is_f64, int64_t n, float noise_threshold, int64_t* h_summary) {
    if (n <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "empty shard");
    urh_arena_reset(ctx);
    UrhClassify cls;
    memset(&cls, 0, sizeof(cls));
    cls.noise_value = 0.0f;
    cls.order = 2;
    cls.thr[0] = noise_threshold;
    const int tol = 9;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int cap = URH_TILE / (tol + 1) + 2;
    UrhTileSummary* tiles;
    uint32_t* staging;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    const unsigned grid = (unsigned)urh_div_up(ntiles, URH_WARPS_PER_BLOCK);
    double f0 = 0.0;
    if (is_f64) {
        const int vec_in = (((uintptr_t)d_mags % 16) == 0) ? 1 : 0;
        URH_LAUNCH(ctx, (k_dense_f32<SrcAbove, double>), grid, URH_WARPS_PER_BLOCK * 32, 0, (const double*)d_mags, n, vec_in, cls, tol, tiles,
                   staging, cap, (int16_t*)nullptr, 0);
        URH_CUDA(ctx, cudaMemcpyAsync(&f0, d_mags, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        const int vec_in = (((uintptr_t)d_mags % 8) == 0) ? 1 : 0;
        URH_LAUNCH(ctx, (k_dense_f32<SrcAbove, float>), grid, URH_WARPS_PER_BLOCK * 32, 0, (const float*)d_mags, n, vec_in, cls, tol, tiles,
                   staging, cap, (int16_t*)nullptr, 0);
        float f32 = 0.f;
        URH_CUDA(ctx, cudaMemcpyAsync(&f32, d_mags, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        f0 = (double)f32;
    }
    ctx->shard_tiles = tiles;
    ctx->shard_staging = staging;
    ctx->shard_cap = cap;
    ctx->shard_n = n;
    ctx->shard_tol = tol;
    URH_CHECK(urh_shard_run_total(ctx, n, tiles, h_summary));   // synchronises
    h_summary[3] = (is_f64 ? (f0 > (double)noise_threshold) : ((float)f0 > noise_threshold)) ? 1 : 0;
    return URH_OK;
}

// segment_messages_from_magnitudes (auto_interpretation.pyx:55-111).  d_mags float32 (is_f64=0) or float64.
// h_segments receives (start, end) pairs, capacity `cap` pairs; *k = number of messages (may exceed cap: call again).
extern "C" int urh_segment_messages(urh_ctx* ctx, const void* d_mags, int is_f64, int64_t n, float noise_threshold,
                                    int64_t* h_segments, int64_t cap, int64_t* k) {
    *k = 0;
    if (n <= 0) return URH_OK;
    int64_t* pos = nullptr;
    int16_t* cl = nullptr;
    int64_t count = 0, last[2];
    float first = 0.f;
    if (is_f64) {
        double f0;
        URH_CUDA(ctx, cudaMemcpyAsync(&f0, d_mags, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CHECK((run_table<SrcAbove, double>(ctx, (const double*)d_mags, n, noise_threshold, 9, &pos, &cl, &count, last)));
        first = (f0 > (double)noise_threshold) ? 1.f : 0.f;
    } else {
        float f0;
        URH_CUDA(ctx, cudaMemcpyAsync(&f0, d_mags, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CHECK((run_table<SrcAbove, float>(ctx, (const float*)d_mags, n, noise_threshold, 9, &pos, &cl, &count, last)));
        first = (f0 > noise_threshold) ? 1.f : 0.f;
    }
    const int rc = urh_segments_from_runs(pos, cl, count, first > 0.f ? 1 : 0, (int)last[0], last[1], n, h_segments, cap, k);
    free(pos);
    free(cl);
    return rc;
}

// get_plateau_lengths (auto_interpretation.pyx:179-208): h_out capacity `cap`; *k = number of plateaus.
extern "C" int urh_plateau_lengths(urh_ctx* ctx, const float* d_rect, int64_t n, float center, int percentage,
                                   uint64_t* h_out, int64_t cap, int64_t* k) {
    *k = 0;
    if (n <= 0) return URH_OK;
    int64_t* pos = nullptr;
    int16_t* cl = nullptr;
    int64_t count = 0, last[2];
    URH_CHECK((run_table<SrcCenter, float>(ctx, d_rect, n, center, 0, &pos, &cl, &count, last)));
    // candidates with tolerance 0 are the run starts (the first one is position 0)
    const uint64_t limit = (uint64_t)percentage * (uint64_t)n / 100;
    uint64_t sum = 0;
    int64_t m = 0;
    for (int64_t j = 1; j < count; j++) {
        // the reference checks `current_sum >= limit` at the top of every sample iteration, i.e. before a
        // boundary at pos[j] can append the run that ends there
        if (sum >= limit) break;
        const uint64_t len = (uint64_t)(pos[j] - pos[j - 1]);
        if (m < cap) h_out[m] = len;
        m++;
        sum += len;
    }
    if (limit == 0) m = 0;
    free(pos);
    free(cl);
    *k = m;
    return URH_OK;
}

// ---- median filter (auto_interpretation.pyx:211-240) ---------------------------------------------------------------------
// window [i, i+k) truncated at the end; values converted to float32 first; result = sorted[k'//2]
__global__ void k_median(const double* __restrict__ x, int64_t n, int k, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float buf[64];
    int kk = k;
    if (i + kk > n) kk = (int)(n - i);
    for (int j = 0; j < kk; j++) {
        const float v = (float)x[i + j];
        int p = j;
        while (p > 0 && buf[p - 1] > v) { buf[p] = buf[p - 1]; p--; }
        buf[p] = v;
    }
    out[i] = buf[kk / 2];
}

extern "C" int urh_median_filter(urh_ctx* ctx, const double* d_x, int64_t n, unsigned int k, float* d_out) {
    if (n <= 0) return URH_OK;
    if (k == 0 || k > 64) URH_FAIL(ctx, URH_ERR_INVALID, "median_filter: k must be in 1..64");
    URH_LAUNCH(ctx, k_median, (unsigned)urh_div_up(n, 128), 128, 0, d_x, n, (int)k, d_out);
    return URH_OK;
}

// ---- arr2decibel (util.pyx:38-48): 10.0f * log10f(re*re + im*im), float32 ------------------------------------------
__global__ void k_decibel(const float2* __restrict__ x, int64_t count, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float2 v = x[i];
        out[i] = __fmul_rn(10.0f, log10f(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))));
    }
}

extern "C" int urh_arr2decibel(urh_ctx* ctx, const float* d_complex, int64_t count, float* d_out) {
    if (count <= 0) return URH_OK;
    const unsigned grid = (unsigned)min((int64_t)ctx->sm_count * 16, urh_div_up(count, 256));
    URH_LAUNCH(ctx, k_decibel, grid, 256, 0, (const float2*)d_complex, count, d_out);
    return URH_OK;
}
