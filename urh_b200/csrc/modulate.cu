// Modulator forward path: ASK / FSK / PSK / OQPSK / GFSK (reference: signal_functions.pyx:56-243,
// modulate_c / __modulate / get_gauss_filtered_freqs_phases / gauss_fir / get_oqpsk_bits).
//
// A call modulates a BATCH of messages that share one parameter set (how every reference caller uses it:
// Modulator.modulate per message from modulate_messages / ContinuousModulator / the simulator, always start=0
// — SURVEY §3.3, H4).  Per message:
//   * FSK: the per-symbol phase corrections are a float32-rounded serial recurrence over the symbols
//     (pyx:121-137) -> one thread per message walks them (double fmod, exactly as the C code does);
//   * GFSK: per-sample frequencies = Gaussian-filtered symbol frequencies (np.convolve 'same'), then a
//     float32-rounded serial phase recurrence over all samples (pyx:220-224) -> one thread per message;
//   * then every sample is independent: t = (float)(i+start)/sample_rate, arg = (float)(2*pi*f*t + phi + corr)
//     evaluated in double, I/Q = a*cosf(arg), a*sinf(arg) with glibc's sinf/cosf restated bit-for-bit
//     (glibc_sincosf.h) and C truncation to the integer output types.
// Bit-exact for ASK/FSK/PSK/OQPSK.  GFSK is tolerance-parity: numpy's float32 convolution runs in OpenBLAS
// sdot, whose summation order depends on the host CPU.
#include "common.cuh"
#include "glibc_sincosf.h"

#include <math.h>

struct ModParams {
    uint32_t sps;
    int mod_type;
    int bps;
    float a, f, phi, sample_rate;
    uint32_t start;
    int out_dtype;
    int nparams;
    float params[256];
};

__device__ __forceinline__ uint32_t symbol_index(const uint8_t* bits, int64_t s, int bps) {
    // bit_array_to_number(bits, end=(s+1)*bps, start=s*bps): MSB first (util.pyx:50-61)
    uint32_t r = 0;
    for (int b = 0; b < bps; b++) r = r * 2u + (uint32_t)bits[s * bps + b];
    return r;
}

// FSK phase corrections, one thread per message (pyx:121-137)
__global__ void k_fsk_corrections(const uint8_t* __restrict__ bits, const int64_t* __restrict__ bit_off,
                                  const int64_t* __restrict__ sym_off, int nmsg, const __grid_constant__ ModParams P,
                                  float* __restrict__ corr) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmsg) return;
    const uint8_t* b = bits + bit_off[m];
    const int64_t nsym = (bit_off[m + 1] - bit_off[m]) / P.bps;
    float* c = corr + sym_off[m];
    if (nsym <= 0) return;
    float prev = 0.0f;
    c[0] = 0.0f;
    uint32_t pidx = symbol_index(b, 0, P.bps);
    const double two_pi = 2.0 * M_PI;
    for (int64_t s = 1; s < nsym; s++) {
        const uint32_t idx = symbol_index(b, s, P.bps);
        const float f = P.params[idx], fp = P.params[pidx];
        if (f != fp) {
            const float t = __fdiv_rn(__ll2float_rn((long long)(s * (int64_t)P.sps + (int64_t)P.start - 1)), P.sample_rate);
            const double v = __dadd_rn((double)prev, __dmul_rn(__dmul_rn(two_pi, (double)__fsub_rn(fp, f)), (double)t));
            prev = (float)fmod(v, two_pi);
        }
        c[s] = prev;
        pidx = idx;
    }
}

// GFSK: Gaussian-filtered per-sample frequencies ('same' convolution of the piecewise-constant symbol frequencies).
// The frequency is constant over a symbol, so c[t] = sum_j freq[t-j] g[j] collapses to one term per symbol the window
// touches: f_sym * (G[hi] - G[lo]) with G the running sum of the taps (double) — ~3 terms instead of 2*sps+1.
__global__ void k_gfsk_freqs(const uint8_t* __restrict__ bits, const int64_t* __restrict__ bit_off,
                             const int64_t* __restrict__ smp_off, int nmsg, const __grid_constant__ ModParams P,
                             const double* __restrict__ gsum /* glen+1 prefix sums */, int glen, float* __restrict__ fp_table) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int m = blockIdx.y; m < nmsg; m += gridDim.y) {   // grid.y is capped at 65535
    const uint8_t* b = bits + bit_off[m];
    const int64_t nsym = (bit_off[m + 1] - bit_off[m]) / P.bps;
    const int64_t nval = nsym * P.sps;
    float* out = fp_table + 2 * smp_off[m];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nval; k += stride) {
        // np.convolve(longer, shorter, 'same'): centred on the longer operand
        const int64_t t = (nval >= glen) ? k + (glen - 1) / 2 : k + (nval - 1) / 2;
        // c[t] = sum over i in [max(0, t-glen+1), min(nval-1, t)] of freq[i] * g[t-i]
        int64_t ilo = t - (glen - 1), ihi = t;
        if (ilo < 0) ilo = 0;
        if (ihi > nval - 1) ihi = nval - 1;
        double acc = 0.0;
        for (int64_t i = ilo; i <= ihi;) {
            const int64_t sidx = i / P.sps;
            int64_t iend = (sidx + 1) * (int64_t)P.sps - 1;  // last sample of this symbol
            if (iend > ihi) iend = ihi;
            // taps j = t - i for i in [i, iend]  ->  j in [t - iend, t - i]
            const double w = gsum[t - i + 1] - gsum[t - iend];
            acc += (double)P.params[symbol_index(b, sidx, P.bps)] * w;
            i = iend + 1;
        }
        out[2 * k] = (float)acc;
    }
    }
}

// GFSK phase recurrence (pyx:220-224): phases[i+1] = float32(2*pi*t[i]*(f[i] - f[i+1]) + phases[i]) is a sequential
// float32 accumulation, so it stays serial per message -- but only the rounding chain: one WARP per message computes the
// 32 increments of a block in parallel (coalesced reads), then folds them in order from registers.
__global__ void __launch_bounds__(128) k_gfsk_phases(const int64_t* __restrict__ bit_off, const int64_t* __restrict__ smp_off, int nmsg,
                                                    const __grid_constant__ ModParams P, float* __restrict__ fp_table) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double two_pi = 2.0 * M_PI;
    for (int64_t m = warp; m < nmsg; m += nwarps) {
        const int64_t nsym = (bit_off[m + 1] - bit_off[m]) / P.bps;
        const int64_t nval = nsym * P.sps;
        if (nval <= 0) continue;
        float* tab = fp_table + 2 * smp_off[m];
        float ph = P.phi;
        if (lane == 0) tab[1] = ph;
        for (int64_t base = 0; base + 1 < nval; base += 32) {
            const int64_t i = base + lane;
            const bool valid = i + 1 < nval;
            double c = 0.0;
            if (valid) {
                const float fcur = tab[2 * i], fnext = tab[2 * (i + 1)];
                // t = np.arange(start, ..., dtype=float32) / sample_rate: float32 index, float32 division
                const float t = __fdiv_rn(__ll2float_rn((long long)i + (long long)P.start), P.sample_rate);
                c = __dmul_rn(__dmul_rn(two_pi, (double)t), (double)__fsub_rn(fcur, fnext));
            }
            const int count = (int)min((int64_t)32, nval - 1 - base);
            float mine = 0.0f;
            for (int l = 0; l < count; l++) {
                const double cl = __shfl_sync(0xffffffffu, c, l);
                ph = (float)__dadd_rn(cl, (double)ph);
                if (lane == l) mine = ph;
            }
            if (valid) tab[2 * (i + 1) + 1] = mine;
        }
    }
}

template <typename OUT>
__device__ __forceinline__ OUT mod_cast(float v);
template <> __device__ __forceinline__ float mod_cast<float>(float v) { return v; }
template <> __device__ __forceinline__ int8_t mod_cast<int8_t>(float v) { return (int8_t)(int)v; }
template <> __device__ __forceinline__ int16_t mod_cast<int16_t>(float v) { return (int16_t)(int)v; }

template <typename OUT>
__global__ void k_modulate(const uint8_t* __restrict__ bits, const int64_t* __restrict__ bit_off,
                           const int64_t* __restrict__ sym_off, const int64_t* __restrict__ smp_off,
                           const int64_t* __restrict__ out_off, int nmsg, const __grid_constant__ ModParams P,
                           const float* __restrict__ corr, const float* __restrict__ fp_table, OUT* __restrict__ out) {
    const double two_pi = 2.0 * M_PI;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int m = blockIdx.y; m < nmsg; m += gridDim.y) {   // grid.y is capped at 65535
    const uint8_t* b = bits + bit_off[m];
    const int64_t nsym = (bit_off[m + 1] - bit_off[m]) / P.bps;
    const int64_t nval = nsym * P.sps;
    OUT* o = out + 2 * out_off[m];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nval; i += stride) {
        const int64_t s = i / P.sps;
        const uint32_t idx = symbol_index(b, s, P.bps);
        float a = P.a, f = P.f, phi = P.phi, pc = 0.0f;
        if (P.mod_type == URH_MOD_ASK) {
            a = P.params[idx];
            if (a == 0.0f) continue;  // output stays zero (pyx:148-150)
        } else if (P.mod_type == URH_MOD_FSK) {
            f = P.params[idx];
            pc = corr[sym_off[m] + s];
        } else if (P.mod_type == URH_MOD_PSK || P.mod_type == URH_MOD_OQPSK) {
            phi = P.params[idx];
        } else if (P.mod_type == URH_MOD_GFSK) {
            f = fp_table[2 * (smp_off[m] + i)];
            phi = fp_table[2 * (smp_off[m] + i) + 1];
        }
        const float t = __fdiv_rn(__ll2float_rn((long long)i + (long long)P.start), P.sample_rate);
        // current_arg = ((((2.0*M_PI)*f)*t) + phi) + phase_correction, double arithmetic, stored as float
        const double argd = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(two_pi, (double)f), (double)t), (double)phi), (double)pc);
        const float arg = (float)argd;
        float sn, cs;
        int ok;
        urh_glibc_sincosf(arg, &sn, &cs, &ok);
        if (!ok) { sn = sinf(arg); cs = cosf(arg); }
        float I = __fmul_rn(a, cs), Q = __fmul_rn(a, sn);
        if (P.mod_type == URH_MOD_OQPSK) {
            // pyx:168-172: Q of the first symbol and I of the last symbol are zeroed
            if (i < P.sps) Q = 0.0f;
            if (i >= nval - P.sps) I = 0.0f;
        }
        o[2 * i] = mod_cast<OUT>(I);
        o[2 * i + 1] = mod_cast<OUT>(Q);
    }
    }
}

// Batch modulate.  d_bits: concatenated bit arrays (uint8, already OQPSK-shuffled if needed);
// h_bit_off[nmsg+1]: bit offsets; h_out_off[nmsg+1]: output SAMPLE offsets (message m occupies
// [h_out_off[m], h_out_off[m+1]) = symbols*sps + pause samples); d_out is zero-filled here.
extern "C" int urh_modulate_batch(urh_ctx* ctx, const uint8_t* d_bits, const int64_t* h_bit_off, const int64_t* h_out_off,
                                  int nmsg, uint32_t samples_per_symbol, int mod_type, const float* h_params, int nparams,
                                  int bits_per_symbol, float carrier_amplitude, float carrier_frequency, float carrier_phase,
                                  float sample_rate, uint32_t start, int out_dtype, const float* h_gauss_fir, int gauss_len,
                                  void* d_out) {
    if (nmsg <= 0) return URH_OK;
    if (mod_type != URH_MOD_ASK && mod_type != URH_MOD_FSK && mod_type != URH_MOD_PSK && mod_type != URH_MOD_GFSK &&
        mod_type != URH_MOD_OQPSK)
        URH_FAIL(ctx, URH_ERR_MODULATION, "unknown modulation type");
    if (mod_type == URH_MOD_OQPSK && bits_per_symbol != 2) URH_FAIL(ctx, URH_ERR_MODULATION, "OQPSK needs bits_per_symbol == 2");
    if (out_dtype != URH_DT_I8 && out_dtype != URH_DT_I16 && out_dtype != URH_DT_F32)
        URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype for modulation");
    if (bits_per_symbol < 1 || bits_per_symbol > 8 || nparams > 256 || nparams < (1 << bits_per_symbol))
        URH_FAIL(ctx, URH_ERR_INVALID, "bits_per_symbol / parameters mismatch");
    urh_arena_reset(ctx);
    ModParams P;
    memset(&P, 0, sizeof(P));
    P.sps = samples_per_symbol; P.mod_type = mod_type; P.bps = bits_per_symbol; P.a = carrier_amplitude;
    P.f = carrier_frequency; P.phi = carrier_phase; P.sample_rate = sample_rate; P.start = start; P.out_dtype = out_dtype;
    P.nparams = nparams;
    memcpy(P.params, h_params, sizeof(float) * nparams);
    // per-message offsets: symbols and modulated (non-pause) samples
    std::vector<int64_t> sym_off(nmsg + 1, 0), smp_off(nmsg + 1, 0);
    int64_t max_samples = 0;
    for (int m = 0; m < nmsg; m++) {
        const int64_t nsym = (h_bit_off[m + 1] - h_bit_off[m]) / bits_per_symbol;
        sym_off[m + 1] = sym_off[m] + nsym;
        smp_off[m + 1] = smp_off[m] + nsym * samples_per_symbol;
        if (nsym * (int64_t)samples_per_symbol > max_samples) max_samples = nsym * samples_per_symbol;
        if (h_out_off[m + 1] - h_out_off[m] < nsym * (int64_t)samples_per_symbol) URH_FAIL(ctx, URH_ERR_INVALID, "output offsets too small");
    }
    int64_t *d_bit_off, *d_sym_off, *d_smp_off, *d_out_off;
    URH_CHECK(urh_arena(ctx, (size_t)nmsg + 1, &d_bit_off));
    URH_CHECK(urh_arena(ctx, (size_t)nmsg + 1, &d_sym_off));
    URH_CHECK(urh_arena(ctx, (size_t)nmsg + 1, &d_smp_off));
    URH_CHECK(urh_arena(ctx, (size_t)nmsg + 1, &d_out_off));
    const size_t ob = (size_t)(nmsg + 1) * sizeof(int64_t);
    URH_CUDA(ctx, cudaMemcpyAsync(d_bit_off, h_bit_off, ob, cudaMemcpyHostToDevice, ctx->stream));
    URH_CUDA(ctx, cudaMemcpyAsync(d_sym_off, sym_off.data(), ob, cudaMemcpyHostToDevice, ctx->stream));
    URH_CUDA(ctx, cudaMemcpyAsync(d_smp_off, smp_off.data(), ob, cudaMemcpyHostToDevice, ctx->stream));
    URH_CUDA(ctx, cudaMemcpyAsync(d_out_off, h_out_off, ob, cudaMemcpyHostToDevice, ctx->stream));
    const size_t elem = out_dtype == URH_DT_F32 ? 4 : (out_dtype == URH_DT_I16 ? 2 : 1);
    URH_CUDA(ctx, cudaMemsetAsync(d_out, 0, (size_t)h_out_off[nmsg] * 2 * elem, ctx->stream));
    float *corr = nullptr, *fp_table = nullptr;
    double* d_gsum = nullptr;
    if (mod_type == URH_MOD_FSK) {
        URH_CHECK(urh_arena(ctx, (size_t)sym_off[nmsg] + 1, &corr));
        URH_LAUNCH(ctx, k_fsk_corrections, (unsigned)urh_div_up(nmsg, 64), 64, 0, d_bits, d_bit_off, d_sym_off, nmsg, P, corr);
    }
    const unsigned gx = (unsigned)max((int64_t)1, min(urh_div_up(max_samples, 256), (int64_t)ctx->sm_count * 8));
    const dim3 grid(gx, (unsigned)min(nmsg, 65535));
    if (mod_type == URH_MOD_GFSK) {
        if (!h_gauss_fir || gauss_len <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "GFSK needs the gaussian filter taps");
        URH_CHECK(urh_arena(ctx, (size_t)smp_off[nmsg] * 2 + 2, &fp_table));
        std::vector<double> gsum((size_t)gauss_len + 1, 0.0);
        for (int j = 0; j < gauss_len; j++) gsum[j + 1] = gsum[j] + (double)h_gauss_fir[j];
        URH_CHECK(urh_arena(ctx, (size_t)gauss_len + 1, &d_gsum));
        URH_CUDA(ctx, cudaMemcpyAsync(d_gsum, gsum.data(), (gauss_len + 1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        URH_LAUNCH(ctx, k_gfsk_freqs, grid, 256, 0, d_bits, d_bit_off, d_smp_off, nmsg, P, d_gsum, gauss_len, fp_table);
        URH_LAUNCH(ctx, k_gfsk_phases, (unsigned)min((int64_t)urh_div_up(nmsg, 4), (int64_t)ctx->sm_count * 16), 128, 0, d_bit_off, d_smp_off, nmsg, P,
                   fp_table);
    }
    if (out_dtype == URH_DT_F32)
        URH_LAUNCH(ctx, k_modulate<float>, grid, 256, 0, d_bits, d_bit_off, d_sym_off, d_smp_off, d_out_off, nmsg, P, corr, fp_table, (float*)d_out);
    else if (out_dtype == URH_DT_I16)
        URH_LAUNCH(ctx, k_modulate<int16_t>, grid, 256, 0, d_bits, d_bit_off, d_sym_off, d_smp_off, d_out_off, nmsg, P, corr, fp_table, (int16_t*)d_out);
    else
        URH_LAUNCH(ctx, k_modulate<int8_t>, grid, 256, 0, d_bits, d_bit_off, d_sym_off, d_smp_off, d_out_off, nmsg, P, corr, fp_table, (int8_t*)d_out);
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the offset vectors above are host temporaries
    return URH_OK;
}
