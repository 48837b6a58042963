// Synthetic capture generator for bench.py / large-size property tests (SURVEY §8d recipe):
// phase-continuous 2-FSK bursts with AWGN and noise-only gaps, generated directly in HBM.
// Not part of the reference's API surface; a measurement utility.
#include "common.cuh"

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// sym_sum[k] = sum_{j<k} b_j with b_j = +-1 (host-computed prefix), sym_bit[k] = b_k
__global__ void k_synth_fsk(float2* __restrict__ out, int64_t n, int64_t global_offset, int sps,
                            const int8_t* __restrict__ sym_bit, const int32_t* __restrict__ sym_sum, double dev_ratio,
                            float amplitude, float sigma, uint64_t seed, int64_t period, int64_t burst,
                            int64_t big_gap_start, int64_t big_gap_end, int64_t tail_start) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t g = i + global_offset;
        const int64_t k = i / sps;
        const int r = (int)(i - k * sps);
        const bool on = (g % period) < burst && !(g >= big_gap_start && g < big_gap_end) && g < tail_start;
        float re = 0.f, im = 0.f;
        if (on) {
            const double m = (double)sps * (double)sym_sum[k] + (double)sym_bit[k] * (double)r;
            double t = m * dev_ratio;
            t -= floor(t);
            float s, c;
            sincospif(2.0f * (float)t, &s, &c);
            re = amplitude * c;
            im = amplitude * s;
        }
        const uint64_t h = splitmix64(seed ^ (uint64_t)g * 0xD6E8FEB86659FD93ull);
        const float u1 = ((float)((h >> 40) + 1)) * (1.0f / 16777217.0f);
        const float u2 = (float)((h >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
        const float rad = sigma * sqrtf(-2.0f * __logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        out[i] = make_float2(re + rad * cs, im + rad * sn);
    }
}

extern "C" int urh_synth_fsk(urh_ctx* ctx, float* d_iq, int64_t n, int64_t global_offset, int sps, const int8_t* d_sym_bit,
                             const int32_t* d_sym_sum, double dev_ratio, float amplitude, float sigma, uint64_t seed,
                             int64_t period, int64_t burst, int64_t big_gap_start, int64_t big_gap_end, int64_t tail_start) {
    if (n <= 0) return URH_OK;
    const int block = 256;
    const unsigned grid = (unsigned)(ctx->sm_count * 16);
    URH_LAUNCH(ctx, k_synth_fsk, grid, block, 0, (float2*)d_iq, n, global_offset, sps, d_sym_bit, d_sym_sum, dev_ratio,
               amplitude, sigma, seed, period, burst, big_gap_start, big_gap_end, tail_start);
    return URH_OK;
}


// M-PSK bursts with AWGN and noise-only gaps for BASELINE configs[4] (symbols are a hash of the GLOBAL symbol index, so any shard of
// the capture can be generated on its own rank): x[g] = a exp(i (2 pi fc g + 2 pi sym(g / sps) / order [+ pi/4 for order 4])) + noise
__global__ void k_synth_psk(float2* __restrict__ out, int64_t n, int64_t global_offset, int sps, int order, double carrier_ratio,
                            float amplitude, float sigma, uint64_t seed, int64_t period, int64_t burst, int64_t tail_start) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t g = i + global_offset;
        const bool on = (g % period) < burst && g < tail_start;
        float re = 0.f, im = 0.f;
        if (on) {
            const uint64_t k = (uint64_t)(g / sps);
            const int sym = (int)(splitmix64(seed * 0x9E3779B97F4A7C15ull + k) % (uint64_t)order);
            double t = (double)g * carrier_ratio;
            t -= floor(t);
            t += (double)sym / (double)order + (order == 4 ? 0.125 : 0.0);
            t -= floor(t);
            float s, c;
            sincospif(2.0f * (float)t, &s, &c);
            re = amplitude * c;
            im = amplitude * s;
        }
        const uint64_t h = splitmix64(seed ^ (uint64_t)g * 0xD6E8FEB86659FD93ull);
        const float u1 = ((float)((h >> 40) + 1)) * (1.0f / 16777217.0f);
        const float u2 = (float)((h >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
        const float rad = sigma * sqrtf(-2.0f * __logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        out[i] = make_float2(re + rad * cs, im + rad * sn);
    }
}

extern "C" int urh_synth_psk(urh_ctx* ctx, float* d_iq, int64_t n, int64_t global_offset, int sps, int order, double carrier_ratio,
                             float amplitude, float sigma, uint64_t seed, int64_t period, int64_t burst, int64_t tail_start) {
    if (n <= 0) return URH_OK;
    if (order != 2 && order != 4) URH_FAIL(ctx, URH_ERR_INVALID, "synth_psk: order 2 or 4");
    URH_LAUNCH(ctx, k_synth_psk, (unsigned)(ctx->sm_count * 16), 256, 0, (float2*)d_iq, n, global_offset, sps, order, carrier_ratio, amplitude,
               sigma, seed, period, burst, tail_start);
    return URH_OK;
}
