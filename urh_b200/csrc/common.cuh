// Shared internals of liburh_b200: context, error handling, scratch arena, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/urh_b200.h"

#define URH_FULL_MASK 0xffffffffu

struct urh_block {
    void* ptr;
    size_t bytes;
};

struct urh_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    cudaStream_t copy_stream[2];
    cudaEvent_t ev_start, ev_stop;
    cudaEvent_t ev_copy[2], ev_comp[2];
    cudaEvent_t ev_k0, ev_k1;  // around the dense kernel when profiling is on
    int profiling;
    int dense_timed;
    // stream timeline (urh_set_profiling(ctx, 2)): events recorded at named points of the sharded step
    cudaEvent_t tl_ev[32];
    const char* tl_name[32];
    int tl_count, tl_ready;
    char err[512];
    int64_t launches;
    // grow-only scratch arena: bump allocation inside a list of blocks, reset at the start of each op
    std::vector<urh_block> arena;
    size_t arena_block;  // current block
    size_t arena_used;   // bytes used in current block
    size_t arena_need;   // bytes requested since last reset (for coalescing)
    // digitizer result (context-owned, separate from the arena so it survives the next op's reset)
    int64_t* pulses;
    size_t pulses_cap_rows;
    int64_t pulses_k;
    // small pinned host mailbox for scalar read-backs
    int64_t* h_mail;
    // pinned staging for the host-pointer entry points
    void* h_stage[2];
    size_t h_stage_bytes;
    // cuFFT plan cache (spectrogram.cu)
    int fft_plan;
    int fft_nfft;
    int64_t fft_batch;
    bool fft_valid;
    // sharded digitizer state between urh_shard_dense and urh_shard_candidates (arena memory)
    void* shard_tiles;
    void* shard_staging;
    int shard_cap, shard_tol;
    int64_t shard_n;
    void* shard_state;
    // NCCL (nccl.cu)
    int64_t costas_stats[3];
    int64_t costas_redone;  // super-chunks the stitch pass had to chain itself
    // cuFFT plans of detect_modulation / cwt_haar: [0] C2C, [1] Z2Z, batch 2, length mod_plan_n
    int mod_plan[2];
    int mod_plan_valid[2];
    int64_t mod_plan_n[2];
    // urh_ppseq_to_bits results (arena)
    int bits_valid;
    int64_t bits_nmsg, bits_total, bits_npos;
    void *bits_ptr, *bits_msg_off, *bits_pauses, *bits_pos;
    const void* center_ts;
    const void* center_x;
    int64_t center_n;
    void* center_prefix;  // tile rank prefix left by urh_afp_demod_stats for urh_center_histogram_tiles (arena)
    void* nccl_comm;
    void* nccl_stage;
    void* nccl_hstage;  // pinned twin of nccl_stage
    int nccl_rank, nccl_world;
    // NVLink peer mailboxes (p2p.cu)
    void* p2p_local;
    void* p2p_peer[8];
    void* p2p_hout;
    int p2p_rank, p2p_world;
    unsigned long long p2p_seq;
    // tilescan.cuh workspace (look-back scans over tile tables)
    void* ts_mem;
    int64_t ts_cap_blocks;
    unsigned long long ts_issued;
    uint32_t ts_epoch;
    // finish.cu / center chain: small device-resident result block and its pinned mirror
    void* step_dev;
    // digitizer exchange state of a sharded capture (finish.cu)
    void* shard_fin;
};

#define URH_CUDA(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,  \
                     cudaGetErrorString(e__));                                                      \
            return (e__ == cudaErrorMemoryAllocation) ? URH_ERR_NOMEM : URH_ERR_CUDA;               \
        }                                                                                           \
    } while (0)

#define URH_CHECK(call)                 \
    do {                                \
        int rc__ = (call);              \
        if (rc__ != URH_OK) return rc__; \
    } while (0)

#define URH_FAIL(ctx, code, ...)                                  \
    do {                                                          \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
        return (code);                                            \
    } while (0)

// Kernel launch with accounting + error check.
#define URH_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
    do {                                                                                  \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                  \
        (ctx)->launches++;                                                                \
        URH_CUDA(ctx, cudaGetLastError());                                                \
    } while (0)

// stream timeline: mark a point of the step (profiling level 2 only; no effect otherwise)
#define URH_TL_MARK(ctx, label) do { if ((ctx)->profiling >= 2 && (ctx)->tl_ready && (ctx)->tl_count < 32) { \
        (ctx)->tl_name[(ctx)->tl_count] = (label); cudaEventRecord((ctx)->tl_ev[(ctx)->tl_count++], (ctx)->stream); } } while (0)
#define URH_TL_RESET(ctx) do { (ctx)->tl_count = 0; } while (0)
// record events around the dense kernel when profiling is enabled
#define URH_PROF_BEGIN(ctx) do { if ((ctx)->profiling) cudaEventRecord((ctx)->ev_k0, (ctx)->stream); } while (0)
#define URH_PROF_END(ctx) do { if ((ctx)->profiling) { cudaEventRecord((ctx)->ev_k1, (ctx)->stream); (ctx)->dense_timed = 1; } } while (0)

// ---- arena ----------------------------------------------------------------------------------------
void urh_arena_reset(urh_ctx* ctx);
int urh_arena_alloc(urh_ctx* ctx, size_t bytes, void** out);
template <typename T>
static inline int urh_arena(urh_ctx* ctx, size_t count, T** out) {
    void* p = nullptr;
    int rc = urh_arena_alloc(ctx, count * sizeof(T), &p);
    *out = (T*)p;
    return rc;
}
int urh_ensure_pulses(urh_ctx* ctx, size_t rows);
void urh_release_mod_plans(urh_ctx* ctx);  // modulation.cu
int urh_ensure_stage(urh_ctx* ctx, size_t bytes);

// read back `count` int64 scalars from device (synchronises the ctx stream)
int urh_read_i64(urh_ctx* ctx, const int64_t* d_src, int count, int64_t* h_out);

__host__ __device__ static inline int64_t urh_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// sample size in bytes of one IQ pair for dtype
static inline int urh_iq_bytes(int dtype) {
    switch (dtype) {
        case URH_DT_I8:
        case URH_DT_U8: return 2;
        case URH_DT_I16:
        case URH_DT_U16: return 4;
        case URH_DT_F32: return 8;
        default: return 0;
    }
}

// NOISE sentinel per modulation (signal_functions.pyx:31-44)
static inline float urh_noise_value(int mod_type) {
    switch (mod_type) {
        case URH_MOD_ASK: return 0.0f;
        case URH_MOD_FSK:
        case URH_MOD_PSK:
        case URH_MOD_OQPSK: return -4.0f;
        case URH_MOD_QAM: return 0.0f * -4.0f;  // NOISE_ASK * NOISE_FSK_PSK = -0.0
        default: return 0.0f;
    }
}
