// Generic in-place device scan (any associative, not necessarily commutative operator) used by the
// sparse stages of the digitizer / segmenter.  Three-phase reduce-then-scan, recursive on the block
// aggregates.  The arrays scanned here are run/candidate tables — orders of magnitude smaller than the
// sample stream — so simplicity wins over single-pass look-back.
#pragma once
#include "common.cuh"

namespace urhscan {

constexpr int BLOCK = 256;
constexpr int ITEMS = 8;
constexpr int CHUNK = BLOCK * ITEMS;

// Phase 1: each block scans its chunk in place (inclusive or exclusive) and writes the chunk aggregate.
template <typename T, typename Op>
__global__ void __launch_bounds__(BLOCK) k_scan_chunks(T* data, T* aggregates, int64_t n, Op op, T identity,
                                                       int exclusive) {
    __shared__ T s_part[BLOCK];
    const int64_t base = (int64_t)blockIdx.x * CHUNK + (int64_t)threadIdx.x * ITEMS;
    T v[ITEMS];
    T acc = identity;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int64_t idx = base + i;
        v[i] = (idx < n) ? data[idx] : identity;
        acc = op(acc, v[i]);
    }
    s_part[threadIdx.x] = acc;
    __syncthreads();
    // Hillis–Steele inclusive scan over the per-thread aggregates (order-preserving)
    for (int off = 1; off < BLOCK; off <<= 1) {
        T other = identity;
        const bool take = threadIdx.x >= off;
        if (take) other = s_part[threadIdx.x - off];
        __syncthreads();
        if (take) s_part[threadIdx.x] = op(other, s_part[threadIdx.x]);
        __syncthreads();
    }
    T run = (threadIdx.x == 0) ? identity : s_part[threadIdx.x - 1];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int64_t idx = base + i;
        const T incl = op(run, v[i]);
        if (idx < n) data[idx] = exclusive ? run : incl;
        run = incl;
    }
    if (threadIdx.x == BLOCK - 1 && aggregates) aggregates[blockIdx.x] = s_part[BLOCK - 1];
}

// Phase 3: fold the (exclusive) prefix of the preceding chunks into every element of a chunk.
template <typename T, typename Op>
__global__ void __launch_bounds__(BLOCK) k_scan_apply(T* data, const T* chunk_prefix, int64_t n, Op op) {
    const T pre = chunk_prefix[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * CHUNK;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int64_t idx = base + (int64_t)i * BLOCK + threadIdx.x;
        if (idx < n) data[idx] = op(pre, data[idx]);
    }
}

// In-place scan of data[0..n).  If d_total is non-null the reduction of all elements is written there.
template <typename T, typename Op>
int device_scan(urh_ctx* ctx, T* data, int64_t n, Op op, T identity, bool exclusive, T* d_total) {
    if (n <= 0) {
        if (d_total) {
            URH_CUDA(ctx, cudaMemcpyAsync(d_total, &identity, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
            URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        return URH_OK;
    }
    const int64_t chunks = urh_div_up(n, CHUNK);
    T* agg = nullptr;
    URH_CHECK(urh_arena<T>(ctx, (size_t)chunks + 1, &agg));
    URH_LAUNCH(ctx, (k_scan_chunks<T, Op>), (unsigned)chunks, BLOCK, 0, data, agg, n, op, identity, exclusive ? 1 : 0);
    if (chunks > 1) {
        // exclusive scan of the aggregates; its total is the grand total
        URH_CHECK((device_scan<T, Op>(ctx, agg, chunks, op, identity, true, d_total)));
        URH_LAUNCH(ctx, (k_scan_apply<T, Op>), (unsigned)chunks, BLOCK, 0, data, agg, n, op);
    } else if (d_total) {
        URH_CUDA(ctx, cudaMemcpyAsync(d_total, agg, sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return URH_OK;
}

struct AddI64 {
    __device__ __forceinline__ int64_t operator()(int64_t a, int64_t b) const { return a + b; }
};

}  // namespace urhscan
