// Metadata all-gather over NVLink peer memory: the sharded digitizer / center detection exchange a few
// int64 per rank several times per step.  Through NCCL each of those costs a staging copy, a collective launch and a copy
// back; here ONE tiny kernel per rank stores its payload straight into every peer's mailbox (cudaIpc-mapped device
// memory, stores travel over NVLink / NVSwitch), then spins on its own mailbox until every peer's entry carries the
// current sequence number and writes the gathered payloads into mapped host memory.
//   mailbox[slot][sender] = {seq, 6 x 8 bytes payload}, slot = seq mod RING.  Every exchange is a rendezvous of all ranks,
//   so a rank is never more than one exchange ahead of a peer: a slot is reused only RING exchanges later.
//   Payload stores, __threadfence_system(), then the seq store; the reader spins on seq (volatile), fences, reads.
//   The spin has a time-out (about 2 s): a missing peer turns into an error code, never into a hung GPU.
#include "common.cuh"

#define P2P_RING 64
#define P2P_MAXW 8
#define P2P_WORDS 6   // 48 bytes of payload

struct P2pSlot {
    unsigned long long seq;
    unsigned long long data[P2P_WORDS];
    unsigned long long pad;
};

struct P2pArgs {
    P2pSlot* peer[P2P_MAXW];
    P2pSlot* local;
    unsigned long long data[P2P_WORDS];
    unsigned long long seq;
    long long timeout;
    int rank, world, slot;
};

// ---- device-resident exchanges (stream-ordered, no host in between) ------------------------------------------------------
// The sharded digitizer / center chain gather a few dozen bytes per rank between kernels, five times per step, and sum one small
// histogram.  A NCCL collective costs tens of microseconds each at 8 ranks; these kernels do the same exchange through the
// mailboxes: every rank stores its payload into every peer's slot (NVLink stores), fence, flag; then waits for the peers' flags
// and copies the gathered payloads to the caller's device buffer.  A time-out raises the error word in mapped host memory
// (checked by the caller at its next synchronisation) instead of hanging the GPU.
#define P2P_DEV_BYTES 240
struct P2pDevSlot {
    unsigned long long seq, pad;
    unsigned long long data[P2P_DEV_BYTES / 8];
};
#define P2P_RED_RING 4
#define P2P_RED_WORDS 6000   // == CEN_MAX_BINS (center.cu)
struct P2pRedSlot {
    unsigned long long seq, pad;
    unsigned long long data[P2P_RED_WORDS];
};
#define P2P_HOST_BYTES ((size_t)P2P_RING * P2P_MAXW * sizeof(P2pSlot))
#define P2P_DEV_OFFSET P2P_HOST_BYTES
#define P2P_RED_OFFSET (P2P_DEV_OFFSET + (size_t)P2P_RING * P2P_MAXW * sizeof(P2pDevSlot))
#define P2P_TOTAL_BYTES (P2P_RED_OFFSET + (size_t)P2P_RED_RING * P2P_MAXW * sizeof(P2pRedSlot))

struct P2pDevArgs {
    char* peer[P2P_MAXW];   // base of every rank's mailbox allocation (peer[rank] == local)
    unsigned long long seq;
    long long timeout;
    int rank, world;
};

// one warp per peer: store my payload into its slot, flag it; wait for that peer's flag in my mailbox, copy its payload out
__global__ void __launch_bounds__(32 * P2P_MAXW) k_p2p_allgather_dev(P2pDevArgs a, const unsigned long long* __restrict__ d_send,
                                                                    unsigned long long* __restrict__ d_recv, int words,
                                                                    unsigned long long* __restrict__ err) {
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (r >= a.world) return;
    const int slot = (int)(a.seq % P2P_RING);
    volatile P2pDevSlot* dst = (volatile P2pDevSlot*)(a.peer[r] + P2P_DEV_OFFSET) + slot * P2P_MAXW + a.rank;
    if (lane < words) dst->data[lane] = d_send[lane];
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        dst->seq = a.seq;
        __threadfence_system();
    }
    volatile P2pDevSlot* src = (volatile P2pDevSlot*)(a.peer[a.rank] + P2P_DEV_OFFSET) + slot * P2P_MAXW + r;
    if (lane == 0) {
        const long long t0 = clock64();
        while (src->seq != a.seq) {
            if (clock64() - t0 > a.timeout) { *err = 1ull; break; }
        }
    }
    __syncwarp();
    __threadfence_system();
    if (lane < words) d_recv[(size_t)r * words + lane] = src->data[lane];
}

// sum over ranks of min(*d_count, max_words) uint64 words: block b sends d_in to peer b; every block then waits for all the
// peers and sums its share of the words over the senders.  d_out must not alias d_in (other blocks are still sending it).
__global__ void __launch_bounds__(256) k_p2p_allreduce_dev(P2pDevArgs a, const unsigned long long* __restrict__ d_in,
                                                          unsigned long long* __restrict__ d_out, const long long* __restrict__ d_count,
                                                          int max_words, unsigned long long* __restrict__ err) {
    __shared__ int s_fail;
    long long cnt = d_count ? *d_count : (long long)max_words;
    if (cnt < 0) cnt = 0;
    if (cnt > max_words) cnt = max_words;
    const int slot = (int)(a.seq % P2P_RED_RING);
    const int b = blockIdx.x;
    if (threadIdx.x == 0) s_fail = 0;
    if (b < a.world) {
        volatile P2pRedSlot* dst = (volatile P2pRedSlot*)(a.peer[b] + P2P_RED_OFFSET) + slot * P2P_MAXW + a.rank;
        for (int i = threadIdx.x; i < (int)cnt; i += blockDim.x) dst->data[i] = d_in[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            dst->seq = a.seq;
            __threadfence_system();
        }
    }
    volatile P2pRedSlot* mine = (volatile P2pRedSlot*)(a.peer[a.rank] + P2P_RED_OFFSET) + slot * P2P_MAXW;
    if (threadIdx.x < a.world) {
        const long long t0 = clock64();
        while (mine[threadIdx.x].seq != a.seq) {
            if (clock64() - t0 > a.timeout) { *err = 1ull; s_fail = 1; break; }
        }
    }
    __syncthreads();
    __threadfence_system();
    if (s_fail) return;
    for (int i = b * blockDim.x + threadIdx.x; i < (int)cnt; i += gridDim.x * blockDim.x) {
        unsigned long long acc = 0ull;
        for (int r = 0; r < a.world; r++) acc += mine[r].data[i];
        d_out[i] = acc;
    }
}

__global__ void k_p2p_allgather(P2pArgs a, unsigned long long* __restrict__ hout) {
    const int r = threadIdx.x;
    if (r >= a.world) return;
    volatile P2pSlot* dst = a.peer[r] + a.slot * P2P_MAXW + a.rank;
#pragma unroll
    for (int i = 0; i < P2P_WORDS; i++) dst->data[i] = a.data[i];
    __threadfence_system();
    dst->seq = a.seq;
    __threadfence_system();   // push the flag out now: this kernel keeps running (spinning) for a while
    volatile P2pSlot* src = a.local + a.slot * P2P_MAXW + r;
    const long long t0 = clock64();
    while (src->seq != a.seq) {
        if (clock64() - t0 > a.timeout) {
            hout[P2P_MAXW * P2P_WORDS] = 1ull;   // error flag
            return;
        }
    }
    __threadfence_system();
#pragma unroll
    for (int i = 0; i < P2P_WORDS; i++) hout[r * P2P_WORDS + i] = src->data[i];
}

// Step 1 on every rank: allocate the mailbox, return its IPC handle (64 bytes) for the launcher plumbing to distribute.
extern "C" int urh_p2p_create(urh_ctx* ctx, char* out_handle64) {
    if (!ctx->p2p_local) {
        URH_CUDA(ctx, cudaMalloc(&ctx->p2p_local, P2P_TOTAL_BYTES));
        URH_CUDA(ctx, cudaMemset(ctx->p2p_local, 0, P2P_TOTAL_BYTES));
    }
    cudaIpcMemHandle_t h;
    URH_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->p2p_local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out_handle64, &h, 64);
    return URH_OK;
}

// Step 2: map every peer's mailbox.  handles = world x 64 bytes in rank order.
extern "C" int urh_p2p_open(urh_ctx* ctx, const char* handles, int rank, int world) {
    if (world < 1 || world > P2P_MAXW || rank < 0 || rank >= world) URH_FAIL(ctx, URH_ERR_INVALID, "p2p: world must be 1..8");
    if (!ctx->p2p_local) URH_FAIL(ctx, URH_ERR_INVALID, "urh_p2p_create must precede urh_p2p_open");
    for (int r = 0; r < world; r++) {
        if (r == rank) { ctx->p2p_peer[r] = ctx->p2p_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        URH_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->p2p_peer[r] = p;
    }
    if (!ctx->p2p_hout) {
        URH_CUDA(ctx, cudaHostAlloc(&ctx->p2p_hout, (P2P_MAXW * P2P_WORDS + 8) * sizeof(unsigned long long), cudaHostAllocMapped));
        memset(ctx->p2p_hout, 0, (P2P_MAXW * P2P_WORDS + 8) * sizeof(unsigned long long));
    }
    ctx->p2p_rank = rank;
    ctx->p2p_world = world;
    // p2p_seq is NOT reset: sequence numbers stay monotonic per context, so slots left over from an earlier session (or from a
    // timed-out exchange) can never satisfy a later wait.  All ranks open together (dist.init_p2p), so their counters stay equal.
    return URH_OK;
}

extern "C" int urh_p2p_close(urh_ctx* ctx) {
    for (int r = 0; r < ctx->p2p_world; r++)
        if (r != ctx->p2p_rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    ctx->p2p_world = 0;
    return URH_OK;
}

// all-gather of bytes_per_rank <= 48 host bytes per rank; h_recv = world x bytes_per_rank
extern "C" int urh_p2p_allgather_host(urh_ctx* ctx, const void* h_send, void* h_recv, size_t bytes_per_rank) {
    if (ctx->p2p_world <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "p2p mailboxes are not open");
    if (bytes_per_rank == 0 || bytes_per_rank > P2P_WORDS * sizeof(unsigned long long)) URH_FAIL(ctx, URH_ERR_INVALID, "p2p payload is 1..48 bytes");
    P2pArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < ctx->p2p_world; r++) a.peer[r] = (P2pSlot*)ctx->p2p_peer[r];
    a.local = (P2pSlot*)ctx->p2p_local;
    memcpy(a.data, h_send, bytes_per_rank);
    a.seq = ++ctx->p2p_seq;
    a.slot = (int)(a.seq % P2P_RING);
    a.rank = ctx->p2p_rank;
    a.world = ctx->p2p_world;
    a.timeout = 4000000000ll;   // ~2 s of SM clocks
    unsigned long long* hout = (unsigned long long*)ctx->p2p_hout;
    hout[P2P_MAXW * P2P_WORDS] = 0ull;
    unsigned long long* d_hout = nullptr;
    URH_CUDA(ctx, cudaHostGetDevicePointer((void**)&d_hout, hout, 0));
    URH_LAUNCH(ctx, k_p2p_allgather, 1, 32, 0, a, d_hout);
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hout[P2P_MAXW * P2P_WORDS]) {
        // the ranks may now disagree about the sequence: the mailboxes are unusable until re-opened; the caller falls back to NCCL
        urh_p2p_close(ctx);
        URH_FAIL(ctx, URH_ERR_CUDA, "p2p all-gather timed out waiting for a peer (exchange %llu); p2p closed", a.seq);
    }
    for (int r = 0; r < ctx->p2p_world; r++) memcpy((char*)h_recv + (size_t)r * bytes_per_rank, hout + r * P2P_WORDS, bytes_per_rank);
    return URH_OK;
}

static int p2p_dev_args(urh_ctx* ctx, P2pDevArgs* a) {
    if (ctx->p2p_world <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "p2p mailboxes are not open");
    memset(a, 0, sizeof(*a));
    for (int r = 0; r < ctx->p2p_world; r++) a->peer[r] = (char*)ctx->p2p_peer[r];
    a->seq = ++ctx->p2p_seq;
    a->rank = ctx->p2p_rank;
    a->world = ctx->p2p_world;
    a->timeout = 10000000000ll;   // ~5 s of SM clocks
    return URH_OK;
}

static unsigned long long* p2p_err_word(urh_ctx* ctx) {
    unsigned long long* d = nullptr;
    cudaHostGetDevicePointer((void**)&d, ctx->p2p_hout, 0);
    return d + P2P_MAXW * P2P_WORDS + 1;
}

// the device exchanges may be used for this payload on this context
bool urh_p2p_usable(urh_ctx* ctx, size_t bytes_per_rank) {
    return ctx->p2p_world > 1 && ctx->p2p_world == ctx->nccl_world && ctx->p2p_rank == ctx->nccl_rank && bytes_per_rank > 0 &&
           bytes_per_rank <= P2P_DEV_BYTES && (bytes_per_rank & 7) == 0;
}

// all-gather of bytes_per_rank (a multiple of 8, <= 240) DEVICE bytes per rank into d_recv (world x bytes_per_rank), enqueued on the
// context stream; no synchronisation.  urh_p2p_check reports a time-out after the caller's next synchronisation.
extern "C" int urh_p2p_allgather_dev(urh_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank) {
    if (bytes_per_rank == 0 || bytes_per_rank > P2P_DEV_BYTES || (bytes_per_rank & 7)) URH_FAIL(ctx, URH_ERR_INVALID, "p2p device payload is 8..240 bytes, a multiple of 8");
    P2pDevArgs a;
    URH_CHECK(p2p_dev_args(ctx, &a));
    URH_LAUNCH(ctx, k_p2p_allgather_dev, 1, 32 * P2P_MAXW, 0, a, (const unsigned long long*)d_send, (unsigned long long*)d_recv,
               (int)(bytes_per_rank / 8), p2p_err_word(ctx));
    return URH_OK;
}

// d_out[i] = sum over ranks of d_in[i], i < min(*d_count, max_words) (d_count: device pointer or NULL = max_words; max_words <= 6000;
// the ranks must agree on the count); d_out must not alias d_in.  Enqueued on the context stream.
extern "C" int urh_p2p_allreduce_u64_dev(urh_ctx* ctx, const void* d_in, void* d_out, const int64_t* d_count, int max_words) {
    if (max_words <= 0 || max_words > P2P_RED_WORDS) URH_FAIL(ctx, URH_ERR_INVALID, "p2p device reduction holds 1..6000 words");
    if (d_in == d_out) URH_FAIL(ctx, URH_ERR_INVALID, "p2p device reduction must not run in place");
    P2pDevArgs a;
    URH_CHECK(p2p_dev_args(ctx, &a));
    URH_LAUNCH(ctx, k_p2p_allreduce_dev, P2P_MAXW, 256, 0, a, (const unsigned long long*)d_in, (unsigned long long*)d_out,
               (const long long*)d_count, max_words, p2p_err_word(ctx));
    return URH_OK;
}

// after a synchronisation: did a device exchange time out?  (then the mailboxes are closed: the ranks may disagree about the sequence)
extern "C" int urh_p2p_check(urh_ctx* ctx) {
    if (!ctx->p2p_hout) return URH_OK;
    unsigned long long* h = (unsigned long long*)ctx->p2p_hout + P2P_MAXW * P2P_WORDS + 1;
    if (*h) {
        *h = 0ull;
        urh_p2p_close(ctx);
        URH_FAIL(ctx, URH_ERR_CUDA, "p2p device exchange timed out waiting for a peer; p2p closed");
    }
    return URH_OK;
}
