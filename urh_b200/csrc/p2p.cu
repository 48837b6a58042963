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
        URH_CUDA(ctx, cudaMalloc(&ctx->p2p_local, (size_t)P2P_RING * P2P_MAXW * sizeof(P2pSlot)));
        URH_CUDA(ctx, cudaMemset(ctx->p2p_local, 0, (size_t)P2P_RING * P2P_MAXW * sizeof(P2pSlot)));
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
    if (!ctx->p2p_hout) URH_CUDA(ctx, cudaHostAlloc(&ctx->p2p_hout, (P2P_MAXW * P2P_WORDS + 8) * sizeof(unsigned long long), cudaHostAllocMapped));
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
