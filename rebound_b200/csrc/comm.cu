// comm.cu -- the exchange step of the sharded hot path, inside the engine.
//
// Rank r of W owns the contiguous index block [N*r/W, N*(r+1)/W) (rebcu_set_shard): force, kick/drift and boundary
// kernels touch only that block, and between the drift and the force evaluation every rank needs the other blocks'
// new positions.  That all-gather of x,y,z (24 B per particle) is the one real exchange of the path -- the role the
// reference's MPI build gives to reb_communication_mpi_distribute_particles inside the force call
// (src/gravity.c:58-61, src/communication_mpi.c:98-181) and to its essential-tree exchange (:354-438).  A collision
// search also needs the other blocks' velocities, an open-boundary removal every field (the compaction shifts
// particles across block borders); engine_exchange() is told which.
//
// Two transports, chosen when the communicator is created:
//   NCCL    one communicator per handle (ncclCommInitRank from a unique id for one process per GPU, ncclCommInitAll
//           for several GPUs driven by the threads of one process).  Equal blocks: one in-place ncclAllGather per
//           field; ragged blocks: equal-sized slots of a staging buffer + one ncclAllGather per array + device copies
//           into place.  All calls of an exchange form one NCCL group
//           and are enqueued on the handle's own stream, i.e. ordered behind the drift and in front of the force
//           kernel with no host synchronisation.
//   LOCAL   handles of ONE process that may share a device (NCCL refuses two ranks on one GPU): every handle pulls the
//           owners' blocks with peer copies on its own stream, fenced by events and two host barriers.  This is what
//           lets the sharded kernels be tested on a single-GPU box, and the fallback when NCCL cannot initialise.
// The transport moves 64-bit words (bit-exact also for the pointer-tag fields).
#include "engine.cuh"
#include <nccl.h>
#include <pthread.h>
#include <stdlib.h>

struct LocalGroup {
    int n = 0;
    rebcu_handle* hs[REBCU_MAX_RANKS] = {};
    cudaEvent_t ready[REBCU_MAX_RANKS] = {}, pulled[REBCU_MAX_RANKS] = {};
    pthread_barrier_t bar;
    int refs = 0;
    pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
};

struct EngineComm {
    int kind = 0;                  // 1 NCCL, 2 LOCAL
    ncclComm_t nccl = nullptr;
    LocalGroup* grp = nullptr;
    uint64_t bytes = 0, calls = 0; // words received by this rank x 8, exchanges
    char* stage = nullptr; size_t stage_bytes = 0;   // NCCL, ragged ranges: equal-sized slots for one ncclAllGather per array
};

static int nccl_fail(rebcu_handle* h, ncclResult_t r, const char* where) {
    char buf[480];
    snprintf(buf, sizeof(buf), "NCCL error %d (%s) at %s", (int)r, ncclGetErrorString(r), where);
    return rebcu_fail(h, REBCU_ERR_CUDA, buf);
}
#define NCCL_TRY(h, expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) return nccl_fail((h), _r, #expr); } while (0)

static int exchange_fields(int need, int* fields) {
    int n = 0;
    if (need & REBCU_EXCHANGE_ALL) { for (int k = 0; k < F_COUNT; k++) fields[n++] = k; return n; }
    fields[n++] = F_X; fields[n++] = F_Y; fields[n++] = F_Z;
    if (need & REBCU_EXCHANGE_VELOCITIES) { fields[n++] = F_VX; fields[n++] = F_VY; fields[n++] = F_VZ; }
    return n;
}

// All-gathers rank-owned ranges of device arrays.  Array a has elements of `bytes[a]` bytes; rank r owns the elements
// [bounds[r], bounds[r+1]) of every array (same bounds for all arrays of a call); afterwards every rank holds every
// owner's range.  In place.  Equal ranges use ncclAllGather directly; ragged ranges go through equal-sized staging slots
// (one ncclAllGather per array + device copies into place); all NCCL calls of one exchange form one group.
static int gather_ranges(rebcu_handle* h, void** ptrs, const int* bytes, int n_arrays, const uint64_t* bounds) {
    EngineComm* C = h->comm;
    const int W = h->world, me = h->rank;
    if (W <= 1 || bounds[W] == bounds[0]) return REBCU_OK;
    if (n_arrays > F_COUNT) return rebcu_fail(h, REBCU_ERR_ARG, "too many arrays in one exchange");
    LaunchScope ls(h, TC_EXCHANGE, 0);
    C->calls++;
    bool even = bounds[0] == 0;
    for (int r = 0; r < W; r++) if (bounds[r + 1] - bounds[r] != bounds[1] - bounds[0]) even = false;
    for (int a = 0; a < n_arrays; a++) C->bytes += (uint64_t)bytes[a] * ((bounds[W] - bounds[0]) - (bounds[me + 1] - bounds[me]));
    if (C->kind == 1) {
        size_t slot[F_COUNT] = {}, stage_off[F_COUNT] = {};
        if (!even) {
            uint64_t longest = 0;
            for (int r = 0; r < W; r++) longest = (bounds[r + 1] - bounds[r] > longest) ? bounds[r + 1] - bounds[r] : longest;
            size_t need = 0;
            for (int a = 0; a < n_arrays; a++) { slot[a] = ((size_t)longest * bytes[a] + 255) & ~(size_t)255; stage_off[a] = need; need += slot[a] * W; }
            if (C->stage_bytes < need) {
                CU_TRY(h, cudaStreamSynchronize(h->stream));
                cudaFree(C->stage); C->stage = nullptr; C->stage_bytes = 0;
                CU_TRY(h, cudaMalloc(&C->stage, need + need / 8));
                C->stage_bytes = need + need / 8;
            }
        }
        NCCL_TRY(h, ncclGroupStart());
        for (int a = 0; a < n_arrays; a++) {
            char* base = (char*)ptrs[a];
            const uint64_t eb = (uint64_t)bytes[a];
            if (even) {
                NCCL_TRY(h, ncclAllGather(base + bounds[me] * eb, base, (bounds[1] - bounds[0]) * eb, ncclChar, C->nccl, h->stream));
            } else {
                // ragged ranges: every rank's range goes into an equal-sized slot of a staging buffer, ONE ncclAllGather
                // moves the slots, device copies put them in place.  (One ncclBroadcast per owner, the first version,
                // reached 150 GB/s on 8 GPUs where the all-gather reaches several times that.)
                char* st = C->stage + stage_off[a];
                if (bounds[me + 1] > bounds[me])
                    CU_TRY(h, cudaMemcpyAsync(st + (size_t)me * slot[a], base + bounds[me] * eb, (bounds[me + 1] - bounds[me]) * eb, cudaMemcpyDeviceToDevice, h->stream));
                NCCL_TRY(h, ncclAllGather(st + (size_t)me * slot[a], st, slot[a], ncclChar, C->nccl, h->stream));
            }
        }
        NCCL_TRY(h, ncclGroupEnd());
        if (!even)
            for (int a = 0; a < n_arrays; a++) {
                char* base = (char*)ptrs[a];
                const uint64_t eb = (uint64_t)bytes[a];
                const char* st = C->stage + stage_off[a];
                for (int r = 0; r < W; r++)
                    if (r != me && bounds[r + 1] > bounds[r])
                        CU_TRY(h, cudaMemcpyAsync(base + bounds[r] * eb, st + (size_t)r * slot[a], (bounds[r + 1] - bounds[r]) * eb, cudaMemcpyDeviceToDevice, h->stream));
            }
        return REBCU_OK;
    }
    LocalGroup* G = C->grp;
    CU_TRY(h, cudaEventRecord(G->ready[me], h->stream));
    for (int a = 0; a < n_arrays; a++) h->comm_view[a] = (uint64_t*)ptrs[a];      // what the peers read from this handle
    h->comm_view_n = n_arrays;
    pthread_barrier_wait(&G->bar);
    for (int r = 0; r < W; r++) {
        if (r == me || bounds[r + 1] == bounds[r]) continue;
        rebcu_handle* peer = G->hs[r];
        CU_TRY(h, cudaStreamWaitEvent(h->stream, G->ready[r], 0));
        for (int a = 0; a < n_arrays; a++) {
            const uint64_t eb = (uint64_t)bytes[a];
            CU_TRY(h, cudaMemcpyPeerAsync((char*)ptrs[a] + bounds[r] * eb, h->device, (const char*)peer->comm_view[a] + bounds[r] * eb, peer->device,
                                          (bounds[r + 1] - bounds[r]) * eb, h->stream));
        }
    }
    CU_TRY(h, cudaEventRecord(G->pulled[me], h->stream));
    pthread_barrier_wait(&G->bar);
    // an owner must not overwrite its range (next kick/drift, next build) before every peer has pulled it
    for (int r = 0; r < W; r++) if (r != me) CU_TRY(h, cudaStreamWaitEvent(h->stream, G->pulled[r], 0));
    return REBCU_OK;
}

int comm_exchange(rebcu_handle* h, int need) {
    int fields[F_COUNT], bytes[F_COUNT];
    const int nf = exchange_fields(need, fields);
    void* ptrs[F_COUNT];
    for (int k = 0; k < nf; k++) { ptrs[k] = h->tag(fields[k]); bytes[k] = 8; }
    uint64_t bounds[REBCU_MAX_RANKS + 1];
    for (int r = 0; r <= h->world; r++) bounds[r] = h->N * (uint64_t)r / (uint64_t)h->world;
    return gather_ranges(h, ptrs, bytes, nf, bounds);
}

int comm_gather_ranges(rebcu_handle* h, void** ptrs, const int* bytes, int n_arrays, const uint64_t* bounds) {
    if (!h->comm) return rebcu_fail(h, REBCU_ERR_ARG, "no communicator");
    return gather_ranges(h, ptrs, bytes, n_arrays, bounds);
}

static void comm_release(rebcu_handle* h) {
    EngineComm* C = h->comm;
    if (!C) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (C->kind == 1 && C->nccl) ncclCommDestroy(C->nccl);
    cudaFree(C->stage);
    if (C->kind == 2 && C->grp) {
        LocalGroup* G = C->grp;
        pthread_mutex_lock(&G->lock);
        const int left = --G->refs;
        pthread_mutex_unlock(&G->lock);
        if (left == 0) {
            for (int r = 0; r < G->n; r++) { if (G->ready[r]) cudaEventDestroy(G->ready[r]); if (G->pulled[r]) cudaEventDestroy(G->pulled[r]); }
            pthread_barrier_destroy(&G->bar);
            delete G;
        }
    }
    delete C;
    h->comm = nullptr;
    h->rank = 0; h->world = 1;
}

void comm_free(rebcu_handle* h) { comm_release(h); }

extern "C" {

int rebcu_comm_unique_id(void* out128) {
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return REBCU_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out128, &id, sizeof(id));
    return REBCU_OK;
}

int rebcu_comm_init_rank(rebcu_handle* h, const void* id128, int rank, int world) {
    if (world < 1 || world > REBCU_MAX_RANKS || rank < 0 || rank >= world) return rebcu_fail(h, REBCU_ERR_ARG, "invalid rank / world");
    comm_release(h);
    CU_TRY(h, cudaSetDevice(h->device));
    h->rank = rank; h->world = world;
    if (world == 1) return REBCU_OK;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    EngineComm* C = new EngineComm();
    C->kind = 1;
    ncclResult_t r = ncclCommInitRank(&C->nccl, world, id, rank);
    if (r != ncclSuccess) { delete C; h->rank = 0; h->world = 1; return nccl_fail(h, r, "ncclCommInitRank"); }
    h->comm = C;
    return REBCU_OK;
}

int rebcu_comm_init_all(rebcu_handle** hs, int n, int transport) {
    if (n < 1 || n > REBCU_MAX_RANKS) return REBCU_ERR_ARG;
    for (int r = 0; r < n; r++) { comm_release(hs[r]); hs[r]->rank = r; hs[r]->world = n; }
    if (n == 1) return REBCU_OK;
    bool distinct = true;
    for (int r = 0; r < n; r++) for (int q = 0; q < r; q++) if (hs[r]->device == hs[q]->device) distinct = false;
    if (transport == REBCU_TRANSPORT_AUTO) {
        const char* e = getenv("REBOUND_B200_COMM");
        transport = (e && strcmp(e, "local") == 0) ? REBCU_TRANSPORT_LOCAL : (distinct ? REBCU_TRANSPORT_NCCL : REBCU_TRANSPORT_LOCAL);
    }
    if (transport == REBCU_TRANSPORT_NCCL) {
        if (!distinct) return rebcu_fail(hs[0], REBCU_ERR_ARG, "NCCL needs one device per rank; use the LOCAL transport for handles that share a device");
        ncclComm_t comms[REBCU_MAX_RANKS];
        int devs[REBCU_MAX_RANKS];
        for (int r = 0; r < n; r++) devs[r] = hs[r]->device;
        ncclResult_t res = ncclCommInitAll(comms, n, devs);
        if (res != ncclSuccess) { for (int r = 0; r < n; r++) { hs[r]->rank = 0; hs[r]->world = 1; } return nccl_fail(hs[0], res, "ncclCommInitAll"); }
        for (int r = 0; r < n; r++) { EngineComm* C = new EngineComm(); C->kind = 1; C->nccl = comms[r]; hs[r]->comm = C; }
        return REBCU_OK;
    }
    LocalGroup* G = new LocalGroup();
    G->n = n; G->refs = n;
    pthread_barrier_init(&G->bar, nullptr, (unsigned)n);
    for (int r = 0; r < n; r++) {
        G->hs[r] = hs[r];
        cudaSetDevice(hs[r]->device);
        if (cudaEventCreateWithFlags(&G->ready[r], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&G->pulled[r], cudaEventDisableTiming) != cudaSuccess) return rebcu_fail(hs[0], REBCU_ERR_CUDA, "event creation failed");
        for (int q = 0; q < n; q++) {
            if (hs[q]->device == hs[r]->device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, hs[r]->device, hs[q]->device);
            if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(hs[q]->device, 0); if (e != cudaSuccess) cudaGetLastError(); }
        }
        EngineComm* C = new EngineComm(); C->kind = 2; C->grp = G; hs[r]->comm = C;
    }
    return REBCU_OK;
}

int rebcu_comm_destroy(rebcu_handle* h) { comm_release(h); return REBCU_OK; }

int rebcu_comm_stats(const rebcu_handle* h, uint64_t* bytes_received, uint64_t* exchanges, int* transport) {
    const EngineComm* C = h->comm;
    *bytes_received = C ? C->bytes : 0;
    *exchanges = C ? C->calls : 0;
    if (transport) *transport = C ? C->kind : 0;
    return REBCU_OK;
}

}  // extern "C"
