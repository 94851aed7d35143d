// context.cu -- handle life cycle, residency (AoS <-> SoA), timing instrumentation.
#include "engine.cuh"
#include <math.h>
#include <stdlib.h>

int rebcu_fail(rebcu_handle* h, int code, const char* msg) {
    if (h) { strncpy(h->err, msg, sizeof(h->err) - 1); h->err[sizeof(h->err) - 1] = 0; }
    return code;
}
int rebcu_cuda_fail(rebcu_handle* h, cudaError_t e, const char* where) {
    char buf[480];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
    return rebcu_fail(h, REBCU_ERR_CUDA, buf);
}

LaunchScope::LaunchScope(rebcu_handle* h_, int cls, int n_launches) : h(h_), idx(-1) {
    h->launches += n_launches;
    if (!h->timing) return;
    TimedRange r; r.cls = cls;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
        if (!h->event_pool.empty()) { *e = h->event_pool.back(); h->event_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(r.a, h->stream);
    h->ranges.push_back(r);
    idx = (int)h->ranges.size() - 1;
}
LaunchScope::~LaunchScope() {
    if (idx >= 0) cudaEventRecord(h->ranges[idx].b, h->stream);
}

extern "C" {

int rebcu_version(void) { return 100; }

int rebcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

rebcu_handle* rebcu_create(int device, void* stream) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    rebcu_handle* h = new rebcu_handle();
    h->device = device;
    if (stream) { h->stream = (cudaStream_t)stream; h->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return nullptr; }
        h->own_stream = true;
    }
    bool ok = cudaMalloc(&h->ghosts_dev, sizeof(GhostShifts)) == cudaSuccess
           && cudaMalloc(&h->scratch, 64 * sizeof(double)) == cudaSuccess
           && cudaMalloc(&h->counters, (16 + REBCU_MAX_RANKS) * sizeof(unsigned long long)) == cudaSuccess
           && cudaMalloc(&h->scratch_big, 2 * 7 * 256 * sizeof(double)) == cudaSuccess
           && cudaMallocHost(&h->pinned, 32 * sizeof(unsigned long long)) == cudaSuccess
           && cudaMallocHost(&h->ghost_ring, GHOST_RING * sizeof(GhostShifts)) == cudaSuccess;
    for (int i = 0; ok && i < GHOST_RING; i++) ok = cudaEventCreateWithFlags(&h->ghost_ring_ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { rebcu_destroy(h); return nullptr; }
    return h;
}

void rebcu_destroy(rebcu_handle* h) {
    if (!h) return;
    if (h->group) group_destroy(h);          // leader: stops the workers and releases the other ranks' handles
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    comm_free(h);
    tree_free(h);
    cudaFree(h->soa); cudaFree(h->aos); cudaFree(h->ghosts_dev); cudaFree(h->scratch); cudaFree(h->counters); cudaFree(h->scratch_big);
    cudaFree(h->col_count); cudaFree(h->col_off); cudaFree(h->col_list); cudaFree(h->col_scan_tmp); cudaFree(h->col_slots); cudaFree(h->col_map);
    cudaFree(h->compact_tmp); cudaFree(h->compact_buf); cudaFree(h->compact_flag); cudaFree(h->compact_pos);
    cudaFree(h->tp_hist);
    cudaFree(h->diag_partial);
    cudaFree(h->row_buf);
    cudaFree(h->resolve_buf);
    cudaFree(h->gravity_cs);
    for (int k = 0; k < AUX_STREAMS; k++) if (h->aux[k]) cudaStreamDestroy(h->aux[k]);
    for (int k = 0; k < 3; k++) if (h->aux_ev[k]) cudaEventDestroy(h->aux_ev[k]);
    for (int k = 0; k < 3 * PIPE_RANGES; k++) if (h->pipe_ev[k]) cudaEventDestroy(h->pipe_ev[k]);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->pairs_host) cudaFreeHost(h->pairs_host);
    if (h->ghost_ring) cudaFreeHost(h->ghost_ring);
    for (int i = 0; i < GHOST_RING; i++) if (h->ghost_ring_ev[i]) cudaEventDestroy(h->ghost_ring_ev[i]);
    for (auto& r : h->ranges) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto& e : h->event_pool) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* rebcu_last_error(const rebcu_handle* h) { return h ? h->err : "invalid handle (no CUDA device?)"; }
void* rebcu_stream(const rebcu_handle* h) { return (void*)h->stream; }

int rebcu_synchronize(rebcu_handle* h) {
    if (group_active(h)) return group_run(h, [](rebcu_handle* s, int) { return rebcu_synchronize(s); });
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return REBCU_OK;
}

int rebcu_host_register(void* ptr, uint64_t bytes) {
    return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? REBCU_OK : REBCU_ERR_CUDA;
}
int rebcu_host_unregister(void* ptr) {
    return cudaHostUnregister(ptr) == cudaSuccess ? REBCU_OK : REBCU_ERR_CUDA;
}

uint64_t rebcu_N(const rebcu_handle* h) { return h->N; }

void* rebcu_device_field(rebcu_handle* h, int field) {
    if (!h->resident || field < 0 || field >= F_COUNT) return nullptr;
    return (void*)h->f(field);
}

int rebcu_exchange_request(const rebcu_handle* h) { return h->exchange_need; }

uint64_t rebcu_launch_count(const rebcu_handle* h) { return h->launches; }

int rebcu_timing_enable(rebcu_handle* h, int on) { h->timing = on != 0; return REBCU_OK; }

int rebcu_timing_reset(rebcu_handle* h) {
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    for (auto& r : h->ranges) { h->event_pool.push_back(r.a); h->event_pool.push_back(r.b); }
    h->ranges.clear();
    return REBCU_OK;
}

int rebcu_timing_read(rebcu_handle* h, double* ms_out, uint64_t* launches_out, int n_classes) {
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n_classes; i++) { ms_out[i] = 0; if (launches_out) launches_out[i] = 0; }
    for (auto& r : h->ranges) {
        float ms = 0;
        CU_TRY(h, cudaEventElapsedTime(&ms, r.a, r.b));
        if (r.cls < n_classes) { ms_out[r.cls] += ms; if (launches_out) launches_out[r.cls]++; }
    }
    return REBCU_OK;
}

int rebcu_set_shard(rebcu_handle* h, int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) return rebcu_fail(h, REBCU_ERR_ARG, "invalid shard");
    h->rank = rank; h->world = world;
    return REBCU_OK;
}
void rebcu_shard_range(const rebcu_handle* h, uint64_t* b, uint64_t* e) { engine_shard(h, b, e); }

}  // extern "C"

void engine_shard(const rebcu_handle* h, uint64_t* b, uint64_t* e) {
    *b = h->N * (uint64_t)h->rank / (uint64_t)h->world;
    *e = h->N * (uint64_t)(h->rank + 1) / (uint64_t)h->world;
}

// Grows the resident buffers to hold n particles (contents are NOT preserved).
int engine_reserve(rebcu_handle* h, uint64_t n) {
    if (n <= h->cap) return REBCU_OK;
    uint64_t cap = ((n + 1023) / 1024) * 1024;   // 8 KiB-aligned array starts
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->soa); cudaFree(h->aos);
    h->soa = nullptr; h->aos = nullptr; h->cap = 0; h->resident = false;
    CU_TRY(h, cudaMalloc(&h->soa, cap * F_COUNT * sizeof(double)));
    CU_TRY(h, cudaMalloc(&h->aos, cap * sizeof(rebcu_particle)));
    h->cap = cap;
    return REBCU_OK;
}

// ---- ghost boxes: src/boundary.c:145-201, evaluated on the host with the host libm fmod ---------
static rebcu_vec6d ghostbox_host(const rebcu_config* c, int i, int j, int k) {
    const double bx = c->root_size * (double)c->N_root_x;
    const double by = c->root_size * (double)c->N_root_y;
    const double bz = c->root_size * (double)c->N_root_z;
    rebcu_vec6d gb = {0, 0, 0, 0, 0, 0};
    if (c->boundary == REBCU_BOUNDARY_OPEN || c->boundary == REBCU_BOUNDARY_PERIODIC) {
        gb.x = bx * (double)i; gb.y = by * (double)j; gb.z = bz * (double)k;
    } else if (c->boundary == REBCU_BOUNDARY_SHEAR) {
        gb.vy = -1.5 * (double)i * c->OMEGA * bx;
        double shift;
        if (i == 0) shift = -fmod(gb.vy * c->t, by);
        else if (i > 0) shift = -fmod(gb.vy * c->t - by / 2., by) - by / 2.;
        else shift = -fmod(gb.vy * c->t + by / 2., by) + by / 2.;
        gb.x = bx * (double)i; gb.y = by * (double)j - shift; gb.z = bz * (double)k;
    }
    return gb;
}

// Offsets for the rings [-gx,gx] x [-gy,gy] x [-gz,gz] in the reference's loop order (x outermost).
// More boxes than REBCU_MAX_GHOST: out->n = -1 (the callers report REBCU_ERR_ARG through engine_upload_ghosts).
void engine_ghost_shifts(const rebcu_config* c, int gx, int gy, int gz, GhostShifts* out) {
    if ((long long)(2 * gx + 1) * (2 * gy + 1) * (2 * gz + 1) > REBCU_MAX_GHOST || gx < 0 || gy < 0 || gz < 0) { out->n = -1; return; }
    int n = 0;
    for (int i = -gx; i <= gx; i++)
        for (int j = -gy; j <= gy; j++)
            for (int k = -gz; k <= gz; k++)
                if (n < REBCU_MAX_GHOST) out->gb[n++] = ghostbox_host(c, i, j, k);
    out->n = n;
}

int engine_upload_ghosts(rebcu_handle* h, const GhostShifts* g) {
    if (g->n < 0) return rebcu_fail(h, REBCU_ERR_ARG, "too many ghost boxes: (2*N_ghost_x+1)(2*N_ghost_y+1)(2*N_ghost_z+1) must not exceed 729");
    // Skip the copy when the device already holds these offsets (always the case without a shear
    // boundary); otherwise stage through a pinned ring so that no stream synchronisation is needed.
    const size_t bytes = offsetof(GhostShifts, gb) + sizeof(rebcu_vec6d) * (size_t)g->n;
    if (h->ghosts_valid && h->ghosts_host.n == g->n && memcmp(&h->ghosts_host, g, bytes) == 0) return REBCU_OK;
    memcpy(&h->ghosts_host, g, bytes);
    const int slot = h->ghost_ring_next;
    h->ghost_ring_next = (slot + 1) % GHOST_RING;
    if (h->ghost_ring_used[slot]) CU_TRY(h, cudaEventSynchronize(h->ghost_ring_ev[slot]));
    memcpy(&h->ghost_ring[slot], g, bytes);
    CU_TRY(h, cudaMemcpyAsync(h->ghosts_dev, &h->ghost_ring[slot], bytes, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaEventRecord(h->ghost_ring_ev[slot], h->stream));
    h->ghost_ring_used[slot] = true;
    h->ghosts_valid = true;
    return REBCU_OK;
}

// ---- AoS <-> SoA -----------------------------------------------------------------------------
// One warp moves 32 records = 3584 contiguous bytes of AoS through shared memory so that both the
// AoS side (16-byte vectors) and the SoA side (8-byte words, 14 arrays) are fully coalesced.
constexpr int PACK_THREADS = 256;
constexpr int WORDS = 14;   // 8-byte words per record

__global__ void __launch_bounds__(PACK_THREADS) unpack_kernel(const uint64_t* __restrict__ aos, uint64_t* __restrict__ soa,
                                                              uint64_t cap, uint64_t n) {
    __shared__ uint64_t tile[PACK_THREADS * WORDS];
    const uint64_t base = (uint64_t)blockIdx.x * PACK_THREADS;
    const uint64_t cnt = min((uint64_t)PACK_THREADS, n - base);
    const uint64_t words = cnt * WORDS;
    for (uint64_t w = threadIdx.x; w < words; w += PACK_THREADS) tile[w] = aos[base * WORDS + w];
    __syncthreads();
    if (threadIdx.x < cnt) {
#pragma unroll
        for (int k = 0; k < WORDS; k++) soa[(uint64_t)k * cap + base + threadIdx.x] = tile[threadIdx.x * WORDS + k];
    }
}

__global__ void __launch_bounds__(PACK_THREADS) pack_kernel(uint64_t* __restrict__ aos, const uint64_t* __restrict__ soa,
                                                            uint64_t cap, uint64_t n) {
    __shared__ uint64_t tile[PACK_THREADS * WORDS];
    const uint64_t base = (uint64_t)blockIdx.x * PACK_THREADS;
    const uint64_t cnt = min((uint64_t)PACK_THREADS, n - base);
    if (threadIdx.x < cnt) {
#pragma unroll
        for (int k = 0; k < WORDS; k++) tile[threadIdx.x * WORDS + k] = soa[(uint64_t)k * cap + base + threadIdx.x];
    }
    __syncthreads();
    const uint64_t words = cnt * WORDS;
    for (uint64_t w = threadIdx.x; w < words; w += PACK_THREADS) aos[base * WORDS + w] = tile[w];
}

// Range versions on an arbitrary stream, for the chunk-pipelined host-buffer path (integrate.cu).
// Runs the caller's exchange callback with the set of fields it has to gather (rebcu_exchange_request).
int engine_exchange(rebcu_handle* h, int need) {
    if (h->comm) return comm_exchange(h, need);        // native transport (comm.cu)
    if (!h->exchange) return REBCU_OK;
    h->exchange_need = need;
    h->exchange(h->exchange_user);
    h->exchange_need = REBCU_EXCHANGE_POSITIONS;
    return REBCU_OK;
}

// b must be a multiple of PACK_THREADS.  The copy runs on s_copy, the AoS<->SoA kernel on s_kernel, chained by
// `ev`: a copy stream then carries nothing but back-to-back DMA transfers and never waits for an SM to free up.
int engine_upload_range(rebcu_handle* h, cudaStream_t s_copy, cudaEvent_t ev, cudaStream_t s_kernel, const rebcu_particle* particles, uint64_t b, uint64_t e) {
    if (e <= b) return REBCU_OK;
    if (particles) CU_TRY(h, cudaMemcpyAsync(h->aos + b, particles + b, (e - b) * sizeof(rebcu_particle), cudaMemcpyHostToDevice, s_copy));
    if (s_copy != s_kernel) { CU_TRY(h, cudaEventRecord(ev, s_copy)); CU_TRY(h, cudaStreamWaitEvent(s_kernel, ev, 0)); }
    h->launches++;
    unpack_kernel<<<div_up(e - b, PACK_THREADS), PACK_THREADS, 0, s_kernel>>>((const uint64_t*)(h->aos + b), (uint64_t*)h->soa + b, h->cap, e - b);
    CU_TRY(h, cudaGetLastError());
    return REBCU_OK;
}

int engine_download_range(rebcu_handle* h, cudaStream_t s_kernel, cudaEvent_t ev, cudaStream_t s_copy, rebcu_particle* particles, uint64_t b, uint64_t e) {
    if (e <= b) return REBCU_OK;
    h->launches++;
    pack_kernel<<<div_up(e - b, PACK_THREADS), PACK_THREADS, 0, s_kernel>>>((uint64_t*)(h->aos + b), (const uint64_t*)h->soa + b, h->cap, e - b);
    CU_TRY(h, cudaGetLastError());
    if (s_copy != s_kernel) { CU_TRY(h, cudaEventRecord(ev, s_kernel)); CU_TRY(h, cudaStreamWaitEvent(s_copy, ev, 0)); }
    if (particles) CU_TRY(h, cudaMemcpyAsync(particles + b, h->aos + b, (e - b) * sizeof(rebcu_particle), cudaMemcpyDeviceToHost, s_copy));
    return REBCU_OK;
}

extern "C" {

int rebcu_upload(rebcu_handle* h, const rebcu_particle* particles, uint64_t N) {
    if (group_active(h)) return group_upload(h, particles, N);
    CU_TRY(h, cudaSetDevice(h->device));
    int err = engine_reserve(h, N);
    if (err) return err;
    h->N = N;
    h->resident = true;
    h->tree.built_for_n = -1;
    if (N == 0) return REBCU_OK;
    CU_TRY(h, cudaMemcpyAsync(h->aos, particles, N * sizeof(rebcu_particle), cudaMemcpyHostToDevice, h->stream));
    {
        LaunchScope ls(h, TC_PACK);
        unpack_kernel<<<div_up(N, PACK_THREADS), PACK_THREADS, 0, h->stream>>>((const uint64_t*)h->aos, (uint64_t*)h->soa, h->cap, N);
    }
    CU_TRY(h, cudaGetLastError());
    return REBCU_OK;
}

int rebcu_download(rebcu_handle* h, rebcu_particle* particles, uint64_t N) {
    if (group_active(h)) return group_download(h, particles, N);
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    if (N < h->N) return rebcu_fail(h, REBCU_ERR_CAPACITY, "host particle buffer too small");
    if (h->N == 0) return REBCU_OK;
    {
        LaunchScope ls(h, TC_PACK);
        pack_kernel<<<div_up(h->N, PACK_THREADS), PACK_THREADS, 0, h->stream>>>((uint64_t*)h->aos, (const uint64_t*)h->soa, h->cap, h->N);
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(particles, h->aos, h->N * sizeof(rebcu_particle), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return REBCU_OK;
}

int rebcu_download_acc(rebcu_handle* h, rebcu_particle* particles, uint64_t N) {
    // The staging AoS still holds the uploaded records, so a full pack + copy returns the caller's
    // x,v,m,r and pointer fields unchanged and the new ax,ay,az.
    return rebcu_download(h, particles, N);
}

}  // extern "C"
