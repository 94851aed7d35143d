// group.cu -- several GPUs behind ONE handle: what lets a program that only knows the reference's C API
// (reb_simulation_integrate on one struct reb_simulation, one thread) use all the GPUs of a node.
//
// rebcu_create_group(devices, n) creates one engine handle per device, joins them with the native exchange
// (rebcu_comm_init_all: NCCL for distinct devices, the LOCAL transport when a device appears twice) and starts one
// worker thread per handle.  It returns the LEADER handle (rank 0).  Every hot-path entry point of the C ABI called on
// the leader fans out: each worker makes the same call on its own handle -- its target block [N r/W, N (r+1)/W), the
// exchange inside the call -- and the leader returns when all are done.  Host buffers are shared by construction (one
// process), so every worker moves only ITS block of r->particles over PCIe (rebcu_upload_shard / rebcu_download_shard);
// sharded collision lists are merged segment by segment into the reference's serial order.
// The reference's counterpart is its MPI build (src/communication_mpi.c), where the caller has to run one process per
// domain; here the caller's program does not change (the shim creates a group when REBOUND_B200_DEVICES names several
// devices, shim_common.c).
#include "engine.cuh"
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct rebcu_group {
    int n = 0;
    rebcu_handle* hs[REBCU_MAX_RANKS] = {};
    std::thread threads[REBCU_MAX_RANKS];
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    uint64_t epoch = 0;
    int pending = 0;
    bool stop = false;
    std::function<int(rebcu_handle*, int)> job;
    int results[REBCU_MAX_RANKS] = {};
};

static thread_local bool tls_group_worker = false;

bool group_active(const rebcu_handle* h) { return h && h->group && !tls_group_worker; }

static void group_worker(rebcu_group* g, int r) {
    tls_group_worker = true;
    cudaSetDevice(g->hs[r]->device);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(rebcu_handle*, int)> fn;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_start.wait(lk, [&] { return g->stop || g->epoch != seen; });
            if (g->stop) return;
            seen = g->epoch;
            fn = g->job;
        }
        const int res = fn(g->hs[r], r);
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->results[r] = res;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// Runs fn(handle of rank r, r) on every worker; returns the first non-zero result (rank order) and copies that rank's
// error text to the leader.
int group_run(rebcu_handle* leader, const std::function<int(rebcu_handle*, int)>& fn) {
    rebcu_group* g = leader->group;
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->job = fn;
        g->pending = g->n;
        g->epoch++;
        g->cv_start.notify_all();
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
    }
    for (int r = 0; r < g->n; r++)
        if (g->results[r] != 0) {
            if (r != 0) { strncpy(leader->err, g->hs[r]->err, sizeof(leader->err) - 1); leader->err[sizeof(leader->err) - 1] = 0; }
            return g->results[r];
        }
    return 0;
}

int group_size(const rebcu_handle* h) { return h->group ? h->group->n : 1; }
rebcu_handle* group_member(const rebcu_handle* h, int r) { return h->group->hs[r]; }

void group_destroy(rebcu_handle* leader) {
    rebcu_group* g = leader->group;
    if (!g) return;
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->stop = true;
        g->cv_start.notify_all();
    }
    for (int r = 0; r < g->n; r++) if (g->threads[r].joinable()) g->threads[r].join();
    leader->group = nullptr;
    for (int r = 1; r < g->n; r++) rebcu_destroy(g->hs[r]);
    delete g;
}

// ---- the calls that need more than "everybody does the same" -------------------------------------------------------
// Whole-array upload / download on the leader = every worker moves its own block.
int group_upload(rebcu_handle* leader, const rebcu_particle* particles, uint64_t N) {
    const int W = group_size(leader);
    return group_run(leader, [=](rebcu_handle* h, int r) {
        const uint64_t b = N * (uint64_t)r / (uint64_t)W;
        return rebcu_upload_shard(h, particles + b, N);
    });
}

int group_download(rebcu_handle* leader, rebcu_particle* particles, uint64_t N) {
    if (!leader->resident) return rebcu_fail(leader, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    if (N < leader->N) return rebcu_fail(leader, REBCU_ERR_CAPACITY, "host particle buffer too small");
    return group_run(leader, [=](rebcu_handle* h, int) {
        uint64_t b, e; rebcu_shard_range(h, &b, &e);
        return rebcu_download_shard(h, particles + b, e - b);
    });
}

// Same call with a private copy of the config on every rank; the leader's copy is handed back.
int group_cfg_call(rebcu_handle* leader, rebcu_config* cfg, const std::function<int(rebcu_handle*, rebcu_config*)>& call) {
    rebcu_config copies[REBCU_MAX_RANKS];
    const int W = group_size(leader);
    for (int r = 0; r < W; r++) copies[r] = *cfg;
    rebcu_config* cp = copies;
    const int err = group_run(leader, [&call, cp](rebcu_handle* h, int r) { return call(h, &cp[r]); });
    *cfg = copies[0];
    return err;
}

// The complete collision list in the reference's serial order from the ranks' lists: segment by segment (ghost-box
// major for DIRECT / LINE, a single segment for TREE / LINETREE), ranks in order inside a segment.
int group_collisions_fetch(rebcu_handle* leader, rebcu_collision* out, uint64_t cap, uint64_t* n_found) {
    const int W = group_size(leader);
    std::vector<std::vector<rebcu_collision>> lists(W);
    std::vector<std::vector<uint64_t>> segs(W);
    auto* L = &lists; auto* S = &segs;
    const int err = group_run(leader, [L, S](rebcu_handle* h, int r) {
        uint64_t n = 0;
        int e = rebcu_collisions_fetch(h, nullptr, 0, &n);
        if (e) return e;
        (*L)[r].resize(n);
        if (n && (e = rebcu_collisions_fetch(h, (*L)[r].data(), n, &n))) return e;
        uint64_t counts[64], ns = 0;
        if ((e = rebcu_collisions_segments(h, counts, 64, &ns))) return e;
        (*S)[r].assign(counts, counts + ns);
        return 0;
    });
    if (err) return err;
    size_t n_seg = 0;
    for (int r = 0; r < W; r++) n_seg = segs[r].size() > n_seg ? segs[r].size() : n_seg;
    std::vector<size_t> pos(W, 0);
    uint64_t total = 0;
    for (size_t s = 0; s < n_seg; s++)
        for (int r = 0; r < W; r++) {
            const uint64_t cnt = s < segs[r].size() ? segs[r][s] : 0;
            for (uint64_t k = 0; k < cnt; k++, total++) if (out && total < cap) out[total] = lists[r][pos[r] + k];
            pos[r] += cnt;
        }
    *n_found = total;            // (the leader is rank 0's own handle: its col_n stays the count of rank 0's list)
    return REBCU_OK;
}

extern "C" {

rebcu_handle* rebcu_create_group(const int* devices, int n) {
    if (n < 1 || n > REBCU_MAX_RANKS) return nullptr;
    if (n == 1) return rebcu_create(devices[0], nullptr);
    rebcu_group* g = new rebcu_group();
    g->n = n;
    for (int r = 0; r < n; r++) {
        g->hs[r] = rebcu_create(devices[r], nullptr);
        if (!g->hs[r]) { for (int q = 0; q < r; q++) rebcu_destroy(g->hs[q]); delete g; return nullptr; }
    }
    if (rebcu_comm_init_all(g->hs, n, REBCU_TRANSPORT_AUTO) != REBCU_OK) {
        // e.g. NCCL could not initialise: peer copies still work inside one process
        if (rebcu_comm_init_all(g->hs, n, REBCU_TRANSPORT_LOCAL) != REBCU_OK) { for (int q = 0; q < n; q++) rebcu_destroy(g->hs[q]); delete g; return nullptr; }
    }
    for (int r = 0; r < n; r++) g->threads[r] = std::thread(group_worker, g, r);
    g->hs[0]->group = g;
    return g->hs[0];
}

int rebcu_group_size(const rebcu_handle* h) { return group_size(h); }

}  // extern "C"
