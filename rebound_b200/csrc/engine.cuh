// engine.cuh -- shared declarations of the B200 engine behind include/rebound_b200.h.
//
// Data layout in HBM (per handle = per simulation):
//   soa      14 contiguous arrays of `cap` 8-byte words: x y z vx vy vz ax ay az m r | name ap sim
//            (struct reb_particle, src/rebound.h:86-104, transposed).  Kernels touch only the arrays
//            they need: force kernels read x,y,z,m (32 B/particle) and write ax,ay,az; kick/drift
//            streams x,v,a in and x,v out; the three tag arrays only move in pack/unpack/compaction.
//   aos      staging copy of the caller's AoS (112 B/particle) for the PCIe transfers.
//   tree_*   sorted keys / permutation / pre-order cell arrays (see tree.cu).
//   col_*    collision candidate counts and the output list (see collision.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stddef.h>
#include <vector>
#include "../../include/rebound_b200.h"

enum { F_X = 0, F_Y, F_Z, F_VX, F_VY, F_VZ, F_AX, F_AY, F_AZ, F_M, F_R, F_NAME, F_AP, F_SIM, F_COUNT };
enum { TC_DIRECT = 0, TC_KICKDRIFT, TC_TREEBUILD, TC_TREEWALK, TC_COLLISION, TC_BOUNDARY, TC_PACK, TC_EXCHANGE, TC_COUNT };

#define REBCU_MAX_GHOST 729   // (2*4+1)^3
#define GHOST_RING 8
#define AUX_STREAMS 6
#define PIPE_RANGES 130
#define REBCU_MAX_RANKS 16

struct EngineComm;           // comm.cu: NCCL communicator or in-process peer group
struct rebcu_group;          // group.cu: several GPUs behind one (leader) handle

struct GhostShifts {          // ghost-box offsets, computed on the host exactly as src/boundary.c:145-201
    int n;
    rebcu_vec6d gb[REBCU_MAX_GHOST];
};

struct TimedRange { cudaEvent_t a, b; int cls; };

struct TreeBuffers {
    uint64_t cap_n = 0;       // particle capacity of the per-particle arrays
    uint64_t cap_cells = 0;
    uint64_t* keys = nullptr;      // [cap_n] 64-bit keys (rootbox | 3 bits/level), unsorted
    uint64_t* keys_sorted = nullptr;
    uint32_t* perm = nullptr;      // [cap_n] sorted position -> particle index
    uint32_t* perm_in = nullptr;
    int32_t* lcp = nullptr;        // [cap_n+1] common prefix (levels) of sorted neighbours k-1,k ; -1 across root boxes
    uint32_t* cell_off = nullptr;  // [cap_n+1] exclusive scan of cells opened per sorted particle
    uint32_t* cell_cnt = nullptr;
    rebcu_treecell* cells = nullptr; // [cap_cells] pre-order
    int32_t* parent = nullptr;     // [cap_cells]
    uint32_t* ready = nullptr;     // [cap_cells] children-done counters for the moment pass
    double4* walk_pos = nullptr;   // [cap_cells] (mx,my,mz,m) packed for the walk
    double4* walk_geo = nullptr;   // [cap_cells] (x,y,z,w) packed for the collision walk
    int2* walk_meta = nullptr;     // [cap_cells] int4 (pt, skip, depth, rootbox)
    int2* walk_meta2 = nullptr;    // [cap_cells] (leaf: pt >= 0 | internal: -(depth+1), skip): all the gravity walk needs
    double4* walk_rec = nullptr;   // [walk_rec_cap] (mx,my,mz, meta2 as 64 bits): everything the traversal decides on, one 32-byte load
    double* walk_m = nullptr;      // [walk_rec_cap] cell mass, read only for accepted cells
    uint64_t walk_rec_cap = 0;
    double4* col_rec = nullptr;    // [col_rec_cap] collision-walk records: leaf (x,y,z | pt, skip), internal (cx,cy,cz | -(depth+1), skip)
    uint64_t col_rec_cap = 0;
    double* quad = nullptr; uint64_t quad_cap = 0;   // [6][quad_cap] mxx mxy mxz myy myz mzz (QUADRUPOLE builds only)
    bool has_quad = false;         // the current tree carries quadrupole moments
    void* sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
    void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
    uint64_t n_cells = 0;
    uint32_t* shard_list = nullptr; // [cap_n] sorted positions owned by this rank (multi-GPU walk)
    int* flags = nullptr;          // device error flags [8]
    int built_for_n = -1;
    bool prefix_ok = true;         // sorting on a key prefix has not hit a long tie run yet (see tree_build)
    // sharded build (tree_build_sharded): every rank sorts and builds the subtrees of its own key range
    bool rec_ready = false;        // walk_rec / walk_m already hold the complete tree (no walk_pack needed)
    bool complete = true;          // the build's cell arrays (walk_pos, walk_geo, walk_meta) hold the whole tree
    bool shard_ok = true;          // this simulation has not needed the replicated build yet (very deep trees)
    int shard_mode = 2;            // 0 never, 1 whenever possible, 2 when N is large enough to pay (rebcu_set_sharded_build)
    uint64_t* sh_keys = nullptr; uint32_t* sh_idx = nullptr; uint64_t sh_cap_n = 0;    // compacted (key, index) of this rank's key range
    uint32_t* sh_hist = nullptr;   // [n_buckets] particles per bucket
    uint64_t* sh_pstart = nullptr; // [n_buckets+1] sorted position where a bucket starts
    uint64_t* sh_tab = nullptr;    // [n_buckets + 8 W] gathered per-rank status words + per-bucket (first, end) local cell offsets
    uint32_t* sh_level = nullptr;  // 3 level arrays (count, cells, start) of the top tree
    long long* sh_delta = nullptr; // [n_buckets] global cell index of a bucket's subtree minus its local offset
    int4* sh_top = nullptr;        // [n_top] top cells: index, depth, skip, count
    uint32_t* sh_info = nullptr;   // small result block (see tree.cu)
    uint64_t sh_buckets_cap = 0;
    uint64_t last_build_cells_local = 0;
    uint64_t sh_pk[REBCU_MAX_RANKS + 1] = {};   // sorted positions where the ranks' key ranges start (last sharded build)
    double* acc_sorted = nullptr; uint64_t acc_sorted_cap = 0;   // [3][cap] accelerations in sorted order (key-range walk of a sharded run)
};

struct rebcu_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t N = 0, cap = 0;
    double* soa = nullptr;
    rebcu_particle* aos = nullptr;
    bool resident = false;
    int rank = 0, world = 1;
    char err[512] = {0};
    uint64_t launches = 0;
    bool timing = false;
    std::vector<TimedRange> ranges;
    std::vector<cudaEvent_t> event_pool;
    GhostShifts* ghosts_dev = nullptr;   // device copy
    GhostShifts ghosts_host;             // what ghosts_dev holds (valid if ghosts_valid)
    bool ghosts_valid = false;
    GhostShifts* ghost_ring = nullptr;   // pinned staging ring
    cudaEvent_t ghost_ring_ev[GHOST_RING] = {};
    bool ghost_ring_used[GHOST_RING] = {};
    int ghost_ring_next = 0;
    TreeBuffers tree;
    // collision buffers
    uint32_t* col_count = nullptr; uint32_t* col_off = nullptr; uint64_t col_cap_n = 0;
    rebcu_collision* col_list = nullptr; uint64_t col_cap = 0; uint64_t col_n = 0;
    void* col_scan_tmp = nullptr; size_t col_scan_tmp_bytes = 0;
    uint64_t* col_slots = nullptr; uint64_t col_slots_cap = 0;   // parked hits of the single-traversal tree search
    // r->map / r->N_map / r->N_targets of the collision search (rebcu_set_collision_subset)
    uint32_t* col_map = nullptr; uint64_t col_map_cap = 0, col_map_n = 0, col_map_max = 0; bool col_map_on = false;
    uint64_t col_targets = REBCU_SIZE_MAX;
    // small device scratch
    double* scratch = nullptr;            // 64 doubles
    unsigned long long* counters = nullptr; // 16 counters
    void* compact_tmp = nullptr; size_t compact_tmp_bytes = 0; double* compact_buf = nullptr; uint64_t compact_cap = 0;
    uint32_t* compact_flag = nullptr; uint32_t* compact_pos = nullptr;
    double* scratch_big = nullptr;        // 2 x 7 x 256 doubles: massive-body snapshots of the fused test-particle step
    int tp_phase = 0;
    cudaStream_t aux[AUX_STREAMS] = {};         // copy/compute overlap streams of the chunk-pipelined host path
    cudaEvent_t aux_ev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_ev[3 * PIPE_RANGES] = {};  // per range: [3i] upload done, [3i+1] kernels done, [3i+2] download done (trace)
    double* gravity_cs = nullptr; uint64_t gravity_cs_cap = 0; bool gravity_cs_valid = false;   // r->gravity_cs of the last COMPENSATED evaluation: x[cap], y[cap], z[cap]
    // device-side hard-sphere resolve (resolve.cu)
    bool resolve_on = false; rebcu_restitution resolve_rest = {0, 0, 1.0, 0, 0, 0, 1}; double resolve_min_v = 0;
    unsigned int resolve_seed = 0; double resolve_plog = 0; uint64_t resolve_log_n = 0; int resolve_rounds = 0;
    uint32_t* resolve_buf = nullptr; uint64_t resolve_cap = 0;
    void* pairs_host = nullptr; uint64_t pairs_host_cap = 0;     // pinned staging of the exact (host-assisted) resolve
    double* row_buf = nullptr; uint64_t row_cap = 0;        // state + term buffer of the massive-row path (testparticle_type 1)
    double* diag_partial = nullptr; uint64_t diag_cap = 0;  // per-block partial sums of the diagnostics
    double* tp_hist = nullptr; uint64_t tp_hist_cap = 0;   // per-step snapshots of the massive bodies
    void (*exchange)(void*) = nullptr;    // multi-GPU position exchange hook (see rebcu_set_exchange_callback)
    void* exchange_user = nullptr;
    int exchange_need = REBCU_EXCHANGE_POSITIONS;   // what the running exchange callback must gather
    rebcu_group* group = nullptr;         // set on the leader handle of a multi-GPU group (group.cu)
    EngineComm* comm = nullptr;           // native exchange (rebcu_comm_init_rank / rebcu_comm_init_all); takes precedence over the callback
    uint64_t* comm_view[F_COUNT] = {};    // LOCAL transport: the arrays the peers pull from during the running exchange
    int comm_view_n = 0;
    int full_check_rank = 0, full_check_world = 0;  // real shard while a full-range boundary check runs (else world 0)
    uint64_t col_seg_n = 0, col_seg_stride = 0;     // local collision list: segments (ghost boxes) x projectiles per segment
    int (*collision_hook)(void*) = nullptr;   // called after each step's collision search (host resolve)
    void* collision_hook_user = nullptr;
    const volatile int* interrupt = nullptr;  // rebcu_set_interrupt_flag: the caller's reb_sigint
    // pinned staging for small host<->device exchanges
    unsigned long long* pinned = nullptr;  // 32 words

    inline double* f(int field) const { return soa + (size_t)field * cap; }
    inline uint64_t* tag(int field) const { return (uint64_t*)(soa + (size_t)field * cap); }
};

int rebcu_fail(rebcu_handle* h, int code, const char* msg);

// ---- multi-GPU group (group.cu): calls on the leader fan out to one worker thread per device ----
#include <functional>
bool group_active(const rebcu_handle* h);       // h leads a group and the caller is not one of its workers
int group_run(rebcu_handle* leader, const std::function<int(rebcu_handle*, int)>& fn);
int group_cfg_call(rebcu_handle* leader, rebcu_config* cfg, const std::function<int(rebcu_handle*, rebcu_config*)>& call);
int group_upload(rebcu_handle* leader, const rebcu_particle* particles, uint64_t N);
int group_download(rebcu_handle* leader, rebcu_particle* particles, uint64_t N);
int group_collisions_fetch(rebcu_handle* leader, rebcu_collision* out, uint64_t cap, uint64_t* n_found);
int group_size(const rebcu_handle* h);
void group_destroy(rebcu_handle* leader);
#define GROUP_UNSUPPORTED(h, what) do { if (group_active(h)) return rebcu_fail((h), REBCU_ERR_ARG, what " is not available on a multi-GPU group handle"); } while (0)
int rebcu_cuda_fail(rebcu_handle* h, cudaError_t e, const char* where);

#define CU_TRY(h, expr)                                                       \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return rebcu_cuda_fail((h), _e, #expr);        \
    } while (0)

// Timing bracket: records an event pair around a region when timing is enabled, and counts launches.
struct LaunchScope {
    rebcu_handle* h; int idx;
    LaunchScope(rebcu_handle* h_, int cls, int n_launches = 1);
    ~LaunchScope();
};

// internal entry points (one per translation unit)
int engine_reserve(rebcu_handle* h, uint64_t n);
int engine_exchange(rebcu_handle* h, int need);
int comm_exchange(rebcu_handle* h, int need);
int comm_gather_ranges(rebcu_handle* h, void** ptrs, const int* bytes, int n_arrays, const uint64_t* bounds);
void comm_free(rebcu_handle* h);
int boundary_check_full(rebcu_handle* h, rebcu_config* c);
int boundary_open_probe(rebcu_handle* h, const rebcu_config* c, bool* any);
int collision_resolve_device(rebcu_handle* h, const rebcu_config* c);
int tree_shard_list(rebcu_handle* h, const uint32_t** list, uint64_t* n_work);
void engine_ghost_shifts(const rebcu_config* c, int gx, int gy, int gz, GhostShifts* out);
int engine_upload_ghosts(rebcu_handle* h, const GhostShifts* g);
int direct_gravity(rebcu_handle* h, const rebcu_config* c);
int zero_acceleration(rebcu_handle* h);
int leapfrog_step(rebcu_handle* h, rebcu_config* c, bool fuse_ok);
int leapfrog_step_ex(rebcu_handle* h, rebcu_config* c, bool carry_in, bool carry_out, bool write_acc);
int sei_step(rebcu_handle* h, rebcu_config* c);
// Chunk-pipelined reb_simulation_steps on a host AoS for the fused test-particle step; returns 1 if the
// configuration is not eligible (caller falls back to upload / steps / download), 0 on success, <0 on error.
int tp_steps_resident(rebcu_handle* h, rebcu_config* c, uint64_t n_steps);
int tp_steps_host_pipelined(rebcu_handle* h, rebcu_config* c, rebcu_particle* particles, uint64_t N, uint64_t n_steps);
int engine_upload_range(rebcu_handle* h, cudaStream_t s_copy, cudaEvent_t ev, cudaStream_t s_kernel, const rebcu_particle* particles, uint64_t b, uint64_t e);
int engine_download_range(rebcu_handle* h, cudaStream_t s_kernel, cudaEvent_t ev, cudaStream_t s_copy, rebcu_particle* particles, uint64_t b, uint64_t e);
int boundary_check(rebcu_handle* h, rebcu_config* c);
int tree_build(rebcu_handle* h, const rebcu_config* c);
int tree_gravity(rebcu_handle* h, rebcu_config* c);
int tree_export(rebcu_handle* h);
int collision_search(rebcu_handle* h, const rebcu_config* c);
int update_acceleration(rebcu_handle* h, rebcu_config* c);
void engine_shard(const rebcu_handle* h, uint64_t* b, uint64_t* e);
void tree_free(rebcu_handle* h);

// ---- strict IEEE helpers: never contracted into FMA, correctly rounded sqrt and divide --------
__device__ __forceinline__ double s_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double s_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double s_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double s_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double s_sqrt(double a) { return __dsqrt_rn(a); }

static inline unsigned int div_up(uint64_t a, uint64_t b) { return (unsigned int)((a + b - 1) / b); }
