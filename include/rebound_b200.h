/*
 * rebound_b200.h -- C ABI of the B200-native force / collision / kick-drift hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes only.  Every hot-path
 * entry point names the reference function (file:line in hannorein/rebound v5.0.0,
 * paths relative to the reference root) whose behaviour it replaces.  The shim
 * translation units under rebound_b200/shim/ forward the reference's own symbols
 * (reb_gravity_basic_calculate_acceleration, ...) to these functions, see
 * INTEGRATION.md.
 *
 * The same structs are used by the CPU oracle (oracle/oracle.c, prefix orc_) and by
 * the reference harness (oracle/ref_harness.c, prefix refh_) so that the parity
 * tests can drive all three through one ctypes binding.
 *
 * All arithmetic is IEEE binary64.  All functions returning int return 0 on success
 * and a negative REBCU_ERR_* code on failure; rebcu_last_error() gives the message
 * (the text equals the reference's reb_simulation_error() text where one exists).
 */
#ifndef REBOUND_B200_H
#define REBOUND_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REBCU_SIZE_MAX UINT64_MAX

/* Layout of struct reb_particle (src/rebound.h:86-104): 11 doubles + 3 pointers = 112 B. */
typedef struct rebcu_particle {
    double x, y, z;
    double vx, vy, vz;
    double ax, ay, az;
    double m;
    double r;
    uint64_t name;   /* const char*            -- carried through untouched */
    uint64_t ap;     /* void*                  -- carried through untouched */
    uint64_t sim;    /* struct reb_simulation* -- carried through untouched */
} rebcu_particle;

/* Layout of struct reb_vec6d (src/rebound.h:134-141). */
typedef struct rebcu_vec6d { double x, y, z, vx, vy, vz; } rebcu_vec6d;

/* Layout of struct reb_collision (src/rebound.h:144-149): 72 B. */
typedef struct rebcu_collision {
    uint64_t p1;
    uint64_t p2;
    rebcu_vec6d gb;
    uint64_t ri;
} rebcu_collision;

/* Enum values are the reference's (src/rebound.h:372-393). */
enum { REBCU_COLLISION_NONE = 0, REBCU_COLLISION_DIRECT = 1, REBCU_COLLISION_TREE = 2,
       REBCU_COLLISION_LINE = 4, REBCU_COLLISION_LINETREE = 5 };
enum { REBCU_BOUNDARY_NONE = 0, REBCU_BOUNDARY_OPEN = 1, REBCU_BOUNDARY_PERIODIC = 2, REBCU_BOUNDARY_SHEAR = 3 };
enum { REBCU_GRAVITY_NONE = 0, REBCU_GRAVITY_BASIC = 1, REBCU_GRAVITY_COMPENSATED = 2, REBCU_GRAVITY_TREE = 3 };
enum { REBCU_IGNORE_TERMS_NONE = 0, REBCU_IGNORE_TERMS_BETWEEN_0_AND_1 = 1, REBCU_IGNORE_TERMS_INVOLVING_0 = 2 };
/* Integrators are selected by name in the reference (src/simulation.c:217-238). */
enum { REBCU_INTEGRATOR_NONE = 0, REBCU_INTEGRATOR_LEAPFROG = 1, REBCU_INTEGRATOR_SEI = 2 };
/* Arithmetic mode of the direct-summation and tree-walk kernels:
 *   STRICT: per-particle ascending-j accumulation, no FMA contraction, IEEE sqrt and divide
 *           => bit-identical to the reference C build (-std=c99, src/Makefile.defs:5).
 *   FAST:   FMA + rsqrt/Newton, j split across lanes; <= 1e-12 relative of STRICT. */
enum { REBCU_MODE_STRICT = 0, REBCU_MODE_FAST = 1 };

/* The scalar fields of struct reb_simulation (src/rebound.h:238-413) the hot path reads.
 * The shim fills this from `r` before each call and writes back t / dt_last_done /
 * gravity_ignore_terms after it. */
typedef struct rebcu_config {
    double t;
    double G;
    double softening;
    double OMEGA;
    double OMEGAZ;
    double dt;
    double dt_last_done;
    double opening_angle2;
    double root_size;
    uint64_t N_active;            /* REBCU_SIZE_MAX: all particles are active */
    int32_t testparticle_type;
    int32_t gravity_ignore_terms;
    int32_t N_root_x, N_root_y, N_root_z;
    int32_t N_ghost_x, N_ghost_y, N_ghost_z;
    int32_t boundary;
    int32_t gravity;
    int32_t collision;
    int32_t integrator;
    int32_t leapfrog_order;       /* struct reb_integrator_leapfrog_state.order (integrator_leapfrog.h:30-32) */
    int32_t mode;                 /* REBCU_MODE_* */
    int32_t quadrupole;           /* 1: the reference was compiled with -DQUADRUPOLE (src/tree.c:148-198, 293-303):
                                   * tree cells carry the mass quadrupole tensor and accepted cells apply it */
} rebcu_config;

/* One cell of the octree in depth-first pre-order (octants ascending), the order in which the
 * reference's recursive functions visit its pointer tree (src/tree.h:34-56, src/tree.c:285-289).
 * `skip` is the pre-order index of the first cell after this cell's subtree. */
typedef struct rebcu_treecell {
    double x, y, z, w;        /* geometric centre and width (tree.c:87-102) */
    double m, mx, my, mz;     /* mass and centre of mass    (tree.c:147-207) */
    int32_t pt;               /* particle index (leaf) or -(number of particles) (tree.c:112,127,129) */
    int32_t skip;
    int32_t depth;            /* 0 = root cell */
    int32_t rootbox;
} rebcu_treecell;

enum {
    REBCU_INTERRUPTED = 1,            /* rebcu_steps stopped early on the caller's interrupt flag (not an error) */
    REBCU_OK = 0,
    REBCU_ERR_CUDA = -1,              /* CUDA runtime failure */
    REBCU_ERR_ARG = -2,               /* invalid argument */
    REBCU_ERR_ROOT_SIZE = -3,         /* tree.c:255-258 */
    REBCU_ERR_OUTSIDE_BOX = -4,       /* tree.c:265-268 */
    REBCU_ERR_NONFINITE = -5,         /* tree.c:66-69 */
    REBCU_ERR_SAME_COORDINATES = -6,  /* tree.c:119-123 */
    REBCU_ERR_LEAPFROG_ORDER = -7,    /* integrator_leapfrog.c:204-206 */
    REBCU_ERR_CAPACITY = -8,          /* caller-provided output buffer too small */
    REBCU_ERR_CELL_SIZE_ZERO = -9,    /* tree.c:107-111 */
    REBCU_ERR_NOT_RESIDENT = -10      /* resident call without a prior rebcu_upload */
};

typedef struct rebcu_handle rebcu_handle;

/* ---- life cycle ------------------------------------------------------------------------- */
/* One handle per simulation (the reference keeps all state in the caller-owned struct
 * reb_simulation, src/simulation.c:98; the struct is size-frozen, so device state lives here).
 * `stream` is a cudaStream_t (NULL = a private non-blocking stream). */
rebcu_handle* rebcu_create(int device, void* stream);
void rebcu_destroy(rebcu_handle* h);
const char* rebcu_last_error(const rebcu_handle* h);
int rebcu_version(void);
int rebcu_device_count(void);
void* rebcu_stream(const rebcu_handle* h);
int rebcu_synchronize(rebcu_handle* h);
/* Pin / unpin a caller-owned host buffer (e.g. r->particles) for full-rate PCIe copies. */
int rebcu_host_register(void* ptr, uint64_t bytes);
int rebcu_host_unregister(void* ptr);

/* ---- residency (r->particles AoS <-> device SoA) ---------------------------------------- */
/* Modelled on the reference's is_synchronized / did_modify_particles protocol
 * (src/simulation.c:633-637, src/particle.c:75,345,371). */
int rebcu_upload(rebcu_handle* h, const rebcu_particle* particles, uint64_t N);
int rebcu_download(rebcu_handle* h, rebcu_particle* particles, uint64_t N);
/* Only ax,ay,az are written back (what a gravity routine produces). */
int rebcu_download_acc(rebcu_handle* h, rebcu_particle* particles, uint64_t N);
/* The Kahan compensation terms of the last REB_GRAVITY_COMPENSATED force evaluation, laid out as the
 * reference's r->gravity_cs (struct reb_vec3d[N], src/gravity.c:293-306): IAS15 reads them
 * (src/integrator_ias15.c:337-343).  Bit-identical in STRICT mode. */
int rebcu_download_gravity_cs(rebcu_handle* h, double* out_xyz, uint64_t N);
uint64_t rebcu_N(const rebcu_handle* h);
/* Device pointer of one resident SoA field (0..10 = x,y,z,vx,vy,vz,ax,ay,az,m,r as doubles;
 * 11..13 = name, ap, sim as 64-bit words), for torch.distributed plumbing; length rebcu_N(h). */
void* rebcu_device_field(rebcu_handle* h, int field);

/* ---- resident hot path ------------------------------------------------------------------- */
/* reb_simulation_update_acceleration, src/simulation.c:640-689 (NONE/BASIC/COMPENSATED/TREE).
 * For TREE this includes the boundary check at gravity.c:56, so N may change. */
int rebcu_update_acceleration(rebcu_handle* h, rebcu_config* cfg);
/* reb_integrator_leapfrog_step (integrator_leapfrog.c:97-209) / reb_integrator_sei_step
 * (integrator_sei.c:86-117), selected by cfg->integrator.  Advances cfg->t, sets dt_last_done. */
int rebcu_integrator_step(rebcu_handle* h, rebcu_config* cfg);
/* reb_gravity_basic_calculate_and_apply_jerk, src/gravity.c:850-924 (called by EOS, integrator_eos.c:101-103,
 * right after a force evaluation): kicks the velocities with the gradient term of the modified-kick schemes,
 * from the resident positions and accelerations; `v` is the reference's argument.  Uses cfg->G, N_active,
 * testparticle_type, gravity_ignore_terms.  Bit-identical to the reference's serial build. */
int rebcu_apply_jerk(rebcu_handle* h, const rebcu_config* cfg, double v);
/* reb_boundary_check, src/boundary.c:35-141.  OPEN removes particles (order preserving,
 * N_active decremented as in particle.c:364-366): cfg->N_active and rebcu_N() change. */
int rebcu_boundary_check(rebcu_handle* h, rebcu_config* cfg);
/* The search part of reb_collision_search (src/collision.c:49-331): fills a device list in the
 * serial build's order and copies it to `out` (capacity `cap` entries); *n_found is the full count.
 * The shuffle and the resolve loop (collision.c:336-404) stay with the caller. */
int rebcu_collision_search(rebcu_handle* h, const rebcu_config* cfg,
                           rebcu_collision* out, uint64_t cap, uint64_t* n_found);
/* r->map / r->N_map / r->N_targets for the following searches (src/rebound.h:257-258,344; read at
 * src/collision.c:53-58; MERCURIUS and TRACE set them around their encounter steps,
 * integrator_mercurius.c:429, integrator_trace.c:820,1047,1093).  map == NULL: every particle is a
 * projectile; N_targets == REBCU_SIZE_MAX: as many targets as projectiles.  As in the reference, DIRECT
 * sends projectile and target slots through the map, LINE ignores N_targets, and TREE / LINETREE only cut
 * the projectile loop to the first N_map particles.  The map is copied; (NULL, 0, REBCU_SIZE_MAX) resets. */
int rebcu_set_collision_subset(rebcu_handle* h, const uint64_t* map, uint64_t N_map, uint64_t N_targets);
/* reb_simulation_steps for a simulation without host callbacks (src/simulation.c:504-603):
 * n x { integrator step; boundary check; collision search }.  With cfg->collision != NONE the
 * collision list of the LAST step is left on the device (rebcu_collisions_fetch). */
int rebcu_steps(rebcu_handle* h, rebcu_config* cfg, uint64_t n_steps);
int rebcu_collisions_fetch(rebcu_handle* h, rebcu_collision* out, uint64_t cap, uint64_t* n_found);
/* The reference polls its global `reb_sigint` inside its long loops and leaves them when a second Ctrl-C arrives
 * (`if (reb_sigint > 1) return;`, src/gravity.c:196, src/collision.c:75,232; declared in src/rebound.c:193).
 * rebcu_steps tests *flag > 1 before every step: the step in progress is completed, cfg->t stands at the last
 * completed step, and the call returns REBCU_INTERRUPTED.  NULL (default) = never. */
int rebcu_set_interrupt_flag(rebcu_handle* h, const volatile int* flag);
/* Hook run after each step's collision search inside rebcu_steps: the place of the shuffle + resolve
 * loop (collision.c:336-404), which stays on the host because r->collision_resolve is a user callback.
 * The callback may call rebcu_collisions_fetch / rebcu_download / rebcu_upload; non-zero aborts. */
int rebcu_set_collision_callback(rebcu_handle* h, int (*cb)(void* user), void* user);

/* ---- tree inspection (parity tests; reb_tree_construct + reb_tree_calculate_gravity_data,
 *      src/tree.c:254-271, 209-229) --------------------------------------------------------- */
int rebcu_tree_build(rebcu_handle* h, const rebcu_config* cfg);
uint64_t rebcu_tree_cell_count(const rebcu_handle* h);
int rebcu_tree_fetch(rebcu_handle* h, rebcu_treecell* out, uint64_t cap);

/* ---- host-buffer drop-ins (what the shim calls when the data lives in r->particles) ------ */
/* reb_gravity_basic_calculate_acceleration        src/gravity.c:167-282
 * reb_gravity_compensated_calculate_acceleration  src/gravity.c:284-531
 * reb_gravity_tree_calculate_acceleration         src/gravity.c:47-106
 * Upload x,y,z,m -> kernel -> write ax,ay,az into the caller's AoS.  cfg->gravity selects. */
int rebcu_gravity_host(rebcu_handle* h, rebcu_config* cfg, rebcu_particle* particles, uint64_t* N);
/* reb_gravity_basic_calculate_and_apply_jerk on a host AoS (x, v, a, m in; v out). */
int rebcu_jerk_host(rebcu_handle* h, const rebcu_config* cfg, rebcu_particle* particles, uint64_t N, double v);
/* reb_collision_search (search part) on a host AoS. */
int rebcu_collision_search_host(rebcu_handle* h, const rebcu_config* cfg,
                                const rebcu_particle* particles, uint64_t N,
                                rebcu_collision* out, uint64_t cap, uint64_t* n_found);
/* reb_simulation_steps on a host AoS: upload, n steps resident, download. */
int rebcu_steps_host(rebcu_handle* h, rebcu_config* cfg, rebcu_particle* particles, uint64_t* N,
                     uint64_t n_steps);

/* ---- device-side hard-sphere resolve (NOT bit-identical, see csrc/resolve.cu) ---- */
/* reb_collision_resolve_hardsphere (src/collision.c:573-665) applied on the device to the list of the last
 * collision search, in the reference's shuffled order (rand_r, collision.c:337-342) with the sequential semantics of
 * its loop (collision.c:351-404).  The rotation uses the device's atan2/sin/cos and the restitution law its pow, so
 * velocities agree with the reference to a few ulp per resolved collision, not bit for bit -- which is why the
 * drop-in librebound keeps resolving on the host.  Restitution laws (the reference takes a user callback,
 * rebound.h:403-409; a device needs a closed form):
 *   CONSTANT   eps = a
 *   POWERLAW   eps = clamp(a * pow(fabs(v) * b, c), lo, hi)     e.g. Bridges et al. 1984: a=0.32, b=100, c=-0.234, [0,1] */
#define REBCU_RESTITUTION_CONSTANT 0
#define REBCU_RESTITUTION_POWERLAW 1
typedef struct rebcu_restitution {
    int32_t kind;
    int32_t pad_;
    double a, b, c, lo, hi;
} rebcu_restitution;
/* enable != 0: rebcu_steps resolves every step's list on the device instead of calling the collision callback.
 * restitution NULL = perfectly elastic (the reference's default when coefficient_of_restitution is NULL).
 * rand_seed = r->rand_seed.  Resets the statistics below. */
int rebcu_set_device_resolve(rebcu_handle* h, int enable, const rebcu_restitution* restitution,
                             double minimum_collision_velocity, unsigned int rand_seed);
/* Resolve the list left by the last rebcu_collision_search now (what rebcu_steps does per step when enabled). */
int rebcu_collision_resolve(rebcu_handle* h, const rebcu_config* cfg);
/* r->collisions_plog, r->collisions_log_n, r->rand_seed as the device resolver advanced them; rounds of the last call. */
int rebcu_collision_stats(const rebcu_handle* h, double* plog, uint64_t* log_n, unsigned int* rand_seed, int* rounds_last);

/* ---- exact collision resolve: the device keeps the order, the caller does the arithmetic ---------------
 * The resolve loop of reb_collision_search (src/collision.c:336-404) applied to the list of the last search WITHOUT
 * bringing the particles to the host: the rand_r shuffle, the sequential semantics (conflict-free rounds) and the two
 * early exits of reb_collision_resolve_hardsphere (no overlap / not approaching, :598,602; exact IEEE arithmetic) run
 * on the device; the collisions that pass them are handed to `fn` in batches whose pairs share no particle, with the
 * current state of both particles (s1, s2 = x y z vx vy vz m r).  fn fills v1, v2 (new velocities), plog_term (its
 * contribution to r->collisions_plog) and logged (1 if it counted the collision in r->collisions_log_n); it may work
 * on the batch in parallel.  The drop-in's fn runs the reference's own reb_collision_resolve_hardsphere -- with the
 * user's coefficient_of_restitution callback -- on a two-particle scratch simulation, so velocities, collisions_plog
 * (summed in the sequential loop's order) and collisions_log_n are the reference's bits.  Only resolvers that touch
 * nothing but the two particles and never remove one qualify (the built-in hard-sphere resolver does).
 * *rand_seed is advanced as r->rand_seed would be; *plog / *log_n are updated in place. */
typedef struct rebcu_resolve_pair {
    uint64_t k;                 /* position in the shuffled list */
    uint64_t p1, p2;
    rebcu_vec6d gb;
    double s1[8], s2[8];
    double v1[3], v2[3];
    double plog_term;
    uint64_t logged;
} rebcu_resolve_pair;
typedef int (*rebcu_pair_resolver)(void* user, rebcu_resolve_pair* pairs, uint64_t n);
int rebcu_collision_resolve_pairs(rebcu_handle* h, unsigned int* rand_seed, rebcu_pair_resolver fn, void* user,
                                  double* plog, uint64_t* log_n, int* rounds);

/* ---- multi-GPU sharding ------------------------------------------------------- */
/* Rank `rank` of `world` owns the contiguous i-block [N*rank/world, N*(rank+1)/world): the direct
 * and tree force kernels and kick/drift touch only that block; positions of the other blocks are
 * refreshed by the caller's all-gather on rebcu_device_field(0..2) between drift and force. */
int rebcu_set_shard(rebcu_handle* h, int rank, int world);
/* Called inside every integrator step between the drift and the force evaluation (where the
 * reference's MPI build calls reb_communication_mpi_distribute_particles, gravity.c:58-61): the caller
 * all-gathers the x,y,z fields of the other ranks' blocks (torch.distributed / NCCL on this stream). */
int rebcu_set_exchange_callback(rebcu_handle* h, void (*cb)(void* user), void* user);
/* Which fields the engine needs gathered by the exchange callback it is calling right now:
 *   POSITIONS   x,y,z                  between drift and force (every step)
 *   VELOCITIES  vx,vy,vz in addition   before a collision search (the overlap test reads the target's
 *                                      velocity, src/collision.c:101-106)
 *   ALL         all 14 fields          before an open-boundary compaction: particles change owner when
 *                                      indices shift (reb_simulation_remove_particle, src/particle.c:313-352)
 * Outside a callback the value is POSITIONS. */
#define REBCU_EXCHANGE_POSITIONS 1
#define REBCU_EXCHANGE_VELOCITIES 2
#define REBCU_EXCHANGE_ALL 4
int rebcu_exchange_request(const rebcu_handle* h);
void rebcu_shard_range(const rebcu_handle* h, uint64_t* begin, uint64_t* end);
/* Sharded collision search: each rank searches for the projectiles of its own block against ALL particles
 * (replicated positions / tree), so its list is the reference's serial list restricted to those
 * projectiles.  The serial order is ghost box, projectile, target for DIRECT/LINE (src/collision.c:64-124)
 * and projectile-major for TREE/LINETREE (:197-269); the complete list is therefore, segment by segment,
 * the concatenation over ranks.  counts[s] = number of local entries in segment s (one segment per ghost
 * box of the innermost ring for DIRECT/LINE, a single segment otherwise). */
int rebcu_collisions_segments(rebcu_handle* h, uint64_t* counts, uint64_t cap, uint64_t* n_segments);

/* ---- several GPUs behind one handle ------------------------------------------------------------------------------
 * One engine handle per device, joined by the native exchange below and driven by one worker thread each; the returned
 * LEADER handle is used like any other handle: every hot-path call on it (upload, download, update_acceleration,
 * integrator_step, boundary_check, steps, steps_host, gravity_host, collision_search, apply_jerk, exit_check, ...)
 * fans out to the ranks, each working on its target block and moving only its block of the host array over PCIe.
 * This is how a single-threaded program written against the reference's C API -- one struct reb_simulation, one call
 * to reb_simulation_integrate -- shards over the GPUs of a node: the drop-in creates a group when the environment
 * variable REBOUND_B200_DEVICES names several devices ("0-7", "0,1,2,3").  A device may be listed more than once
 * (the ranks then exchange through peer copies instead of NCCL).  Not available on a group: the diagnostics
 * (rebcu_energy / _com / _angular_momentum), rebcu_download_gravity_cs, the device-side resolve, the tree
 * inspection calls and a host collision callback inside rebcu_steps.  rebcu_destroy on the leader releases everything. */
rebcu_handle* rebcu_create_group(const int* devices, int n);
int rebcu_group_size(const rebcu_handle* h);

/* ---- native exchange: NCCL (or in-process peer copies) inside the engine ------------------------------------
 * With a communicator attached, the engine performs the exchange itself wherever it would have called the exchange
 * callback: in-place all-gather of the owners' blocks of the requested fields, enqueued on the handle's stream between
 * the drift and the force kernels (the place of reb_communication_mpi_distribute_particles, src/gravity.c:58-61).
 *   one process per GPU:    rank 0 calls rebcu_comm_unique_id, ships the 128 bytes to the other ranks by any means
 *                           (torch.distributed, MPI, a file), every rank calls rebcu_comm_init_rank.
 *   one process, many GPUs: rebcu_comm_init_all on the handles (one per device); every handle must then be driven
 *                           by its own host thread, all threads making the same sequence of calls.
 * Both also set the shard (rebcu_set_shard).  transport: NCCL needs distinct devices; LOCAL (peer copies fenced by
 * events and host barriers) also works for handles that share a device -- how sharded runs are tested on one GPU. */
#define REBCU_TRANSPORT_AUTO 0
#define REBCU_TRANSPORT_NCCL 1
#define REBCU_TRANSPORT_LOCAL 2
int rebcu_comm_unique_id(void* out128);
/* The exchange on demand: afterwards every rank holds the owners' current values of the fields `need` names
 * (REBCU_EXCHANGE_* mask; e.g. ALL before rank 0 downloads a complete state). */
int rebcu_exchange(rebcu_handle* h, int need);
/* Sharded residency: each rank's host memory holds only ITS block of r->particles (rebcu_shard_range of N_total), the
 * way a rank of the reference's MPI build owns only its particles.  Upload = own block over PCIe + one exchange of all
 * fields over NVLink; download = own block.  (rebcu_upload / rebcu_download move the whole array on every rank.) */
int rebcu_upload_shard(rebcu_handle* h, const rebcu_particle* block, uint64_t N_total);
int rebcu_download_shard(rebcu_handle* h, rebcu_particle* block, uint64_t cap);
int rebcu_comm_init_rank(rebcu_handle* h, const void* id128, int rank, int world);
int rebcu_comm_init_all(rebcu_handle** handles, int n, int transport);
int rebcu_comm_destroy(rebcu_handle* h);
/* Tree builds of a sharded run (several ranks, native exchange): 0 = every rank builds the whole tree (replicated),
 * 1 = every rank sorts and builds only the subtrees of its own key range and the ranks all-gather the traversal records
 * (csrc/tree.cu: tree_build_sharded; falls back to the replicated build for quadrupole trees, trees deeper than 480
 * levels and collision searches), 2 (default) = 1 when N >= 2^18.  The accelerations are the same bits either way. */
int rebcu_set_sharded_build(rebcu_handle* h, int mode);
/* Bytes this rank received through the exchange and the number of exchanges since the communicator was created;
 * *transport = REBCU_TRANSPORT_* in use (0: none).  Exchange time is timing class 7 of rebcu_timing_read. */
int rebcu_comm_stats(const rebcu_handle* h, uint64_t* bytes_received, uint64_t* exchanges, int* transport);

/* ---- diagnostics on the resident state ----------------------------------------- */
/* reb_simulation_energy, src/tools.c:108-162: out3 = {kinetic, potential, kinetic + potential}
 * (r->energy_offset is host state and is not added).  Uses cfg->G, N_active, testparticle_type.
 * The reference sums into one scalar in index order; this is a compensated parallel sum: equal to
 * ~1e-13 relative, not bit for bit. */
int rebcu_energy(rebcu_handle* h, const rebcu_config* cfg, double* out3);
/* reb_simulation_com, src/tools.c:376-408: out10 = {m, x, y, z, vx, vy, vz, ax, ay, az} of the centre of
 * mass (plain sums if the total mass is not positive). */
int rebcu_com(rebcu_handle* h, double* out10);
/* reb_simulation_angular_momentum, src/tools.c:164-174. */
int rebcu_angular_momentum(rebcu_handle* h, double* out3);

/* The exit conditions run_heartbeat tests after every step (src/simulation.c:242-272): *escape = some particle
 * is farther than exit_max_distance from the origin (r->status = REB_STATUS_ESCAPE), *encounter = some pair is
 * closer than exit_min_distance (REB_STATUS_ENCOUNTER).  A distance of 0 switches its check off, as in the
 * reference.  Pure predicates in the reference's expression order: exact. */
int rebcu_exit_check(rebcu_handle* h, double exit_max_distance, double exit_min_distance, int* escape, int* encounter);

/* ---- instrumentation ----------------------------------------------------------------------- */
/* Sustained FP64 FMA rate of the device in TFLOP/s (2 flop per DFMA), timed with CUDA events: the
 * measured peak bench.py reports the FP64-bound kernels against. */
int rebcu_measure_fp64_peak(rebcu_handle* h, double* tflops);
/* Number of kernels this handle launched since creation (bench.py's gpu_launches). */
uint64_t rebcu_launch_count(const rebcu_handle* h);
/* Device time in ms (CUDA events on the handle's stream) accumulated per kernel class since the
 * last reset; classes: 0 direct force, 1 kick/drift, 2 tree build, 3 tree walk, 4 collision,
 * 5 boundary, 6 pack/unpack.  Enabled by rebcu_timing_enable (adds event records per launch). */
int rebcu_timing_enable(rebcu_handle* h, int on);
int rebcu_timing_read(rebcu_handle* h, double* ms_out, uint64_t* launches_out, int n_classes);
int rebcu_timing_reset(rebcu_handle* h);
/* Work counters of the tree walk on the current tree (call right after a TREE force evaluation with the same cfg);
 * what bench.py derives its FP64 roofline figures from.  out6 = {interactions of the per-particle criterion
 * (src/tree.c:284: what the reference and the STRICT walk evaluate) summed over this rank's particles and ghost boxes,
 * cells those walks visit, list entries summed over the groups of the last FAST group walk (each entry is evaluated by
 * 32 lanes), cells its traversals tested, number of groups, cells in the tree}. */
int rebcu_tree_walk_stats(rebcu_handle* h, const rebcu_config* cfg, uint64_t* out6);
/* Device self-test of the STRICT kernels' branch-free sqrt/divide against __dsqrt_rn/__ddiv_rn on
 * n_samples pseudo-random operand pairs: result4 = {sqrt mismatches, divide mismatches (both must be 0),
 * sqrt / divide operands of the ordinary families that were sent to the generic path}. */
int rebcu_selftest_math(rebcu_handle* h, uint64_t n_samples, uint64_t seed, uint64_t* result4);
/* Device self-tests of the hand-written primitives behind the tree build (csrc/primitives.cuh): a stable LSD
 * radix sort of host (key, value) pairs by the low `bits` key bits, and an exclusive scan, both in place. */
int rebcu_selftest_sort(rebcu_handle* h, uint64_t* keys, uint32_t* vals, uint64_t n, int bits);
int rebcu_selftest_scan(rebcu_handle* h, uint32_t* values, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* REBOUND_B200_H */
