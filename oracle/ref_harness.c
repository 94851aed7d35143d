/*
 * ref_harness.c -- TEST INFRASTRUCTURE.  Thin plain-pointer wrapper around the UNMODIFIED
 * reference (hannorein/rebound v5.0.0), compiled together with the reference's own sources
 * where they lie under /root/reference/src by oracle/Makefile into oracle/_ref/.
 *
 * Nothing in the product path may link or call this.  It exists to
 *   (1) pin the CPU restatement in oracle/oracle.c against the real reference,
 *   (2) generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
 *   (3) serve as the "reference" CPU baseline in bench.py --impl reference.
 *
 * Every function takes the same rebcu_config / rebcu_particle structs as the product's C ABI
 * (include/rebound_b200.h) and drives the reference through its public API.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "rebound.h"
#include "tree.h"
#include "boundary.h"
#include "collision.h"
#include "gravity.h"
#include "../include/rebound_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

_Static_assert(sizeof(struct reb_particle) == sizeof(rebcu_particle), "particle layout");
_Static_assert(sizeof(struct reb_collision) == sizeof(rebcu_collision), "collision layout");
_Static_assert(sizeof(struct reb_vec6d) == sizeof(rebcu_vec6d), "vec6d layout");

static char refh_errbuf[512];

const char* refh_last_error(void){ return refh_errbuf; }

int refh_openmp_threads(void){
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void refh_set_threads(int n){
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static struct reb_simulation* make_sim(const rebcu_config* c, const rebcu_particle* p, uint64_t N){
    struct reb_simulation* r = reb_simulation_create();
    r->save_messages = 1;
    r->t = c->t; r->G = c->G; r->softening = c->softening;
    r->OMEGA = c->OMEGA; r->OMEGAZ = c->OMEGAZ;
    r->dt = c->dt; r->dt_last_done = c->dt_last_done;
    r->opening_angle2 = c->opening_angle2;
    r->root_size = c->root_size;
    r->N_active = (c->N_active == REBCU_SIZE_MAX) ? SIZE_MAX : (size_t)c->N_active;
    r->testparticle_type = c->testparticle_type;
    r->gravity_ignore_terms = c->gravity_ignore_terms;
    r->N_root_x = c->N_root_x; r->N_root_y = c->N_root_y; r->N_root_z = c->N_root_z;
    r->N_ghost_x = c->N_ghost_x; r->N_ghost_y = c->N_ghost_y; r->N_ghost_z = c->N_ghost_z;
    r->boundary = c->boundary;
    r->gravity = c->gravity;
    r->collision = c->collision;
    r->rand_seed = 42;
    switch (c->integrator){
        case REBCU_INTEGRATOR_LEAPFROG: {
            struct reb_integrator_leapfrog_state* s = reb_simulation_set_integrator(r, "leapfrog");
            if (c->leapfrog_order) s->order = (unsigned int)c->leapfrog_order;
            break; }
        case REBCU_INTEGRATOR_SEI:
            reb_simulation_set_integrator(r, "sei");
            break;
        default:
            reb_simulation_set_integrator(r, "none");
    }
    // Particles are installed directly (reb_simulation_add would reject particles outside the box,
    // which some parity cases need in order to exercise the hot path's own error handling).
    r->N = N;
    r->N_allocated = N ? N : 1;
    r->particles = calloc(r->N_allocated, sizeof(struct reb_particle));
    memcpy(r->particles, p, N*sizeof(struct reb_particle));
    for (size_t i=0;i<N;i++){ r->particles[i].sim = r; r->particles[i].name = NULL; r->particles[i].ap = NULL; }
    return r;
}

// Returns 0 if no error message is pending, otherwise -1 and copies the text.
static int collect_error(struct reb_simulation* r){
    refh_errbuf[0] = 0;
    int err = 0;
    if (r->messages){
        for (int i=0;i<10;i++){
            if (r->messages[i] && r->messages[i][0]=='e'){
                strncpy(refh_errbuf, r->messages[i]+1, sizeof(refh_errbuf)-1);
                err = -1;
            }
        }
    }
    return err;
}

static void copy_back(struct reb_simulation* r, rebcu_config* c, rebcu_particle* p, uint64_t* N){
    // name/ap/sim of the caller's records are preserved; all 11 doubles are returned.
    for (size_t i=0;i<r->N;i++){
        memcpy(&p[i], &r->particles[i], 11*sizeof(double));
    }
    *N = r->N;
    c->t = r->t;
    c->dt_last_done = r->dt_last_done;
    c->N_active = (r->N_active==SIZE_MAX)?REBCU_SIZE_MAX:r->N_active;
    c->gravity_ignore_terms = r->gravity_ignore_terms;
    c->OMEGAZ = r->OMEGAZ;
}

/* reb_simulation_update_acceleration (simulation.c:640) on the given state. */
int refh_gravity(rebcu_config* c, rebcu_particle* p, uint64_t* N){
    struct reb_simulation* r = make_sim(c, p, *N);
    reb_simulation_update_acceleration(r);
    int err = collect_error(r);
    copy_back(r, c, p, N);
    reb_simulation_free(r);
    return err;
}

/* reb_gravity_basic_calculate_and_apply_jerk (gravity.c:850) on the given positions, velocities and accelerations. */
int refh_apply_jerk(rebcu_config* c, rebcu_particle* p, uint64_t N, double v){
    struct reb_simulation* r = make_sim(c, p, N);
    reb_gravity_basic_calculate_and_apply_jerk(r, v);
    int err = collect_error(r);
    uint64_t n = N;
    copy_back(r, c, p, &n);
    reb_simulation_free(r);
    return err;
}

/* Force evaluation that also returns r->gravity_cs (src/gravity.c:293-306), 3 doubles per particle. */
int refh_gravity_cs(rebcu_config* c, rebcu_particle* p, uint64_t* N, double* cs_out){
    struct reb_simulation* r = make_sim(c, p, *N);
    reb_simulation_update_acceleration(r);
    int err = collect_error(r);
    if (!err && r->gravity_cs) memcpy(cs_out, r->gravity_cs, r->N*sizeof(struct reb_vec3d));
    copy_back(r, c, p, N);
    reb_simulation_free(r);
    return err;
}

/* Repeats the force evaluation n times (timing); returns seconds per evaluation in *sec. */
int refh_gravity_timed(rebcu_config* c, rebcu_particle* p, uint64_t* N, int n_evals, double* sec){
    struct reb_simulation* r = make_sim(c, p, *N);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k=0;k<n_evals;k++) reb_simulation_update_acceleration(r);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *sec = ((t1.tv_sec-t0.tv_sec) + 1e-9*(t1.tv_nsec-t0.tv_nsec))/n_evals;
    int err = collect_error(r);
    copy_back(r, c, p, N);
    reb_simulation_free(r);
    return err;
}

/* reb_boundary_check (boundary.c:35). */
int refh_boundary_check(rebcu_config* c, rebcu_particle* p, uint64_t* N){
    struct reb_simulation* r = make_sim(c, p, *N);
    reb_boundary_check(r);
    int err = collect_error(r);
    copy_back(r, c, p, N);
    reb_simulation_free(r);
    return err;
}

/* One integrator step only (integrator.callbacks.step), no boundary / collision pass. */
int refh_integrator_step(rebcu_config* c, rebcu_particle* p, uint64_t* N){
    struct reb_simulation* r = make_sim(c, p, *N);
    if (r->integrator.callbacks.step) r->integrator.callbacks.step(r, r->integrator.state);
    int err = collect_error(r);
    copy_back(r, c, p, N);
    reb_simulation_free(r);
    return err;
}

static enum REB_COLLISION_RESOLVE_OUTCOME resolve_none(struct reb_simulation* const r, struct reb_collision c){
    (void)r; (void)c;
    return REB_COLLISION_RESOLVE_OUTCOME_REMOVE_NONE;
}

static double restitution_bridges(const struct reb_simulation* const r, double v){
    // examples/shearing_sheet/problem.c:96-103
    (void)r;
    double eps = 0.32*pow(fabs(v)*100.,-0.234);
    if (eps>1) eps=1;
    if (eps<0) eps=0;
    return eps;
}

/* resolve: 0 = record only (no state change), 1 = reference hardsphere (eps=1),
 *          2 = reference hardsphere with the Bridges restitution law of examples/shearing_sheet. */
static void install_resolve(struct reb_simulation* r, int resolve, double minimum_collision_velocity){
    r->minimum_collision_velocity = minimum_collision_velocity;
    switch (resolve){
        case 1: r->collision_resolve = reb_collision_resolve_hardsphere; break;
        case 2: r->collision_resolve = reb_collision_resolve_hardsphere;
                r->coefficient_of_restitution = restitution_bridges; break;
        default: r->collision_resolve = resolve_none;
    }
}

/* Undo the Fisher-Yates-like shuffle of collision.c:337-342 so that the list is returned in the
 * order the search produced it.  seed0 is r->rand_seed before the search. */
static void unshuffle(struct reb_collision* list, size_t n, unsigned int seed0){
    if (n==0) return;
    size_t* swaps = malloc(n*sizeof(size_t));
    unsigned int seed = seed0;
    for (size_t i=0;i<n;i++) swaps[i] = rand_r(&seed)%n;
    for (size_t i=n;i-->0;){
        struct reb_collision c1 = list[i];
        list[i] = list[swaps[i]];
        list[swaps[i]] = c1;
    }
    free(swaps);
}

/* reb_collision_search (collision.c:49) with a no-op resolve; list returned in search order. */
int refh_collision_search(rebcu_config* c, rebcu_particle* p, uint64_t N,
                          rebcu_collision* out, uint64_t cap, uint64_t* n_found){
    struct reb_simulation* r = make_sim(c, p, N);
    install_resolve(r, 0, 0.);
    unsigned int seed0 = r->rand_seed;
    reb_collision_search(r);
    int err = collect_error(r);
    unshuffle(r->collisions, r->N_collisions, seed0);
    *n_found = r->N_collisions;
    size_t ncopy = r->N_collisions < cap ? r->N_collisions : cap;
    if (out && ncopy) memcpy(out, r->collisions, ncopy*sizeof(struct reb_collision));
    reb_simulation_free(r);
    return err;
}

/* reb_collision_search with r->map / r->N_map / r->N_targets set (rebound.h:257-258,344). */
int refh_collision_search_subset(rebcu_config* c, rebcu_particle* p, uint64_t N,
                                 const uint64_t* map, uint64_t N_map, uint64_t N_targets,
                                 rebcu_collision* out, uint64_t cap, uint64_t* n_found){
    struct reb_simulation* r = make_sim(c, p, N);
    install_resolve(r, 0, 0.);
    r->map = (size_t*)map;                 /* not owned by the simulation (rebound.h:258) */
    r->N_map = map ? (size_t)N_map : 0;
    r->N_targets = (size_t)N_targets;
    unsigned int seed0 = r->rand_seed;
    reb_collision_search(r);
    int err = collect_error(r);
    unshuffle(r->collisions, r->N_collisions, seed0);
    *n_found = r->N_collisions;
    size_t ncopy = r->N_collisions < cap ? r->N_collisions : cap;
    if (out && ncopy) memcpy(out, r->collisions, ncopy*sizeof(struct reb_collision));
    r->map = NULL;
    reb_simulation_free(r);
    return err;
}

/* reb_simulation_steps (simulation.c:504).  aux[0] <- collisions_log_n, aux[1] <- collisions_plog,
 * aux[2] <- wall seconds spent in reb_simulation_steps. */
int refh_steps(rebcu_config* c, rebcu_particle* p, uint64_t* N, uint64_t n_steps,
               int resolve, double minimum_collision_velocity, double* aux){
    struct reb_simulation* r = make_sim(c, p, *N);
    install_resolve(r, resolve, minimum_collision_velocity);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    reb_simulation_steps(r, n_steps);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    int err = collect_error(r);
    if (aux){
        aux[0] = (double)r->collisions_log_n;
        aux[1] = r->collisions_plog;
        aux[2] = (t1.tv_sec-t0.tv_sec) + 1e-9*(t1.tv_nsec-t0.tv_nsec);
    }
    copy_back(r, c, p, N);
    reb_simulation_free(r);
    return err;
}

/* The exit checks of run_heartbeat (simulation.c:242-272; static): reb_simulation_integrate runs the heartbeat once
 * before its loop (:392) and leaves as soon as the status is non-negative, so integrating to the current time
 * evaluates exactly those checks.  Returns r->status (4 escape, 3 encounter, 0 = REB_STATUS_SUCCESS: neither). */
int refh_exit_check(rebcu_config* c, rebcu_particle* p, uint64_t N, double exit_max_distance, double exit_min_distance){
    struct reb_simulation* r = make_sim(c, p, N);
    r->exit_max_distance = exit_max_distance;
    r->exit_min_distance = exit_min_distance;
    r->integrator.callbacks.step = NULL;        /* should a step run after all, it does nothing */
    r->gravity = REB_GRAVITY_NONE; r->collision = REB_COLLISION_NONE; r->boundary = REB_BOUNDARY_NONE;
    r->exact_finish_time = 1;
    int status = (int)reb_simulation_integrate(r, r->t);
    reb_simulation_free(r);
    return status;
}

/* reb_simulation_energy (tools.c:108). */
double refh_energy(rebcu_config* c, rebcu_particle* p, uint64_t N){
    struct reb_simulation* r = make_sim(c, p, N);
    double e = reb_simulation_energy(r);
    reb_simulation_free(r);
    return e;
}

/* reb_simulation_com (tools.c:410): out = {m,x,y,z,vx,vy,vz,ax,ay,az}. */
void refh_com(rebcu_config* c, rebcu_particle* p, uint64_t N, double* out){
    struct reb_simulation* r = make_sim(c, p, N);
    struct reb_particle com = reb_simulation_com(r);
    out[0]=com.m; out[1]=com.x; out[2]=com.y; out[3]=com.z; out[4]=com.vx; out[5]=com.vy; out[6]=com.vz;
    out[7]=com.ax; out[8]=com.ay; out[9]=com.az;
    reb_simulation_free(r);
}

/* reb_simulation_angular_momentum (tools.c:164). */
void refh_angular_momentum(rebcu_config* c, rebcu_particle* p, uint64_t N, double* out){
    struct reb_simulation* r = make_sim(c, p, N);
    struct reb_vec3d L = reb_simulation_angular_momentum(r);
    out[0]=L.x; out[1]=L.y; out[2]=L.z;
    reb_simulation_free(r);
}

static size_t dump_cell(const struct reb_treecell* node, int depth, int rootbox,
                        rebcu_treecell* out, uint64_t cap, size_t idx){
    size_t me = idx;
    if (me < cap){
        out[me].x = node->x; out[me].y = node->y; out[me].z = node->z; out[me].w = node->w;
        out[me].m = node->m; out[me].mx = node->mx; out[me].my = node->my; out[me].mz = node->mz;
        out[me].pt = node->pt; out[me].depth = depth; out[me].rootbox = rootbox;
    }
    idx++;
    for (int o=0;o<8;o++){
        if (node->oct[o]) idx = dump_cell(node->oct[o], depth+1, rootbox, out, cap, idx);
    }
    if (me < cap) out[me].skip = (int32_t)idx;
    return idx;
}

/* reb_tree_construct + reb_tree_calculate_gravity_data (tree.c:254, 209), dumped in depth-first
 * pre-order, root boxes ascending, octants ascending. */
int refh_tree_dump(rebcu_config* c, rebcu_particle* p, uint64_t N,
                   rebcu_treecell* out, uint64_t cap, uint64_t* n_cells){
    struct reb_simulation* r = make_sim(c, p, N);
    reb_tree_construct(r);
    int err = collect_error(r);
    size_t idx = 0;
    if (r->tree_root){
        reb_tree_calculate_gravity_data(r);
        size_t N_root = r->N_root_x*r->N_root_y*r->N_root_z;
        for (size_t i=0;i<N_root;i++){
            if (r->tree_root[i]) idx = dump_cell(r->tree_root[i], 0, (int)i, out, cap, idx);
        }
        reb_tree_delete(r);
    }
    *n_cells = idx;
    reb_simulation_free(r);
    return err;
}

/* Initial-condition generators of the reference, so that fixtures can be produced from the
 * reference's own recipes (tools.c:463-502, examples/selfgravity_plummer/problem.c:26-45). */
int refh_make_plummer(uint64_t N, double M, double R, unsigned int seed, rebcu_particle* out){
    struct reb_simulation* r = reb_simulation_create();
    r->rand_seed = seed;
    reb_simulation_add_plummer(r, N, M, R);
    reb_simulation_move_to_com(r);
    for (size_t i=0;i<r->N && i<N;i++){
        memset(&out[i], 0, sizeof(rebcu_particle));
        memcpy(&out[i], &r->particles[i], 11*sizeof(double));
    }
    reb_simulation_free(r);
    return 0;
}

/* ---- bounded samples of the large tree workloads (bench.py: cpu_baseline and --impl reference) ------------------
 * A full step of BASELINE.json's C4 (N = 2^24) takes the reference minutes.  A session runs the serial phases of
 * reb_gravity_tree_calculate_acceleration (gravity.c:47-106) once on the FULL problem, each timed -- boundary check,
 * reb_tree_construct, reb_tree_calculate_gravity_data -- and then walks the tree for a SAMPLE of the particles
 * (every stride-th one, all ghost boxes, the reference's own per-particle function and OpenMP schedule), so that
 *   seconds per step ~= t_boundary + t_construct + t_gravity_data + t_walk_sample * N / n_sample + t_delete + t_rest
 * with every term measured on the real 2^24-particle tree.  t_rest = one reb_simulation_steps(r,1) with gravity
 * switched off (drift, kick, drift, boundary check). */
struct refh_session { struct reb_simulation* r; };

void* refh_tree_open(rebcu_config* c, rebcu_particle* p, uint64_t N, double* sec3){
    struct refh_session* s = calloc(1, sizeof(*s));
    s->r = make_sim(c, p, N);
    struct reb_simulation* r = s->r;
    struct timespec t0, t1, t2, t3;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    reb_boundary_check(r);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    reb_tree_construct(r);
    clock_gettime(CLOCK_MONOTONIC, &t2);
    if (collect_error(r) || !r->tree_root){ reb_simulation_free(r); free(s); return NULL; }
    reb_tree_calculate_gravity_data(r);
    clock_gettime(CLOCK_MONOTONIC, &t3);
    sec3[0] = (t1.tv_sec-t0.tv_sec) + 1e-9*(t1.tv_nsec-t0.tv_nsec);
    sec3[1] = (t2.tv_sec-t1.tv_sec) + 1e-9*(t2.tv_nsec-t1.tv_nsec);
    sec3[2] = (t3.tv_sec-t2.tv_sec) + 1e-9*(t3.tv_nsec-t2.tv_nsec);
    return s;
}

uint64_t refh_tree_session_N(void* session){ return ((struct refh_session*)session)->r->N; }

/* The walk of gravity.c:76-99 for particles offset, offset+stride, ...; returns the wall seconds and the count. */
int refh_tree_walk_sample(void* session, uint64_t stride, uint64_t offset, double* sec, uint64_t* n_walked){
    struct reb_simulation* r = ((struct refh_session*)session)->r;
    struct reb_particle* const particles = r->particles;
    const size_t N = r->N;
    if (stride == 0) stride = 1;
    const size_t n_s = offset < N ? (N - offset + stride - 1)/stride : 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel for schedule(guided)
    for (size_t k=0; k<n_s; k++){
        const size_t i = offset + k*stride;
        particles[i].ax = 0; particles[i].ay = 0; particles[i].az = 0;
    }
    for (int gbx=-r->N_ghost_x; gbx<=r->N_ghost_x; gbx++){
        for (int gby=-r->N_ghost_y; gby<=r->N_ghost_y; gby++){
            for (int gbz=-r->N_ghost_z; gbz<=r->N_ghost_z; gbz++){
#pragma omp parallel for schedule(guided)
                for (size_t k=0; k<n_s; k++){
                    const size_t i = offset + k*stride;
                    struct reb_vec6d gb = reb_boundary_get_ghostbox(r, gbx,gby,gbz);
                    gb.x += particles[i].x;
                    gb.y += particles[i].y;
                    gb.z += particles[i].z;
                    reb_tree_calculate_acceleration_for_particle(r, (int)i, gb);
                }
            }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *sec = (t1.tv_sec-t0.tv_sec) + 1e-9*(t1.tv_nsec-t0.tv_nsec);
    *n_walked = n_s;
    return 0;
}

/* Accelerations of the sampled particles (parity spot checks at full size): out = ax,ay,az per sampled particle. */
int refh_tree_sample_acc(void* session, uint64_t stride, uint64_t offset, double* out, uint64_t cap){
    struct reb_simulation* r = ((struct refh_session*)session)->r;
    uint64_t k = 0;
    for (size_t i=offset; i<r->N && k<cap; i+=stride, k++){
        out[3*k] = r->particles[i].ax; out[3*k+1] = r->particles[i].ay; out[3*k+2] = r->particles[i].az;
    }
    return (int)k;
}

/* sec2 = {reb_tree_delete, one reb_simulation_steps(r,1) with gravity and collisions switched off}. */
int refh_tree_close(void* session, double* sec2){
    struct refh_session* s = session;
    struct reb_simulation* r = s->r;
    struct timespec t0, t1, t2;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    reb_tree_delete(r);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    r->gravity = REB_GRAVITY_NONE; r->collision = REB_COLLISION_NONE;
    reb_simulation_steps(r, 1);
    clock_gettime(CLOCK_MONOTONIC, &t2);
    sec2[0] = (t1.tv_sec-t0.tv_sec) + 1e-9*(t1.tv_nsec-t0.tv_nsec);
    sec2[1] = (t2.tv_sec-t1.tv_sec) + 1e-9*(t2.tv_nsec-t1.tv_nsec);
    reb_simulation_free(r);
    free(s);
    return 0;
}

