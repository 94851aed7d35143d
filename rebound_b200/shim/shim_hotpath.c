/*
 * shim_hotpath.c -- the reference's hot-path symbols, forwarding to the CUDA engine.
 *
 * Defines (same names, same signatures as the reference):
 *   reb_gravity_basic_calculate_acceleration        src/gravity.c:167
 *   reb_gravity_compensated_calculate_acceleration  src/gravity.c:284
 *   reb_gravity_tree_calculate_acceleration         src/gravity.c:47
 *   reb_gravity_basic_calculate_and_apply_jerk      src/gravity.c:850
 *   reb_boundary_check                              src/boundary.c:35
 *   reb_collision_search                            src/collision.c:49
 * The reference's own definitions are compiled under the names *_cpuref (-D renames in
 * rebound_b200/shim/Makefile); they are only used for the modes outside the GPU hot path
 * (track_energy_offset, free_particle_ap, unknown collision modes).
 *
 * Default mode: host-authoritative -- every call uploads r->particles, runs on the GPU and writes
 * the result back, so every host hook of the reference keeps working unchanged.
 * REBOUND_B200_RESIDENT=1: see shim_integrators.c.
 */
#include <stdlib.h>
#include <string.h>
#include "shim_common.h"
#include "gravity.h"
#include "boundary.h"
#include "collision.h"
#include "particle.h"

void reb_boundary_check_cpuref(struct reb_simulation* r);
void reb_collision_search_cpuref(struct reb_simulation* const r);

/* ---- gravity ------------------------------------------------------------------------------- */
static void gravity_gpu(struct reb_simulation* r){
    struct shim_state* s = shim_get(r);
    if (!s) return;
    if (s->lazy && !s->host_stale) shim_lazy_open(s, 0);      /* host copy current but read-only (shim_lazy.c): the host path below writes it */
    rebcu_config c;
    shim_fill_config(r, &c);
    if (s->host_stale){
        /* resident mode: called from inside the device-side integrator step is impossible (the step
         * computes forces itself); a direct call while the host is stale works on the device copy. */
        int err = rebcu_update_acceleration(s->h, &c);
        if (shim_report(r, s, err)) return;
        r->N = rebcu_N(s->h);
        r->N_active = (c.N_active==REBCU_SIZE_MAX)?SIZE_MAX:(size_t)c.N_active;
        return;
    }
    uint64_t N = r->N;
    int err = rebcu_gravity_host(s->h, &c, (rebcu_particle*)r->particles, &N);
    s->device_valid = 0;
    if (shim_report(r, s, err)) return;
    if (N != r->N){                       /* tree gravity + open boundary removed particles (gravity.c:56) */
        r->N = N;
        r->did_modify_particles = 1;
    }
    r->N_active = (c.N_active==REBCU_SIZE_MAX)?SIZE_MAX:(size_t)c.N_active;
    if (r->gravity==REB_GRAVITY_COMPENSATED && r->N>0){
        /* r->gravity_cs is public state of the compensated routine (gravity.c:293-306); IAS15 reads it
         * (integrator_ias15.c:337-343) */
        if (r->N_allocated_gravity_cs<r->N){
            r->gravity_cs = realloc(r->gravity_cs, r->N*sizeof(struct reb_vec3d));
            r->N_allocated_gravity_cs = r->N;
        }
        shim_report(r, s, rebcu_download_gravity_cs(s->h, (double*)r->gravity_cs, r->N));
    }
}

void reb_gravity_basic_calculate_acceleration(struct reb_simulation* r){ gravity_gpu(r); }
void reb_gravity_compensated_calculate_acceleration(struct reb_simulation* r){ gravity_gpu(r); }
void reb_gravity_tree_calculate_acceleration(struct reb_simulation* r){
    if (r->boundary==REB_BOUNDARY_OPEN && (r->track_energy_offset || r->free_particle_ap || r->integrator.callbacks.will_remove_particle)){
        /* per-removal host side effects (boundary.c:66-72, particle.c:338-358): run the boundary pass on the host */
        struct shim_state* s = shim_get(r);
        if (!s) return;
        if (shim_to_host(r, s)) return;
        reb_boundary_check_cpuref(r);
        s->device_valid = 0;
    }
    gravity_gpu(r);
}

/* The jerk kick of the modified-kick schemes (gravity.c:850-924), called by EOS right after a force evaluation
 * (integrator_eos.c:101-103): positions, velocities and the accelerations just written go up, velocities come back. */
void reb_gravity_basic_calculate_and_apply_jerk(struct reb_simulation* r, const double v){
    struct shim_state* s = shim_get(r);
    if (!s) return;
    if (s->lazy && !s->host_stale) shim_lazy_open(s, 0);      /* host copy current but read-only (shim_lazy.c): the host path below writes it */
    rebcu_config c;
    shim_fill_config(r, &c);
    if (s->host_stale){
        shim_report(r, s, rebcu_apply_jerk(s->h, &c, v));
        return;
    }
    int err = rebcu_jerk_host(s->h, &c, (rebcu_particle*)r->particles, r->N, v);
    s->device_valid = 0;
    shim_report(r, s, err);
}

/* ---- boundary ------------------------------------------------------------------------------ */
void reb_boundary_check(struct reb_simulation* r){
    if (r->boundary==REB_BOUNDARY_NONE) return;
    struct shim_state* s = shim_get(r);
    if (!s) return;
    if (s->lazy && !s->host_stale) shim_lazy_open(s, 0);      /* host copy current but read-only (shim_lazy.c): the host path below writes it */
    if (r->boundary==REB_BOUNDARY_OPEN && (r->track_energy_offset || r->free_particle_ap || r->integrator.callbacks.will_remove_particle)){
        if (shim_to_host(r, s)) return;
        reb_boundary_check_cpuref(r);
        s->device_valid = 0;
        return;
    }
    rebcu_config c;
    shim_fill_config(r, &c);
    const int resident = s->host_stale;
    if (!resident){
        int err = rebcu_upload(s->h, (const rebcu_particle*)r->particles, r->N);
        if (shim_report(r, s, err)) return;
    }
    int err = rebcu_boundary_check(s->h, &c);
    if (shim_report(r, s, err)) return;
    const uint64_t N = rebcu_N(s->h);
    if (!resident){
        err = rebcu_download(s->h, (rebcu_particle*)r->particles, N);
        if (shim_report(r, s, err)) return;
        s->device_valid = 0;
    }
    if (N != r->N){
        if (N==0 && r->N>0) reb_simulation_warning(r, "Last particle removed.");      /* particle.c:346 */
        r->N = N;
        r->did_modify_particles = resident ? r->did_modify_particles : 1;
        s->uploaded_N = N;
    }
    r->N_active = (c.N_active==REBCU_SIZE_MAX)?SIZE_MAX:(size_t)c.N_active;
}

/* ---- collisions ---------------------------------------------------------------------------- */
/* The resolve loop without bringing the particles home (include/rebound_b200.h: rebcu_collision_resolve_pairs): the
 * engine keeps the shuffled order and hands over, round by round, the collisions whose particles are free; here the
 * reference's OWN reb_collision_resolve_hardsphere (src/collision.c:573-665) -- with the user's coefficient_of_restitution
 * callback -- runs on a two-particle scratch simulation per pair, on all cores (the pairs of a batch share no particle). */
#include <pthread.h>
#include <unistd.h>
struct pair_job { struct reb_simulation* r; rebcu_resolve_pair* pairs; uint64_t begin, end; };
static void* pair_worker(void* arg){
    struct pair_job* job = arg;
    struct reb_simulation rs = *job->r;                 /* scalar settings and callbacks of the caller's simulation */
    struct reb_particle two[2];
    rs.particles = two; rs.N = 2; rs.N_allocated = 2;
    for (uint64_t j=job->begin; j<job->end; j++){
        rebcu_resolve_pair* q = &job->pairs[j];
        memset(two, 0, sizeof(two));
        two[0].x = q->s1[0]; two[0].y = q->s1[1]; two[0].z = q->s1[2]; two[0].vx = q->s1[3]; two[0].vy = q->s1[4]; two[0].vz = q->s1[5];
        two[0].m = q->s1[6]; two[0].r = q->s1[7]; two[0].sim = &rs;
        two[1].x = q->s2[0]; two[1].y = q->s2[1]; two[1].z = q->s2[2]; two[1].vx = q->s2[3]; two[1].vy = q->s2[4]; two[1].vz = q->s2[5];
        two[1].m = q->s2[6]; two[1].r = q->s2[7]; two[1].sim = &rs;
        struct reb_collision c;
        memset(&c, 0, sizeof(c));
        c.p1 = 0; c.p2 = 1;
        memcpy(&c.gb, &q->gb, sizeof(c.gb));
        rs.collisions_plog = 0.; rs.collisions_log_n = 0;
        reb_collision_resolve_hardsphere(&rs, c);
        q->v1[0] = two[0].vx; q->v1[1] = two[0].vy; q->v1[2] = two[0].vz;
        q->v2[0] = two[1].vx; q->v2[1] = two[1].vy; q->v2[2] = two[1].vz;
        q->plog_term = rs.collisions_plog;
        q->logged = rs.collisions_log_n ? 1 : 0;
    }
    return NULL;
}
static int shim_pair_resolver(void* user, rebcu_resolve_pair* pairs, uint64_t n){
    struct reb_simulation* r = user;
    static long cores = 0;
    if (!cores){ cores = sysconf(_SC_NPROCESSORS_ONLN); if (cores < 1) cores = 1; if (cores > 32) cores = 32; }
    int nt = (int)(n/4096);                              /* a thread is worth a few thousand pairs */
    if (nt > cores) nt = (int)cores;
    if (nt <= 1){ struct pair_job job = {r, pairs, 0, n}; pair_worker(&job); return 0; }
    pthread_t th[32]; struct pair_job jobs[32];
    for (int t=0;t<nt;t++){
        jobs[t].r = r; jobs[t].pairs = pairs; jobs[t].begin = n*(uint64_t)t/(uint64_t)nt; jobs[t].end = n*(uint64_t)(t+1)/(uint64_t)nt;
        if (pthread_create(&th[t], NULL, pair_worker, &jobs[t])){ pair_worker(&jobs[t]); th[t] = 0; }
    }
    for (int t=0;t<nt;t++) if (th[t]) pthread_join(th[t], NULL);
    return 0;
}
/* Can this simulation's collisions be resolved without the particles on the host?  The built-in hard-sphere resolver
 * only (it touches nothing but the two particles and never removes one); REBOUND_B200_DEVICE_RESOLVE=0 switches it off. */
int shim_resolve_on_device(const struct reb_simulation* r, const struct shim_state* s){
    static int on = -1;
    if (on < 0){ const char* e = getenv("REBOUND_B200_DEVICE_RESOLVE"); on = (e && e[0]=='0') ? 0 : 1; }
    return on && r->collision_resolve==reb_collision_resolve_hardsphere && rebcu_group_size(s->h)==1;
}
int shim_resolve_pairs(struct reb_simulation* r, struct shim_state* s){
    uint64_t logn = (uint64_t)r->collisions_log_n;
    int err = rebcu_collision_resolve_pairs(s->h, &r->rand_seed, shim_pair_resolver, r, &r->collisions_plog, &logn, NULL);
    r->collisions_log_n = (int64_t)logn;
    return shim_report(r, s, err);
}

void reb_collision_search(struct reb_simulation* const r){
    const int gpu_mode = (r->collision==REB_COLLISION_DIRECT || r->collision==REB_COLLISION_TREE
                          || r->collision==REB_COLLISION_LINE || r->collision==REB_COLLISION_LINETREE);
    if (r->collision==REB_COLLISION_NONE) return;               /* collision.c:52 switch: nothing to do */
    if (!gpu_mode){
        /* outside the GPU path (unknown modes): the reference's routine on host data; a simulation that never
         * used the GPU does not get a device context for this */
        struct shim_state* s0 = shim_find(r);
        if (s0){ if (shim_to_host(r, s0)) return; s0->device_valid = 0; }
        reb_collision_search_cpuref(r);
        return;
    }
    struct shim_state* s = shim_get(r);
    if (!s) return;
    if (s->lazy && !s->host_stale) shim_lazy_open(s, 0);      /* host copy current but read-only (shim_lazy.c): the host path below writes it */
    r->N_collisions = 0;
    rebcu_config c;
    shim_fill_config(r, &c);
    if (!s->host_stale){
        int err = rebcu_upload(s->h, (const rebcu_particle*)r->particles, r->N);
        if (shim_report(r, s, err)) return;
    }
    /* r->map / r->N_map / r->N_targets (collision.c:53-58): MERCURIUS and TRACE search among the particles of a
     * close encounter only.  The subset is forwarded when it is set or was set by the previous search. */
    const int subset = r->map!=NULL || r->N_targets!=SIZE_MAX;
    if (subset || s->subset_set){
        int e2 = rebcu_set_collision_subset(s->h, (const uint64_t*)r->map, r->map ? r->N_map : 0,
                                            r->N_targets==SIZE_MAX ? REBCU_SIZE_MAX : (uint64_t)r->N_targets);
        if (shim_report(r, s, e2)) return;
        s->subset_set = subset;
    }
    uint64_t n_found = 0;
    if (s->host_stale && shim_resolve_on_device(r, s)){
        /* resident simulation + the built-in hard-sphere resolver: the list stays on the device and is resolved there,
         * pair batches travelling to the reference's own resolver and back (no particle array download) */
        int err = rebcu_collision_search(s->h, &c, NULL, 0, &n_found);
        if (shim_report(r, s, err)) return;
        r->N_collisions = 0;                            /* consumed on the device; r->collisions is not filled */
        if (n_found) shim_resolve_pairs(r, s);
        return;
    }
    int err = rebcu_collision_search(s->h, &c, (rebcu_collision*)r->collisions, r->N_allocated_collisions, &n_found);
    if (shim_report(r, s, err)) return;
    if (n_found > r->N_allocated_collisions){
        /* grow as collision.c:108-113 does (doubling from 32) and fetch the full list */
        size_t cap = r->N_allocated_collisions ? r->N_allocated_collisions : 32;
        while (cap < n_found) cap *= 2;
        r->collisions = realloc(r->collisions, sizeof(struct reb_collision)*cap);
        r->N_allocated_collisions = cap;
        err = rebcu_collisions_fetch(s->h, (rebcu_collision*)r->collisions, cap, &n_found);
        if (shim_report(r, s, err)) return;
    }
    r->N_collisions = n_found;
    if (n_found==0) return;
    /* The shuffle and the resolve loop work on r->particles (user callback ABI). */
    if (shim_to_host(r, s)) return;
    s->device_valid = 0;

    /* collision.c:336-342 */
    for (size_t i=0;i<r->N_collisions;i++){
        size_t new = rand_r(&(r->rand_seed))%r->N_collisions;
        struct reb_collision c1 = r->collisions[i];
        r->collisions[i] = r->collisions[new];
        r->collisions[new] = c1;
    }
    /* collision.c:345-404 */
    enum REB_COLLISION_RESOLVE_OUTCOME (*resolve) (struct reb_simulation* const r, struct reb_collision c) = r->collision_resolve;
    if (resolve==NULL) resolve = reb_collision_resolve_halt;
    for (size_t i=0;i<r->N_collisions;i++){
        struct reb_collision col = r->collisions[i];
        if (col.p1 == SIZE_MAX || col.p2 == SIZE_MAX) continue;
        enum REB_COLLISION_RESOLVE_OUTCOME outcome = resolve(r, col);
        for (int which=0; which<2; which++){
            const int bit = which==0 ? REB_COLLISION_RESOLVE_OUTCOME_REMOVE_P1 : REB_COLLISION_RESOLVE_OUTCOME_REMOVE_P2;
            if (!(outcome & bit)) continue;
            const size_t gone = which==0 ? col.p1 : col.p2;
            if (reb_simulation_remove_particle(r, gone)) continue;
            if (which==0 && col.p2 > col.p1 && col.p2!=SIZE_MAX) col.p2--;
            for (size_t j=i+1;j<r->N_collisions;j++){
                struct reb_collision* cp = &(r->collisions[j]);
                if (cp->p1==gone || cp->p2==gone){ cp->p1 = SIZE_MAX; cp->p2 = SIZE_MAX; }
                if (cp->p1 > gone && cp->p1!=SIZE_MAX) cp->p1--;
                if (cp->p2 > gone && cp->p2!=SIZE_MAX) cp->p2--;
            }
        }
    }
}
