/*
 * mock_engine.c -- TEST INFRASTRUCTURE, never shipped and never loaded by the product.
 *
 * A stand-in for rebound_b200/librebound_b200.so that implements the part of the C ABI (include/rebound_b200.h)
 * the drop-in shim calls, on the CPU, by delegating every operation to the oracle (oracle/liboracle.so).  It exists
 * for one purpose: to run the HOST LOGIC of the shim sources in rebound_b200/shim -- residency modes, lazy host/device coherence,
 * the device batches behind reb_simulation_steps / reb_simulation_integrate, the collision-subset and exit-check
 * forwarding, error propagation -- in the CPU test suite (tests/test_hostlogic_cpu.py), where no GPU exists.  The
 * shim objects are the product's own (rebound_b200/_dropin/obj/shim_*.o); only the engine underneath is replaced.
 * Since the oracle is bit-identical to the reference, the mock drop-in must reproduce the reference driver's output
 * bit for bit in every scenario and residency mode; any difference is a bug in the shim.
 *
 * "Device" state = a private copy of the particle array: upload/download are memcpys, so stale-copy bugs in the
 * shim (reading r->particles while the device is ahead, or the reverse) show up exactly as they would on a GPU.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <signal.h>
#include "../../include/rebound_b200.h"

const char* orc_last_error(void);
int orc_gravity(rebcu_config* c, rebcu_particle* p, uint64_t* N);
int orc_gravity_cs(rebcu_config* c, rebcu_particle* p, uint64_t* N, double* cs_out);
int orc_boundary_check(rebcu_config* c, rebcu_particle* p, uint64_t* N);
int orc_integrator_step(rebcu_config* c, rebcu_particle* p, uint64_t* N);
int orc_collision_search_subset(rebcu_config* c, rebcu_particle* p, uint64_t N, const uint64_t* map, uint64_t N_map,
                                uint64_t N_targets, rebcu_collision* out, uint64_t cap, uint64_t* n_found);
int orc_steps(rebcu_config* c, rebcu_particle* p, uint64_t* N, uint64_t n_steps, int resolve, double mcv, double* aux);
int orc_exit_check(rebcu_config* c, rebcu_particle* p, uint64_t N, double exit_max_distance, double exit_min_distance);
int orc_apply_jerk(rebcu_config* c, rebcu_particle* p, uint64_t N, double v);

struct rebcu_handle {
    rebcu_particle* p; uint64_t N, cap; int resident;
    double* cs; uint64_t cs_n; int cs_valid;
    rebcu_collision* col; uint64_t col_n, col_cap;
    uint64_t* map; uint64_t map_n; int map_on; uint64_t n_targets;
    const volatile int* interrupt;
    int (*collision_hook)(void*); void* collision_hook_user;
    char err[512];
    /* statistics the tests read through mock_counters(): how often the shim moved the particle array */
};
static unsigned long long n_uploads, n_downloads, n_step_calls, n_steps_total;
void mock_counters(unsigned long long* out4){ out4[0]=n_uploads; out4[1]=n_downloads; out4[2]=n_step_calls; out4[3]=n_steps_total; }
/* MOCK_ENGINE_STATS=<file>: the counters are written there when the process exits */
static void write_stats(void){
    const char* path = getenv("MOCK_ENGINE_STATS");
    if (!path) return;
    FILE* f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "{\"uploads\": %llu, \"downloads\": %llu, \"step_calls\": %llu, \"steps\": %llu}\n", n_uploads, n_downloads, n_step_calls, n_steps_total);
    fclose(f);
}

static int fail(rebcu_handle* h, int code, const char* msg){ strncpy(h->err, msg, sizeof(h->err)-1); return code; }
static int from_oracle(rebcu_handle* h, int err){ if (err) strncpy(h->err, orc_last_error(), sizeof(h->err)-1); return err; }

rebcu_handle* rebcu_create(int device, void* stream){
    (void)stream;
    if (device != 0) return NULL;
    rebcu_handle* h = calloc(1, sizeof(*h));
    h->n_targets = REBCU_SIZE_MAX;
    static int registered = 0;
    if (!registered){ registered = 1; atexit(write_stats); }
    return h;
}
/* the mock has one "device": a group request degenerates to a single handle */
rebcu_handle* rebcu_create_group(const int* devices, int n){ (void)n; return rebcu_create(devices[0], NULL); }
int rebcu_set_sharded_build(rebcu_handle* h, int mode){ (void)h; (void)mode; return 0; }
void rebcu_destroy(rebcu_handle* h){ if (!h) return; free(h->p); free(h->cs); free(h->col); free(h->map); free(h); }
const char* rebcu_last_error(const rebcu_handle* h){ return h->err; }
int rebcu_host_register(void* ptr, uint64_t bytes){ (void)ptr; (void)bytes; return 0; }
int rebcu_host_unregister(void* ptr){ (void)ptr; return 0; }
uint64_t rebcu_N(const rebcu_handle* h){ return h->N; }
int rebcu_set_interrupt_flag(rebcu_handle* h, const volatile int* flag){ h->interrupt = flag; return 0; }

int rebcu_upload(rebcu_handle* h, const rebcu_particle* particles, uint64_t N){
    if (N > h->cap){ free(h->p); h->cap = N + N/4 + 16; h->p = malloc(h->cap*sizeof(rebcu_particle)); }
    if (N) memcpy(h->p, particles, N*sizeof(rebcu_particle));
    h->N = N; h->resident = 1; h->cs_valid = 0;
    n_uploads++;
    return 0;
}
int rebcu_download(rebcu_handle* h, rebcu_particle* particles, uint64_t N){
    if (!h->resident) return fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    if (N > h->N) return fail(h, REBCU_ERR_CAPACITY, "download beyond N");
    if (N) memcpy(particles, h->p, N*sizeof(rebcu_particle));
    n_downloads++;
    return 0;
}
int rebcu_download_gravity_cs(rebcu_handle* h, double* out_xyz, uint64_t N){
    if (!h->cs_valid || N > h->cs_n) return fail(h, REBCU_ERR_ARG, "no compensated evaluation to read gravity_cs from");
    memcpy(out_xyz, h->cs, 3*N*sizeof(double));
    return 0;
}

int rebcu_update_acceleration(rebcu_handle* h, rebcu_config* cfg){
    if (!h->resident) return fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    if (cfg->gravity == REBCU_GRAVITY_COMPENSATED){
        if (h->cs_n < h->N){ free(h->cs); h->cs = malloc(3*(h->N+1)*sizeof(double)); h->cs_n = h->N; }
        const int err = from_oracle(h, orc_gravity_cs(cfg, h->p, &h->N, h->cs));
        h->cs_valid = !err;
        return err;
    }
    return from_oracle(h, orc_gravity(cfg, h->p, &h->N));
}
int rebcu_gravity_host(rebcu_handle* h, rebcu_config* cfg, rebcu_particle* particles, uint64_t* N){
    int err = rebcu_upload(h, particles, *N);
    if (!err) err = rebcu_update_acceleration(h, cfg);
    if (err) return err;
    *N = h->N;
    return rebcu_download(h, particles, *N);
}
int rebcu_integrator_step(rebcu_handle* h, rebcu_config* cfg){
    if (!h->resident) return fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    n_step_calls++; n_steps_total++;
    return from_oracle(h, orc_integrator_step(cfg, h->p, &h->N));
}
int rebcu_boundary_check(rebcu_handle* h, rebcu_config* cfg){
    if (!h->resident) return fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    return from_oracle(h, orc_boundary_check(cfg, h->p, &h->N));
}
int rebcu_collision_search(rebcu_handle* h, const rebcu_config* cfg, rebcu_collision* out, uint64_t cap, uint64_t* n_found);
int rebcu_set_collision_callback(rebcu_handle* h, int (*cb)(void*), void* user){ h->collision_hook = cb; h->collision_hook_user = user; return 0; }
int rebcu_group_size(const rebcu_handle* h){ (void)h; return 1; }

int rebcu_steps(rebcu_handle* h, rebcu_config* cfg, uint64_t n_steps){
    if (!h->resident) return fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    n_step_calls++;
    for (uint64_t s=0; s<n_steps; s++){
        if (h->interrupt && *h->interrupt > 1) return REBCU_INTERRUPTED;
        int err;
        if (cfg->collision != REBCU_COLLISION_NONE && h->collision_hook){
            /* the engine's step with a collision hook: integrator step, boundary check, search, hook (csrc/api.cu) */
            err = from_oracle(h, orc_integrator_step(cfg, h->p, &h->N));
            if (!err) err = from_oracle(h, orc_boundary_check(cfg, h->p, &h->N));
            uint64_t nf = 0;
            if (!err) err = rebcu_collision_search(h, cfg, NULL, 0, &nf);
            if (!err) err = h->collision_hook(h->collision_hook_user);
        }else{
            err = from_oracle(h, orc_steps(cfg, h->p, &h->N, 1, 0, 0., NULL));
        }
        if (err) return err;
        n_steps_total++;
        {   /* MOCK_SIGINT_AT_STEP=<k>[,<times>]: the user's Ctrl-C arrives while the k-th step of the process runs */
            static long at = -2; static int times = 1;
            if (at == -2){ const char* e = getenv("MOCK_SIGINT_AT_STEP"); at = e ? atol(e) : -1; const char* c = e ? strchr(e, ',') : NULL; if (c) times = atoi(c+1); }
            if (at >= 0 && (long)n_steps_total == at) for (int k=0;k<times;k++) raise(SIGINT);
        }
    }
    return 0;
}
int rebcu_steps_host(rebcu_handle* h, rebcu_config* cfg, rebcu_particle* particles, uint64_t* N, uint64_t n_steps){
    int err = rebcu_upload(h, particles, *N);
    if (!err) err = rebcu_steps(h, cfg, n_steps);
    if (err) return err;
    *N = h->N;
    return rebcu_download(h, particles, *N);
}

int rebcu_set_collision_subset(rebcu_handle* h, const uint64_t* map, uint64_t N_map, uint64_t N_targets){
    h->n_targets = N_targets;
    h->map_on = map != NULL;
    h->map_n = map ? N_map : 0;
    free(h->map); h->map = NULL;
    if (map && N_map){ h->map = malloc(N_map*sizeof(uint64_t)); memcpy(h->map, map, N_map*sizeof(uint64_t)); }
    return 0;
}
int rebcu_collisions_fetch(rebcu_handle* h, rebcu_collision* out, uint64_t cap, uint64_t* n_found){
    *n_found = h->col_n;
    const uint64_t n = h->col_n < cap ? h->col_n : cap;
    if (n && out) memcpy(out, h->col, n*sizeof(rebcu_collision));
    return 0;
}
int rebcu_collision_search(rebcu_handle* h, const rebcu_config* cfg, rebcu_collision* out, uint64_t cap, uint64_t* n_found){
    if (!h->resident) return fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    rebcu_config c = *cfg;
    static const uint64_t empty_map[1] = {0};
    const uint64_t* map = h->map_on ? (h->map ? h->map : empty_map) : NULL;
    uint64_t n = 0;
    if (h->col_cap == 0){ h->col_cap = 1024; h->col = malloc(h->col_cap*sizeof(rebcu_collision)); }
    int err = from_oracle(h, orc_collision_search_subset(&c, h->p, h->N, map, h->map_n, h->n_targets, h->col, h->col_cap, &n));
    if (!err && n > h->col_cap){
        free(h->col); h->col_cap = n + n/4; h->col = malloc(h->col_cap*sizeof(rebcu_collision));
        err = from_oracle(h, orc_collision_search_subset(&c, h->p, h->N, map, h->map_n, h->n_targets, h->col, h->col_cap, &n));
    }
    h->col_n = err ? 0 : n;
    if (err) return err;
    return rebcu_collisions_fetch(h, out, cap, n_found);
}

int rebcu_exit_check(rebcu_handle* h, double exit_max_distance, double exit_min_distance, int* escape, int* encounter){
    if (!h->resident) return fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    rebcu_config c; memset(&c, 0, sizeof(c));
    *escape = orc_exit_check(&c, h->p, h->N, exit_max_distance, 0.) == 4;
    *encounter = orc_exit_check(&c, h->p, h->N, 0., exit_min_distance) == 3;
    return 0;
}
int rebcu_apply_jerk(rebcu_handle* h, const rebcu_config* cfg, double v){
    if (!h->resident) return fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    rebcu_config c = *cfg;
    return from_oracle(h, orc_apply_jerk(&c, h->p, h->N, v));
}
int rebcu_jerk_host(rebcu_handle* h, const rebcu_config* cfg, rebcu_particle* particles, uint64_t N, double v){
    int err = rebcu_upload(h, particles, N);
    if (!err) err = rebcu_apply_jerk(h, cfg, v);
    if (err) return err;
    return rebcu_download(h, particles, N);
}

/* The exact resolve of the engine (csrc/resolve.cu), sequentially: shuffle, then every collision in order -- early exits
 * of the hard-sphere resolver here, the arithmetic in the caller's resolver, one pair per call. */
#include <stdlib.h>
int rebcu_collision_resolve_pairs(rebcu_handle* h, unsigned int* rand_seed, rebcu_pair_resolver fn, void* user,
                                  double* plog, uint64_t* log_n, int* rounds){
    const uint64_t n = h->col_n;
    if (rounds) *rounds = 1;
    for (uint64_t i=0;i<n;i++){
        const uint64_t j = (uint64_t)rand_r(rand_seed)%n;
        rebcu_collision c1 = h->col[i]; h->col[i] = h->col[j]; h->col[j] = c1;
    }
    for (uint64_t i=0;i<n;i++){
        const rebcu_collision c = h->col[i];
        const rebcu_particle* a = &h->p[c.p1]; const rebcu_particle* b = &h->p[c.p2];
        const double x21 = a->x + c.gb.x - b->x, y21 = a->y + c.gb.y - b->y, z21 = a->z + c.gb.z - b->z;
        const double rp = a->r + b->r;
        if (rp*rp < x21*x21 + y21*y21 + z21*z21) continue;
        const double vx21 = a->vx + c.gb.vx - b->vx, vy21 = a->vy + c.gb.vy - b->vy, vz21 = a->vz + c.gb.vz - b->vz;
        if (vx21*x21 + vy21*y21 + vz21*z21 > 0) continue;
        rebcu_resolve_pair q;
        memset(&q, 0, sizeof(q));
        q.k = i; q.p1 = c.p1; q.p2 = c.p2; q.gb = c.gb;
        q.s1[0]=a->x; q.s1[1]=a->y; q.s1[2]=a->z; q.s1[3]=a->vx; q.s1[4]=a->vy; q.s1[5]=a->vz; q.s1[6]=a->m; q.s1[7]=a->r;
        q.s2[0]=b->x; q.s2[1]=b->y; q.s2[2]=b->z; q.s2[3]=b->vx; q.s2[4]=b->vy; q.s2[5]=b->vz; q.s2[6]=b->m; q.s2[7]=b->r;
        q.v1[0]=a->vx; q.v1[1]=a->vy; q.v1[2]=a->vz; q.v2[0]=b->vx; q.v2[1]=b->vy; q.v2[2]=b->vz;
        if (fn(user, &q, 1)) return fail(h, REBCU_ERR_ARG, "the pair resolver reported an error");
        h->p[c.p1].vx = q.v1[0]; h->p[c.p1].vy = q.v1[1]; h->p[c.p1].vz = q.v1[2];
        h->p[c.p2].vx = q.v2[0]; h->p[c.p2].vy = q.v2[1]; h->p[c.p2].vz = q.v2[2];
        if (q.logged){ *plog += q.plog_term; (*log_n)++; }
    }
    h->col_n = 0;
    return 0;
}
