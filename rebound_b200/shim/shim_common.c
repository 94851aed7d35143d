/* shim_common.c -- side table simulation -> device handle, config marshalling, residency helpers. */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shim_common.h"
#include "rebound_internal.h"   /* reb_sigint */
#include "integrator_leapfrog.h"

/* One entry per simulation that has used the replaced hot path.  Entries are individually allocated (their addresses
 * are handed out and must stay valid while the table grows) and released when the simulation is freed
 * (reb_simulation_free / reb_simulation_free_pointers, shim_steps.c), so neither a long parameter sweep nor a large
 * list of simultaneously alive simulations runs into a limit. */
static struct shim_state** table = NULL;
static int table_n = 0, table_cap = 0;
static pthread_mutex_t table_lock = PTHREAD_MUTEX_INITIALIZER;

struct shim_state* shim_get(struct reb_simulation* r){
    pthread_mutex_lock(&table_lock);
    struct shim_state* s = NULL;
    for (int i=0;i<table_n;i++) if (table[i] && table[i]->r==r){ s = table[i]; break; }
    if (!s){
        int slot = -1;
        for (int i=0;i<table_n && slot<0;i++) if (table[i]==NULL) slot = i;
        if (slot<0){
            if (table_n==table_cap){
                const int cap = table_cap ? 2*table_cap : 64;
                struct shim_state** t = realloc(table, cap*sizeof(*t));
                if (t){ table = t; table_cap = cap; }
            }
            if (table_n<table_cap) slot = table_n++;
        }
        if (slot>=0){
            s = calloc(1, sizeof(*s));
            table[slot] = NULL;
            int device = 0;
            const char* env = getenv("REBOUND_B200_DEVICE");
            if (env) device = atoi(env);
            /* REBOUND_B200_DEVICES="0-7" or "0,1,2,3" (a device may repeat): one multi-GPU group handle -- this program's
             * single call to reb_simulation_integrate() then shards its particles over those GPUs (include/rebound_b200.h) */
            int devs[16], n_dev = 0;
            const char* list = getenv("REBOUND_B200_DEVICES");
            if (list){
                const char* q = list;
                while (*q && n_dev < 16){
                    char* end;
                    long a = strtol(q, &end, 10);
                    if (end==q) break;
                    long b = a;
                    if (*end=='-'){ q = end+1; b = strtol(q, &end, 10); if (end==q) break; }
                    for (long d=a; d<=b && n_dev<16; d++) devs[n_dev++] = (int)d;
                    q = end;
                    if (*q==',') q++;
                }
            }
            if (s) s->h = (n_dev > 1) ? rebcu_create_group(devs, n_dev) : rebcu_create(n_dev==1 ? devs[0] : device, NULL);
            /* second Ctrl-C: leave multi-step device calls as the reference leaves its loops (src/rebound.c:193-200) */
            if (s && s->h){ const char* sb = getenv("REBOUND_B200_SHARD_BUILD"); if (sb) rebcu_set_sharded_build(s->h, atoi(sb)); }   /* 0 replicated, 1 per-rank tree builds, 2 automatic */
            if (s && s->h){ s->r = r; rebcu_set_interrupt_flag(s->h, (const volatile int*)&reb_sigint); table[slot] = s; }
            else { free(s); s = NULL; }
        }
    }
    pthread_mutex_unlock(&table_lock);
    if (!s) reb_simulation_error(r, "rebound_b200: no usable CUDA device; the GPU hot path has no CPU fallback.");
    return s;
}

struct shim_state* shim_find(struct reb_simulation* r){
    pthread_mutex_lock(&table_lock);
    struct shim_state* s = NULL;
    for (int i=0;i<table_n;i++) if (table[i] && table[i]->r==r){ s = table[i]; break; }
    pthread_mutex_unlock(&table_lock);
    return s;
}

void shim_forget(struct reb_simulation* r){
    pthread_mutex_lock(&table_lock);
    for (int i=0;i<table_n;i++) if (table[i] && table[i]->r==r){
        shim_lazy_forget(table[i]);
        if (table[i]->pinned_ptr) rebcu_host_unregister(table[i]->pinned_ptr);
        rebcu_destroy(table[i]->h);
        free(table[i]);
        table[i] = NULL;
    }
    pthread_mutex_unlock(&table_lock);
}

int shim_residency(void){
    static int mode = -1;
    if (mode<0){
        const char* env = getenv("REBOUND_B200_RESIDENT");
        mode = (env && env[0]=='0') ? SHIM_HOST_AUTHORITATIVE : (env && env[0]=='1') ? SHIM_RESIDENT : SHIM_AUTO;
    }
    return mode;
}

int shim_resident(const struct reb_simulation* r){
    const int mode = shim_residency();
    if (mode!=SHIM_AUTO) return mode==SHIM_RESIDENT;
    /* Between the steps of reb_simulation_steps / reb_simulation_integrate only these can look at r->particles
     * (run_heartbeat and reb_simulation_step, src/simulation.c:240-274, 514-603); without them nobody sees the host
     * copy before the reb_simulation_synchronize that ends the call (:455, :511). */
    /* exit distances: run_heartbeat's scans are evaluated on the device when nothing moves the particles between the
     * integrator step and the heartbeat (shim_integrators.c) */
    const int exits_ok = (!r->exit_max_distance && !r->exit_min_distance)
                      || (r->boundary==REB_BOUNDARY_NONE && r->collision==REB_COLLISION_NONE);
    /* a heartbeat alone does not end the residency when the particle array can be fetched on demand (shim_lazy.c) */
    return (!r->heartbeat || shim_lazy_possible(r)) && !r->pre_timestep_modifications && !r->post_timestep_modifications
        && exits_ok && !r->display_data && !r->server_data
        && !r->N_odes;     /* user ODEs are integrated on the host after every step and read r->particles (simulation.c:531-556) */
}

void shim_fill_config(const struct reb_simulation* r, rebcu_config* c){
    memset(c, 0, sizeof(*c));
    c->t = r->t; c->G = r->G; c->softening = r->softening;
    c->OMEGA = r->OMEGA; c->OMEGAZ = r->OMEGAZ;
    c->dt = r->dt; c->dt_last_done = r->dt_last_done;
    c->opening_angle2 = r->opening_angle2;
#ifdef QUADRUPOLE
    c->quadrupole = 1;           /* the reference sources of this build carry the quadrupole option (tree.h:43-50) */
#endif
    c->root_size = r->root_size;
    c->N_active = (r->N_active==SIZE_MAX)?REBCU_SIZE_MAX:(uint64_t)r->N_active;
    c->testparticle_type = r->testparticle_type;
    c->gravity_ignore_terms = (int32_t)r->gravity_ignore_terms;
    c->N_root_x = (int32_t)r->N_root_x; c->N_root_y = (int32_t)r->N_root_y; c->N_root_z = (int32_t)r->N_root_z;
    c->N_ghost_x = r->N_ghost_x; c->N_ghost_y = r->N_ghost_y; c->N_ghost_z = r->N_ghost_z;
    c->boundary = (int32_t)r->boundary;
    c->gravity = (int32_t)r->gravity;
    c->collision = (int32_t)r->collision;
    c->integrator = REBCU_INTEGRATOR_NONE;
    if (r->integrator.name && strcmp(r->integrator.name, "leapfrog")==0){
        c->integrator = REBCU_INTEGRATOR_LEAPFROG;
        const struct reb_integrator_leapfrog_state* st = r->integrator.state;
        c->leapfrog_order = st ? (int32_t)st->order : 2;
    }else if (r->integrator.name && strcmp(r->integrator.name, "sei")==0){
        c->integrator = REBCU_INTEGRATOR_SEI;
    }
    static int mode = -1;            /* read once: this runs before every replaced call */
    if (mode < 0){ const char* e = getenv("REBOUND_B200_MODE"); mode = (e && strcmp(e, "fast")==0) ? REBCU_MODE_FAST : REBCU_MODE_STRICT; }
    c->mode = mode;
}

int shim_report(struct reb_simulation* r, struct shim_state* s, int err){
    if (err) reb_simulation_error(r, rebcu_last_error(s->h));
    return err;
}

/* REBOUND_B200_PIN=1: page-lock r->particles so the PCIe copies run at full rate (pageable copies reach
 * about half).  Opt-in because librebound realloc()s the array when particles are added
 * (src/particle.c:53-58); the registration follows the pointer, but the old block is released by libc
 * before this code can unpin it, which is only safe if the run does not grow the array while stepping. */
static void shim_pin(struct reb_simulation* r, struct shim_state* s){
    static int want = -1;
    if (want < 0){ const char* e = getenv("REBOUND_B200_PIN"); want = (e && e[0]=='1') ? 1 : 0; }
    if (!want) return;
    const size_t bytes = r->N_allocated*sizeof(struct reb_particle);
    if (s->pinned_ptr==(void*)r->particles && s->pinned_bytes==bytes) return;
    if (s->pinned_ptr){ rebcu_host_unregister(s->pinned_ptr); s->pinned_ptr = NULL; s->pinned_bytes = 0; }
    if (r->particles && bytes >= (8u<<20) && rebcu_host_register(r->particles, bytes)==0){
        s->pinned_ptr = r->particles; s->pinned_bytes = bytes;
    }
}

int shim_to_device(struct reb_simulation* r, struct shim_state* s){
    /* a protected host copy that the device copy cannot replace (edited / moved / resized array): open it before the upload reads it */
    if (s->lazy && !(shim_resident(r) && s->device_valid && !r->did_modify_particles && s->uploaded_from==r->particles && s->uploaded_N==r->N))
        shim_lazy_open(s, 0);
    shim_pin(r, s);
    /* Only a resident simulation trusts the device copy across calls; otherwise every call uploads. */
    if (shim_resident(r) && s->device_valid && !r->did_modify_particles && s->uploaded_from==r->particles && s->uploaded_N==r->N) return 0;
    int err = rebcu_upload(s->h, (const rebcu_particle*)r->particles, r->N);
    if (err) return shim_report(r, s, err);
    s->uploaded_from = r->particles; s->uploaded_N = r->N;
    s->device_valid = 1;
    s->host_stale = 0;
    return 0;
}

int shim_to_host(struct reb_simulation* r, struct shim_state* s){
    if (s->lazy) shim_lazy_open(s, 1);      /* lifts the protection and downloads if the host is behind */
    if (!s->host_stale) return 0;
    const uint64_t n = rebcu_N(s->h);
    if (n > r->N_allocated) return shim_report(r, s, REBCU_ERR_CAPACITY);
    int err = rebcu_download(s->h, (rebcu_particle*)r->particles, n);
    if (err) return shim_report(r, s, err);
    r->N = n;
    s->uploaded_N = n;
    s->host_stale = 0;
    return 0;
}
