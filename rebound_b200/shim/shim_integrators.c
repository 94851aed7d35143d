/*
 * shim_integrators.c -- the `leapfrog` and `sei` entries of the reference's integrator registry,
 * backed by the fused kick/drift and SEI kernels.
 *
 * Defines the data symbols reb_integrator_leapfrog (src/integrator_leapfrog.c:32-48) and
 * reb_integrator_sei (src/integrator_sei.c:47-63) referenced by the X-macro in rebound.h:205 /
 * simulation.c:222-224.  The reference's own structs are compiled as *_cpuref; its step functions
 * (reb_integrator_leapfrog_step, reb_integrator_sei_step), create/free functions, field descriptor
 * lists and the lf4/lf6/lf8 constants keep their names and are reused here.
 *
 * A step runs entirely on the device when the force is one the engine owns (gravity NONE / BASIC /
 * COMPENSATED / TREE, no variational particles, no additional_forces callback).  Otherwise the
 * reference's host step runs and only its force call lands on the GPU (shim_hotpath.c).
 *
 * REBOUND_B200_RESIDENT=1: particles stay in HBM between steps; r->is_synchronized is cleared and the
 * `synchronize` callback (called by reb_simulation_synchronize at src/simulation.c:331,336,455,511,521,562)
 * brings r->particles up to date -- the protocol of the reference's WHFast integrators.
 */
#include <math.h>
#include <string.h>
#include "shim_common.h"
#include "integrator_leapfrog.h"
#include "integrator_sei.h"

void reb_integrator_leapfrog_step(struct reb_simulation* r, void* state);
void* reb_integrator_leapfrog_create();
void reb_integrator_leapfrog_free(void* p);
extern const struct reb_binarydata_field_descriptor reb_integrator_leapfrog_field_descriptor_list[];
void reb_integrator_sei_step(struct reb_simulation* r, void* state);
void* reb_integrator_sei_create();
void reb_integrator_sei_free(void* p);
extern const struct reb_binarydata_field_descriptor reb_integrator_sei_field_descriptor_list[];

static void sei_step(struct reb_simulation* r, void* state);

static int device_step_possible(const struct reb_simulation* r){
    if (r->N_var || r->additional_forces) return 0;
    switch (r->gravity){
        case REB_GRAVITY_NONE: case REB_GRAVITY_BASIC: case REB_GRAVITY_COMPENSATED: case REB_GRAVITY_TREE: break;
        default: return 0;
    }
    if (r->gravity==REB_GRAVITY_TREE && r->boundary==REB_BOUNDARY_OPEN
        && (r->track_energy_offset || r->free_particle_ap)) return 0;
    return 1;
}

/* The SEI state is a cache the reference fills at the top of its step and serialises with the simulation
 * (struct reb_integrator_sei_state is private to src/integrator_sei.c:33-39: lastdt, sindt, tandt, sindtz, tandtz; the
 * Python package exposes it as sim.integrator.lastdt).  The device step recomputes the same four numbers from OMEGA,
 * OMEGAZ and dt with the same libm calls, so the cache is kept exactly as the reference would leave it
 * (integrator_sei.c:91-101).  Returns 0 in the one case where the device step would NOT do what the reference
 * does: a cache that is valid for this dt but was computed for another OMEGA (the reference keeps using it). */
struct shim_sei_state { double lastdt, sindt, tandt, sindtz, tandtz; };
int shim_prepare_integrator_state(struct reb_simulation* r){
    if (r->integrator.callbacks.step!=sei_step || r->integrator.state==NULL) return 1;
    struct shim_sei_state* sei = r->integrator.state;
    if (sei->lastdt!=r->dt){
        if (r->OMEGAZ==-1) r->OMEGAZ = r->OMEGA;
        sei->sindt = sin(r->OMEGA*(-r->dt/2.));
        sei->tandt = tan(r->OMEGA*(-r->dt/4.));
        sei->sindtz = sin(r->OMEGAZ*(-r->dt/2.));
        sei->tandtz = tan(r->OMEGAZ*(-r->dt/4.));
        sei->lastdt = r->dt;
        return 1;
    }
    const double oz = (r->OMEGAZ==-1) ? r->OMEGA : r->OMEGAZ;       /* what the device step would use */
    return sei->sindt==sin(r->OMEGA*(-r->dt/2.)) && sei->tandt==tan(r->OMEGA*(-r->dt/4.))
        && sei->sindtz==sin(oz*(-r->dt/2.)) && sei->tandtz==tan(oz*(-r->dt/4.)) && r->OMEGAZ!=-1;
}

static void device_step(struct reb_simulation* r, void (*host_step)(struct reb_simulation*, void*), void* state){
    if (!device_step_possible(r) || !shim_prepare_integrator_state(r)){
        struct shim_state* s = shim_find(r);
        if (s){ if (shim_to_host(r, s)) return; s->device_valid = 0; }
        host_step(r, state);
        return;
    }
    struct shim_state* s = shim_get(r);
    if (!s) return;
    rebcu_config c;
    shim_fill_config(r, &c);
    if (shim_to_device(r, s)) return;
    int err = rebcu_integrator_step(s->h, &c);
    if (shim_report(r, s, err)) return;
    r->t = c.t;
    r->dt_last_done = c.dt_last_done;
    r->gravity_ignore_terms = c.gravity_ignore_terms;
    r->OMEGAZ = c.OMEGAZ;
    r->N_active = (c.N_active==REBCU_SIZE_MAX)?SIZE_MAX:(size_t)c.N_active;
    s->host_stale = 1;
    /* run_heartbeat (src/simulation.c:240-274) scans r->particles after every step when an exit distance is set.
     * When nothing else changes the particles between the integrator step and that scan (no boundary, no
     * collisions, no post-timestep hook), the same predicates are evaluated on the device here and the simulation
     * stays resident: the host scan then runs on the state of the last synchronisation, which had passed it.
     * Otherwise such simulations are kept host-current. */
    const int exits = r->exit_max_distance || r->exit_min_distance;
    const int exits_on_device = exits && r->boundary==REB_BOUNDARY_NONE && r->collision==REB_COLLISION_NONE
                                && !r->post_timestep_modifications && !r->heartbeat;
    /* user ODEs are integrated on the host right after this callback returns and read r->particles
     * (src/simulation.c:531-556): in every residency mode the host copy must be current then */
    if (shim_resident(r) && !r->N_odes && (!exits || exits_on_device)){
        r->N = rebcu_N(s->h);            /* tree gravity + open boundary may have removed particles */
        s->uploaded_N = r->N;
        r->is_synchronized = 0;
        if (r->heartbeat && shim_residency()==SHIM_AUTO) shim_lazy_protect(r, s);   /* the heartbeat may look: fetch on demand */
        if (exits){
            int escape = 0, encounter = 0;
            err = rebcu_exit_check(s->h, r->exit_max_distance, r->exit_min_distance, &escape, &encounter);
            if (shim_report(r, s, err)) return;
            if (escape) r->status = REB_STATUS_ESCAPE;             /* simulation.c:251-253 */
            if (encounter) r->status = REB_STATUS_ENCOUNTER;       /* simulation.c:267-269 */
        }
    }else{
        shim_to_host(r, s);
    }
}

static void synchronize(struct reb_simulation* r, void* state){
    (void)state;
    struct shim_state* s = shim_find(r);
    if (!s){ r->is_synchronized = 1; return; }
    if (shim_to_host(r, s)) return;
    r->is_synchronized = 1;
    /* automatic residency: the call is over, the host copy is the truth again (it may be edited without any flag).
     * Explicit residency keeps the device copy across synchronisations (host edits are flagged with
     * r->did_modify_particles) -- except when timestep-modification hooks are installed: the reference synchronises
     * right before it calls them (src/simulation.c:521-524, 562-565) precisely so that they can edit the particles. */
    if (shim_residency()==SHIM_AUTO || r->pre_timestep_modifications || r->post_timestep_modifications) s->device_valid = 0;
}

static void leapfrog_step(struct reb_simulation* r, void* state){ device_step(r, reb_integrator_leapfrog_step, state); }
static void sei_step(struct reb_simulation* r, void* state){ device_step(r, reb_integrator_sei_step, state); }

int shim_is_device_integrator(const struct reb_simulation* r){
    return (r->integrator.callbacks.step==leapfrog_step || r->integrator.callbacks.step==sei_step) && device_step_possible(r);
}

const struct reb_integrator reb_integrator_leapfrog = {
    .documentation = "Leapfrog (drift-kick-drift), orders 2, 4, 6 and 8; kick/drift run as fused CUDA kernels on the resident particle arrays.",
    .step = leapfrog_step,
    .synchronize = synchronize,
    .create = reb_integrator_leapfrog_create,
    .free = reb_integrator_leapfrog_free,
    .field_descriptor_list = reb_integrator_leapfrog_field_descriptor_list,
};

const struct reb_integrator reb_integrator_sei = {
    .documentation = "Symplectic Epicycle Integrator for the shearing sheet (Rein & Tremaine 2011); operators run as CUDA kernels on the resident particle arrays.",
    .step = sei_step,
    .synchronize = synchronize,
    .create = reb_integrator_sei_create,
    .free = reb_integrator_sei_free,
    .field_descriptor_list = reb_integrator_sei_field_descriptor_list,
};

/* Explicit release of the device state of a simulation (optional; everything is freed at exit). */
void reb_b200_release(struct reb_simulation* r){ shim_forget(r); }
