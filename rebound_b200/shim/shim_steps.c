/*
 * shim_steps.c -- reb_simulation_steps (src/simulation.c:504-513) and reb_simulation_integrate
 * (src/simulation.c:465-503) with the whole steps that need nothing from the host run as ONE device batch.
 *
 * The reference advances a simulation one reb_simulation_step at a time; between two steps only these can look at
 * r->particles: the heartbeat and the exit checks (run_heartbeat, :240-274), the Simulationarchive heartbeat
 * (:507), pre/post_timestep_modifications (:521-524, :562-565), a collision resolve callback, a viewer.  When none
 * of them is installed, a run of n steps is observably one operation on (particles, t, dt_last_done, steps_done,
 * walltime), and it is handed to rebcu_steps_host / rebcu_steps: the chunk-pipelined host path or the multi-step
 * launches of the engine instead of n round trips through the per-step callback.  The results are the same bits
 * (tests/test_gpu_dropin.py runs every scenario through here in automatic mode).
 *
 * reb_simulation_integrate keeps its exit logic: the batch stops at least two steps short of tmax and the
 * reference's own loop (reb_check_exit, :290-351: exact_finish_time, the shortened last step, the status codes)
 * finishes the run.  SIGINT is handled as in the reference (:371-372, :346-349): the handler is installed before the
 * batch, batches are cut into pieces, and an interrupt that arrives during a batch ends the run right there with
 * REB_STATUS_SIGINT (the reference's loop would reset reb_sigint on entry, :371, and lose it).
 *
 * Only in automatic residency mode (REBOUND_B200_RESIDENT unset); the explicit modes keep their per-step behaviour.
 */
#include <math.h>
#include <signal.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "shim_common.h"
#include "rebound_internal.h"   /* reb_sigint, reb_sigint_handler */

void reb_simulation_steps_cpuref(struct reb_simulation* const r, size_t N_steps);
enum REB_STATUS reb_simulation_integrate_cpuref(struct reb_simulation* const r, double tmax);

static int batching_disabled(void){
    static int off = -1;
    if (off < 0){ const char* e = getenv("REBOUND_B200_BATCH"); off = (e && e[0]=='0') ? 1 : 0; }
    return off;
}

/* Nothing between two steps can observe or change the simulation, and the step itself runs entirely on the device. */
static int batch_possible(const struct reb_simulation* r){
    if (batching_disabled() || shim_residency()!=SHIM_AUTO) return 0;
    if (!shim_resident(r)) return 0;                               /* heartbeat, timestep hooks, viewer, exit distances with boundary/collisions */
    if (r->heartbeat) return 0;                                    /* called after every step */
    if (r->exit_max_distance || r->exit_min_distance) return 0;    /* checked after every step: keep the per-step path */
    if (r->simulationarchive_filename) return 0;
    if (r->collision!=REB_COLLISION_NONE){
        /* the resolve loop needs the host after every search -- unless it is the built-in hard-sphere resolver, which the
         * engine's exact resolve serves pair batch by pair batch (shim_hotpath.c) */
        struct shim_state* s0 = shim_get((struct reb_simulation*)r);
        if (!s0 || !shim_resolve_on_device(r, s0)) return 0;
        if (r->collision!=REB_COLLISION_DIRECT && r->collision!=REB_COLLISION_TREE && r->collision!=REB_COLLISION_LINE && r->collision!=REB_COLLISION_LINETREE) return 0;
        if (r->map || r->N_targets!=SIZE_MAX) return 0;
    }
    if (r->N_odes || r->N_var || r->additional_forces) return 0;
    if (!shim_is_device_integrator(r)) return 0;
    switch (r->gravity){
        case REB_GRAVITY_NONE: case REB_GRAVITY_BASIC: case REB_GRAVITY_COMPENSATED: case REB_GRAVITY_TREE: break;
        default: return 0;
    }
    if (r->boundary==REB_BOUNDARY_OPEN && (r->track_energy_offset || r->free_particle_ap || r->integrator.callbacks.will_remove_particle)) return 0;
    if (r->N==0 || r->dt==0.) return 0;
    return 1;
}

/* n steps as one device batch on r->particles; the bookkeeping of reb_simulation_step (:514-603) for n steps.
 * Returns 0 on success; on an engine error the message is already queued with reb_simulation_error. */
static int batch_collision_cb(void* user){
    struct reb_simulation* r = user;
    struct shim_state* s = shim_find(r);
    return s ? shim_resolve_pairs(r, s) : -1;
}

static int run_batch(struct reb_simulation* r, size_t n, int pipelined){
    struct shim_state* s = shim_get(r);
    if (!s) return -1;
    if (!shim_prepare_integrator_state(r)) return -2;   /* not a step the device may take: the caller goes step by step */
    if (shim_to_host(r, s)) return -1;                 /* a device copy that is ahead comes home first */
    rebcu_config c;
    shim_fill_config(r, &c);
    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    uint64_t N = r->N;
    int err;
    const int with_collisions = r->collision!=REB_COLLISION_NONE;
    if (with_collisions){ rebcu_set_collision_subset(s->h, NULL, 0, REBCU_SIZE_MAX); s->subset_set = 0; rebcu_set_collision_callback(s->h, batch_collision_cb, r); }
    if (pipelined){
        err = rebcu_steps_host(s->h, &c, (rebcu_particle*)r->particles, &N, n);
    }else{
        err = rebcu_upload(s->h, (const rebcu_particle*)r->particles, r->N);
        if (!err) err = rebcu_steps(s->h, &c, n);
        if (!err){ N = rebcu_N(s->h); err = rebcu_download(s->h, (rebcu_particle*)r->particles, N); }
    }
    if (with_collisions){ rebcu_set_collision_callback(s->h, NULL, NULL); r->N_collisions = 0; }
    int interrupted = 0;
    if (err==REBCU_INTERRUPTED){
        /* a second Ctrl-C stopped the engine between two steps: not an error.  The state of the last completed step
         * comes home (rebcu_steps_host returns before its download) and c.t tells how many steps were done. */
        interrupted = 1;
        N = rebcu_N(s->h);
        err = rebcu_download(s->h, (rebcu_particle*)r->particles, N);
        if (c.dt != 0.){ const double done = floor((c.t - r->t)/c.dt + 0.5); n = (done > 0. && done < (double)n) ? (size_t)done : (done <= 0. ? 0 : n); }
    }
    s->device_valid = 0; s->host_stale = 0;
    if (shim_report(r, s, err)) return -1;
    gettimeofday(&t1, NULL);
    r->t = c.t;
    r->dt_last_done = c.dt_last_done;
    r->gravity_ignore_terms = c.gravity_ignore_terms;
    r->OMEGAZ = c.OMEGAZ;
    r->N_active = (c.N_active==REBCU_SIZE_MAX)?SIZE_MAX:(size_t)c.N_active;
    if (N != r->N){
        if (N==0) reb_simulation_warning(r, "Last particle removed.");      /* particle.c:346 */
        r->N = N;
    }
    r->did_modify_particles = 0;                        /* simulation.c:567 */
    r->is_synchronized = 1;
    /* walltime bookkeeping, simulation.c:588-600 */
    const double el = (double)(t1.tv_sec-t0.tv_sec) + (double)(t1.tv_usec-t0.tv_usec)/1e6;
    r->walltime_last_step = el/(double)(n ? n : 1);
    r->walltime_last_steps_sum += el;
    r->walltime_last_steps_N += n;
    if (r->walltime_last_steps_sum > 0.1){
        r->walltime_last_steps = r->walltime_last_steps_sum/r->walltime_last_steps_N;
        r->walltime_last_steps_sum = 0;
        r->walltime_last_steps_N = 0;
    }
    r->walltime += el;
    r->steps_done += n;                                 /* simulation.c:603 */
    if (interrupted){ if (reb_sigint < 2) reb_sigint = 2; return 1; }
    return 0;
}

/* The device state of a simulation goes with the simulation (reb_simulation_free, src/simulation.c:124-198; the Python
 * package calls it from Simulation.__del__, rebound/simulation.py:154-158): its handle is released first. */
void reb_simulation_free_cpuref(struct reb_simulation* const r);
void reb_simulation_free(struct reb_simulation* const r){
    if (r) shim_forget(r);
    reb_simulation_free_cpuref(r);
}

void reb_simulation_steps(struct reb_simulation* const r, size_t N_steps){
    if (N_steps >= 2 && batch_possible(r)){
        /* run_heartbeat has nothing to do here (no heartbeat, no exit distances); the call ends synchronised.  An
         * engine error is queued with reb_simulation_error; the reference would go on stepping into the same error. */
        const size_t piece = 4096;
        const int one_call = N_steps <= piece;
        size_t left = N_steps;
        while (left){
            const size_t n = left < piece ? left : piece;
            const int rc = run_batch(r, n, one_call);
            if (rc==-2){ reb_simulation_steps_cpuref(r, left); return; }
            if (rc) return;
            left -= n;
            if (left && !batch_possible(r)){ reb_simulation_steps_cpuref(r, left); return; }
        }
        return;
    }
    reb_simulation_steps_cpuref(r, N_steps);
}

enum REB_STATUS reb_simulation_integrate(struct reb_simulation* const r, double tmax){
    if (batch_possible(r) && isfinite(tmax) && tmax != r->t
        && r->status != REB_STATUS_PAUSED && r->status != REB_STATUS_SCREENSHOT){
        const double dt = copysign(r->dt, (tmax > r->t) ? 1.0 : -1.0);      /* simulation.c:377-380 */
        /* whole steps that certainly fit; the last two (and the exit logic) belong to the reference's loop */
        double nf = floor((tmax - r->t)/dt) - 2.;
        if (nf >= 2.){
            r->dt = dt;
            reb_sigint = 0;
            signal(SIGINT, reb_sigint_handler);                             /* simulation.c:371-372 */
            const double piece = 4096.;
            const int one_call = nf <= piece;
            while (nf >= 1. && !reb_sigint){
                const size_t n = (size_t)(nf < piece ? nf : piece);
                if (run_batch(r, n, one_call)) break;
                nf -= (double)n;
                /* t accumulates rounding errors: never run past the point where two whole steps still fit */
                const double left = floor((tmax - r->t)/dt) - 2.;
                if (left < nf) nf = left;
                if (!batch_possible(r)) break;                              /* e.g. the last particle left an open box */
            }
            if (reb_sigint){
                /* what reb_check_exit does with a pending interrupt (simulation.c:346-349) and the end of
                 * reb_simulation_integrate_raw (:455-459): synchronise, report, leave */
                reb_simulation_synchronize(r);
                r->status = REB_STATUS_SIGINT;                  /* reb_sigint stays set, as the reference leaves it */
                return r->status;
            }
        }
    }
    return reb_simulation_integrate_cpuref(r, tmax);
}
