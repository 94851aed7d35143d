/*
 * shim_common.h -- glue shared by the drop-in translation units.
 *
 * The shim files are compiled against the reference's OWN headers (-I<reference>/src, nothing is
 * copied) and linked with the reference's unmodified sources into a librebound whose hot-path
 * symbols forward to the CUDA engine (include/rebound_b200.h).  See INTEGRATION.md.
 *
 * Device state cannot live in struct reb_simulation (its size is frozen: the Python package checks
 * sizeof at import, rebound/simulation.py:1478-1482), so it lives in a side table keyed by the
 * simulation pointer.
 */
#ifndef REBOUND_B200_SHIM_COMMON_H
#define REBOUND_B200_SHIM_COMMON_H

#include "rebound.h"
#include "../../include/rebound_b200.h"

/* shim-internal symbols are not exported (the reference's CI requires every exported symbol of
 * librebound to start with reb_, .github/workflows/c.yml:17-22) */
#pragma GCC visibility push(hidden)

struct shim_state {
    struct reb_simulation* r;
    rebcu_handle* h;
    int device_valid;            /* the device SoA holds the current particle state */
    int host_stale;              /* r->particles is behind the device copy (resident mode, is_synchronized==0) */
    struct reb_particle* uploaded_from;
    size_t uploaded_N;
    void* pinned_ptr;            /* r->particles as registered with cudaHostRegister (REBOUND_B200_PIN=1) */
    size_t pinned_bytes;
    int subset_set;              /* the engine holds a collision subset (r->map / r->N_targets) from the previous search */
    int lazy;                    /* r->particles is page-protected while the device copy is ahead (shim_lazy.c) */
    unsigned long lazy_faults;   /* downloads triggered by somebody touching the protected array */
};

/* Returns the per-simulation state (creating the rebcu handle on first use); NULL + reb_simulation_error
 * if no CUDA device is usable -- there is no CPU fallback for the replaced functions. */
struct shim_state* shim_get(struct reb_simulation* r);
/* The state if this simulation has one, else NULL (never creates a handle, never raises). */
struct shim_state* shim_find(struct reb_simulation* r);
void shim_forget(struct reb_simulation* r);

/* Residency of the particle state between the steps of one reb_simulation_steps / _integrate call:
 *   REBOUND_B200_RESIDENT=0  host-authoritative: every replaced call uploads r->particles and writes the result back
 *   REBOUND_B200_RESIDENT=1  resident: particles stay in HBM, r->is_synchronized = 0 until reb_simulation_synchronize
 *                            (the protocol WHFast uses, src/simulation.c:633-637); the caller flags host edits with
 *                            r->did_modify_particles
 *   unset (default)          automatic: resident inside a call as long as no callback or exit check can observe
 *                            r->particles between its steps; the device copy is dropped at the synchronize that ends the
 *                            call, so host edits between calls need no flag */
enum { SHIM_HOST_AUTHORITATIVE = 0, SHIM_RESIDENT = 1, SHIM_AUTO = 2 };
int shim_residency(void);
int shim_resident(const struct reb_simulation* r);

/* The simulation's integrator is one of the two this library provides (leapfrog, sei) and its step would run on the device. */
int shim_is_device_integrator(const struct reb_simulation* r);
/* Brings the integrator's own state (the SEI cache) to where the reference's step would leave it; 0 = this step has to
 * run through the reference's host step (a stale cache the reference would keep using). */
int shim_prepare_integrator_state(struct reb_simulation* r);

void shim_fill_config(const struct reb_simulation* r, rebcu_config* c);
/* Forwards a rebcu error to reb_simulation_error (src/simulation.c:82-86); returns err. */
int shim_report(struct reb_simulation* r, struct shim_state* s, int err);

/* Make the device copy current: uploads r->particles unless the device already holds this state
 * (same array, same N, r->did_modify_particles not set -- the flag WHFast relies on, rebound.h:247). */
int shim_to_device(struct reb_simulation* r, struct shim_state* s);
/* Make r->particles current: downloads if the device is ahead. */
int shim_to_host(struct reb_simulation* r, struct shim_state* s);

/* Collision resolve without the particles on the host (shim_hotpath.c). */
int shim_resolve_on_device(const struct reb_simulation* r, const struct shim_state* s);
int shim_resolve_pairs(struct reb_simulation* r, struct shim_state* s);

/* Lazy host copy under a heartbeat (shim_lazy.c). */
int shim_lazy_possible(const struct reb_simulation* r);
void shim_lazy_protect(struct reb_simulation* r, struct shim_state* s);
void shim_lazy_open(struct shim_state* s, int keep_device);
void shim_lazy_forget(struct shim_state* s);

#pragma GCC visibility pop
#endif
