/*
 * shim_coherence.c -- keeps the reference's host-side readers coherent with a RESIDENT simulation
 * (REBOUND_B200_RESIDENT=1, r->is_synchronized == 0 between steps): SURVEY.md section 8f-3.
 *
 * Every serialisation path of the reference -- reb_simulation_save_to_file / Simulationarchive snapshots
 * (src/simulationarchive.c:434,488), reb_simulation_copy (src/simulation.c:605-616,709-715),
 * reb_simulation_diff (src/simulation.c:618-705), the server and display copies (src/server.c:325,555) --
 * funnels through reb_binarydata_simulation_to_stream (src/binarydata.c:625).  Wrapping that one function
 * with a lazy device->host copy makes all of them see the current particles; the copy is pure data
 * movement, so a Simulationarchive restart stays bit-exact (test_simulationarchive.py:610-630).
 * The O(N)/O(N^2) diagnostics of src/tools.c that read r->particles get the same treatment.
 * In the default host-authoritative mode r->particles is always current and these wrappers do nothing.
 */
#include "shim_common.h"

void reb_binarydata_simulation_to_stream_cpuref(struct reb_simulation* r, char** bufp, size_t* sizep);
double reb_simulation_energy_cpuref(struct reb_simulation* const r);
struct reb_particle reb_simulation_com_cpuref(struct reb_simulation* r);
struct reb_vec3d reb_simulation_angular_momentum_cpuref(const struct reb_simulation* const r);
void reb_simulation_move_to_com_cpuref(struct reb_simulation* const r);
void reb_simulation_move_to_hel_cpuref(struct reb_simulation* const r);

/* device -> host if the device is ahead; 0 on success */
static int lazy_sync(struct reb_simulation* r){
    struct shim_state* s = shim_find(r);
    if (!s || !s->host_stale) return 0;
    if (shim_to_host(r, s)) return 1;
    r->is_synchronized = 1;
    return 0;
}

void reb_binarydata_simulation_to_stream(struct reb_simulation* r, char** bufp, size_t* sizep){
    lazy_sync(r);
    reb_binarydata_simulation_to_stream_cpuref(r, bufp, sizep);
}

double reb_simulation_energy(struct reb_simulation* const r){
    lazy_sync(r);
    return reb_simulation_energy_cpuref(r);
}

struct reb_particle reb_simulation_com(struct reb_simulation* r){
    lazy_sync(r);
    return reb_simulation_com_cpuref(r);
}

struct reb_vec3d reb_simulation_angular_momentum(const struct reb_simulation* const r){
    lazy_sync((struct reb_simulation*)r);
    return reb_simulation_angular_momentum_cpuref(r);
}

/* These two rewrite every particle on the host: the device copy is stale afterwards. */
static void host_rewrite(struct reb_simulation* r, void (*fn)(struct reb_simulation* const)){
    lazy_sync(r);
    fn(r);
    struct shim_state* s = shim_find(r);
    if (s) s->device_valid = 0;
}
void reb_simulation_move_to_com(struct reb_simulation* const r){ host_rewrite(r, reb_simulation_move_to_com_cpuref); }
void reb_simulation_move_to_hel(struct reb_simulation* const r){ host_rewrite(r, reb_simulation_move_to_hel_cpuref); }
