// api.cu -- the extern "C" hot-path entry points of include/rebound_b200.h that compose the kernels:
// force dispatch, integrator step, whole steps, and the host-buffer drop-ins the shim calls.
#include "engine.cuh"
#include <time.h>

// reb_simulation_update_acceleration, src/simulation.c:640-689 (the GPU-resident gravity modes).
int update_acceleration(rebcu_handle* h, rebcu_config* c) {
    switch (c->gravity) {
        case REBCU_GRAVITY_NONE: return zero_acceleration(h);
        case REBCU_GRAVITY_BASIC:
        case REBCU_GRAVITY_COMPENSATED: return direct_gravity(h, c);
        case REBCU_GRAVITY_TREE: return tree_gravity(h, c);
        default: return rebcu_fail(h, REBCU_ERR_ARG, "Gravity calculation not yet implemented.");
    }
}

static int integrator_step(rebcu_handle* h, rebcu_config* c, bool carry_in, bool carry_out, bool write_acc) {
    switch (c->integrator) {
        case REBCU_INTEGRATOR_LEAPFROG: return leapfrog_step_ex(h, c, carry_in, carry_out, write_acc);
        case REBCU_INTEGRATOR_SEI: return sei_step(h, c);
        case REBCU_INTEGRATOR_NONE: c->t += c->dt; c->dt_last_done = c->dt; return REBCU_OK;   // reb_integrator_none
        default: return rebcu_fail(h, REBCU_ERR_ARG, "Integrator not found.");
    }
}

extern "C" {

int rebcu_update_acceleration(rebcu_handle* h, rebcu_config* cfg) {
    if (group_active(h)) return group_cfg_call(h, cfg, [](rebcu_handle* s, rebcu_config* c) { return rebcu_update_acceleration(s, c); });
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    return update_acceleration(h, cfg);
}

int rebcu_integrator_step(rebcu_handle* h, rebcu_config* cfg) {
    if (group_active(h)) return group_cfg_call(h, cfg, [](rebcu_handle* s, rebcu_config* c) { return rebcu_integrator_step(s, c); });
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    return integrator_step(h, cfg, false, false, true);
}

int rebcu_boundary_check(rebcu_handle* h, rebcu_config* cfg) {
    if (group_active(h)) {
        // the sharded check: every rank must see every particle's position to agree on what left the box
        return group_cfg_call(h, cfg, [](rebcu_handle* s, rebcu_config* c) {
            if (c->boundary == REBCU_BOUNDARY_OPEN) {
                bool any = true;
                int e = boundary_open_probe(s, c, &any);
                if (e || !any) return e;
                if ((e = rebcu_exchange(s, REBCU_EXCHANGE_POSITIONS))) return e;
                return boundary_check_full(s, c);
            }
            return rebcu_boundary_check(s, c);
        });
    }
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    return boundary_check(h, cfg);
}

int rebcu_collisions_fetch(rebcu_handle* h, rebcu_collision* out, uint64_t cap, uint64_t* n_found) {
    if (group_active(h)) return group_collisions_fetch(h, out, cap, n_found);
    *n_found = h->col_n;
    const uint64_t n = h->col_n < cap ? h->col_n : cap;
    if (n && out) {
        CU_TRY(h, cudaMemcpyAsync(out, h->col_list, n * sizeof(rebcu_collision), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
    }
    return REBCU_OK;
}

int rebcu_collision_search(rebcu_handle* h, const rebcu_config* cfg, rebcu_collision* out, uint64_t cap, uint64_t* n_found) {
    if (group_active(h)) {
        const int e = group_run(h, [cfg](rebcu_handle* s, int) { uint64_t n = 0; return rebcu_collision_search(s, cfg, nullptr, 0, &n); });
        if (e) return e;
        return group_collisions_fetch(h, out, cap, n_found);
    }
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    int err = collision_search(h, cfg);
    if (err) return err;
    return rebcu_collisions_fetch(h, out, cap, n_found);
}

// reb_simulation_steps (src/simulation.c:504-513) for a simulation without host callbacks: per step
// integrator (:527-529), boundary check (:575), collision search (:584).  Consecutive leapfrog steps
// share one kick+drift+drift launch when nothing observes the state in between.
int rebcu_steps(rebcu_handle* h, rebcu_config* cfg, uint64_t n_steps) {
    if (group_active(h)) {
        if (h->collision_hook && cfg->collision != REBCU_COLLISION_NONE)
            return rebcu_fail(h, REBCU_ERR_ARG, "a host collision callback inside rebcu_steps is not available on a multi-GPU group handle");
        return group_cfg_call(h, cfg, [n_steps](rebcu_handle* s, rebcu_config* c) { return rebcu_steps(s, c, n_steps); });
    }
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    auto interrupted = [h] { return h->interrupt && *h->interrupt > 1; };
    if (n_steps && interrupted()) return REBCU_INTERRUPTED;
    {
        // few massive bodies + many test particles: the whole batch of steps in two launches
        const int fused = tp_steps_resident(h, cfg, n_steps);
        if (fused <= 0) return fused;
    }
    const bool can_carry = cfg->integrator == REBCU_INTEGRATOR_LEAPFROG && cfg->boundary == REBCU_BOUNDARY_NONE
                        && cfg->collision == REBCU_COLLISION_NONE && h->exchange == nullptr && h->comm == nullptr;
    bool carried = false;
    // REBOUND_B200_STEP_TRACE=1: per step, host wall clock of the phases with the stream drained after each (stderr)
    static const bool trace = getenv("REBOUND_B200_STEP_TRACE") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    auto lap = [&](double& t) { if (!trace) return 0.; cudaStreamSynchronize(h->stream); const double u = now(), d = u - t; t = u; return 1e3 * d; };
    for (uint64_t s = 0; s < n_steps; s++) {
        const bool stop = interrupted();              // finish this step (a carried half-kick must be closed), then leave
        const bool last = (s + 1 == n_steps) || stop;
        const bool carry_out = can_carry && !last;
        double t_lap = trace ? now() : 0.;
        int err = integrator_step(h, cfg, carried, carry_out, last);
        if (err) return err;
        const double ms_step = lap(t_lap);
        carried = carry_out;
        if (h->world > 1 && cfg->boundary == REBCU_BOUNDARY_OPEN) {
            // every rank must see every particle's final position to agree on what left the box -- but only when some
            // block did lose a particle (8 bytes per rank say so)
            bool any = true;
            err = boundary_open_probe(h, cfg, &any);
            if (!err && any) {
                err = engine_exchange(h, REBCU_EXCHANGE_POSITIONS);
                if (!err) err = boundary_check_full(h, cfg);
            }
        } else {
            err = boundary_check(h, cfg);
        }
        if (err) return err;
        const double ms_boundary = lap(t_lap);
        double ms_search = 0., ms_resolve = 0.;
        if (cfg->collision != REBCU_COLLISION_NONE) {
            err = collision_search(h, cfg);
            if (err) return err;
            ms_search = lap(t_lap);
            if (h->resolve_on) {
                err = collision_resolve_device(h, cfg);
                if (err) return err;
            } else if (h->collision_hook) {
                err = h->collision_hook(h->collision_hook_user);
                if (err) return err;
            }
            ms_resolve = lap(t_lap);
        }
        if (trace) fprintf(stderr, "[step %llu] integrator %.2f ms, boundary %.2f, collision search %.2f, resolve %.2f\n",
                           (unsigned long long)s, ms_step, ms_boundary, ms_search, ms_resolve);
        if (stop && s + 1 < n_steps) return REBCU_INTERRUPTED;
    }
    return REBCU_OK;
}

int rebcu_gravity_host(rebcu_handle* h, rebcu_config* cfg, rebcu_particle* particles, uint64_t* N) {
    int err = rebcu_upload(h, particles, *N);
    if (err) return err;
    err = group_active(h) ? rebcu_update_acceleration(h, cfg) : update_acceleration(h, cfg);
    if (err) return err;
    *N = h->N;
    return rebcu_download(h, particles, *N);
}

int rebcu_collision_search_host(rebcu_handle* h, const rebcu_config* cfg, const rebcu_particle* particles, uint64_t N,
                                rebcu_collision* out, uint64_t cap, uint64_t* n_found) {
    int err = rebcu_upload(h, particles, N);
    if (err) return err;
    return rebcu_collision_search(h, cfg, out, cap, n_found);
}

int rebcu_steps_host(rebcu_handle* h, rebcu_config* cfg, rebcu_particle* particles, uint64_t* N, uint64_t n_steps) {
    if (group_active(h)) {
        int err = rebcu_upload(h, particles, *N);
        if (!err) err = rebcu_steps(h, cfg, n_steps);
        if (err) return err;
        *N = h->N;
        return rebcu_download(h, particles, *N);
    }
    CU_TRY(h, cudaSetDevice(h->device));
    // few massive bodies + many test particles: chunked, copies overlapped with the kernels
    const int piped = tp_steps_host_pipelined(h, cfg, particles, *N, n_steps);
    if (piped <= 0) return piped;
    int err = rebcu_upload(h, particles, *N);
    if (err) return err;
    err = rebcu_steps(h, cfg, n_steps);
    if (err) return err;
    *N = h->N;
    return rebcu_download(h, particles, *N);
}

// The exchange on demand (what the engine does by itself between drift and force): after it every rank holds the
// owners' current values of the requested fields.
int rebcu_exchange(rebcu_handle* h, int need) {
    if (group_active(h)) return group_run(h, [need](rebcu_handle* s, int) { return rebcu_exchange(s, need); });
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    if (h->world <= 1) return REBCU_OK;
    return engine_exchange(h, need);
}

// Sharded residency: every rank's host memory holds only its own block of r->particles (as in the reference's MPI
// build, where a rank owns the particles of its root boxes).  Upload: own block over PCIe, everything else over the
// exchange transport (all 14 fields once; afterwards only positions travel per step).  Download: own block only.
int rebcu_upload_shard(rebcu_handle* h, const rebcu_particle* block, uint64_t N_total) {
    CU_TRY(h, cudaSetDevice(h->device));
    int err = engine_reserve(h, N_total);
    if (err) return err;
    h->N = N_total; h->resident = true; h->tree.built_for_n = -1;
    uint64_t b, e; engine_shard(h, &b, &e);
    if (e > b) {
        CU_TRY(h, cudaMemcpyAsync(h->aos + b, block, (e - b) * sizeof(rebcu_particle), cudaMemcpyHostToDevice, h->stream));
        if ((err = engine_upload_range(h, h->stream, nullptr, h->stream, nullptr, b, e))) return err;
    }
    if (h->world > 1) return engine_exchange(h, REBCU_EXCHANGE_ALL);
    return REBCU_OK;
}

int rebcu_download_shard(rebcu_handle* h, rebcu_particle* block, uint64_t cap) {
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    uint64_t b, e; engine_shard(h, &b, &e);
    if (cap < e - b) return rebcu_fail(h, REBCU_ERR_CAPACITY, "host block buffer too small");
    if (e > b) {
        int err = engine_download_range(h, h->stream, nullptr, h->stream, nullptr, b, e);
        if (err) return err;
        CU_TRY(h, cudaMemcpyAsync(block, h->aos + b, (e - b) * sizeof(rebcu_particle), cudaMemcpyDeviceToHost, h->stream));
    }
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return REBCU_OK;
}

int rebcu_set_exchange_callback(rebcu_handle* h, void (*cb)(void*), void* user) {
    h->exchange = cb; h->exchange_user = user;
    return REBCU_OK;
}

int rebcu_set_interrupt_flag(rebcu_handle* h, const volatile int* flag) {
    if (group_active(h)) return group_run(h, [flag](rebcu_handle* s, int) { return rebcu_set_interrupt_flag(s, flag); });
    h->interrupt = flag;
    return REBCU_OK;
}

int rebcu_set_collision_callback(rebcu_handle* h, int (*cb)(void*), void* user) {
    h->collision_hook = cb; h->collision_hook_user = user;
    return REBCU_OK;
}

int rebcu_tree_build(rebcu_handle* h, const rebcu_config* cfg) {
    GROUP_UNSUPPORTED(h, "rebcu_tree_build");
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    return tree_build(h, cfg);
}

uint64_t rebcu_tree_cell_count(const rebcu_handle* h) { return h->tree.n_cells; }

int rebcu_tree_fetch(rebcu_handle* h, rebcu_treecell* out, uint64_t cap) {
    if (cap < h->tree.n_cells) return rebcu_fail(h, REBCU_ERR_CAPACITY, "tree cell buffer too small");
    if (h->tree.n_cells) {
        int err = tree_export(h);
        if (err) return err;
        CU_TRY(h, cudaMemcpyAsync(out, h->tree.cells, h->tree.n_cells * sizeof(rebcu_treecell), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
    }
    return REBCU_OK;
}

}  // extern "C"
