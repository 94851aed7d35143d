// boundary.cu -- reb_boundary_check (src/boundary.c:35-141) on the resident SoA.
//
// SHEAR / PERIODIC: one streaming pass that wraps positions (and shifts y, vy across the radial
// boundary); the time-dependent offsets are computed on the host with libm fmod exactly as
// boundary.c:82-83 and passed by value.
// OPEN: particles outside the box are removed with an order-preserving compaction of all 14 SoA
// arrays (what the reference's repeated reb_simulation_remove_particle shifts amount to,
// particle.c:360-369), N_active is decremented once per removed active particle.
// Bound: HBM (48 B/particle for the wrap; the compaction only runs when something left the box).
#include "engine.cuh"
#include "primitives.cuh"
#include <math.h>

namespace {

struct WrapArgs {
    double *x, *y, *z, *vy;
    double bx, by, bz;
    double offp1, offm1, dvy;
    int shear;
};

__global__ void __launch_bounds__(256) wrap_kernel(WrapArgs a, uint64_t b, uint64_t e) {
    const uint64_t i = b + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= e) return;
    double x = a.x[i], y = a.y[i], z = a.z[i];
    const double hx = s_div(a.bx, 2.), hy = s_div(a.by, 2.), hz = s_div(a.bz, 2.);
    if (a.shear) {
        double vy = a.vy[i];
        bool touched = false;
        while (x > hx) { x = s_sub(x, a.bx); y = s_add(y, a.offp1); vy = s_add(vy, a.dvy); touched = true; }
        while (x < -hx) { x = s_add(x, a.bx); y = s_add(y, a.offm1); vy = s_sub(vy, a.dvy); touched = true; }
        if (touched) a.vy[i] = vy;
    } else {
        while (x > hx) x = s_sub(x, a.bx);
        while (x < -hx) x = s_add(x, a.bx);
    }
    while (y > hy) y = s_sub(y, a.by);
    while (y < -hy) y = s_add(y, a.by);
    while (z > hz) z = s_sub(z, a.bz);
    while (z < -hz) z = s_add(z, a.bz);
    a.x[i] = x; a.y[i] = y; a.z[i] = z;
}

// keep[i] = 1 if particle i stays (boundary.c:47-64); counters[0] += removed, counters[1] += removed actives
__global__ void __launch_bounds__(256) open_flag_kernel(const double* x, const double* y, const double* z, uint64_t n, uint64_t Na,
                                                        double hx, double hy, double hz, uint32_t* keep, unsigned long long* counters) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    bool out = false;
    if (i < n) {
        const double px = x[i], py = y[i], pz = z[i];
        out = px > hx || px < -hx || py > hy || py < -hy || pz > hz || pz < -hz;
        keep[i] = out ? 0u : 1u;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, out);
    const unsigned int ma = __ballot_sync(0xffffffffu, out && i < Na);
    if ((threadIdx.x & 31) == 0 && m) {
        atomicAdd(&counters[0], (unsigned long long)__popc(m));
        if (ma) atomicAdd(&counters[1], (unsigned long long)__popc(ma));
    }
}

__global__ void __launch_bounds__(256) compact_kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t cap,
                                                      const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const uint64_t d = pos[i];
#pragma unroll
    for (int k = 0; k < F_COUNT; k++) dst[(uint64_t)k * cap + d] = src[(uint64_t)k * cap + i];
}

}  // namespace

int boundary_check(rebcu_handle* h, rebcu_config* c) {
    if (c->boundary == REBCU_BOUNDARY_NONE || h->N == 0) return REBCU_OK;
    const double bx = c->root_size * (double)c->N_root_x;
    const double by = c->root_size * (double)c->N_root_y;
    const double bz = c->root_size * (double)c->N_root_z;
    const uint64_t N = h->N;
    if (c->boundary == REBCU_BOUNDARY_SHEAR || c->boundary == REBCU_BOUNDARY_PERIODIC) {
        WrapArgs a;
        a.x = h->f(F_X); a.y = h->f(F_Y); a.z = h->f(F_Z); a.vy = h->f(F_VY);
        a.bx = bx; a.by = by; a.bz = bz;
        a.shear = c->boundary == REBCU_BOUNDARY_SHEAR;
        a.offp1 = a.offm1 = a.dvy = 0;
        if (a.shear) {
            const double OMEGA = c->OMEGA;
            a.offp1 = -fmod(-1.5 * OMEGA * bx * c->t + by / 2., by) - by / 2.;   // boundary.c:82
            a.offm1 = -fmod(1.5 * OMEGA * bx * c->t - by / 2., by) + by / 2.;    // boundary.c:83
            a.dvy = 3. / 2. * OMEGA * bx;                                        // boundary.c:91
        }
        uint64_t b, e; engine_shard(h, &b, &e);
        if (e > b) {
            LaunchScope ls(h, TC_BOUNDARY);
            wrap_kernel<<<div_up(e - b, 256), 256, 0, h->stream>>>(a, b, e);
        }
        CU_TRY(h, cudaGetLastError());
        return REBCU_OK;
    }
    // OPEN
    if (h->compact_cap < h->cap) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->compact_flag); cudaFree(h->compact_pos); cudaFree(h->compact_buf); cudaFree(h->compact_tmp);
        h->compact_flag = h->compact_pos = nullptr; h->compact_buf = nullptr; h->compact_tmp = nullptr; h->compact_cap = 0;
        CU_TRY(h, cudaMalloc(&h->compact_flag, h->cap * sizeof(uint32_t)));
        CU_TRY(h, cudaMalloc(&h->compact_pos, h->cap * sizeof(uint32_t)));
        const size_t tmp = prim::scan_scratch_words(h->cap) * sizeof(uint32_t);
        h->compact_tmp_bytes = tmp;
        CU_TRY(h, cudaMalloc(&h->compact_tmp, tmp));
        h->compact_cap = h->cap;
    }
    const uint64_t Na = (c->N_active == REBCU_SIZE_MAX) ? 0 : c->N_active;
    CU_TRY(h, cudaMemsetAsync(h->counters, 0, 2 * sizeof(unsigned long long), h->stream));
    {
        LaunchScope ls(h, TC_BOUNDARY);
        open_flag_kernel<<<div_up(N, 256), 256, 0, h->stream>>>(h->f(F_X), h->f(F_Y), h->f(F_Z), N, Na, bx / 2., by / 2., bz / 2.,
                                                              h->compact_flag, h->counters);
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(h->pinned, h->counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const uint64_t removed = h->pinned[0];
    uint64_t removed_active = h->pinned[1];
    if (removed == 0) return REBCU_OK;
    if (h->full_check_world > 1) {
        // Sharded: every rank found the same particles outside (positions are replicated), but it only owns the
        // velocities / accelerations / tags of its own block, and the compaction moves particles across block
        // borders.  Gather every field from its owner first; all ranks then compact identical arrays.
        if (!h->exchange && !h->comm) return rebcu_fail(h, REBCU_ERR_ARG, "open-boundary removal while sharded over several GPUs needs the exchange callback");
        h->rank = h->full_check_rank; h->world = h->full_check_world;
        const int xerr = engine_exchange(h, REBCU_EXCHANGE_ALL);
        h->rank = 0; h->world = 1;
        if (xerr) return xerr;
    }
    // order-preserving compaction into a second SoA block, then swap
    if (!h->compact_buf) CU_TRY(h, cudaMalloc(&h->compact_buf, h->cap * F_COUNT * sizeof(double)));
    {
        LaunchScope ls(h, TC_BOUNDARY, 2);
        prim::exclusive_scan_u32(h->stream, h->compact_flag, h->compact_pos, N, (uint32_t*)h->compact_tmp);
        compact_kernel<<<div_up(N, 256), 256, 0, h->stream>>>((const uint64_t*)h->soa, (uint64_t*)h->compact_buf, h->cap,
                                                            h->compact_flag, h->compact_pos, N);
    }
    CU_TRY(h, cudaGetLastError());
    double* t = h->soa; h->soa = h->compact_buf; h->compact_buf = t;
    h->N = N - removed;
    h->tree.built_for_n = -1;
    if (c->N_active != REBCU_SIZE_MAX) {
        // particle.c:336-343: removing the very last particle does not touch N_active
        if (h->N == 0 && N - 1 < c->N_active && removed_active > 0) removed_active--;
        c->N_active -= removed_active;
    }
    return REBCU_OK;
}

// Sharded runs, open boundary, after a step: did ANY rank's block lose a particle?  Every rank flags its own block
// (whose positions it holds) and the ranks gather the W counts -- 8 bytes each instead of all positions.  Only when the
// answer is yes does the caller exchange positions and run the full check.
int boundary_open_probe(rebcu_handle* h, const rebcu_config* c, bool* any) {
    *any = true;
    if (!h->comm || h->world <= 1) return REBCU_OK;            // no native transport: the caller takes the full path
    const uint64_t N = h->N;
    *any = false;
    if (N == 0) return REBCU_OK;
    if (h->compact_cap < h->cap) { *any = true; return REBCU_OK; }      // buffers of the full check not there yet: let it allocate them
    const double bx = c->root_size * (double)c->N_root_x, by = c->root_size * (double)c->N_root_y, bz = c->root_size * (double)c->N_root_z;
    uint64_t b, e; engine_shard(h, &b, &e);
    unsigned long long* tab = h->counters + 16;                 // REBCU_MAX_RANKS words behind the 16 counters
    CU_TRY(h, cudaMemsetAsync(h->counters, 0, 2 * sizeof(unsigned long long), h->stream));
    if (e > b) {
        LaunchScope ls(h, TC_BOUNDARY);
        open_flag_kernel<<<div_up(e - b, 256), 256, 0, h->stream>>>(h->f(F_X) + b, h->f(F_Y) + b, h->f(F_Z) + b, e - b, 0, bx / 2., by / 2., bz / 2.,
                                                                  h->compact_flag + b, h->counters);
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(tab + h->rank, h->counters, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, h->stream));
    uint64_t bounds[REBCU_MAX_RANKS + 1];
    for (int r = 0; r <= h->world; r++) bounds[r] = (uint64_t)r;
    void* ptr = tab; int bytes = 8;
    int err = comm_gather_ranges(h, &ptr, &bytes, 1, bounds);
    if (err) return err;
    CU_TRY(h, cudaMemcpyAsync(h->pinned, tab, (size_t)h->world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    for (int r = 0; r < h->world; r++) if (h->pinned[r]) *any = true;
    return REBCU_OK;
}

// Sharded runs: every rank checks EVERY particle (the positions are replicated after the exchange), so that
// wraps and removals are applied identically everywhere.
int boundary_check_full(rebcu_handle* h, rebcu_config* c) {
    const int rank = h->rank, world = h->world;
    h->full_check_rank = rank; h->full_check_world = world;
    h->rank = 0; h->world = 1;
    const int err = boundary_check(h, c);
    h->rank = rank; h->world = world;
    h->full_check_world = 0;
    return err;
}
