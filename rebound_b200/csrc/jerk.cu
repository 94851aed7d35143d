// jerk.cu -- reb_gravity_basic_calculate_and_apply_jerk (src/gravity.c:850-924): the velocity kick from the
// gradient of the accelerations that the modified-kick splitting schemes apply (EOS, integrator_eos.c:101-103),
// evaluated on the positions and the accelerations of the preceding force evaluation.
//
// The reference walks the pairs (i, j < i) in serial order and SCATTERS into both velocities, so the order of the
// additions is part of the result.  Seen from one particle k that order is simple: first its own row (partners
// p < k, k on the i-side of the pair, ascending p), then -- as i runs on -- the back reaction of every later row
// (p > k, k on the j-side, ascending p): one sum over ascending p.  One thread per particle therefore carries its
// velocity through all sources in index order with strictly rounded arithmetic and reproduces the serial build
// bit for bit (the oracle checks this formulation on the CPU as well, oracle.c: orc_apply_jerk_gather).
// Which pairs exist (gravity.c:857-858, 861-863, 892-896, 912):
//   p < k: (i=k, j=p)  p >= startj, and k >= starti if k is massive (the test-particle rows have no starti)
//   p > k: (i=p, j=k)  k >= startj, and p >= starti if p is massive, testparticle_type != 0 otherwise
// Note that a test-particle row runs over ALL j < i, test particles included (gravity.c:896).
// Bound: FP64 pipe (one sqrt and three divisions per pair); sources staged through shared memory, 56 B each.
#include "engine.cuh"

namespace {

struct JerkArgs {
    const double *x, *y, *z, *ax, *ay, *az, *m;
    double *vx, *vy, *vz;
    uint32_t n, n_active, starti, startj, type;
    uint32_t ib, nloc;            // this rank's block of particles
    double vG2;                   // 2.*v*G (gravity.c:876)
};

constexpr int JT = 128;

__global__ void __launch_bounds__(JT) jerk_kernel(JerkArgs a) {
    __shared__ double sx[JT], sy[JT], sz[JT], sax[JT], say[JT], saz[JT], sm[JT];
    const uint32_t kl = blockIdx.x * JT + threadIdx.x;
    const uint32_t k = a.ib + kl;
    const bool valid = kl < a.nloc;
    double xk = 0, yk = 0, zk = 0, axk = 0, ayk = 0, azk = 0, wx = 0, wy = 0, wz = 0;
    if (valid) {
        xk = a.x[k]; yk = a.y[k]; zk = a.z[k]; axk = a.ax[k]; ayk = a.ay[k]; azk = a.az[k];
        wx = a.vx[k]; wy = a.vy[k]; wz = a.vz[k];
    }
    const bool k_massive = k < a.n_active;
    const bool row_ok = !(k_massive && k < a.starti);      // k may be the i-side of a pair
    const bool col_ok = k >= a.startj;                     // k may be the j-side of a pair
    for (uint32_t t0 = (a.startj / JT) * JT; t0 < a.n; t0 += JT) {
        __syncthreads();
        const uint32_t q0 = t0 + threadIdx.x;
        if (q0 < a.n) {
            sx[threadIdx.x] = a.x[q0]; sy[threadIdx.x] = a.y[q0]; sz[threadIdx.x] = a.z[q0];
            sax[threadIdx.x] = a.ax[q0]; say[threadIdx.x] = a.ay[q0]; saz[threadIdx.x] = a.az[q0];
            sm[threadIdx.x] = a.m[q0];
        }
        __syncthreads();
        if (!valid) continue;
        const int qn = (int)min((uint32_t)JT, a.n - t0);
        for (int qq = 0; qq < qn; qq++) {
            const uint32_t q = t0 + qq;
            if (q < a.startj || q == k) continue;
            const bool k_is_i = q < k;
            if (k_is_i) { if (!row_ok) continue; }
            else { if (!col_ok) continue; if (q < a.n_active ? q < a.starti : a.type == 0) continue; }
            // d = x_i - x_j, da = a_i - a_j with (i, j) = (max, min) of (k, q)   (gravity.c:865-871)
            const double px = sx[qq], py = sy[qq], pz = sz[qq], pax = sax[qq], pay = say[qq], paz = saz[qq];
            const double dx = k_is_i ? s_sub(xk, px) : s_sub(px, xk), dy = k_is_i ? s_sub(yk, py) : s_sub(py, yk),
                         dz = k_is_i ? s_sub(zk, pz) : s_sub(pz, zk);
            const double dax = k_is_i ? s_sub(axk, pax) : s_sub(pax, axk), day = k_is_i ? s_sub(ayk, pay) : s_sub(pay, ayk),
                         daz = k_is_i ? s_sub(azk, paz) : s_sub(paz, azk);
            const double dr = s_sqrt(s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz)));
            const double alphasum = s_add(s_add(s_mul(dax, dx), s_mul(day, dy)), s_mul(daz, dz));
            const double pf2 = s_div(a.vG2, s_mul(s_mul(dr, dr), dr));                         // gravity.c:876
            const double pf1 = s_div(s_mul(s_div(s_mul(alphasum, pf2), dr), 3.), dr);          // gravity.c:879
            const double pf1o = s_mul(pf1, sm[qq]), pf2o = s_mul(pf2, sm[qq]);                 // the OTHER particle's mass
            if (k_is_i) {        // gravity.c:882-884, 909-911
                wx = s_add(wx, s_sub(s_mul(dx, pf1o), s_mul(dax, pf2o)));
                wy = s_add(wy, s_sub(s_mul(dy, pf1o), s_mul(day, pf2o)));
                wz = s_add(wz, s_sub(s_mul(dz, pf1o), s_mul(daz, pf2o)));
            } else {             // gravity.c:885-887, 915-917
                wx = s_add(wx, s_sub(s_mul(dax, pf2o), s_mul(dx, pf1o)));
                wy = s_add(wy, s_sub(s_mul(day, pf2o), s_mul(dy, pf1o)));
                wz = s_add(wz, s_sub(s_mul(daz, pf2o), s_mul(dz, pf1o)));
            }
        }
    }
    if (valid) { a.vx[k] = wx; a.vy[k] = wy; a.vz[k] = wz; }
}

int apply_jerk(rebcu_handle* h, const rebcu_config* c, double v) {
    const uint64_t n = h->N;
    if (n < 2) return REBCU_OK;
    if (n >= (1ull << 31)) return rebcu_fail(h, REBCU_ERR_ARG, "jerk supports N < 2^31");
    // sharded: the kick needs the accelerations of every block, which only their owners hold
    if (h->world > 1) { const int xerr = engine_exchange(h, REBCU_EXCHANGE_ALL); if (xerr) return xerr; }
    uint64_t ib, ie; engine_shard(h, &ib, &ie);
    if (ie == ib) return REBCU_OK;
    JerkArgs a;
    a.x = h->f(F_X); a.y = h->f(F_Y); a.z = h->f(F_Z); a.ax = h->f(F_AX); a.ay = h->f(F_AY); a.az = h->f(F_AZ); a.m = h->f(F_M);
    a.vx = h->f(F_VX); a.vy = h->f(F_VY); a.vz = h->f(F_VZ);
    a.n = (uint32_t)n;
    a.n_active = (uint32_t)((c->N_active == REBCU_SIZE_MAX) ? n : (c->N_active < n ? c->N_active : n));
    a.starti = (c->gravity_ignore_terms == REBCU_IGNORE_TERMS_NONE) ? 1 : 2;             // gravity.c:857
    a.startj = (c->gravity_ignore_terms == REBCU_IGNORE_TERMS_INVOLVING_0) ? 1 : 0;      // gravity.c:858
    a.type = (uint32_t)c->testparticle_type;
    a.ib = (uint32_t)ib; a.nloc = (uint32_t)(ie - ib);
    a.vG2 = 2. * v * c->G;
    {
        LaunchScope ls(h, TC_DIRECT);
        jerk_kernel<<<div_up(a.nloc, JT), JT, 0, h->stream>>>(a);
    }
    CU_TRY(h, cudaGetLastError());
    return REBCU_OK;
}

}  // namespace

extern "C" {

int rebcu_apply_jerk(rebcu_handle* h, const rebcu_config* cfg, double v) {
    if (group_active(h)) return group_run(h, [cfg, v](rebcu_handle* s, int) { return rebcu_apply_jerk(s, cfg, v); });
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    return apply_jerk(h, cfg, v);
}

int rebcu_jerk_host(rebcu_handle* h, const rebcu_config* cfg, rebcu_particle* particles, uint64_t N, double v) {
    int err = rebcu_upload(h, particles, N);
    if (err) return err;
    err = group_active(h) ? rebcu_apply_jerk(h, cfg, v) : apply_jerk(h, cfg, v);
    if (err) return err;
    return rebcu_download(h, particles, N);
}

}  // extern "C"
