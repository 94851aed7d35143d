// integrate.cu -- leapfrog kick/drift and SEI operators on the resident SoA.
//
// Replaces drift / kick / reb_integrator_leapfrog_step (src/integrator_leapfrog.c:72-93, 97-209) and
// operator_H012 / operator_phi1 / reb_integrator_sei_step (src/integrator_sei.c:126-174, 86-117).
//
// All arithmetic is strict (separately rounded multiply and add, as the reference's -std=c99 build),
// so trajectories are bit-identical as long as the accelerations are.
// Bound: HBM.  Algorithmic bytes per particle: drift 72 (x,v in; x out), kick+drift 120
// (x,v,a in; x,v out), fused test-particle step 96 (x,v in; x,v out; accelerations stay in registers).
#include "engine.cuh"
#include "strict_math.cuh"
#include "fast_math.cuh"
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

namespace {

// Composition coefficients, integrator_leapfrog.c:68-70.
const double LF4 = 0.675603595979828817023843904485;
const double LF6[5] = {0.1867, 0.5554970237124784, 0.1294669489134754, -0.843265623387734, 0.9432033015235604};
const double LF8[9] = {0.128865979381443, 0.581514087105251, -0.410175371469850, 0.1851469357165877, -0.4095523434208514,
                       0.1444059410800120, 0.2783355003936797, 0.3149566839162949, -0.6269948254051343979};

// Drift / kick coefficients of one step, each product formed as the reference writes it
// (integrator_leapfrog.c:102-203).  Returns the number of kicks (drifts = kicks + 1) or -1.
int lf_schedule(int order, double dt, double* drift, double* kick) {
    if (order == 2) { drift[0] = dt * 0.5; kick[0] = dt; drift[1] = dt * 0.5; return 1; }
    if (order == 4) {
        drift[0] = dt * LF4;          kick[0] = dt * 2. * LF4;
        drift[1] = dt * (0.5 - LF4);  kick[1] = dt * (1. - 4. * LF4);
        drift[2] = dt * (0.5 - LF4);  kick[2] = dt * 2. * LF4;
        drift[3] = dt * LF4;
        return 3;
    }
    if (order == 6 || order == 8) {
        const double* a = (order == 6) ? LF6 : LF8;
        const int s = (order == 6) ? 5 : 9;
        const int nk = 2 * s - 1;
        for (int k = 0; k < nk; k++) kick[k] = dt * a[(k < s) ? k : (2 * s - 2 - k)];
        drift[0] = dt * a[0] * 0.5;
        for (int k = 1; k < nk; k++) { const int lo = (k < s) ? (k - 1) : (2 * s - 2 - k); drift[k] = dt * (a[lo] + a[lo + 1]) * 0.5; }
        drift[nk] = dt * a[0] * 0.5;
        return nk;
    }
    return -1;
}

struct Soa {
    double *x, *y, *z, *vx, *vy, *vz, *ax, *ay, *az, *m;
};
Soa soa_of(const rebcu_handle* h) {
    return Soa{h->f(F_X), h->f(F_Y), h->f(F_Z), h->f(F_VX), h->f(F_VY), h->f(F_VZ), h->f(F_AX), h->f(F_AY), h->f(F_AZ), h->f(F_M)};
}

// x += d*v   (integrator_leapfrog.c:77-79)
__global__ void __launch_bounds__(256) drift_kernel(Soa s, double d, uint64_t b, uint64_t e) {
    const uint64_t i = b + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= e) return;
    s.x[i] = s_add(s.x[i], s_mul(d, s.vx[i]));
    s.y[i] = s_add(s.y[i], s_mul(d, s.vy[i]));
    s.z[i] = s_add(s.z[i], s_mul(d, s.vz[i]));
}

// v += k*a ; x += d1*v ; [x += d2*v]   (kick :89-91 followed by one or two drifts)
__global__ void __launch_bounds__(256) kick_drift_kernel(Soa s, double k, double d1, int has_d2, double d2, uint64_t b, uint64_t e) {
    const uint64_t i = b + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= e) return;
    const double vx = s_add(s.vx[i], s_mul(k, s.ax[i]));
    const double vy = s_add(s.vy[i], s_mul(k, s.ay[i]));
    const double vz = s_add(s.vz[i], s_mul(k, s.az[i]));
    double x = s_add(s.x[i], s_mul(d1, vx));
    double y = s_add(s.y[i], s_mul(d1, vy));
    double z = s_add(s.z[i], s_mul(d1, vz));
    if (has_d2) { x = s_add(x, s_mul(d2, vx)); y = s_add(y, s_mul(d2, vy)); z = s_add(z, s_mul(d2, vz)); }
    s.vx[i] = vx; s.vy[i] = vy; s.vz[i] = vz;
    s.x[i] = x; s.y[i] = y; s.z[i] = z;
}

// ---- fused leapfrog step for "few massive bodies + many test particles" (testparticle_type 0) ----
// Every CTA keeps the N_active massive bodies in shared memory and advances them itself (the same
// strictly rounded operations in every CTA, so all copies agree bit for bit); each thread then does
// drift -> force from the massive bodies -> kick -> drift for its own particle with the
// accelerations never leaving registers.  One launch = one whole leapfrog step, 96 B/particle.
constexpr int TP_MAX_ACTIVE = 256;
constexpr int TP_BLOCK = 128;

struct TpArgs {
    Soa s;
    const double* act_in;   // snapshot of the massive bodies before the step: 7 arrays of Na (x y z vx vy vz m)
    double* act_out;        // snapshot after the step
    uint64_t N; int Na;
    uint64_t i_begin;       // first particle of this launch (chunk-pipelined host path), normally 0; N = end
    double d0; int has_d0;  // leading drift (absent if the previous launch already applied it)
    double k, d1, d2; int has_d2;
    double G, soft2;
    double gbx, gby, gbz;   // offset of ghost box (0,0,0), added as the reference does (gravity.c:222-224)
    int kahan, fast, write_acc;
    int windowed;           // G lies inside the window of the branch-free sqrt/divide (strict_math.cuh)
};

// Force of the massive bodies (shared memory) on one particle, strict arithmetic.
//   WINDOWED: branch-free sqrt/divide (strict_math.cuh); returns the running window key, the caller falls
//             back to the generic variant (WINDOWED=false) if it reaches STRICT_WINDOW_LIMIT.
//   SELF:     the particle may be one of the sources (first warp only) and must skip itself.
template <bool KAHAN, bool WINDOWED, bool SELF>
__device__ __forceinline__ unsigned tp_force_strict(const double4* src, int Na, int self, double xi, double yi, double zi,
                                                    double G, double soft2, double& ax, double& ay, double& az) {
    double cx = 0, cy = 0, cz = 0;
    ax = ay = az = 0;
    const double negG = -G;
    unsigned wmax = 0;
#pragma unroll 2
    for (int j = 0; j < Na; j++) {
        const double4 sj = src[j];
        const double dx = s_sub(xi, sj.x), dy = s_sub(yi, sj.y), dz = s_sub(zi, sj.z);
        const double r2 = s_add(s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz)), soft2);
        const double r = WINDOWED ? fsqrt_rn_w(r2) : s_sqrt(r2);
        double p;
        if (!KAHAN) {
            const double b = s_mul(s_mul(r, r), r);                       // gravity.c:226
            p = s_mul(WINDOWED ? fdiv_rn_w(negG, b) : s_div(negG, b), sj.w);
        } else {
            const double b = s_mul(r2, r);                                // gravity.c:320-321
            p = s_mul(-(WINDOWED ? fdiv_rn_w(G, b) : s_div(G, b)), sj.w);
        }
        if (SELF && j == self) continue;
        if (WINDOWED) wmax = max(wmax, strict_window_key(r2));
        if (!KAHAN) {
            ax = s_add(ax, s_mul(p, dx)); ay = s_add(ay, s_mul(p, dy)); az = s_add(az, s_mul(p, dz));
        } else {
            double y, t;
            y = s_sub(s_mul(p, dx), cx); t = s_add(ax, y); cx = s_sub(s_sub(t, ax), y); ax = t;
            y = s_sub(s_mul(p, dy), cy); t = s_add(ay, y); cy = s_sub(s_sub(t, ay), y); ay = t;
            y = s_sub(s_mul(p, dz), cz); t = s_add(az, y); cz = s_sub(s_sub(t, az), y); az = t;
        }
    }
    return wmax;
}

// premul: the source records already hold -G*m in .w (tp_multistep_kernel writes them that way once per step and CTA,
// which takes the multiplication out of the pair term: the same product, hence the same bits)
__device__ __forceinline__ void tp_force_fast(const double4* src, int Na, int self, double xi, double yi, double zi,
                                              double G, double soft2, bool kahan, double& ax, double& ay, double& az, bool premul = false) {
    double cx = 0, cy = 0, cz = 0;
    ax = ay = az = 0;
    const double negG = premul ? 1.0 : -G;
    for (int j = 0; j < Na; j++) {
        if (j == self) continue;
        const double4 sj = src[j];
        const double dx = xi - sj.x, dy = yi - sj.y, dz = zi - sj.z;
        const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, soft2)));
        const double p = fast_m_over_r3(r2, premul ? sj.w : negG * sj.w);
        if (!kahan) { ax = fma(p, dx, ax); ay = fma(p, dy, ay); az = fma(p, dz, az); }
        else {
            double y, t;
            y = fma(p, dx, -cx); t = ax + y; cx = (t - ax) - y; ax = t;
            y = fma(p, dy, -cy); t = ay + y; cy = (t - ay) - y; ay = t;
            y = fma(p, dz, -cz); t = az + y; cz = (t - az) - y; az = t;
        }
    }
}

template <bool FAST, bool KAHAN>
__global__ void __launch_bounds__(TP_BLOCK) tp_leapfrog_kernel(const TpArgs a) {
    __shared__ double4 src[TP_MAX_ACTIVE];
    const int Na = a.Na;
    for (int j = threadIdx.x; j < Na; j += TP_BLOCK) {
        double x = a.act_in[0 * Na + j], y = a.act_in[1 * Na + j], z = a.act_in[2 * Na + j];
        if (a.has_d0) {
            x = s_add(x, s_mul(a.d0, a.act_in[3 * Na + j]));
            y = s_add(y, s_mul(a.d0, a.act_in[4 * Na + j]));
            z = s_add(z, s_mul(a.d0, a.act_in[5 * Na + j]));
        }
        src[j] = make_double4(x, y, z, a.act_in[6 * Na + j]);
    }
    __syncthreads();
    const uint64_t i = a.i_begin + (uint64_t)blockIdx.x * TP_BLOCK + threadIdx.x;
    if (i >= a.N) return;
    double x = a.s.x[i], y = a.s.y[i], z = a.s.z[i];
    double vx = a.s.vx[i], vy = a.s.vy[i], vz = a.s.vz[i];
    if (a.has_d0) { x = s_add(x, s_mul(a.d0, vx)); y = s_add(y, s_mul(a.d0, vy)); z = s_add(z, s_mul(a.d0, vz)); }
    double ax, ay, az;
    {
        double xi = x, yi = y, zi = z;
        if (!KAHAN) { xi = s_add(a.gbx, x); yi = s_add(a.gby, y); zi = s_add(a.gbz, z); }
        const int self = (i < (uint64_t)Na) ? (int)i : -1;
        const bool warp_has_self = (i - (threadIdx.x & 31)) < (uint64_t)Na;        // warp-uniform
        if (FAST) {
            tp_force_fast(src, Na, self, xi, yi, zi, a.G, a.soft2, KAHAN, ax, ay, az);
        } else {
            unsigned w = STRICT_WINDOW_LIMIT;
            if (a.windowed) {
                if (warp_has_self) w = tp_force_strict<KAHAN, true, true>(src, Na, self, xi, yi, zi, a.G, a.soft2, ax, ay, az);
                else w = tp_force_strict<KAHAN, true, false>(src, Na, self, xi, yi, zi, a.G, a.soft2, ax, ay, az);
            }
            if (w >= STRICT_WINDOW_LIMIT) tp_force_strict<KAHAN, false, true>(src, Na, self, xi, yi, zi, a.G, a.soft2, ax, ay, az);
        }
    }
    vx = s_add(vx, s_mul(a.k, ax)); vy = s_add(vy, s_mul(a.k, ay)); vz = s_add(vz, s_mul(a.k, az));
    x = s_add(x, s_mul(a.d1, vx)); y = s_add(y, s_mul(a.d1, vy)); z = s_add(z, s_mul(a.d1, vz));
    if (a.has_d2) { x = s_add(x, s_mul(a.d2, vx)); y = s_add(y, s_mul(a.d2, vy)); z = s_add(z, s_mul(a.d2, vz)); }
    a.s.x[i] = x; a.s.y[i] = y; a.s.z[i] = z;
    a.s.vx[i] = vx; a.s.vy[i] = vy; a.s.vz[i] = vz;
    if (a.write_acc) { a.s.ax[i] = ax; a.s.ay[i] = ay; a.s.az[i] = az; }
    if (i < (uint64_t)Na) {
        a.act_out[0 * Na + i] = x; a.act_out[1 * Na + i] = y; a.act_out[2 * Na + i] = z;
        a.act_out[3 * Na + i] = vx; a.act_out[4 * Na + i] = vy; a.act_out[5 * Na + i] = vz;
        a.act_out[6 * Na + i] = a.s.m[i];
    }
}

// ---- many steps per launch --------------------------------------------------------------------
// With testparticle_type 0 the massive bodies do not feel the test particles, so their whole trajectory over
// n_steps can be computed first by ONE small CTA (tp_history_kernel) and recorded step by step (hist[st] = the
// state the step kernel would receive as act_in: before the leading drift for st = 0, after the merged trailing
// drift otherwise).  tp_multistep_kernel then advances every test particle through all n_steps in a single
// launch with x, v held in registers: per particle 48 B are read and 72 B written once per call instead of
// per step, and the launch count is independent of n_steps.  Every particle executes exactly the operation
// sequence of tp_leapfrog_kernel, so the results are bit-identical.
struct TpMultiArgs {
    Soa s;
    double* hist;           // (n_steps+1) x 7 x Na
    uint64_t i_begin, i_end; int Na;
    uint64_t n_steps;
    double d0, k, d1, d2;
    double G, soft2, gbx, gby, gbz;
    int windowed, fast, kahan;
};

template <bool FAST, bool KAHAN>
__device__ __forceinline__ void tp_force_any(const double4* src, int Na, int self, bool warp_has_self, double x, double y, double z,
                                             const TpMultiArgs& a, double& ax, double& ay, double& az, bool premul = false) {
    double xi = x, yi = y, zi = z;
    if (!KAHAN) { xi = s_add(a.gbx, x); yi = s_add(a.gby, y); zi = s_add(a.gbz, z); }
    if (FAST) { tp_force_fast(src, Na, self, xi, yi, zi, a.G, a.soft2, KAHAN, ax, ay, az, premul); return; }
    unsigned w = STRICT_WINDOW_LIMIT;
    if (a.windowed) {
        if (warp_has_self) w = tp_force_strict<KAHAN, true, true>(src, Na, self, xi, yi, zi, a.G, a.soft2, ax, ay, az);
        else w = tp_force_strict<KAHAN, true, false>(src, Na, self, xi, yi, zi, a.G, a.soft2, ax, ay, az);
    }
    if (w >= STRICT_WINDOW_LIMIT) tp_force_strict<KAHAN, false, true>(src, Na, self, xi, yi, zi, a.G, a.soft2, ax, ay, az);
}

// One CTA: the massive bodies alone, all steps; records hist[0..n_steps] and writes their final state.
template <bool FAST, bool KAHAN>
__global__ void __launch_bounds__(TP_MAX_ACTIVE) tp_history_kernel(const TpMultiArgs a) {
    __shared__ double4 src[TP_MAX_ACTIVE];
    const int j = threadIdx.x, Na = a.Na;
    const bool live = j < Na;
    double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, m = 0, ax = 0, ay = 0, az = 0;
    if (live) { x = a.s.x[j]; y = a.s.y[j]; z = a.s.z[j]; vx = a.s.vx[j]; vy = a.s.vy[j]; vz = a.s.vz[j]; m = a.s.m[j]; }
    for (uint64_t st = 0; st < a.n_steps; st++) {
        if (live) {
            double* hs = a.hist + st * 7 * Na;
            hs[0 * Na + j] = x; hs[1 * Na + j] = y; hs[2 * Na + j] = z; hs[3 * Na + j] = vx; hs[4 * Na + j] = vy; hs[5 * Na + j] = vz; hs[6 * Na + j] = m;
            if (st == 0) { x = s_add(x, s_mul(a.d0, vx)); y = s_add(y, s_mul(a.d0, vy)); z = s_add(z, s_mul(a.d0, vz)); }
            src[j] = make_double4(x, y, z, m);
        }
        __syncthreads();
        if (live) {
            tp_force_any<FAST, KAHAN>(src, Na, j, true, x, y, z, a, ax, ay, az);
            vx = s_add(vx, s_mul(a.k, ax)); vy = s_add(vy, s_mul(a.k, ay)); vz = s_add(vz, s_mul(a.k, az));
            x = s_add(x, s_mul(a.d1, vx)); y = s_add(y, s_mul(a.d1, vy)); z = s_add(z, s_mul(a.d1, vz));
            if (st + 1 < a.n_steps) { x = s_add(x, s_mul(a.d2, vx)); y = s_add(y, s_mul(a.d2, vy)); z = s_add(z, s_mul(a.d2, vz)); }
        }
        __syncthreads();
    }
    if (live) {
        double* hs = a.hist + a.n_steps * 7 * Na;
        hs[0 * Na + j] = x; hs[1 * Na + j] = y; hs[2 * Na + j] = z; hs[3 * Na + j] = vx; hs[4 * Na + j] = vy; hs[5 * Na + j] = vz; hs[6 * Na + j] = m;
        a.s.x[j] = x; a.s.y[j] = y; a.s.z[j] = z; a.s.vx[j] = vx; a.s.vy[j] = vy; a.s.vz[j] = vz;
        a.s.ax[j] = ax; a.s.ay[j] = ay; a.s.az[j] = az;
    }
}

// Pair-parallel variant for Na <= TP_PAIR_MAX: thread (i,j) evaluates the pair prefactor and separation,
// thread i then accumulates its row in ascending j with exactly the operations of tp_force_strict /
// tp_force_fast, so the bits equal tp_history_kernel's while the ~300-cycle sqrt/divide chains of one step
// run side by side instead of back to back (the history is on the critical path of every call).
constexpr int TP_PAIR_MAX = 16;

template <bool FAST, bool KAHAN>
__global__ void __launch_bounds__(TP_PAIR_MAX * TP_PAIR_MAX) tp_history_pair_kernel(const TpMultiArgs a) {
    __shared__ double4 src[TP_PAIR_MAX];
    __shared__ double4 term[TP_PAIR_MAX * TP_PAIR_MAX];      // (p, dx, dy, dz) of pair (i,j)
    const int t = threadIdx.x, Na = a.Na;
    const bool live = t < Na;                                 // owner of body t
    const int pi = t / Na, pj = t - pi * Na;
    const bool pair = t < Na * Na && pi != pj;
    double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, m = 0, ax = 0, ay = 0, az = 0;
    if (live) { x = a.s.x[t]; y = a.s.y[t]; z = a.s.z[t]; vx = a.s.vx[t]; vy = a.s.vy[t]; vz = a.s.vz[t]; m = a.s.m[t]; }
    const double negG = -a.G;
    for (uint64_t st = 0; st < a.n_steps; st++) {
        if (live) {
            double* hs = a.hist + st * 7 * Na;
            hs[0 * Na + t] = x; hs[1 * Na + t] = y; hs[2 * Na + t] = z; hs[3 * Na + t] = vx; hs[4 * Na + t] = vy; hs[5 * Na + t] = vz; hs[6 * Na + t] = m;
            if (st == 0) { x = s_add(x, s_mul(a.d0, vx)); y = s_add(y, s_mul(a.d0, vy)); z = s_add(z, s_mul(a.d0, vz)); }
            src[t] = make_double4(x, y, z, m);
        }
        __syncthreads();
        if (pair) {
            const double4 si = src[pi], sj = src[pj];
            double xi = si.x, yi = si.y, zi = si.z;
            if (!KAHAN) { xi = s_add(a.gbx, xi); yi = s_add(a.gby, yi); zi = s_add(a.gbz, zi); }
            double p, dx, dy, dz;
            if (FAST) {
                dx = xi - sj.x; dy = yi - sj.y; dz = zi - sj.z;
                const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, a.soft2)));
                p = fast_m_over_r3(r2, negG * sj.w);
            } else {
                dx = s_sub(xi, sj.x); dy = s_sub(yi, sj.y); dz = s_sub(zi, sj.z);
                const double r2 = s_add(s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz)), a.soft2);
                const bool w = a.windowed && strict_window_key(r2) < STRICT_WINDOW_LIMIT;
                const double r = w ? fsqrt_rn_w(r2) : s_sqrt(r2);
                if (!KAHAN) {
                    const double b = s_mul(s_mul(r, r), r);
                    p = s_mul(w ? fdiv_rn_w(negG, b) : s_div(negG, b), sj.w);
                } else {
                    const double b = s_mul(r2, r);
                    p = s_mul(-(w ? fdiv_rn_w(a.G, b) : s_div(a.G, b)), sj.w);
                }
            }
            term[t] = make_double4(p, dx, dy, dz);
        }
        __syncthreads();
        if (live) {
            double cx = 0, cy = 0, cz = 0;
            ax = ay = az = 0;
            for (int j = 0; j < Na; j++) {
                if (j == t) continue;
                const double4 q = term[t * Na + j];
                if (FAST) {
                    if (!KAHAN) { ax = fma(q.x, q.y, ax); ay = fma(q.x, q.z, ay); az = fma(q.x, q.w, az); }
                    else {
                        double yy, tt;
                        yy = fma(q.x, q.y, -cx); tt = ax + yy; cx = (tt - ax) - yy; ax = tt;
                        yy = fma(q.x, q.z, -cy); tt = ay + yy; cy = (tt - ay) - yy; ay = tt;
                        yy = fma(q.x, q.w, -cz); tt = az + yy; cz = (tt - az) - yy; az = tt;
                    }
                } else if (!KAHAN) {
                    ax = s_add(ax, s_mul(q.x, q.y)); ay = s_add(ay, s_mul(q.x, q.z)); az = s_add(az, s_mul(q.x, q.w));
                } else {
                    double yy, tt;
                    yy = s_sub(s_mul(q.x, q.y), cx); tt = s_add(ax, yy); cx = s_sub(s_sub(tt, ax), yy); ax = tt;
                    yy = s_sub(s_mul(q.x, q.z), cy); tt = s_add(ay, yy); cy = s_sub(s_sub(tt, ay), yy); ay = tt;
                    yy = s_sub(s_mul(q.x, q.w), cz); tt = s_add(az, yy); cz = s_sub(s_sub(tt, az), yy); az = tt;
                }
            }
            vx = s_add(vx, s_mul(a.k, ax)); vy = s_add(vy, s_mul(a.k, ay)); vz = s_add(vz, s_mul(a.k, az));
            x = s_add(x, s_mul(a.d1, vx)); y = s_add(y, s_mul(a.d1, vy)); z = s_add(z, s_mul(a.d1, vz));
            if (st + 1 < a.n_steps) { x = s_add(x, s_mul(a.d2, vx)); y = s_add(y, s_mul(a.d2, vy)); z = s_add(z, s_mul(a.d2, vz)); }
        }
    }
    if (live) {
        double* hs = a.hist + a.n_steps * 7 * Na;
        hs[0 * Na + t] = x; hs[1 * Na + t] = y; hs[2 * Na + t] = z; hs[3 * Na + t] = vx; hs[4 * Na + t] = vy; hs[5 * Na + t] = vz; hs[6 * Na + t] = m;
        a.s.x[t] = x; a.s.y[t] = y; a.s.z[t] = z; a.s.vx[t] = vx; a.s.vy[t] = vy; a.s.vz[t] = vz;
        a.s.ax[t] = ax; a.s.ay[t] = ay; a.s.az[t] = az;
    }
}

// Test particles i in [max(i_begin, Na), i_end): all steps in one launch, massive bodies replayed from hist.
template <bool FAST, bool KAHAN>
__global__ void __launch_bounds__(TP_BLOCK) tp_multistep_kernel(const TpMultiArgs a) {
    __shared__ double4 src[2][TP_MAX_ACTIVE];
    const int Na = a.Na;
    const uint64_t i = a.i_begin + (uint64_t)blockIdx.x * TP_BLOCK + threadIdx.x;
    const bool live = i < a.i_end && i >= (uint64_t)Na;
    double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0;
    if (live) { x = a.s.x[i]; y = a.s.y[i]; z = a.s.z[i]; vx = a.s.vx[i]; vy = a.s.vy[i]; vz = a.s.vz[i]; }
    for (uint64_t st = 0; st < a.n_steps; st++) {
        double4* buf = src[st & 1];
        const double* hs = a.hist + st * 7 * Na;
        for (int j = threadIdx.x; j < Na; j += TP_BLOCK) {
            double sx = hs[0 * Na + j], sy = hs[1 * Na + j], sz = hs[2 * Na + j];
            if (st == 0) {
                sx = s_add(sx, s_mul(a.d0, hs[3 * Na + j])); sy = s_add(sy, s_mul(a.d0, hs[4 * Na + j])); sz = s_add(sz, s_mul(a.d0, hs[5 * Na + j]));
            }
            buf[j] = make_double4(sx, sy, sz, FAST ? -a.G * hs[6 * Na + j] : hs[6 * Na + j]);
        }
        __syncthreads();      // one barrier per step: the other buffer is only rewritten two steps later
        if (live) {
            if (st == 0) { x = s_add(x, s_mul(a.d0, vx)); y = s_add(y, s_mul(a.d0, vy)); z = s_add(z, s_mul(a.d0, vz)); }
            tp_force_any<FAST, KAHAN>(buf, Na, -1, false, x, y, z, a, ax, ay, az, FAST);
            vx = s_add(vx, s_mul(a.k, ax)); vy = s_add(vy, s_mul(a.k, ay)); vz = s_add(vz, s_mul(a.k, az));
            x = s_add(x, s_mul(a.d1, vx)); y = s_add(y, s_mul(a.d1, vy)); z = s_add(z, s_mul(a.d1, vz));
            if (st + 1 < a.n_steps) { x = s_add(x, s_mul(a.d2, vx)); y = s_add(y, s_mul(a.d2, vy)); z = s_add(z, s_mul(a.d2, vz)); }
        }
    }
    if (live) {
        a.s.x[i] = x; a.s.y[i] = y; a.s.z[i] = z; a.s.vx[i] = vx; a.s.vy[i] = vy; a.s.vz[i] = vz;
        a.s.ax[i] = ax; a.s.ay[i] = ay; a.s.az[i] = az;
    }
}

__global__ void tp_snapshot_kernel(Soa s, double* act, int Na) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Na) return;
    act[0 * Na + j] = s.x[j]; act[1 * Na + j] = s.y[j]; act[2 * Na + j] = s.z[j];
    act[3 * Na + j] = s.vx[j]; act[4 * Na + j] = s.vy[j]; act[5 * Na + j] = s.vz[j];
    act[6 * Na + j] = s.m[j];
}

// ---- SEI -------------------------------------------------------------------------------------
struct SeiConsts { double sindt, tandt, sindtz, tandtz, OMEGA, OMEGAZ, dt; };

// operator_H012, integrator_sei.c:126-157 (expression order preserved)
__device__ __forceinline__ void sei_h012(const SeiConsts& c, double& x, double& y, double& z, double& vx, double& vy, double& vz) {
    const double zx = s_mul(z, c.OMEGAZ);
    const double zy = vz;
    const double zt1 = s_sub(zx, s_mul(c.tandtz, zy));
    const double zyt = s_add(s_mul(c.sindtz, zt1), zy);
    const double zxt = s_sub(zt1, s_mul(c.tandtz, zyt));
    z = s_div(zxt, c.OMEGAZ);
    vz = zyt;
    const double aO = s_add(s_mul(2., vy), s_mul(s_mul(4., x), c.OMEGA));
    const double bO = s_sub(s_mul(y, c.OMEGA), s_mul(2., vx));
    const double ys = s_div(s_sub(s_mul(y, c.OMEGA), bO), 2.);
    const double xs = s_sub(s_mul(x, c.OMEGA), aO);
    const double xst1 = s_sub(xs, s_mul(c.tandt, ys));
    const double yst = s_add(s_mul(c.sindt, xst1), ys);
    const double xst = s_sub(xst1, s_mul(c.tandt, yst));
    x = s_div(s_add(xst, aO), c.OMEGA);
    y = s_sub(s_div(s_add(s_mul(yst, 2.), bO), c.OMEGA), s_mul(s_mul(3. / 4., aO), c.dt));
    vx = yst;
    vy = s_sub(s_mul(-xst, 2.), s_mul(3. / 2., aO));
}

// phase 0: H012 ; phase 1: phi1 (integrator_sei.c:168-174) then H012
__global__ void __launch_bounds__(256) sei_kernel(Soa s, SeiConsts c, int phase, uint64_t b, uint64_t e) {
    const uint64_t i = b + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= e) return;
    double x = s.x[i], y = s.y[i], z = s.z[i], vx = s.vx[i], vy = s.vy[i], vz = s.vz[i];
    if (phase == 1) {
        vx = s_add(vx, s_mul(s.ax[i], c.dt));
        vy = s_add(vy, s_mul(s.ay[i], c.dt));
        vz = s_add(vz, s_mul(s.az[i], c.dt));
    }
    sei_h012(c, x, y, z, vx, vy, vz);
    s.x[i] = x; s.y[i] = y; s.z[i] = z; s.vx[i] = vx; s.vy[i] = vy; s.vz[i] = vz;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// Leapfrog.  `carry_in`: the leading drift of this step was already applied by the previous step's
// last launch; `carry_out`: also apply the leading drift of the NEXT step (same dt) in this step's
// last launch.  Both are only used by rebcu_steps when nothing observes the state between steps.
// -------------------------------------------------------------------------------------------------
static bool tp_path_ok(const rebcu_handle* h, const rebcu_config* c) {
    const uint64_t Na = (c->N_active == REBCU_SIZE_MAX) ? h->N : c->N_active;
    return h->world == 1 && Na > 0 && Na <= TP_MAX_ACTIVE && Na < h->N && c->testparticle_type == 0
        && (c->gravity == REBCU_GRAVITY_BASIC || c->gravity == REBCU_GRAVITY_COMPENSATED)
        && c->N_ghost_x == 0 && c->N_ghost_y == 0 && c->N_ghost_z == 0
        && (c->leapfrog_order == 2 || c->leapfrog_order == 0);
}

int leapfrog_step_ex(rebcu_handle* h, rebcu_config* c, bool carry_in, bool carry_out, bool write_acc) {
    c->gravity_ignore_terms = REBCU_IGNORE_TERMS_NONE;        // integrator_leapfrog.c:98
    double drift[20], kick[20];
    const int order = c->leapfrog_order ? c->leapfrog_order : 2;
    const int nk = lf_schedule(order, c->dt, drift, kick);
    if (nk < 0) return rebcu_fail(h, REBCU_ERR_LEAPFROG_ORDER, "Leapfrog order not supported.");
    uint64_t b, e; engine_shard(h, &b, &e);
    const uint64_t n = e - b;
    Soa s = soa_of(h);

    if (tp_path_ok(h, c)) {
        const int Na = (int)c->N_active;
        double* act = h->scratch_big;
        if (!act) return rebcu_fail(h, REBCU_ERR_CUDA, "scratch missing");
        double* act_in = act + (size_t)h->tp_phase * 7 * TP_MAX_ACTIVE;
        double* act_out = act + (size_t)(1 - h->tp_phase) * 7 * TP_MAX_ACTIVE;
        if (!carry_in) {
            LaunchScope ls(h, TC_KICKDRIFT);
            tp_snapshot_kernel<<<div_up(Na, 128), 128, 0, h->stream>>>(s, act_in, Na);
        }
        TpArgs a;
        a.s = s; a.act_in = act_in; a.act_out = act_out; a.N = h->N; a.Na = Na; a.i_begin = 0;
        a.d0 = drift[0]; a.has_d0 = carry_in ? 0 : 1;
        a.k = kick[0]; a.d1 = drift[1]; a.d2 = drift[0]; a.has_d2 = carry_out ? 1 : 0;
        a.G = c->G; a.soft2 = c->softening * c->softening;
        { GhostShifts g0; engine_ghost_shifts(c, 0, 0, 0, &g0); a.gbx = g0.gb[0].x; a.gby = g0.gb[0].y; a.gbz = g0.gb[0].z; }
        a.kahan = c->gravity == REBCU_GRAVITY_COMPENSATED; a.fast = c->mode == REBCU_MODE_FAST;
        a.write_acc = write_acc ? 1 : 0;
        {
            LaunchScope ls(h, TC_DIRECT);
            const unsigned int nb = div_up(h->N, TP_BLOCK);
            a.windowed = strict_window_ok(c->G) ? 1 : 0;
            if (a.fast) { if (a.kahan) tp_leapfrog_kernel<true, true><<<nb, TP_BLOCK, 0, h->stream>>>(a); else tp_leapfrog_kernel<true, false><<<nb, TP_BLOCK, 0, h->stream>>>(a); }
            else { if (a.kahan) tp_leapfrog_kernel<false, true><<<nb, TP_BLOCK, 0, h->stream>>>(a); else tp_leapfrog_kernel<false, false><<<nb, TP_BLOCK, 0, h->stream>>>(a); }
        }
        CU_TRY(h, cudaGetLastError());
        h->tp_phase = 1 - h->tp_phase;
        c->t += drift[0]; c->t += drift[1];
        c->dt_last_done = c->dt;
        return REBCU_OK;
    }

    for (int k = 0; k < nk; k++) {
        if (k == 0 && !carry_in && n) {
            LaunchScope ls(h, TC_KICKDRIFT);
            drift_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(s, drift[0], b, e);
        }
        c->t += drift[k];
        int err = engine_exchange(h, REBCU_EXCHANGE_POSITIONS);
        if (err) return err;
        err = update_acceleration(h, c);
        if (err) return err;
        engine_shard(h, &b, &e);   // N may shrink (tree gravity + open boundary)
        s = soa_of(h);
        const bool last = (k == nk - 1);
        if (e > b) {
            LaunchScope ls(h, TC_KICKDRIFT);
            kick_drift_kernel<<<div_up(e - b, 256), 256, 0, h->stream>>>(s, kick[k], drift[k + 1], (last && carry_out) ? 1 : 0, drift[0], b, e);
        }
    }
    CU_TRY(h, cudaGetLastError());
    c->t += drift[nk];
    c->dt_last_done = c->dt;
    return REBCU_OK;
}

static TpMultiArgs tp_multi_args(rebcu_handle* h, const rebcu_config* c, uint64_t n_steps, const double* drift, const double* kick) {
    TpMultiArgs a;
    a.s = soa_of(h); a.hist = h->tp_hist; a.i_begin = 0; a.i_end = h->N; a.Na = (int)c->N_active; a.n_steps = n_steps;
    a.d0 = drift[0]; a.k = kick[0]; a.d1 = drift[1]; a.d2 = drift[0];
    a.G = c->G; a.soft2 = c->softening * c->softening;
    GhostShifts g0; engine_ghost_shifts(c, 0, 0, 0, &g0);
    a.gbx = g0.gb[0].x; a.gby = g0.gb[0].y; a.gbz = g0.gb[0].z;
    a.windowed = strict_window_ok(c->G) ? 1 : 0;
    a.fast = c->mode == REBCU_MODE_FAST; a.kahan = c->gravity == REBCU_GRAVITY_COMPENSATED;
    return a;
}
static void tp_launch_history(const TpMultiArgs& a, cudaStream_t s) {
    if (a.Na <= TP_PAIR_MAX) {
        const int nt = ((max(a.Na * a.Na, 32) + 31) / 32) * 32;
        if (a.fast) { if (a.kahan) tp_history_pair_kernel<true, true><<<1, nt, 0, s>>>(a); else tp_history_pair_kernel<true, false><<<1, nt, 0, s>>>(a); }
        else { if (a.kahan) tp_history_pair_kernel<false, true><<<1, nt, 0, s>>>(a); else tp_history_pair_kernel<false, false><<<1, nt, 0, s>>>(a); }
        return;
    }
    if (a.fast) { if (a.kahan) tp_history_kernel<true, true><<<1, TP_MAX_ACTIVE, 0, s>>>(a); else tp_history_kernel<true, false><<<1, TP_MAX_ACTIVE, 0, s>>>(a); }
    else { if (a.kahan) tp_history_kernel<false, true><<<1, TP_MAX_ACTIVE, 0, s>>>(a); else tp_history_kernel<false, false><<<1, TP_MAX_ACTIVE, 0, s>>>(a); }
}
static void tp_launch_multistep(const TpMultiArgs& a, cudaStream_t s) {
    const unsigned int nb = div_up(a.i_end - a.i_begin, TP_BLOCK);
    if (a.fast) { if (a.kahan) tp_multistep_kernel<true, true><<<nb, TP_BLOCK, 0, s>>>(a); else tp_multistep_kernel<true, false><<<nb, TP_BLOCK, 0, s>>>(a); }
    else { if (a.kahan) tp_multistep_kernel<false, true><<<nb, TP_BLOCK, 0, s>>>(a); else tp_multistep_kernel<false, false><<<nb, TP_BLOCK, 0, s>>>(a); }
}

static int tp_reserve_history(rebcu_handle* h, uint64_t n_steps, int Na) {
    const uint64_t hist_need = (n_steps + 1) * 7 * (uint64_t)Na;
    if (h->tp_hist_cap < hist_need) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->tp_hist); h->tp_hist = nullptr; h->tp_hist_cap = 0;
        CU_TRY(h, cudaMalloc(&h->tp_hist, hist_need * sizeof(double)));
        h->tp_hist_cap = hist_need;
    }
    return REBCU_OK;
}

// Resident version: reb_simulation_steps for the fused test-particle configuration in two launches
// (massive-body history, then every test particle through all steps).  Returns 1 if not eligible.
int tp_steps_resident(rebcu_handle* h, rebcu_config* c, uint64_t n_steps) {
    if (!(tp_path_ok(h, c) && c->integrator == REBCU_INTEGRATOR_LEAPFROG && c->boundary == REBCU_BOUNDARY_NONE
          && c->collision == REBCU_COLLISION_NONE && h->exchange == nullptr && n_steps >= 2 && n_steps <= 65536)) return 1;
    const int Na = (int)c->N_active;
    int err = tp_reserve_history(h, n_steps, Na);
    if (err) return err;
    c->gravity_ignore_terms = REBCU_IGNORE_TERMS_NONE;
    double drift[20], kick[20];
    lf_schedule(2, c->dt, drift, kick);
    TpMultiArgs a = tp_multi_args(h, c, n_steps, drift, kick);
    {
        LaunchScope ls(h, TC_KICKDRIFT);
        tp_launch_history(a, h->stream);
    }
    {
        LaunchScope ls(h, TC_DIRECT);
        tp_launch_multistep(a, h->stream);
    }
    CU_TRY(h, cudaGetLastError());
    for (uint64_t st = 0; st < n_steps; st++) { c->t += drift[0]; c->t += drift[1]; }
    c->dt_last_done = c->dt;
    return REBCU_OK;
}

// -------------------------------------------------------------------------------------------------
// reb_simulation_steps on a HOST particle array for the fused test-particle configuration, pipelined:
// the particles are cut into chunks; each chunk is uploaded, advanced through ALL n_steps and downloaded
// on one of two streams, so the PCIe copies of one chunk overlap the kernels of another (test particles of
// type 0 do not interact).  The massive bodies live in chunk 0, whose launches record their state after
// every step (tp_hist); the later chunks replay that history, so every chunk sees, step by step, exactly
// the massive-body positions the single-launch-per-step path would give it.  Bit-identical results.
// -------------------------------------------------------------------------------------------------
int tp_steps_host_pipelined(rebcu_handle* h, rebcu_config* c, rebcu_particle* particles, uint64_t N, uint64_t n_steps) {
    const uint64_t N_saved = h->N;
    h->N = N;
    const bool ok = tp_path_ok(h, c) && c->integrator == REBCU_INTEGRATOR_LEAPFROG && c->boundary == REBCU_BOUNDARY_NONE
                 && c->collision == REBCU_COLLISION_NONE && h->exchange == nullptr && !h->timing
                 && N >= (1u << 18) && n_steps >= 1 && n_steps <= 65536;
    h->N = N_saved;
    if (!ok) return 1;
    int err = engine_reserve(h, N);
    if (err) return err;
    h->N = N; h->resident = true; h->tree.built_for_n = -1;
    const int Na = (int)c->N_active;
    for (int k = 0; k < AUX_STREAMS; k++) if (!h->aux[k]) CU_TRY(h, cudaStreamCreateWithFlags(&h->aux[k], cudaStreamNonBlocking));
    for (int k = 0; k < 3; k++) if (!h->aux_ev[k]) CU_TRY(h, cudaEventCreateWithFlags(&h->aux_ev[k], cudaEventDisableTiming));
    if ((err = tp_reserve_history(h, n_steps, Na))) return err;
    c->gravity_ignore_terms = REBCU_IGNORE_TERMS_NONE;
    double drift[20], kick[20];
    lf_schedule(2, c->dt, drift, kick);
    TpMultiArgs a = tp_multi_args(h, c, n_steps, drift, kick);

    // everything queued on the handle's stream so far must be finished before the aux streams touch the buffers
    CU_TRY(h, cudaEventRecord(h->aux_ev[0], h->stream));
    for (int k = 0; k < AUX_STREAMS; k++) CU_TRY(h, cudaStreamWaitEvent(h->aux[k], h->aux_ev[0], 0));

    // Three-stage pipeline: aux[0] carries nothing but the H2D copies, aux[1] nothing but the D2H copies, and
    // aux[2..] the kernels (AoS->SoA, all steps, SoA->AoS) round-robin, chained by per-range events -- so both
    // DMA engines stream back to back (measured 2.3 ms for 112 MB each way, full duplex) and kernels of
    // neighbouring ranges overlap to fill each other's tail waves.
    // range 0 = the first 256 particles (holds the massive bodies): their history starts as early as possible
    int n_chunks = 24;
    if (const char* e = getenv("REBOUND_B200_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= PIPE_RANGES - 2) n_chunks = v; }
    const uint64_t head = 256;
    uint64_t cs = (N - head + n_chunks - 1) / n_chunks;
    cs = ((cs + 255) / 256) * 256;
    cudaStream_t s_up = h->aux[0], s_down = h->aux[1];
    const bool trace = getenv("REBOUND_B200_PIPE_TRACE") != nullptr;     // prints a per-range timeline to stderr
    cudaEvent_t ev_t0 = nullptr;
    if (trace) { CU_TRY(h, cudaEventCreate(&ev_t0)); CU_TRY(h, cudaEventRecord(ev_t0, s_up)); }
    int idx = 0;
    for (uint64_t b = 0; b < N; idx++) {
        if (idx >= PIPE_RANGES) return rebcu_fail(h, REBCU_ERR_ARG, "pipelined host path: too many ranges");
        const uint64_t e = (idx == 0) ? head : ((b + cs < N) ? b + cs : N);
        cudaEvent_t& ev_up = h->pipe_ev[3 * idx];
        cudaEvent_t& ev_k = h->pipe_ev[3 * idx + 1];
        cudaEvent_t& ev_dn = h->pipe_ev[3 * idx + 2];
        const unsigned evf = trace ? cudaEventDefault : cudaEventDisableTiming;
        if (trace && ev_up) { cudaEventDestroy(ev_up); cudaEventDestroy(ev_k); ev_up = ev_k = nullptr; }
        if (!ev_up) CU_TRY(h, cudaEventCreateWithFlags(&ev_up, evf));
        if (!ev_k) CU_TRY(h, cudaEventCreateWithFlags(&ev_k, evf));
        if (trace && !ev_dn) CU_TRY(h, cudaEventCreateWithFlags(&ev_dn, evf));
        cudaStream_t s = h->aux[2 + idx % (AUX_STREAMS - 2)];
        if ((err = engine_upload_range(h, s_up, ev_up, s, particles, b, e))) return err;
        if (idx == 0) {
            h->launches++;
            tp_launch_history(a, s);
            CU_TRY(h, cudaEventRecord(h->aux_ev[1], s));
        } else {
            CU_TRY(h, cudaStreamWaitEvent(s, h->aux_ev[1], 0));      // massive-body history complete
        }
        a.i_begin = b; a.i_end = e;
        h->launches++;
        tp_launch_multistep(a, s);
        CU_TRY(h, cudaGetLastError());
        if ((err = engine_download_range(h, s, ev_k, s_down, particles, b, e))) return err;
        if (trace) CU_TRY(h, cudaEventRecord(ev_dn, s_down));
        b = e;
    }
    for (int k = 0; k < AUX_STREAMS; k++) CU_TRY(h, cudaStreamSynchronize(h->aux[k]));
    if (trace) {
        for (int i = 0; i < idx; i++) {
            float tu = 0, tk = 0, td = 0;
            cudaEventElapsedTime(&tu, ev_t0, h->pipe_ev[3 * i]);
            cudaEventElapsedTime(&tk, ev_t0, h->pipe_ev[3 * i + 1]);
            cudaEventElapsedTime(&td, ev_t0, h->pipe_ev[3 * i + 2]);
            fprintf(stderr, "[pipe] range %2d  upload done %.3f  kernels done %.3f  download done %.3f ms\n", i, tu, tk, td);
            cudaEventDestroy(h->pipe_ev[3 * i]); cudaEventDestroy(h->pipe_ev[3 * i + 1]); cudaEventDestroy(h->pipe_ev[3 * i + 2]);
            h->pipe_ev[3 * i] = h->pipe_ev[3 * i + 1] = h->pipe_ev[3 * i + 2] = nullptr;
        }
        cudaEventDestroy(ev_t0);
    }
    for (uint64_t st = 0; st < n_steps; st++) { c->t += drift[0]; c->t += drift[1]; }
    c->dt_last_done = c->dt;
    return REBCU_OK;
}

int leapfrog_step(rebcu_handle* h, rebcu_config* c, bool fuse_ok) {
    (void)fuse_ok;
    return leapfrog_step_ex(h, c, false, false, true);
}

int sei_step(rebcu_handle* h, rebcu_config* c) {
    c->gravity_ignore_terms = REBCU_IGNORE_TERMS_NONE;        // integrator_sei.c:88
    if (c->OMEGAZ == -1) c->OMEGAZ = c->OMEGA;                // integrator_sei.c:93-95
    SeiConsts k;
    k.sindt = sin(c->OMEGA * (-c->dt / 2.));
    k.tandt = tan(c->OMEGA * (-c->dt / 4.));
    k.sindtz = sin(c->OMEGAZ * (-c->dt / 2.));
    k.tandtz = tan(c->OMEGAZ * (-c->dt / 4.));
    k.OMEGA = c->OMEGA; k.OMEGAZ = c->OMEGAZ; k.dt = c->dt;
    uint64_t b, e; engine_shard(h, &b, &e);
    if (e > b) {
        LaunchScope ls(h, TC_KICKDRIFT);
        sei_kernel<<<div_up(e - b, 256), 256, 0, h->stream>>>(soa_of(h), k, 0, b, e);
    }
    c->t += c->dt / 2.;
    int err = engine_exchange(h, REBCU_EXCHANGE_POSITIONS);
    if (err) return err;
    err = update_acceleration(h, c);
    if (err) return err;
    engine_shard(h, &b, &e);
    if (e > b) {
        LaunchScope ls(h, TC_KICKDRIFT);
        sei_kernel<<<div_up(e - b, 256), 256, 0, h->stream>>>(soa_of(h), k, 1, b, e);
    }
    CU_TRY(h, cudaGetLastError());
    c->t += c->dt / 2.;
    c->dt_last_done = c->dt;
    return REBCU_OK;
}
