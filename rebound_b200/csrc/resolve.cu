// resolve.cu -- hard-sphere collision resolve on the device, for callers of the C ABI that
// accept a documented relaxation of the parity bar.
//
// What the reference does (src/collision.c:336-404, 573-665): shuffle the list with rand_r, then walk it IN ORDER,
// each collision reading and changing the velocities of its two particles -- so a collision sees the outcome of
// every earlier collision that involves one of its particles.
//
// Here the same order is kept with conflict-free rounds: in a round every pending collision that is the earliest
// pending one of BOTH its particles is resolved (no two such collisions share a particle, and everything that had to
// happen before them already has); the globally earliest pending collision always qualifies, so the rounds terminate.
// The result is the one of the sequential loop up to the arithmetic inside a single resolve.
//
// The relaxation: the resolver rotates with atan2 / sin / cos and restitution laws use pow.  CUDA's libm does not
// return glibc's bits for these (and glibc picks FMA / non-FMA variants per CPU), so velocities after a resolved
// collision agree with the reference to a few ulp, not bit for bit; collisions_plog is a compensated parallel sum.
// The drop-in librebound therefore keeps resolving on the host with the reference's own functions; this entry point is
// for resident simulations driven through the C ABI (rebcu_set_device_resolve), where it removes the per-step
// download + serial host loop (C5: ~7e5 list entries per step).
#include "engine.cuh"
#include "strict_math.cuh"
#include <stdlib.h>
#include <math.h>
#include <time.h>
#include <vector>
#include <algorithm>

namespace {

constexpr uint32_t NONE = 0xffffffffu;

struct ResolveArgs {
    double *x, *y, *z, *vx, *vy, *vz;
    const double *m, *r;
    const rebcu_collision* list;
    const uint32_t* order;        // processing position k -> list index (the rand_r shuffle)
    uint32_t n;
    uint32_t* first;              // per particle: earliest pending processing position
    uint8_t* done;                // per processing position
    double* plog_term;            // per processing position
    unsigned long long* counters; // [0] pending after this round, [1] resolved (collisions_log_n)
    rebcu_restitution rest;
    double min_v;
};

__global__ void __launch_bounds__(256) claim_kernel(ResolveArgs A) {
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k >= A.n || A.done[k]) return;
    const rebcu_collision c = A.list[A.order[k]];
    atomicMin(&A.first[c.p1], k);
    atomicMin(&A.first[c.p2], k);
}

__device__ __forceinline__ double restitution(const rebcu_restitution& R, double v) {
    if (R.kind == REBCU_RESTITUTION_CONSTANT) return R.a;
    double eps = R.a * pow(fabs(v) * R.b, R.c);      // e.g. Bridges et al.: 0.32*pow(fabs(v)*100., -0.234)
    if (eps > R.hi) eps = R.hi;
    if (eps < R.lo) eps = R.lo;
    return eps;
}

// reb_collision_resolve_hardsphere, src/collision.c:573-665 (same expressions, device libm)
__global__ void __launch_bounds__(256) resolve_kernel(ResolveArgs A) {
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k >= A.n || A.done[k]) return;
    const rebcu_collision c = A.list[A.order[k]];
    const uint32_t p1 = (uint32_t)c.p1, p2 = (uint32_t)c.p2;
    if (A.first[p1] != k || A.first[p2] != k) { atomicAdd(&A.counters[0], 1ull); return; }     // an earlier collision is pending
    A.done[k] = 1;
    double term = 0.;
    const double x21 = A.x[p1] + c.gb.x - A.x[p2];
    const double y21 = A.y[p1] + c.gb.y - A.y[p2];
    const double z21 = A.z[p1] + c.gb.z - A.z[p2];
    const double r1 = A.r[p1], r2 = A.r[p2], m1 = A.m[p1], m2 = A.m[p2];
    const double rp = r1 + r2;
    const double v1x = A.vx[p1], v1y = A.vy[p1], v1z = A.vz[p1];
    const double v2x = A.vx[p2], v2y = A.vy[p2], v2z = A.vz[p2];
    const double oldvyouter = (x21 > 0) ? v1y : v2y;
    bool act = !(rp * rp < x21 * x21 + y21 * y21 + z21 * z21);
    const double vx21 = v1x + c.gb.vx - v2x;
    const double vy21 = v1y + c.gb.vy - v2y;
    const double vz21 = v1z + c.gb.vz - v2z;
    if (act && vx21 * x21 + vy21 * y21 + vz21 * z21 > 0) act = false;        // not approaching
    if (act) {
        const double theta = atan2(z21, y21);
        const double stheta = sin(theta), ctheta = cos(theta);
        const double vy21n = ctheta * vy21 + stheta * vz21;
        const double y21n = ctheta * y21 + stheta * z21;
        const double phi = atan2(y21n, x21);
        const double cphi = cos(phi), sphi = sin(phi);
        const double vx21nn = cphi * vx21 + sphi * vy21n;
        const double eps = restitution(A.rest, vx21nn);
        double dvx2 = -(1.0 + eps) * vx21nn;
        const double minr = (r1 > r2) ? r2 : r1;
        const double maxr = (r1 < r2) ? r2 : r1;
        double mindv = minr * A.min_v;
        const double rr = sqrt(x21 * x21 + y21 * y21 + z21 * z21);
        mindv *= 1. - (rr - maxr) / minr;
        if (mindv > maxr * A.min_v) mindv = maxr * A.min_v;
        if (dvx2 < mindv) dvx2 = mindv;
        const double dvx2n = cphi * dvx2;
        const double dvy2n = sphi * dvx2;
        const double dvy2nn = ctheta * dvy2n;
        const double dvz2nn = stheta * dvy2n;
        const double p2pf = m1 / (m1 + m2);
        const double n2x = v2x - p2pf * dvx2n, n2y = v2y - p2pf * dvy2nn, n2z = v2z - p2pf * dvz2nn;
        const double p1pf = m2 / (m1 + m2);
        const double n1x = v1x + p1pf * dvx2n, n1y = v1y + p1pf * dvy2nn, n1z = v1z + p1pf * dvz2nn;
        A.vx[p2] = n2x; A.vy[p2] = n2y; A.vz[p2] = n2z;
        A.vx[p1] = n1x; A.vy[p1] = n1y; A.vz[p1] = n1z;
        term = (x21 > 0) ? -fabs(x21) * (oldvyouter - n1y) * m1 : -fabs(x21) * (oldvyouter - n2y) * m2;
        atomicAdd(&A.counters[1], 1ull);
    }
    A.plog_term[k] = term;
}

// the particles of still pending collisions get a fresh "earliest" slot for the next round
__global__ void __launch_bounds__(256) release_kernel(ResolveArgs A) {
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k >= A.n) return;
    const rebcu_collision c = A.list[A.order[k]];
    A.first[c.p1] = NONE;
    A.first[c.p2] = NONE;
}

__global__ void __launch_bounds__(256) plog_partial_kernel(const double* __restrict__ t, uint32_t n, double* __restrict__ partial) {
    __shared__ double sm[8];
    double s = 0, e = 0;
    for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) { const double y = t[k] - e, u = s + y; e = (u - s) - y; s = u; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double v = 0; for (int w = 0; w < 8; w++) v += sm[w]; partial[blockIdx.x] = v; }
}

}  // namespace

// ---- exact resolve: the device keeps the order, the CALLER does the arithmetic ----------------------------------------
// For the drop-in, which owes the reference's bits.  The transcendental functions of the resolver (atan2 / sin / cos in
// reb_collision_resolve_hardsphere, pow in a user's restitution law) cannot be reproduced on the device, but everything
// around them can: the rand_r shuffle, the conflict-free rounds that keep the sequential semantics, and the two early
// exits of the resolver (no overlap / not approaching: src/collision.c:598,602 -- plain IEEE arithmetic, evaluated here
// with the reference's expression order).  Only the collisions that pass them travel to the host, as small records with
// the current state of their two particles; the caller's resolver (the reference's own function, run on a two-particle
// scratch simulation by the shim, in parallel -- the pairs of a round share no particle) returns the new velocities, which
// are scattered back.  Per step this moves a few tens of MB instead of the whole particle array both ways, and the host
// loop runs on all cores instead of one.  collisions_plog is summed by the caller's order (processing position), so it
// matches the sequential loop bit for bit as well.
struct PairsArgs {
    ResolveArgs R;
    rebcu_resolve_pair* pairs; unsigned int* n_pairs; unsigned int cap;      // 264-byte records, filled here, completed by the caller
};

__global__ void __launch_bounds__(256) ready_kernel(PairsArgs P) {
    const ResolveArgs& A = P.R;
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k >= A.n || A.done[k]) return;
    const rebcu_collision c = A.list[A.order[k]];
    const uint32_t p1 = (uint32_t)c.p1, p2 = (uint32_t)c.p2;
    if (A.first[p1] != k || A.first[p2] != k) { atomicAdd(&A.counters[0], 1ull); return; }     // an earlier collision is pending
    A.done[k] = 1;
    const double x1 = A.x[p1], y1 = A.y[p1], z1 = A.z[p1], x2 = A.x[p2], y2 = A.y[p2], z2 = A.z[p2];
    const double x21 = s_sub(s_add(x1, c.gb.x), x2), y21 = s_sub(s_add(y1, c.gb.y), y2), z21 = s_sub(s_add(z1, c.gb.z), z2);
    const double r1 = A.r[p1], r2 = A.r[p2];
    const double rp = s_add(r1, r2);
    const double d2 = s_add(s_add(s_mul(x21, x21), s_mul(y21, y21)), s_mul(z21, z21));
    if (s_mul(rp, rp) < d2) return;                                                          // collision.c:598
    const double v1x = A.vx[p1], v1y = A.vy[p1], v1z = A.vz[p1], v2x = A.vx[p2], v2y = A.vy[p2], v2z = A.vz[p2];
    const double vx21 = s_sub(s_add(v1x, c.gb.vx), v2x), vy21 = s_sub(s_add(v1y, c.gb.vy), v2y), vz21 = s_sub(s_add(v1z, c.gb.vz), v2z);
    if (s_add(s_add(s_mul(vx21, x21), s_mul(vy21, y21)), s_mul(vz21, z21)) > 0) return;      // collision.c:602: not approaching
    const unsigned int slot = atomicAdd(P.n_pairs, 1u);
    if (slot >= P.cap) return;                            // cannot happen: cap >= N/2 pairs share no particle
    rebcu_resolve_pair& q = P.pairs[slot];
    q.k = k; q.p1 = p1; q.p2 = p2; q.gb = c.gb;
    q.s1[0] = x1; q.s1[1] = y1; q.s1[2] = z1; q.s1[3] = v1x; q.s1[4] = v1y; q.s1[5] = v1z; q.s1[6] = A.m[p1]; q.s1[7] = r1;
    q.s2[0] = x2; q.s2[1] = y2; q.s2[2] = z2; q.s2[3] = v2x; q.s2[4] = v2y; q.s2[5] = v2z; q.s2[6] = A.m[p2]; q.s2[7] = r2;
    q.v1[0] = v1x; q.v1[1] = v1y; q.v1[2] = v1z; q.v2[0] = v2x; q.v2[1] = v2y; q.v2[2] = v2z;
    q.plog_term = 0.; q.logged = 0;
}

__global__ void __launch_bounds__(256) apply_pairs_kernel(ResolveArgs A, const rebcu_resolve_pair* __restrict__ pairs, unsigned int n) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const rebcu_resolve_pair& q = pairs[j];
    A.vx[q.p1] = q.v1[0]; A.vy[q.p1] = q.v1[1]; A.vz[q.p1] = q.v1[2];
    A.vx[q.p2] = q.v2[0]; A.vy[q.p2] = q.v2[1]; A.vz[q.p2] = q.v2[2];
}

// Resolves the list the last collision search left on the device.  Returns the number of rounds in *rounds (may be null).
int collision_resolve_device(rebcu_handle* h, const rebcu_config* c) {
    (void)c;
    const uint64_t n = h->col_n;
    if (n == 0) return REBCU_OK;
    if (n >= NONE) return rebcu_fail(h, REBCU_ERR_ARG, "device resolve supports fewer than 2^32 list entries");
    // the rand_r shuffle of the reference, applied to the processing order (collision.c:337-342)
    std::vector<uint32_t> order(n);
    for (uint64_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t j = (uint64_t)rand_r(&h->resolve_seed) % n;
        const uint32_t t = order[i]; order[i] = order[j]; order[j] = t;
    }
    const uint64_t words = n /*order*/ + h->cap /*first*/ + (n + 3) / 4 /*done*/ + 2 * n /*plog terms*/ + 2 * 1024 /*partials*/ + 64;
    if (h->resolve_cap < words) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->resolve_buf); h->resolve_buf = nullptr; h->resolve_cap = 0;
        CU_TRY(h, cudaMalloc(&h->resolve_buf, (words + words / 4) * sizeof(uint32_t)));
        h->resolve_cap = words + words / 4;
    }
    uint32_t* base = h->resolve_buf;
    ResolveArgs A;
    A.plog_term = (double*)base;                       // 8-byte aligned first
    double* partial = A.plog_term + n;
    A.order = base + 2 * n + 2 * 1024;
    A.first = (uint32_t*)A.order + n;
    A.done = (uint8_t*)(A.first + h->cap);
    A.x = h->f(F_X); A.y = h->f(F_Y); A.z = h->f(F_Z); A.vx = h->f(F_VX); A.vy = h->f(F_VY); A.vz = h->f(F_VZ);
    A.m = h->f(F_M); A.r = h->f(F_R);
    A.list = h->col_list; A.n = (uint32_t)n;
    A.counters = h->counters + 8;
    A.rest = h->resolve_rest; A.min_v = h->resolve_min_v;
    CU_TRY(h, cudaMemcpyAsync((void*)A.order, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemsetAsync(A.first, 0xff, h->cap * sizeof(uint32_t), h->stream));
    CU_TRY(h, cudaMemsetAsync(A.done, 0, n, h->stream));
    CU_TRY(h, cudaMemsetAsync(A.counters, 0, 2 * sizeof(unsigned long long), h->stream));
    const unsigned nb = div_up(n, 256);
    unsigned long long* pin = h->pinned + 20;
    int rounds = 0;
    for (;;) {
        {
            LaunchScope ls(h, TC_COLLISION, 3);
            claim_kernel<<<nb, 256, 0, h->stream>>>(A);
            resolve_kernel<<<nb, 256, 0, h->stream>>>(A);
            release_kernel<<<nb, 256, 0, h->stream>>>(A);
        }
        rounds++;
        CU_TRY(h, cudaMemcpyAsync(pin, A.counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(h, cudaMemsetAsync(A.counters, 0, sizeof(unsigned long long), h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        if (pin[0] == 0) break;
        if (rounds > 100000) return rebcu_fail(h, REBCU_ERR_CUDA, "device resolve did not terminate");
    }
    h->resolve_log_n += pin[1];
    h->resolve_rounds = rounds;
    // collisions_plog: compensated parallel sum of the per-collision terms
    const unsigned pb = nb < 1024 ? nb : 1024;
    {
        LaunchScope ls(h, TC_COLLISION);
        plog_partial_kernel<<<pb, 256, 0, h->stream>>>(A.plog_term, A.n, partial);
    }
    CU_TRY(h, cudaGetLastError());
    std::vector<double> part(pb);
    CU_TRY(h, cudaMemcpyAsync(part.data(), partial, pb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    double s = 0, e = 0;
    for (unsigned k = 0; k < pb; k++) { const double y = part[k] - e, u = s + y; e = (u - s) - y; s = u; }
    h->resolve_plog += s;
    h->col_n = 0;                                       // consumed
    return REBCU_OK;
}

extern "C" {

int rebcu_set_device_resolve(rebcu_handle* h, int enable, const rebcu_restitution* restitution, double minimum_collision_velocity,
                             unsigned int rand_seed) {
    if (enable) GROUP_UNSUPPORTED(h, "the device-side collision resolve");
    h->resolve_on = enable != 0;
    if (restitution) h->resolve_rest = *restitution;
    else { h->resolve_rest.kind = REBCU_RESTITUTION_CONSTANT; h->resolve_rest.a = 1.0; h->resolve_rest.b = h->resolve_rest.c = 0; h->resolve_rest.lo = 0; h->resolve_rest.hi = 1; }
    h->resolve_min_v = minimum_collision_velocity;
    h->resolve_seed = rand_seed;
    h->resolve_plog = 0; h->resolve_log_n = 0; h->resolve_rounds = 0;
    return REBCU_OK;
}

int rebcu_collision_resolve(rebcu_handle* h, const rebcu_config* cfg) {
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    if (h->world > 1) return rebcu_fail(h, REBCU_ERR_ARG, "device resolve while sharded over several GPUs is not implemented");
    return collision_resolve_device(h, cfg);
}

int rebcu_collision_resolve_pairs(rebcu_handle* h, unsigned int* rand_seed, rebcu_pair_resolver fn, void* user,
                                  double* plog, uint64_t* log_n, int* rounds_out) {
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    GROUP_UNSUPPORTED(h, "rebcu_collision_resolve_pairs");
    CU_TRY(h, cudaSetDevice(h->device));
    if (h->world > 1) return rebcu_fail(h, REBCU_ERR_ARG, "collision resolve while sharded over several GPUs is not implemented");
    const uint64_t n = h->col_n;
    if (rounds_out) *rounds_out = 0;
    if (n == 0) return REBCU_OK;
    if (n >= NONE) return rebcu_fail(h, REBCU_ERR_ARG, "device resolve supports fewer than 2^32 list entries");
    static const bool trace = getenv("REBOUND_B200_RESOLVE_TRACE") != nullptr;      // per-call timing breakdown on stderr
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    const double t_begin = now();
    double t_fn = 0, t_wait = 0;
    uint64_t n_pairs_total = 0;
    std::vector<uint32_t> order(n);
    for (uint64_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    for (uint64_t i = 0; i < n; i++) {                     // collision.c:337-342
        const uint64_t j = (uint64_t)rand_r(rand_seed) % n;
        const uint32_t t = order[i]; order[i] = order[j]; order[j] = t;
    }
    const uint64_t pair_cap = (n < h->N / 2 + 1) ? n : h->N / 2 + 1;
    const uint64_t words = n /*order*/ + h->cap /*first*/ + (n + 3) / 4 /*done*/ + 64 + pair_cap * sizeof(rebcu_resolve_pair) / 4;
    if (h->resolve_cap < words) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->resolve_buf); h->resolve_buf = nullptr; h->resolve_cap = 0;
        CU_TRY(h, cudaMalloc(&h->resolve_buf, (words + words / 4) * sizeof(uint32_t)));
        h->resolve_cap = words + words / 4;
    }
    if (h->pairs_host_cap < pair_cap) {
        if (h->pairs_host) cudaFreeHost(h->pairs_host);
        h->pairs_host = nullptr; h->pairs_host_cap = 0;
        CU_TRY(h, cudaMallocHost(&h->pairs_host, (pair_cap + pair_cap / 4 + 16) * sizeof(rebcu_resolve_pair)));
        h->pairs_host_cap = pair_cap + pair_cap / 4 + 16;
    }
    rebcu_resolve_pair* out_host = (rebcu_resolve_pair*)h->pairs_host;
    uint32_t* base = h->resolve_buf;
    PairsArgs P;
    ResolveArgs& A = P.R;
    P.pairs = (rebcu_resolve_pair*)base;                   // 8-byte aligned first
    rebcu_resolve_pair* out_dev = P.pairs;
    A.order = (uint32_t*)(out_dev + pair_cap);
    A.first = (uint32_t*)A.order + n;
    A.done = (uint8_t*)(A.first + h->cap);
    A.plog_term = nullptr;
    A.x = h->f(F_X); A.y = h->f(F_Y); A.z = h->f(F_Z); A.vx = h->f(F_VX); A.vy = h->f(F_VY); A.vz = h->f(F_VZ);
    A.m = h->f(F_M); A.r = h->f(F_R);
    A.list = h->col_list; A.n = (uint32_t)n;
    A.counters = h->counters + 8;
    A.rest = h->resolve_rest; A.min_v = 0;
    P.n_pairs = (unsigned int*)(h->counters + 10);
    P.cap = (unsigned int)pair_cap;
    CU_TRY(h, cudaMemcpyAsync((void*)A.order, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemsetAsync(A.first, 0xff, h->cap * sizeof(uint32_t), h->stream));
    CU_TRY(h, cudaMemsetAsync(A.done, 0, n, h->stream));
    CU_TRY(h, cudaMemsetAsync(h->counters + 8, 0, 3 * sizeof(unsigned long long), h->stream));
    const double t_setup = now();
    std::vector<double> terms(n, 0.);          // per processing position: the term of a logged collision
    std::vector<uint8_t> logged(n, 0);
    const unsigned nb = div_up(n, 256);
    unsigned long long* pin = h->pinned + 20;
    int rounds = 0;
    for (;;) {
        {
            LaunchScope ls(h, TC_COLLISION, 2);
            claim_kernel<<<nb, 256, 0, h->stream>>>(A);
            ready_kernel<<<nb, 256, 0, h->stream>>>(P);
        }
        rounds++;
        CU_TRY(h, cudaMemcpyAsync(pin, h->counters + 8, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(h, cudaMemsetAsync(h->counters + 8, 0, 3 * sizeof(unsigned long long), h->stream));
        { const double t0 = now(); CU_TRY(h, cudaStreamSynchronize(h->stream)); t_wait += now() - t0; }
        const unsigned long long pending = pin[0];
        const unsigned int np = (unsigned int)(pin[2] & 0xffffffffull);
        if (np > pair_cap) return rebcu_fail(h, REBCU_ERR_CAPACITY, "pair buffer overflow in the exact resolve");
        if (np) {
            CU_TRY(h, cudaMemcpyAsync(out_host, out_dev, np * sizeof(rebcu_resolve_pair), cudaMemcpyDeviceToHost, h->stream));
            { const double t0 = now(); CU_TRY(h, cudaStreamSynchronize(h->stream)); t_wait += now() - t0; }
            n_pairs_total += np;
            const double t_fn0 = now();
            const int ferr = fn(user, out_host, np);
            t_fn += now() - t_fn0;
            if (ferr) return rebcu_fail(h, REBCU_ERR_ARG, "the pair resolver reported an error");
            for (unsigned int j = 0; j < np; j++) if (out_host[j].logged) { terms[out_host[j].k] = out_host[j].plog_term; logged[out_host[j].k] = 1; }
            CU_TRY(h, cudaMemcpyAsync(out_dev, out_host, np * sizeof(rebcu_resolve_pair), cudaMemcpyHostToDevice, h->stream));
            LaunchScope ls(h, TC_COLLISION, 1);
            apply_pairs_kernel<<<div_up(np, 256), 256, 0, h->stream>>>(A, out_dev, np);
        }
        {
            LaunchScope ls(h, TC_COLLISION, 1);
            release_kernel<<<nb, 256, 0, h->stream>>>(A);
        }
        CU_TRY(h, cudaGetLastError());
        if (pending == 0) break;
        if (rounds > 100000) return rebcu_fail(h, REBCU_ERR_CUDA, "exact resolve did not terminate");
    }
    // collisions_plog += term, in the order of the sequential loop (collision.c:351, 655-661)
    double pl = *plog;
    uint64_t n_logged = 0;
    for (uint64_t k = 0; k < n; k++) if (logged[k]) { pl += terms[k]; n_logged++; }
    *plog = pl;
    *log_n += n_logged;
    h->resolve_rounds = rounds;
    if (rounds_out) *rounds_out = rounds;
    h->col_n = 0;                                           // consumed
    if (trace) fprintf(stderr, "[resolve] list %llu pairs %llu rounds %d: setup %.2f ms, device waits %.2f ms, host resolver %.2f ms, total %.2f ms\n",
                       (unsigned long long)n, (unsigned long long)n_pairs_total, rounds, 1e3 * (t_setup - t_begin), 1e3 * t_wait, 1e3 * t_fn, 1e3 * (now() - t_begin));
    return REBCU_OK;
}

int rebcu_collision_stats(const rebcu_handle* h, double* plog, uint64_t* log_n, unsigned int* rand_seed, int* rounds_last) {
    if (plog) *plog = h->resolve_plog;
    if (log_n) *log_n = h->resolve_log_n;
    if (rand_seed) *rand_seed = h->resolve_seed;
    if (rounds_last) *rounds_last = h->resolve_rounds;
    return REBCU_OK;
}

}  // extern "C"
