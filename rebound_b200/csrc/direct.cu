// direct.cu -- FP64 direct-summation gravity.
//
// Replaces reb_gravity_basic_calculate_acceleration (src/gravity.c:167-282) and
// reb_gravity_compensated_calculate_acceleration (src/gravity.c:284-531).
//
// Gather form: particle i accumulates its own sum over its sources j in ascending order, which is
// what the reference's OpenMP build does (gravity.c:216-232, 309-414) and what its serial build
// produces bit for bit when there are no ghost boxes.
//   sources(i) = [0, N_active)  U  ([N_active, N) if i is active and testparticle_type==1)
//   minus j==i and the pairs excluded by gravity_ignore_terms (gravity.c:219-221).
//
// STRICT kernels: one thread per i, j-tiles of (x,y,z,m) staged in shared memory (next tile
// prefetched into registers), arithmetic in the reference's expression order with
// __dadd_rn/__dmul_rn/__dsqrt_rn/__ddiv_rn => bit-identical accelerations.
// FAST kernels: 32 i per block, the j range split across the warps of the block, FMA + rsqrt;
// partial sums are combined in a fixed warp order (deterministic, <=1e-12 relative of STRICT).
//
// Bound: FP64 pipe.  HBM traffic is 32 B read per source per CTA tile + 24 B written per particle.
#include "engine.cuh"
#include "strict_math.cuh"
#include "fast_math.cuh"
#include <stdlib.h>

namespace {

constexpr uint64_t NO_SKIP = ~0ull;

struct DirectArgs {
    const double* x; const double* y; const double* z; const double* m;
    double* ax; double* ay; double* az;
    uint64_t N, Na;
    uint64_t i_begin, i_end;
    int type, terms;
    double G, soft2;
    const GhostShifts* ghosts;   // device
    int use_ghosts;              // 0: compensated (no shift at all, gravity.c:315-317)
    int windowed;                // G inside the window of the branch-free sqrt/divide (strict_math.cuh)
    double* csx; double* csy; double* csz;   // COMPENSATED: final Kahan compensation per particle (r->gravity_cs, gravity.c:297), else null
};

// Per-particle source range and the (at most two) excluded source indices.
__device__ __forceinline__ void source_set(const DirectArgs& a, uint64_t i, uint64_t& ns, uint64_t& skip0, uint64_t& skip1) {
    ns = (i < a.Na && a.type) ? a.N : a.Na;
    skip0 = i;
    skip1 = NO_SKIP;
    if (a.terms == REBCU_IGNORE_TERMS_BETWEEN_0_AND_1) { if (i < 2) skip1 = 1 - i; }
    else if (a.terms == REBCU_IGNORE_TERMS_INVOLVING_0) { if (i == 0) ns = 0; else skip1 = 0; }
}

// ------------------------------------------------------------------------------------------------
// STRICT
// ------------------------------------------------------------------------------------------------
// Generic-path recomputation of one particle (sources read straight from global memory); only runs for
// particles whose branch-free pass met an operand outside the fast range of fsqrt_rn / fdiv_rn.
template <bool KAHAN>
__device__ __forceinline__ void direct_slow(const DirectArgs& a, uint64_t i, uint64_t ns, uint64_t skip0, uint64_t skip1,
                                         double pxi, double pyi, double pzi, double& sx, double& sy, double& sz,
                                         double& cx, double& cy, double& cz) {
    sx = sy = sz = 0;
    cx = cy = cz = 0;
    const double negG = -a.G;
    const int ngb = a.use_ghosts ? a.ghosts->n : 1;
    for (int g = 0; g < ngb; g++) {
        double xi = pxi, yi = pyi, zi = pzi;
        if (a.use_ghosts) { xi = s_add(a.ghosts->gb[g].x, pxi); yi = s_add(a.ghosts->gb[g].y, pyi); zi = s_add(a.ghosts->gb[g].z, pzi); }
        for (uint64_t j = 0; j < ns; j++) {
            if (j == skip0 || j == skip1) continue;
            const double dx = s_sub(xi, a.x[j]), dy = s_sub(yi, a.y[j]), dz = s_sub(zi, a.z[j]);
            const double r2 = s_add(s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz)), a.soft2);
            const double r = s_sqrt(r2);
            if (!KAHAN) {
                const double p = s_mul(s_div(negG, s_mul(s_mul(r, r), r)), a.m[j]);
                sx = s_add(sx, s_mul(p, dx)); sy = s_add(sy, s_mul(p, dy)); sz = s_add(sz, s_mul(p, dz));
            } else {
                const double p = s_mul(-s_div(a.G, s_mul(r2, r)), a.m[j]);
                double y, t;
                y = s_sub(s_mul(p, dx), cx); t = s_add(sx, y); cx = s_sub(s_sub(t, sx), y); sx = t;
                y = s_sub(s_mul(p, dy), cy); t = s_add(sy, y); cy = s_sub(s_sub(t, sy), y); sy = t;
                y = s_sub(s_mul(p, dz), cz); t = s_add(sz, y); cz = s_sub(s_sub(t, sz), y); sz = t;
            }
        }
    }
}

struct StrictAcc { double sx, sy, sz, cx, cy, cz; unsigned wmax; };

// One shared-memory tile of sources against one particle.
//   WINDOWED: branch-free sqrt/divide (strict_math.cuh), the window key of the used terms is tracked in acc.wmax
//   PRED:     per-term predicate (source range end, the particle itself, ignored pairs); tiles a whole warp
//             can use unconditionally run with PRED=false and carry no integer work in the loop.
template <bool KAHAN, bool WINDOWED, bool PRED>
__device__ __forceinline__ void strict_tile(const double4* __restrict__ tile, int jn, uint64_t t0, uint64_t ns, uint64_t skip0,
                                            uint64_t skip1, double xi, double yi, double zi, double G, double soft2, StrictAcc& A) {
    const double negG = -G;
#pragma unroll 4
    for (int jj = 0; jj < jn; jj++) {
        const double4 s = tile[jj];
        const double dx = s_sub(xi, s.x);
        const double dy = s_sub(yi, s.y);
        const double dz = s_sub(zi, s.z);
        const double r2 = s_add(s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz)), soft2);
        const double r = WINDOWED ? fsqrt_rn_w(r2) : s_sqrt(r2);
        double p;
        if (!KAHAN) {
            const double b = s_mul(s_mul(r, r), r);          // prefact = -G/(_r*_r*_r)*particles[j].m   (gravity.c:226)
            p = s_mul(WINDOWED ? fdiv_rn_w(negG, b) : s_div(negG, b), s.w);
        } else {
            const double b = s_mul(r2, r);                   // prefact = G/(r2*r); prefactj = -prefact*m_j (gravity.c:320-321)
            p = s_mul(-(WINDOWED ? fdiv_rn_w(G, b) : s_div(G, b)), s.w);
        }
        if (PRED) {
            const uint64_t j = t0 + jj;
            if (!((j < ns) & (j != skip0) & (j != skip1))) continue;
        }
        if (WINDOWED) A.wmax = max(A.wmax, strict_window_key(r2));
        if (!KAHAN) {
            A.sx = s_add(A.sx, s_mul(p, dx));
            A.sy = s_add(A.sy, s_mul(p, dy));
            A.sz = s_add(A.sz, s_mul(p, dz));
        } else {
            double y, t;
            y = s_sub(s_mul(p, dx), A.cx); t = s_add(A.sx, y); A.cx = s_sub(s_sub(t, A.sx), y); A.sx = t;
            y = s_sub(s_mul(p, dy), A.cy); t = s_add(A.sy, y); A.cy = s_sub(s_sub(t, A.sy), y); A.sy = t;
            y = s_sub(s_mul(p, dz), A.cz); t = s_add(A.sz, y); A.cz = s_sub(s_sub(t, A.sz), y); A.sz = t;
        }
    }
}

// The unconditional windowed tile with U source terms advanced in lock step (see strict_math.cuh); the
// accumulation itself stays in ascending source order.
template <bool KAHAN, int U>
__device__ __forceinline__ void strict_tile_lockstep(const double4* __restrict__ tile, int jn, double xi, double yi, double zi,
                                                     double G, double soft2, StrictAcc& A) {
    const double negG = -G;
    int jj = 0;
    for (; jj + U <= jn; jj += U) {
        double dx[U], dy[U], dz[U], r2[U], r[U], b[U], q[U], m[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const double4 s = tile[jj + u];
            dx[u] = s_sub(xi, s.x); dy[u] = s_sub(yi, s.y); dz[u] = s_sub(zi, s.z); m[u] = s.w;
        }
#pragma unroll
        for (int u = 0; u < U; u++) r2[u] = s_add(s_add(s_add(s_mul(dx[u], dx[u]), s_mul(dy[u], dy[u])), s_mul(dz[u], dz[u])), soft2);
#pragma unroll
        for (int u = 0; u < U; u++) A.wmax = max(A.wmax, strict_window_key(r2[u]));
        fsqrt_rn_w_vec<U>(r2, r);
#pragma unroll
        for (int u = 0; u < U; u++) b[u] = KAHAN ? s_mul(r2[u], r[u]) : s_mul(s_mul(r[u], r[u]), r[u]);
        fdiv_rn_w_vec<U>(KAHAN ? G : negG, b, q);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const double p = KAHAN ? s_mul(-q[u], m[u]) : s_mul(q[u], m[u]);
            if (!KAHAN) {
                A.sx = s_add(A.sx, s_mul(p, dx[u]));
                A.sy = s_add(A.sy, s_mul(p, dy[u]));
                A.sz = s_add(A.sz, s_mul(p, dz[u]));
            } else {
                double y, t;
                y = s_sub(s_mul(p, dx[u]), A.cx); t = s_add(A.sx, y); A.cx = s_sub(s_sub(t, A.sx), y); A.sx = t;
                y = s_sub(s_mul(p, dy[u]), A.cy); t = s_add(A.sy, y); A.cy = s_sub(s_sub(t, A.sy), y); A.sy = t;
                y = s_sub(s_mul(p, dz[u]), A.cz); t = s_add(A.sz, y); A.cz = s_sub(s_sub(t, A.sz), y); A.sz = t;
            }
        }
    }
    if (jj < jn) strict_tile<KAHAN, true, false>(tile + jj, jn - jj, 0, ~0ull, NO_SKIP, NO_SKIP, xi, yi, zi, G, soft2, A);
}

template <bool KAHAN, int BLOCK, int JPT>
__global__ void __launch_bounds__(BLOCK) direct_strict_kernel(const DirectArgs a) {
    constexpr int TJ = BLOCK * JPT;
    __shared__ double4 tile[TJ];

    const uint64_t i0 = a.i_begin + (uint64_t)blockIdx.x * BLOCK;
    const uint64_t i = i0 + threadIdx.x;
    const bool valid = i < a.i_end;
    // block-uniform upper bound of the source range
    const uint64_t ns_blk = (a.type && i0 < a.Na) ? a.N : a.Na;
    uint64_t ns = ns_blk, skip0 = NO_SKIP, skip1 = NO_SKIP;      // idle lanes behave like "clean" lanes and never store
    double pxi = 0, pyi = 0, pzi = 0;
    if (valid) {
        source_set(a, i, ns, skip0, skip1);
        pxi = a.x[i]; pyi = a.y[i]; pzi = a.z[i];
    }
    StrictAcc A = {0, 0, 0, 0, 0, 0, 0};   // running sums, Kahan compensation (r->gravity_cs[i]), window key

    const int ngb = a.use_ghosts ? a.ghosts->n : 1;
    for (int g = 0; g < ngb; g++) {
        double xi = pxi, yi = pyi, zi = pzi;
        if (a.use_ghosts) {
            xi = s_add(a.ghosts->gb[g].x, pxi);
            yi = s_add(a.ghosts->gb[g].y, pyi);
            zi = s_add(a.ghosts->gb[g].z, pzi);
        }
        // prefetch tile 0
        double4 pre[JPT];
#pragma unroll
        for (int k = 0; k < JPT; k++) {
            const uint64_t j = (uint64_t)k * BLOCK + threadIdx.x;
            pre[k] = (j < ns_blk) ? make_double4(a.x[j], a.y[j], a.z[j], a.m[j]) : make_double4(0, 0, 0, 0);
        }
        for (uint64_t t0 = 0; t0 < ns_blk; t0 += TJ) {
            __syncthreads();   // previous tile fully consumed
#pragma unroll
            for (int k = 0; k < JPT; k++) tile[k * BLOCK + threadIdx.x] = pre[k];
            __syncthreads();
            const uint64_t tn = t0 + TJ;
            if (tn < ns_blk) {
#pragma unroll
                for (int k = 0; k < JPT; k++) {
                    const uint64_t j = tn + (uint64_t)k * BLOCK + threadIdx.x;
                    pre[k] = (j < ns_blk) ? make_double4(a.x[j], a.y[j], a.z[j], a.m[j]) : make_double4(0, 0, 0, 0);
                }
            }
            const int jn = (ns_blk - t0 < (uint64_t)TJ) ? (int)(ns_blk - t0) : TJ;
            const uint64_t t1 = t0 + (uint64_t)jn;
            const bool lane_clean = (t1 <= ns) && (skip0 < t0 || skip0 >= t1) && (skip1 < t0 || skip1 >= t1);
            const bool clean = __all_sync(0xffffffffu, lane_clean);      // warp-uniform
            if (a.windowed) {
                if (clean) strict_tile_lockstep<KAHAN, 4>(tile, jn, xi, yi, zi, a.G, a.soft2, A);
                else strict_tile<KAHAN, true, true>(tile, jn, t0, ns, skip0, skip1, xi, yi, zi, a.G, a.soft2, A);
            } else {
                strict_tile<KAHAN, false, true>(tile, jn, t0, ns, skip0, skip1, xi, yi, zi, a.G, a.soft2, A);
            }
        }
        __syncthreads();
    }
    if (valid) {
        if (A.wmax >= STRICT_WINDOW_LIMIT) direct_slow<KAHAN>(a, i, ns, skip0, skip1, pxi, pyi, pzi, A.sx, A.sy, A.sz, A.cx, A.cy, A.cz);
        a.ax[i] = A.sx; a.ay[i] = A.sy; a.az[i] = A.sz;
        if (KAHAN && a.csx) { a.csx[i] = A.cx; a.csy[i] = A.cy; a.csz[i] = A.cz; }
    }
}

// Strict kernel for MID-SIZE problems (too few particles to give every scheduler several warps with one
// thread per particle: C1, N = 16384, is 512 warps for 592 schedulers).  A CTA still owns 32 particles (lane =
// particle in every warp) but spreads the sources of a tile over SPLIT_W-1 producer warps: producer w evaluates
// the pair prefactor and the products p*dx, p*dy, p*dz for its SPLIT_T sources -- all of the sqrt/divide work -- and
// parks them in shared memory; warp 0 then adds the parked products to the running sums in ascending source
// order, exactly the additions (or Kahan updates) of the one-thread-per-particle kernel.  Only the 3 (12) adds per
// term are sequential, the other ~33 FP64 instructions run on 7x more warps; the bits do not change.
// SPLIT_W warps per CTA: 1 adder + SPLIT_W-1 producers; SPLIT_T sources per producer per tile (advanced in lock step)
template <bool KAHAN, int SPLIT_W, int SPLIT_T>
__global__ void __launch_bounds__(32 * SPLIT_W) direct_strict_split_kernel(const DirectArgs a) {
    constexpr int SPLIT_TJ = (SPLIT_W - 1) * SPLIT_T;
    __shared__ double term[2][SPLIT_TJ][3][32];
    __shared__ double4 src[2][SPLIT_TJ];
    __shared__ unsigned wflag[SPLIT_W][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t i0 = a.i_begin + (uint64_t)blockIdx.x * 32;
    const uint64_t i = i0 + lane;
    const bool valid = i < a.i_end;
    const uint64_t ns_blk = (a.type && i0 < a.Na) ? a.N : a.Na;
    uint64_t ns = 0, skip0 = NO_SKIP, skip1 = NO_SKIP;
    double pxi = 0, pyi = 0, pzi = 0;
    if (valid) {
        source_set(a, i, ns, skip0, skip1);
        pxi = a.x[i]; pyi = a.y[i]; pzi = a.z[i];
    }
    double sx = 0, sy = 0, sz = 0, cx = 0, cy = 0, cz = 0;
    unsigned wmax = 0;
    const double negG = -a.G;
    const uint64_t n_tiles = (ns_blk + SPLIT_TJ - 1) / SPLIT_TJ;
    const int ngb = a.use_ghosts ? a.ghosts->n : 1;
    for (int g = 0; g < ngb; g++) {
        double xi = pxi, yi = pyi, zi = pzi;
        if (a.use_ghosts) {
            xi = s_add(a.ghosts->gb[g].x, pxi);
            yi = s_add(a.ghosts->gb[g].y, pyi);
            zi = s_add(a.ghosts->gb[g].z, pzi);
        }
        __syncthreads();                         // the previous ghost box is fully consumed
        for (int jj = threadIdx.x; jj < SPLIT_TJ; jj += 32 * SPLIT_W) {
            const uint64_t j = jj;
            src[0][jj] = (j < ns_blk) ? make_double4(a.x[j], a.y[j], a.z[j], a.m[j]) : make_double4(0, 0, 0, 0);
        }
        __syncthreads();
        // iteration k: producers evaluate tile k into buffer k&1, the adder consumes tile k-1, warp 0 stages the
        // sources of tile k+1; one barrier per iteration
        for (uint64_t k = 0; k <= n_tiles; k++) {
            const int b = (int)(k & 1);
            if (w > 0) {
                if (k < n_tiles) {
                    const int j0 = (w - 1) * SPLIT_T;
                    double dx[SPLIT_T], dy[SPLIT_T], dz[SPLIT_T], r2[SPLIT_T], r[SPLIT_T], bb[SPLIT_T], q[SPLIT_T], m[SPLIT_T];
#pragma unroll
                    for (int u = 0; u < SPLIT_T; u++) {
                        const double4 sj = src[b][j0 + u];
                        dx[u] = s_sub(xi, sj.x); dy[u] = s_sub(yi, sj.y); dz[u] = s_sub(zi, sj.z); m[u] = sj.w;
                    }
#pragma unroll
                    for (int u = 0; u < SPLIT_T; u++) r2[u] = s_add(s_add(s_add(s_mul(dx[u], dx[u]), s_mul(dy[u], dy[u])), s_mul(dz[u], dz[u])), a.soft2);
#pragma unroll
                    for (int u = 0; u < SPLIT_T; u++) {
                        const uint64_t j = k * SPLIT_TJ + j0 + u;
                        if ((j < ns) & (j != skip0) & (j != skip1)) wmax = max(wmax, strict_window_key(r2[u]));
                    }
                    fsqrt_rn_w_vec<SPLIT_T>(r2, r);
#pragma unroll
                    for (int u = 0; u < SPLIT_T; u++) bb[u] = KAHAN ? s_mul(r2[u], r[u]) : s_mul(s_mul(r[u], r[u]), r[u]);
                    fdiv_rn_w_vec<SPLIT_T>(KAHAN ? a.G : negG, bb, q);
#pragma unroll
                    for (int u = 0; u < SPLIT_T; u++) {
                        const double p = KAHAN ? s_mul(-q[u], m[u]) : s_mul(q[u], m[u]);
                        term[b][j0 + u][0][lane] = s_mul(p, dx[u]);
                        term[b][j0 + u][1][lane] = s_mul(p, dy[u]);
                        term[b][j0 + u][2][lane] = s_mul(p, dz[u]);
                    }
                }
            } else {
                if (k + 1 < n_tiles) {
                    for (int jj = lane; jj < SPLIT_TJ; jj += 32) {
                        const uint64_t j = (k + 1) * SPLIT_TJ + jj;
                        src[b ^ 1][jj] = (j < ns_blk) ? make_double4(a.x[j], a.y[j], a.z[j], a.m[j]) : make_double4(0, 0, 0, 0);
                    }
                }
                if (k >= 1) {
                    const uint64_t t0 = (k - 1) * SPLIT_TJ;
#pragma unroll 4
                    for (int jj = 0; jj < SPLIT_TJ; jj++) {
                        const uint64_t j = t0 + jj;
                        if (!((j < ns) & (j != skip0) & (j != skip1))) continue;
                        const double tx = term[b ^ 1][jj][0][lane], ty = term[b ^ 1][jj][1][lane], tz = term[b ^ 1][jj][2][lane];
                        if (!KAHAN) {
                            sx = s_add(sx, tx); sy = s_add(sy, ty); sz = s_add(sz, tz);
                        } else {
                            double y, t;
                            y = s_sub(tx, cx); t = s_add(sx, y); cx = s_sub(s_sub(t, sx), y); sx = t;
                            y = s_sub(ty, cy); t = s_add(sy, y); cy = s_sub(s_sub(t, sy), y); sy = t;
                            y = s_sub(tz, cz); t = s_add(sz, y); cz = s_sub(s_sub(t, sz), y); sz = t;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    wflag[w][lane] = wmax;
    __syncthreads();
    if (w == 0 && valid) {
        unsigned wm = 0;
#pragma unroll
        for (int q = 1; q < SPLIT_W; q++) wm = max(wm, wflag[q][lane]);
        if (wm >= STRICT_WINDOW_LIMIT) direct_slow<KAHAN>(a, i, ns, skip0, skip1, pxi, pyi, pzi, sx, sy, sz, cx, cy, cz);
        a.ax[i] = sx; a.ay[i] = sy; a.az[i] = sz;
        if (KAHAN && a.csx) { a.csx[i] = cx; a.csy[i] = cy; a.csz[i] = cz; }
    }
}

// ------------------------------------------------------------------------------------------------
// FAST
// ------------------------------------------------------------------------------------------------
// Block = 32*IPT particles x W warps: every lane owns IPT particles (register tiling: one shared-memory read of a
// source serves IPT pair terms), warp w handles source chunks w, w+W, ... of 32 sources each, staged in a private
// shared-memory slab (double buffered through registers), and the W partial sums are combined in warp order.
// The pair term is fast_math.cuh's: 16 FP64 instructions (25 with the Kahan update), -G folded into the staged mass.
// Chunks that every particle of the block uses in full -- inside the common source range, not holding any of the
// block's own particles nor a source excluded by gravity_ignore_terms -- run a loop without any per-pair predicate;
// the few remaining chunks (the block's diagonal, the end of the range) select mass 0 / r2 1 for the excluded pairs.
template <bool KAHAN>
__device__ __forceinline__ void fast_accumulate(double f, double dx, double dy, double dz, double& sx, double& sy, double& sz,
                                                double& cx, double& cy, double& cz) {
    if (!KAHAN) {
        sx = fma(f, dx, sx); sy = fma(f, dy, sy); sz = fma(f, dz, sz);
    } else {
        double y, t;
        y = fma(f, dx, -cx); t = sx + y; cx = (t - sx) - y; sx = t;
        y = fma(f, dy, -cy); t = sy + y; cy = (t - sy) - y; sy = t;
        y = fma(f, dz, -cz); t = sz + y; cz = (t - sz) - y; sz = t;
    }
}

template <bool KAHAN, int W, int IPT>
__global__ void __launch_bounds__(32 * W) direct_fast_kernel(const DirectArgs a) {
    constexpr int NR = KAHAN ? 6 : 3;
    __shared__ double4 slab[W][32];
    __shared__ double red[W][IPT][NR][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t i0 = a.i_begin + (uint64_t)blockIdx.x * (32 * IPT);
    uint64_t i[IPT], ns[IPT], skip0[IPT], skip1[IPT];
    bool valid[IPT];
    double pxi[IPT], pyi[IPT], pzi[IPT];
    double sx[IPT], sy[IPT], sz[IPT], cx[IPT], cy[IPT], cz[IPT];
#pragma unroll
    for (int q = 0; q < IPT; q++) {
        i[q] = i0 + (uint64_t)q * 32 + lane;
        valid[q] = i[q] < a.i_end;
        ns[q] = 0; skip0[q] = NO_SKIP; skip1[q] = NO_SKIP;
        pxi[q] = pyi[q] = pzi[q] = 0;
        if (valid[q]) {
            source_set(a, i[q], ns[q], skip0[q], skip1[q]);
            pxi[q] = a.x[i[q]]; pyi[q] = a.y[i[q]]; pzi[q] = a.z[i[q]];
        }
        sx[q] = sy[q] = sz[q] = cx[q] = cy[q] = cz[q] = 0;
    }
    const uint64_t i_last = ((i0 + 32 * IPT < a.i_end) ? i0 + 32 * IPT : a.i_end) - 1;
    const uint64_t ns_blk = (a.type && i0 < a.Na) ? a.N : a.Na;          // the longest source range in the block
    uint64_t ns_lo = (a.type && i_last < a.Na) ? a.N : a.Na;              // the range every particle of the block uses
    if (a.terms == REBCU_IGNORE_TERMS_INVOLVING_0 && i0 == 0) ns_lo = 0;  // particle 0 has no sources at all
    const bool has_terms = a.terms != REBCU_IGNORE_TERMS_NONE;
    const double negG = -a.G;

    const int ngb = a.use_ghosts ? a.ghosts->n : 1;
    for (int g = 0; g < ngb; g++) {
        double xi[IPT], yi[IPT], zi[IPT];
#pragma unroll
        for (int q = 0; q < IPT; q++) {
            xi[q] = pxi[q]; yi[q] = pyi[q]; zi[q] = pzi[q];
            if (a.use_ghosts) { xi[q] += a.ghosts->gb[g].x; yi[q] += a.ghosts->gb[g].y; zi[q] += a.ghosts->gb[g].z; }
        }
        uint64_t t0 = (uint64_t)w * 32;
        double4 pre = make_double4(0, 0, 0, 0);
        if (t0 + lane < ns_blk) { const uint64_t j = t0 + lane; pre = make_double4(a.x[j], a.y[j], a.z[j], negG * a.m[j]); }
        for (; t0 < ns_blk; t0 += 32 * W) {
            __syncwarp();
            slab[w][lane] = pre;
            __syncwarp();
            const uint64_t tn = t0 + 32 * W + lane;
            pre = make_double4(0, 0, 0, 0);
            if (tn < ns_blk) pre = make_double4(a.x[tn], a.y[tn], a.z[tn], negG * a.m[tn]);
            const bool plain = (t0 + 32 <= ns_lo) && (t0 + 32 <= i0 || t0 > i_last) && !(has_terms && t0 < 2);
            if (plain) {
#pragma unroll 4
                for (int jj = 0; jj < 32; jj++) {
                    const double4 s = slab[w][jj];
#pragma unroll
                    for (int q = 0; q < IPT; q++) {
                        const double dx = xi[q] - s.x, dy = yi[q] - s.y, dz = zi[q] - s.z;
                        const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, a.soft2)));
                        const double f = fast_m_over_r3(r2, s.w);
                        fast_accumulate<KAHAN>(f, dx, dy, dz, sx[q], sy[q], sz[q], cx[q], cy[q], cz[q]);
                    }
                }
            } else {
                const int jn = (ns_blk - t0 < 32ull) ? (int)(ns_blk - t0) : 32;
                for (int jj = 0; jj < jn; jj++) {
                    const uint64_t j = t0 + jj;
                    const double4 s = slab[w][jj];
#pragma unroll
                    for (int q = 0; q < IPT; q++) {
                        const bool ok = (j < ns[q]) & (j != skip0[q]) & (j != skip1[q]);
                        const double dx = xi[q] - s.x, dy = yi[q] - s.y, dz = zi[q] - s.z;
                        const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, a.soft2)));
                        const double f = fast_m_over_r3(ok ? r2 : 1.0, ok ? s.w : 0.0);
                        fast_accumulate<KAHAN>(f, dx, dy, dz, sx[q], sy[q], sz[q], cx[q], cy[q], cz[q]);
                    }
                }
            }
        }
    }
    // combine the W partial sums in warp order (fixed => deterministic)
#pragma unroll
    for (int q = 0; q < IPT; q++) {
        red[w][q][0][lane] = sx[q]; red[w][q][1][lane] = sy[q]; red[w][q][2][lane] = sz[q];
        if (KAHAN) { red[w][q][3][lane] = cx[q]; red[w][q][4][lane] = cy[q]; red[w][q][5][lane] = cz[q]; }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < IPT; q++) {
        if ((q % W) != w || !valid[q]) continue;
        double tx = 0, ty = 0, tz = 0, ex = 0, ey = 0, ez = 0;
        for (int k = 0; k < W; k++) {
            // two-sum of the partials; the per-warp Kahan residuals are folded into the error term
            double v, t, bb;
            v = red[k][q][0][lane]; t = tx + v; bb = t - tx; ex += (tx - (t - bb)) + (v - bb) - (KAHAN ? red[k][q][3][lane] : 0.0); tx = t;
            v = red[k][q][1][lane]; t = ty + v; bb = t - ty; ey += (ty - (t - bb)) + (v - bb) - (KAHAN ? red[k][q][4][lane] : 0.0); ty = t;
            v = red[k][q][2][lane]; t = tz + v; bb = t - tz; ez += (tz - (t - bb)) + (v - bb) - (KAHAN ? red[k][q][5][lane] : 0.0); tz = t;
        }
        const double fx = tx + ex, fy = ty + ey, fz = tz + ez;
        a.ax[i[q]] = fx; a.ay[i[q]] = fy; a.az[i[q]] = fz;
        // what the last addition lost, in the sign convention of a Kahan compensation (true sum = a - cs)
        if (KAHAN && a.csx) { a.csx[i[q]] = (fx - tx) - ex; a.csy[i[q]] = (fy - ty) - ey; a.csz[i[q]] = (fz - tz) - ez; }
    }
}

// ------------------------------------------------------------------------------------------------
// Active rows of testparticle_type 1: a handful of massive particles that feel EVERY particle.
// ------------------------------------------------------------------------------------------------
// With one thread per target such a row is a serial loop over N sources (137 ms per evaluation at N_active = 10,
// N = 2^20: slower than the reference's CPU).  STRICT: the terms p*dx, p*dy, p*dz of all (row, source) pairs are
// evaluated by one thread per SOURCE (all the sqrt/divide work, fully parallel) into a term buffer, then one warp
// per row adds them in ascending source order -- the same additions as direct_strict_kernel, so the same bits; what
// remains serial is one dependent add (four with Kahan) per source.  FAST: per-CTA partial sums, fixed-order final sum.
constexpr int ROW_MAX = 256;

struct RowArgs {
    DirectArgs a;
    uint64_t row0; int n_rows;          // rows [row0, row0 + n_rows), all < N_active
    uint64_t j0, j1;                    // source chunk of this launch
    double* terms; uint64_t chunk_cap;  // [chunk_cap/2 source pairs][n_rows][3][2]
    double* state;                      // [n_rows][6] running sx sy sz cx cy cz between chunks
    double gbx, gby, gbz;               // ghost box (0,0,0) shift, added as the reference does (BASIC only)
    int first, last;                    // first / last chunk
};

template <bool KAHAN>
__global__ void __launch_bounds__(128) row_terms_kernel(const RowArgs R) {
    __shared__ double4 rows[ROW_MAX];
    const DirectArgs& a = R.a;
    for (int r = threadIdx.x; r < R.n_rows; r += 128) {
        const uint64_t i = R.row0 + r;
        double xi = a.x[i], yi = a.y[i], zi = a.z[i];
        if (a.use_ghosts) { xi = s_add(R.gbx, xi); yi = s_add(R.gby, yi); zi = s_add(R.gbz, zi); }
        rows[r] = make_double4(xi, yi, zi, 0.);
    }
    __syncthreads();
    const uint64_t j = R.j0 + (uint64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= R.j1) return;
    const double xj = a.x[j], yj = a.y[j], zj = a.z[j], mj = a.m[j];
    const double negG = -a.G;
    const uint64_t col = j - R.j0;
    for (int r = 0; r < R.n_rows; r++) {
        const double4 ri = rows[r];
        const double dx = s_sub(ri.x, xj), dy = s_sub(ri.y, yj), dz = s_sub(ri.z, zj);
        const double r2 = s_add(s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz)), a.soft2);
        const unsigned key = strict_window_key(r2);
        double p;
        if (a.windowed && key < STRICT_WINDOW_LIMIT) {
            const double d = fsqrt_rn_w(r2);
            p = KAHAN ? s_mul(-fdiv_rn_w(a.G, s_mul(r2, d)), mj) : s_mul(fdiv_rn_w(negG, s_mul(s_mul(d, d), d)), mj);
        } else {
            const double d = s_sqrt(r2);          // correctly rounded either way: no recomputation needed
            p = KAHAN ? s_mul(-s_div(a.G, s_mul(r2, d)), mj) : s_mul(s_div(negG, s_mul(s_mul(d, d), d)), mj);
        }
        // [source pair][row][component][2]: the two sources of a pair are adjacent, so the ordered sum moves 16 bytes
        // per chain and pair, and the chains of a pair are contiguous (a warp-wide access touches four cache lines)
        double* t = R.terms + ((col >> 1) * (uint64_t)(3 * R.n_rows) + 3 * r) * 2 + (col & 1);
        t[0] = s_mul(p, dx); t[2] = s_mul(p, dy); t[4] = s_mul(p, dz);
    }
}

// Ordered sums.  Lane (row, component): one warp carries the 30 chains of 10 rows in lock step, so one DADD
// instruction per source advances all of them.
constexpr int ROWS_PER_WARP = 10;
constexpr int ROW_BATCH = 16;
constexpr int ROW_STAGES = 8;

template <bool KAHAN>
__global__ void __launch_bounds__(32) row_sum_kernel(const RowArgs R) {
    const DirectArgs& a = R.a;
    const int lane = threadIdx.x;
    const int rl = lane / 3, c = lane - 3 * rl;
    const int r = blockIdx.x * ROWS_PER_WARP + rl;
    if (rl >= ROWS_PER_WARP || r >= R.n_rows) return;
    const uint64_t i = R.row0 + r;
    uint64_t ns, skip0, skip1;
    source_set(a, i, ns, skip0, skip1);
    double s = 0, e = 0;
    if (!R.first) { s = R.state[(uint64_t)r * 6 + c]; e = R.state[(uint64_t)r * 6 + 3 + c]; }
    const uint64_t S = (uint64_t)3 * R.n_rows;
    const double* t = R.terms + ((uint64_t)r * 3 + c) * 2;       // + (source / 2) * 2S + (source & 1)
    const uint64_t cnt = R.j1 - R.j0;
    // The chain is one dependent add per source; everything else must stay off its critical path.  The terms of one
    // source are contiguous over the chains, so a warp-wide 8-byte access touches two cache lines; they are streamed
    // into a shared-memory ring with cp.async (no registers held while in flight): ROW_STAGES-1 batches of 16 sources
    // are under way while one batch is added, which covers the L2 round trip (~700 cycles) of a single warp.
    __shared__ double2 ring[ROW_STAGES][ROW_BATCH / 2][32];
    const uint64_t n_batches = (cnt + ROW_BATCH - 1) / ROW_BATCH;
    auto issue = [&](uint64_t bt) {
        if (bt < n_batches) {
            const double* src = t + bt * ROW_BATCH * S;          // (bt*ROW_BATCH/2) pairs * 2S doubles; the buffer is padded
            const int st = (int)(bt % ROW_STAGES);
#pragma unroll
            for (int u = 0; u < ROW_BATCH / 2; u++) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&ring[st][u][lane]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (uint64_t)u * 2 * S) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (uint64_t bt = 0; bt + 1 < ROW_STAGES; bt++) issue(bt);
    for (uint64_t bt = 0; bt < n_batches; bt++) {
        issue(bt + ROW_STAGES - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(ROW_STAGES - 1) : "memory");
        const int st = (int)(bt % ROW_STAGES);
        const uint64_t b0 = bt * ROW_BATCH;
        const uint64_t jb = R.j0 + b0, je = jb + ROW_BATCH;
        const bool clean = (b0 + ROW_BATCH <= cnt) && (je <= ns) && (skip0 < jb || skip0 >= je) && (skip1 < jb || skip1 >= je);
        double v[ROW_BATCH];
#pragma unroll
        for (int u = 0; u < ROW_BATCH / 2; u++) { const double2 w2 = ring[st][u][lane]; v[2 * u] = w2.x; v[2 * u + 1] = w2.y; }
        if (clean) {
#pragma unroll
            for (int u = 0; u < ROW_BATCH; u++) {
                if (!KAHAN) s = s_add(s, v[u]);
                else { const double y = s_sub(v[u], e), w = s_add(s, y); e = s_sub(s_sub(w, s), y); s = w; }
            }
        } else {
#pragma unroll
            for (int u = 0; u < ROW_BATCH; u++) {
                const uint64_t j = jb + u;
                if ((b0 + u < cnt) & (j < ns) & (j != skip0) & (j != skip1)) {
                    if (!KAHAN) s = s_add(s, v[u]);
                    else { const double y = s_sub(v[u], e), w = s_add(s, y); e = s_sub(s_sub(w, s), y); s = w; }
                }
            }
        }
    }
    if (R.last) {
        double* out = (c == 0) ? a.ax : (c == 1) ? a.ay : a.az;
        out[i] = s;
        if (KAHAN && a.csx) { double* cs = (c == 0) ? a.csx : (c == 1) ? a.csy : a.csz; cs[i] = e; }
    } else {
        R.state[(uint64_t)r * 6 + c] = s; R.state[(uint64_t)r * 6 + 3 + c] = e;
    }
}

// FAST: CTA b owns sources [j0 + 1024 b, +1024); for every row each thread sums its 8 sources, the CTA reduces in a
// fixed order and writes one partial per (row, CTA); row_fast_final_kernel adds the partials in CTA order.
template <bool KAHAN>
__global__ void __launch_bounds__(128) row_fast_kernel(const RowArgs R, double* __restrict__ partial, unsigned n_cta) {
    __shared__ double4 rows[ROW_MAX];
    __shared__ double red[4][3];
    const DirectArgs& a = R.a;
    for (int r = threadIdx.x; r < R.n_rows; r += 128) {
        const uint64_t i = R.row0 + r;
        double xi = a.x[i], yi = a.y[i], zi = a.z[i];
        if (a.use_ghosts) { xi += R.gbx; yi += R.gby; zi += R.gbz; }
        rows[r] = make_double4(xi, yi, zi, 0.);
    }
    double4 src[8];
    uint64_t jj[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        jj[k] = R.j0 + (uint64_t)blockIdx.x * 1024 + (uint64_t)k * 128 + threadIdx.x;
        src[k] = (jj[k] < R.j1) ? make_double4(a.x[jj[k]], a.y[jj[k]], a.z[jj[k]], -a.G * a.m[jj[k]]) : make_double4(0, 0, 0, 0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int r = 0; r < R.n_rows; r++) {
        const uint64_t i = R.row0 + r;
        uint64_t ns, skip0, skip1;
        source_set(a, i, ns, skip0, skip1);
        const double4 ri = rows[r];
        double px = 0, py = 0, pz = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double dx = ri.x - src[k].x, dy = ri.y - src[k].y, dz = ri.z - src[k].z;
            const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, a.soft2)));
            const bool ok = (jj[k] < R.j1) & (jj[k] < ns) & (jj[k] != skip0) & (jj[k] != skip1);
            const double p = fast_m_over_r3(ok ? r2 : 1.0, ok ? src[k].w : 0.0);
            px = fma(p, dx, px); py = fma(p, dy, py); pz = fma(p, dz, pz);
        }
        for (int o = 16; o > 0; o >>= 1) {
            px += __shfl_down_sync(0xffffffffu, px, o); py += __shfl_down_sync(0xffffffffu, py, o); pz += __shfl_down_sync(0xffffffffu, pz, o);
        }
        __syncthreads();
        if (lane == 0) { red[w][0] = px; red[w][1] = py; red[w][2] = pz; }
        __syncthreads();
        if (threadIdx.x < 3) {
            const int c = threadIdx.x;
            partial[((uint64_t)r * 3 + c) * n_cta + blockIdx.x] = ((red[0][c] + red[1][c]) + red[2][c]) + red[3][c];
        }
    }
}

template <bool KAHAN>
__global__ void __launch_bounds__(128) row_fast_final_kernel(const RowArgs R, const double* __restrict__ partial, unsigned n_cta) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= R.n_rows * 3) return;
    const int r = t / 3, c = t - 3 * r;
    const double* p = partial + (uint64_t)t * n_cta;
    double s = 0, e = 0;
    for (unsigned k = 0; k < n_cta; k++) { const double y = p[k] - e, u = s + y; e = (u - s) - y; s = u; }
    const uint64_t i = R.row0 + r;
    double* out = (c == 0) ? R.a.ax : (c == 1) ? R.a.ay : R.a.az;
    out[i] = s;
    if (KAHAN && R.a.csx) { double* cs = (c == 0) ? R.a.csx : (c == 1) ? R.a.csy : R.a.csz; cs[i] = e; }
}

__global__ void zero3_kernel(double* ax, double* ay, double* az, uint64_t b, uint64_t e) {
    const uint64_t i = b + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < e) { ax[i] = 0; ay[i] = 0; az[i] = 0; }
}

template <bool KAHAN>
void launch_strict(rebcu_handle* h, const DirectArgs& a, uint64_t n_i) {
    // Spread small problems over all 148 SMs: one warp per CTA until there are >= 2 CTAs per SM.
    // Mid-size problems: too few particles to keep the FP64 pipe busy with one thread per particle, so the sources
    // of a tile are spread over producer warps (direct_strict_split_kernel).  Measured on B200, Plummer sphere,
    // split vs one-thread-per-particle: BASIC N=1024 0.036 vs 0.114 ms, 4096 0.13 vs 0.44, 16384 1.04 vs 1.75,
    // 32768 3.58 vs 4.26, 49152 8.2 vs 7.2 (crossover); COMPENSATED 16384 1.67 vs 2.11, 24576 4.1 vs 2.6 (the
    // Kahan update makes the adder's dependent chain four times longer, so it crosses over earlier).
    const uint64_t split_max = KAHAN ? 20480 : 40960;
    if (a.windowed && a.Na >= 256 && n_i < split_max) {
        // 1 adder + 3 producer warps, 8 sources per producer in lock step: fastest of the (W, T) shapes that fit
        // 48 KB of static shared memory (N = 16384: (4,8) 1.06 ms, (4,4) 1.11, (8,4) 1.18, (6,6) 1.33, (12,2) 1.47,
        // (8,2) 1.84)
        direct_strict_split_kernel<KAHAN, 4, 8><<<div_up(n_i, 32), 128, 0, h->stream>>>(a);
        return;
    }
    if (n_i >= 148ull * 128 * 2) direct_strict_kernel<KAHAN, 128, 1><<<div_up(n_i, 128), 128, 0, h->stream>>>(a);
    else if (n_i >= 148ull * 64 * 2) direct_strict_kernel<KAHAN, 64, 2><<<div_up(n_i, 64), 64, 0, h->stream>>>(a);
    else direct_strict_kernel<KAHAN, 32, 4><<<div_up(n_i, 32), 32, 0, h->stream>>>(a);
}

template <bool KAHAN, int IPT>
void launch_fast_ipt(rebcu_handle* h, const DirectArgs& a, uint64_t n_i) {
    const unsigned int blocks = div_up(n_i, 32 * IPT);
    // aim for >= 16 warps per SM
    const uint64_t want = (148ull * 16 + blocks - 1) / blocks;
    if (want >= 8 && a.Na >= 2048) direct_fast_kernel<KAHAN, 8, IPT><<<blocks, 256, 0, h->stream>>>(a);
    else if (want >= 4 && a.Na >= 1024) direct_fast_kernel<KAHAN, 4, IPT><<<blocks, 128, 0, h->stream>>>(a);
    else if (want >= 2 && a.Na >= 512) direct_fast_kernel<KAHAN, 2, IPT><<<blocks, 64, 0, h->stream>>>(a);
    else direct_fast_kernel<KAHAN, 1, IPT><<<blocks, 32, 0, h->stream>>>(a);
}

template <bool KAHAN>
void launch_fast(rebcu_handle* h, const DirectArgs& a, uint64_t n_i) {
    // two particles per lane once there are enough particles to fill the machine that way (REBOUND_B200_FAST_IPT=1|2 forces)
    static const int forced = [] { const char* e = getenv("REBOUND_B200_FAST_IPT"); return e ? atoi(e) : 0; }();
    // (C1, N = 16384: 0.37 ms per evaluation with two particles per lane, 0.44 with one -- one broadcast LDS pair per
    // pair term keeps the shared-memory pipe as busy as the FP64 pipe)
    const bool two = forced ? forced == 2 : n_i >= 148ull * 64;
    if (two) launch_fast_ipt<KAHAN, 2>(h, a, n_i); else launch_fast_ipt<KAHAN, 1>(h, a, n_i);
}

}  // namespace

int zero_acceleration(rebcu_handle* h) {
    uint64_t b, e; engine_shard(h, &b, &e);
    if (e <= b) return REBCU_OK;
    LaunchScope ls(h, TC_KICKDRIFT);
    zero3_kernel<<<div_up(e - b, 256), 256, 0, h->stream>>>(h->f(F_AX), h->f(F_AY), h->f(F_AZ), b, e);
    CU_TRY(h, cudaGetLastError());
    return REBCU_OK;
}

int direct_gravity(rebcu_handle* h, const rebcu_config* c) {
    const uint64_t N = h->N;
    if (N == 0) return REBCU_OK;
    DirectArgs a;
    a.x = h->f(F_X); a.y = h->f(F_Y); a.z = h->f(F_Z); a.m = h->f(F_M);
    a.ax = h->f(F_AX); a.ay = h->f(F_AY); a.az = h->f(F_AZ);
    a.N = N;
    a.Na = (c->N_active == REBCU_SIZE_MAX) ? N : (c->N_active < N ? c->N_active : N);
    engine_shard(h, &a.i_begin, &a.i_end);
    a.type = c->testparticle_type;
    a.terms = c->gravity_ignore_terms;
    a.G = c->G;
    a.soft2 = c->softening * c->softening;
    a.ghosts = h->ghosts_dev;
    const bool kahan = c->gravity == REBCU_GRAVITY_COMPENSATED;
    a.use_ghosts = kahan ? 0 : 1;
    a.windowed = strict_window_ok(c->G) ? 1 : 0;
    int ghosts_n = 1;
    rebcu_vec6d gb0 = {0, 0, 0, 0, 0, 0};
    if (!kahan) {
        GhostShifts g;
        engine_ghost_shifts(c, c->N_ghost_x, c->N_ghost_y, c->N_ghost_z, &g);
        int err = engine_upload_ghosts(h, &g);
        if (err) return err;
        ghosts_n = g.n; gb0 = g.gb[0];
    }
    a.csx = a.csy = a.csz = nullptr;
    h->gravity_cs_valid = false;
    if (kahan) {
        if (h->gravity_cs_cap < h->cap) {
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            cudaFree(h->gravity_cs); h->gravity_cs = nullptr; h->gravity_cs_cap = 0;
            CU_TRY(h, cudaMalloc(&h->gravity_cs, 3 * h->cap * sizeof(double)));
            h->gravity_cs_cap = h->cap;
        }
        a.csx = h->gravity_cs; a.csy = h->gravity_cs + h->gravity_cs_cap; a.csz = h->gravity_cs + 2 * h->gravity_cs_cap;
        h->gravity_cs_valid = true;
    }
    // testparticle_type 1 with few massive particles among many: the massive rows see all N sources and go through
    // the row path; the regular kernels then only handle the test particles (sources = the massive ones).
    const bool fast_mode = c->mode == REBCU_MODE_FAST;
    // (FAST: any number of massive rows up to 8192, in tiles of ROW_MAX rows; STRICT: up to ROW_MAX rows, more go to the
    // split kernel below)
    if (a.type && a.Na < N && a.Na >= 1 && a.Na <= (uint64_t)(fast_mode ? 8192 : ROW_MAX) && N >= 4096 && (kahan || ghosts_n == 1)) {
        const uint64_t rows_begin = a.i_begin, rows_end = a.i_end < a.Na ? a.i_end : a.Na;
        for (uint64_t r0 = rows_begin; r0 < rows_end; r0 += ROW_MAX) {
            const uint64_t r1 = (r0 + ROW_MAX < rows_end) ? r0 + ROW_MAX : rows_end;
            RowArgs R;
            R.a = a; R.row0 = r0; R.n_rows = (int)(r1 - r0);
            R.gbx = gb0.x; R.gby = gb0.y; R.gbz = gb0.z;
            const bool fast = fast_mode;
            // chunk of sources per pass: the term buffer (24 B per row and source) stays L2 resident (<= 48 MB), so the
            // ordered sum reads it back at L2 latency
            uint64_t chunk = (48ull << 20) / (8ull * 3 * (uint64_t)R.n_rows);
            if (const char* e = getenv("REBOUND_B200_ROWCHUNK")) { const uint64_t v = strtoull(e, nullptr, 10); if (v >= 1024) chunk = v; }   // tests: force several chunks
            chunk = (chunk / 1024) * 1024;
            if (chunk > N) chunk = ((N + 1023) / 1024) * 1024;
            const uint64_t need = (fast ? (uint64_t)R.n_rows * 3 * ((N + 1023) / 1024) : (uint64_t)R.n_rows * 3 * (chunk + 64)) + 64;
            if (h->row_cap < need + 6 * ROW_MAX) {
                CU_TRY(h, cudaStreamSynchronize(h->stream));
                cudaFree(h->row_buf); h->row_buf = nullptr; h->row_cap = 0;
                CU_TRY(h, cudaMalloc(&h->row_buf, (need + 6 * ROW_MAX) * sizeof(double)));
                h->row_cap = need + 6 * ROW_MAX;
            }
            R.state = h->row_buf; R.terms = h->row_buf + 6 * ROW_MAX; R.chunk_cap = chunk;
            if (fast) {
                const unsigned n_cta = (unsigned)((N + 1023) / 1024);
                R.j0 = 0; R.j1 = N; R.first = R.last = 1;
                LaunchScope ls(h, TC_DIRECT, 2);
                if (kahan) { row_fast_kernel<true><<<n_cta, 128, 0, h->stream>>>(R, R.terms, n_cta); row_fast_final_kernel<true><<<div_up(R.n_rows * 3, 128), 128, 0, h->stream>>>(R, R.terms, n_cta); }
                else { row_fast_kernel<false><<<n_cta, 128, 0, h->stream>>>(R, R.terms, n_cta); row_fast_final_kernel<false><<<div_up(R.n_rows * 3, 128), 128, 0, h->stream>>>(R, R.terms, n_cta); }
            } else {
                for (uint64_t j0 = 0; j0 < N; j0 += chunk) {
                    R.j0 = j0; R.j1 = (j0 + chunk < N) ? j0 + chunk : N;
                    R.first = (j0 == 0); R.last = (R.j1 == N);
                    LaunchScope ls(h, TC_DIRECT, 2);
                    if (kahan) { row_terms_kernel<true><<<div_up(R.j1 - R.j0, 128), 128, 0, h->stream>>>(R); row_sum_kernel<true><<<div_up(R.n_rows, ROWS_PER_WARP), 32, 0, h->stream>>>(R); }
                    else { row_terms_kernel<false><<<div_up(R.j1 - R.j0, 128), 128, 0, h->stream>>>(R); row_sum_kernel<false><<<div_up(R.n_rows, ROWS_PER_WARP), 32, 0, h->stream>>>(R); }
                }
            }
            CU_TRY(h, cudaGetLastError());
        }
        if (a.i_begin < a.Na) a.i_begin = a.Na < a.i_end ? a.Na : a.i_end;
    } else if (a.type && a.Na < N && N >= 4096 && a.i_begin < a.Na && c->mode != REBCU_MODE_FAST && a.windowed) {
        // More massive rows than the row path takes: each still sees all N sources, which one thread per target would
        // walk serially.  The producer/adder split kernel spreads a row block's sources over three producer warps and
        // leaves one dependent add per source on the adder (same bits); the test particles follow separately.
        DirectArgs ar = a;
        ar.i_end = a.i_end < a.Na ? a.i_end : a.Na;
        {
            LaunchScope ls(h, TC_DIRECT);
            if (kahan) direct_strict_split_kernel<true, 4, 8><<<div_up(ar.i_end - ar.i_begin, 32), 128, 0, h->stream>>>(ar);
            else direct_strict_split_kernel<false, 4, 8><<<div_up(ar.i_end - ar.i_begin, 32), 128, 0, h->stream>>>(ar);
        }
        CU_TRY(h, cudaGetLastError());
        a.i_begin = ar.i_end;
    }
    const uint64_t n_i = a.i_end - a.i_begin;
    if (n_i == 0) return REBCU_OK;
    {
        LaunchScope ls(h, TC_DIRECT);
        if (c->mode == REBCU_MODE_FAST) { if (kahan) launch_fast<true>(h, a, n_i); else launch_fast<false>(h, a, n_i); }
        else { if (kahan) launch_strict<true>(h, a, n_i); else launch_strict<false>(h, a, n_i); }
    }
    CU_TRY(h, cudaGetLastError());
    return REBCU_OK;
}

namespace {
__global__ void __launch_bounds__(256) interleave3_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                          double* __restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) { out[3 * i] = x[i]; out[3 * i + 1] = y[i]; out[3 * i + 2] = z[i]; }
}
}  // namespace

// r->gravity_cs of the last COMPENSATED force evaluation as struct reb_vec3d[N] (x,y,z interleaved).
extern "C" int rebcu_download_gravity_cs(rebcu_handle* h, double* out_xyz, uint64_t N) {
    GROUP_UNSUPPORTED(h, "rebcu_download_gravity_cs");
    if (!h->resident || !h->gravity_cs_valid) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no compensation terms: the last force evaluation was not REB_GRAVITY_COMPENSATED");
    if (N < h->N) return rebcu_fail(h, REBCU_ERR_CAPACITY, "gravity_cs buffer too small");
    if (h->N == 0) return REBCU_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    double* stage = (double*)h->aos;            // AoS staging block: 112 B per particle >= 24 B needed
    {
        LaunchScope ls(h, TC_PACK);
        interleave3_kernel<<<div_up(h->N, 256), 256, 0, h->stream>>>(h->gravity_cs, h->gravity_cs + h->gravity_cs_cap,
                                                                     h->gravity_cs + 2 * h->gravity_cs_cap, stage, h->N);
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(out_xyz, stage, 3 * h->N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return REBCU_OK;
}
