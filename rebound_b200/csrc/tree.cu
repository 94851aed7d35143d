// tree.cu -- octree build, multipole (mass / centre-of-mass) pass and Barnes-Hut walk.
//
// Replaces reb_tree_construct / reb_tree_add_particle_to_cell (src/tree.c:254-271, 80-134),
// reb_tree_calculate_gravity_data (src/tree.c:147-229), reb_tree_calculate_acceleration_for_particle
// (src/tree.c:275-328), reb_tree_delete (src/tree.c:231-252, nothing to do here: arena buffers) and
// reb_gravity_tree_calculate_acceleration (src/gravity.c:47-106).
//
// The reference inserts particles one by one into a pointer octree with one particle per leaf and
// unbounded depth.  Its topology does not depend on the insertion order: a cell exists iff it
// contains >= 1 particle and its parent contains >= 2.  So the identical tree is built in bulk:
//   1. keys      per particle: root box (particle.c:119-126) and the octant path obtained by REPLAYING
//                the reference's `p < centre` comparisons with its exact centre recurrence
//                (centre = parent centre +- w/4, tree.c:99-102, which rounds when root_size is not a
//                power of two) -- octant bit 1 = low side, so ascending key order = the reference's
//                depth-first octant order.  64-bit key = root box | 3 bits per level.
//   2. sort      stable LSD radix sort of (key, index), hand-written (primitives.cuh).
//   3. ties      particles agreeing in all key levels are ordered (and duplicates detected,
//                tree.c:119-123) by continuing the exact descent pairwise.
//   4. lcp       common path length of sorted neighbours; cells opened at sorted position k are the
//                internal cells at depths (lcp[k], lcp[k+1]] plus the leaf at 1+max(lcp[k],lcp[k+1]).
//   5. cells     emitted in depth-first pre-order (= sorted order) with the reference's geometry,
//                pt = particle index or -(particle count), and skip = first cell after the subtree.
//   6. moments   bottom-up with per-cell child counters; each parent combines its children in octant
//                order with the reference's expression (tree.c:162-179), so m, mx, my, mz are bitwise.
//   7. walk      one thread per particle in key order (spatially coherent warps), stackless pre-order
//                traversal with per-lane opening decisions: every particle sees exactly the
//                reference's interaction list in the reference's order (ghost boxes, root boxes,
//                octants ascending) => bitwise accelerations in STRICT mode.
// Bounds: build steps are HBM-streaming (keys 24 B in / 12 B out per particle, sort ~8 passes x 24 B,
// cells ~1.5/particle x 80 B out); the walk is FP64-pipe / L2-latency bound.
#include "engine.cuh"
#include "strict_math.cuh"
#include "fast_math.cuh"
#include "primitives.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int MAX_DESCENT = 1200;   // below 2^-1074 x root_size a cell has width zero (tree.c:107-111)
constexpr int W_TABLE = 64;
constexpr int TIE_RUN_MAX = 16;     // longest run of equal key prefixes the prefix-sorted build accepts

struct TreeParams {
    double root_size;
    int Nx, Ny, Nz;
    int L0;        // octant levels stored in the key
    int prefix;    // L0 was cut short to save sort passes: tie runs longer than TIE_RUN_MAX abort the build (flags[4])
    int rbits;     // bits of root box index above the path
    uint64_t n;
    // sharded build (tree_build_sharded): neighbours in different buckets (key >> bshift) count as sharing `clamp` levels,
    // and the cells of a bucket's subtree are emitted at their GLOBAL pre-order index = local offset + bdelta[bucket]
    int clamp;     // -1: plain build
    int bshift;
    const long long* bdelta;
};

struct Geo { double cx, cy, cz, w; };

// Root box of a particle and that root cell's centre: particle.c:119-126 and tree.c:87-97.
__device__ __forceinline__ int root_cell(const TreeParams& P, double x, double y, double z, Geo& g) {
    const double rs = P.root_size;
    const double bx = s_mul(rs, (double)P.Nx), by = s_mul(rs, (double)P.Ny), bz = s_mul(rs, (double)P.Nz);
    const int i = ((int)floor(s_div(s_add(x, s_div(bx, 2.)), rs)) + P.Nx) % P.Nx;
    const int j = ((int)floor(s_div(s_add(y, s_div(by, 2.)), rs)) + P.Ny) % P.Ny;
    const int k = ((int)floor(s_div(s_add(z, s_div(bz, 2.)), rs)) + P.Nz) % P.Nz;
    g.w = rs;
    g.cx = s_add(s_div(-bx, 2.), s_mul(rs, s_add(0.5, (double)i)));
    g.cy = s_add(s_div(-by, 2.), s_mul(rs, s_add(0.5, (double)j)));
    g.cz = s_add(s_div(-bz, 2.), s_mul(rs, s_add(0.5, (double)k)));
    return (k * P.Ny + j) * P.Nx + i;
}

// Octant of (x,y,z) in cell g (tree.c:136-142), then g becomes that child cell (tree.c:99-102).
__device__ __forceinline__ int descend(Geo& g, double x, double y, double z) {
    const int o = (x < g.cx ? 1 : 0) | (y < g.cy ? 2 : 0) | (z < g.cz ? 4 : 0);
    g.w = s_div(g.w, 2.);
    const double q = s_div(g.w, 2.);
    g.cx = s_add(g.cx, (o & 1) ? -q : q);
    g.cy = s_add(g.cy, (o & 2) ? -q : q);
    g.cz = s_add(g.cz, (o & 4) ? -q : q);
    return o;
}

// flags: [0] min index outside box, [1] min index non-finite, [2] min (larger index) of a duplicate pair,
//        [3] cell size zero
__global__ void __launch_bounds__(256) key_kernel(TreeParams P, const double* __restrict__ x, const double* __restrict__ y,
                                                  const double* __restrict__ z, uint64_t* __restrict__ keys,
                                                  uint32_t* __restrict__ idx, int* flags) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= P.n) return;
    const double px = x[i], py = y[i], pz = z[i];
    idx[i] = (uint32_t)i;
    const double rs = P.root_size;
    // tree.c:265-268 then tree.c:66-69
    if (fabs(px) > s_div(s_mul(rs, (double)P.Nx), 2.) || fabs(py) > s_div(s_mul(rs, (double)P.Ny), 2.) ||
        fabs(pz) > s_div(s_mul(rs, (double)P.Nz), 2.)) { atomicMin(&flags[0], (int)i); keys[i] = ~0ull; return; }
    if (!isfinite(px) || !isfinite(py) || !isfinite(pz)) { atomicMin(&flags[1], (int)i); keys[i] = ~0ull; return; }
    Geo g;
    const int rb = root_cell(P, px, py, pz, g);
    uint64_t key = (uint64_t)rb;
    for (int l = 0; l < P.L0; l++) key = (key << 3) | (uint64_t)descend(g, px, py, pz);
    keys[i] = key;
}

// Number of common octant levels of two particles known to share all L0 key levels (or -2 if identical
// coordinates); *less = a sorts before b.
__device__ int deep_common(const TreeParams& P, double ax, double ay, double az, double bx, double by, double bz,
                           bool* less, int* flags) {
    if (ax == bx && ay == by && az == bz) { *less = false; return -2; }
    Geo g;
    root_cell(P, ax, ay, az, g);
    int c = 0;
    for (; c < MAX_DESCENT; c++) {
        Geo ga = g;
        const int oa = descend(ga, ax, ay, az);
        Geo gb = g;
        const int ob = descend(gb, bx, by, bz);
        if (oa != ob) { *less = oa < ob; return c; }
        g = ga;
        if (!(g.w > 0.0)) { atomicMin(&flags[3], 1); break; }
    }
    *less = false;
    return c;
}

// Orders runs of equal keys by the deeper path (insertion sort by the thread owning the run start).
__global__ void __launch_bounds__(256) tie_kernel(TreeParams P, const uint64_t* __restrict__ keys, uint32_t* __restrict__ perm,
                                                  const double* __restrict__ x, const double* __restrict__ y,
                                                  const double* __restrict__ z, int* flags) {
    const uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (k + 1 >= P.n) return;
    if (keys[k] != keys[k + 1]) return;
    if (k > 0 && keys[k - 1] == keys[k]) return;     // not the run start
    uint64_t e = k + 1;
    while (e + 1 < P.n && keys[e + 1] == keys[k]) {
        e++;
        if (P.prefix && e - k >= TIE_RUN_MAX) { atomicMin(&flags[4], 1); return; }      // too clustered for this prefix
    }
    for (uint64_t a = k + 1; a <= e; a++) {
        const uint32_t pa = perm[a];
        const double ax = x[pa], ay = y[pa], az = z[pa];
        uint64_t b = a;
        while (b > k) {
            const uint32_t pb = perm[b - 1];
            bool less;
            const int c = deep_common(P, ax, ay, az, x[pb], y[pb], z[pb], &less, flags);
            if (c == -2) { atomicMin(&flags[2], (int)max(pa, pb)); break; }
            if (!less) break;
            perm[b] = pb;
            b--;
        }
        perm[b] = pa;
    }
}

// lcp[k] = common octant levels of sorted particles k-1 and k (same root box), -1 across root boxes and
// at both ends (k = 0 and k = n).
__global__ void __launch_bounds__(256) lcp_kernel(TreeParams P, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                                                  const double* __restrict__ x, const double* __restrict__ y,
                                                  const double* __restrict__ z, int32_t* __restrict__ lcp, int* flags) {
    const uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (k > P.n) return;
    if (k == 0 || k == P.n) { lcp[k] = P.clamp; return; }
    const uint64_t a = keys[k - 1], b = keys[k];
    int c;
    if (a != b) {
        const uint64_t d = a ^ b;
        const int hb = 63 - __clzll((long long)d);
        c = (hb >= 3 * P.L0) ? -1 : (P.L0 - 1 - hb / 3);
    } else {
        const uint32_t pa = perm[k - 1], pb = perm[k];
        bool less;
        c = deep_common(P, x[pa], y[pa], z[pa], x[pb], y[pb], z[pb], &less, flags);
        if (c == -2) { atomicMin(&flags[2], (int)max(pa, pb)); c = P.L0; }
        else c += 0;
    }
    if (c > 480) atomicMin(&flags[5], 1);            // beyond the depth the sharded build carries (normal-range w^2 only)
    lcp[k] = c > P.clamp ? c : P.clamp;
}

__global__ void __launch_bounds__(256) count_kernel(uint64_t n, const int32_t* __restrict__ lcp, uint32_t* __restrict__ cnt) {
    const uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const int d = lcp[k + 1] - lcp[k];
    cnt[k] = (uint32_t)((d > 0 ? d : 0) + 1);
}

struct CellArrays {
    double4* pos;    // mx,my,mz,m
    double4* geo;    // x,y,z,w
    int4* meta;      // pt, skip, depth, rootbox
    int32_t* parent;
    uint32_t* ready; // children still missing
    int2* meta2;     // (leaf: pt | internal: -(depth+1), skip) -- the 8 bytes the gravity walk reads per visit
    double* quad;    // null, or mxx mxy mxz myy myz mzz as six arrays of quad_stride cells (QUADRUPOLE, tree.c:148-198)
    uint64_t quad_stride;
};

// Emits the cells opened at sorted position k (geometry, pt, skip, depth, rootbox; leaf moments).
__global__ void __launch_bounds__(128) emit_kernel(TreeParams P, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                                                   const int32_t* __restrict__ lcp, const uint32_t* __restrict__ off,
                                                   const double* __restrict__ x, const double* __restrict__ y,
                                                   const double* __restrict__ z, const double* __restrict__ m, CellArrays C) {
    const uint64_t k = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    if (k >= P.n) return;
    const int lo = lcp[k], hi = lcp[k + 1];
    const int n_int = hi > lo ? hi - lo : 0;
    const int leaf_depth = 1 + (lo > hi ? lo : hi);
    const uint64_t key = keys[k];
    const long long delta = P.bdelta ? P.bdelta[key >> P.bshift] : 0ll;
    const uint32_t base = (uint32_t)((long long)off[k] + delta);
    const uint32_t p = perm[k];
    const double px = x[p], py = y[p], pz = z[p];
    Geo g;
    const int rb = root_cell(P, px, py, pz, g);
    for (int d = 0; d <= leaf_depth; d++) {
        if (d > 0) descend(g, px, py, pz);
        const bool internal = (d > lo && d <= hi);
        if (!internal && d != leaf_depth) continue;
        const uint32_t c = internal ? base + (uint32_t)(d - lo - 1) : base + (uint32_t)n_int;
        C.geo[c] = make_double4(g.cx, g.cy, g.cz, g.w);
        if (!internal) {
            C.meta[c] = make_int4((int)p, (int)c + 1, d, rb);
            C.meta2[c] = make_int2((int)p, (int)c + 1);
            C.pos[c] = make_double4(px, py, pz, m[p]);      // tree.c:201-205
            C.ready[c] = 0;
            continue;
        }
        // last sorted position whose path shares the first d levels with particle k
        uint64_t e;
        if (d <= P.L0) {
            const int sh = 3 * (P.L0 - d);
            const uint64_t pref = key >> sh;
            uint64_t step = 1, loj = k + 1;         // loj is known to be inside (lcp[k+1] >= d)
            uint64_t hij = P.n;                      // first position known to be outside (exclusive bound)
            while (loj + step < P.n && (keys[loj + step] >> sh) == pref) { loj += step; step <<= 1; }
            if (loj + step < hij) hij = loj + step;
            while (loj + 1 < hij) {                   // invariant: loj inside, hij outside or n
                const uint64_t mid = loj + (hij - loj) / 2;
                if ((keys[mid] >> sh) == pref) loj = mid; else hij = mid;
            }
            e = loj;
        } else {
            e = k + 1;
            while (e + 1 < P.n && lcp[e + 1] >= d) e++;
        }
        const int skip = (int)((long long)off[e + 1] + delta);
        C.meta[c] = make_int4(-(int)(e - k + 1), skip, d, rb);
        C.meta2[c] = make_int2(-(d + 1), skip);
    }
}

// Every internal cell adopts its children (the first child follows it, the next one starts where the
// previous subtree ends) and counts them.
__global__ void __launch_bounds__(256) adopt_kernel(uint64_t c_begin, uint64_t n_cells, CellArrays C) {
    const uint64_t c = c_begin + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= c_begin + n_cells) return;
    const int4 mt = C.meta[c];
    if (mt.z == 0) C.parent[c] = -1;
    if (mt.x >= 0) return;
    uint32_t nch = 0;
    for (int ch = (int)c + 1; ch < mt.y; ch = C.meta[ch].y) { C.parent[ch] = (int)c; nch++; }
    C.ready[c] = nch;
}

// Bottom-up moments: the thread that delivers the last child of a cell computes that cell
// (tree.c:156-179: children in octant order, sum of mx*m, then divide by the total mass).
__global__ void __launch_bounds__(256) moment_kernel(uint64_t c_begin, uint64_t n_cells, CellArrays C) {
    const uint64_t c0 = c_begin + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (c0 >= c_begin + n_cells) return;
    if (C.meta[c0].x < 0) return;            // start from leaves only
    int c = C.parent[c0];
    while (c >= 0) {
        __threadfence();
        if (atomicSub(&C.ready[c], 1u) != 1u) return;    // siblings still pending
        __threadfence();
        const int end = C.meta[c].y;
        double mm = 0., mx = 0., my = 0., mz = 0.;
        for (int ch = c + 1; ch < end; ch = C.meta[ch].y) {
            const volatile double4* q = (const volatile double4*)&C.pos[ch];
            const double dx = q->x, dy = q->y, dz = q->z, dm = q->w;
            mx = s_add(mx, s_mul(dx, dm));
            my = s_add(my, s_mul(dy, dm));
            mz = s_add(mz, s_mul(dz, dm));
            mm = s_add(mm, dm);
        }
        if (mm > 0) { mx = s_div(mx, mm); my = s_div(my, mm); mz = s_div(mz, mm); }
        volatile double4* o = (volatile double4*)&C.pos[c];
        o->x = mx; o->y = my; o->z = mz; o->w = mm;
        if (C.quad) {
            // tree.c:180-197 (Hernquist 1987): children in octant order, each term  d.mxx + d_m*(3*qx*qx - qr2)  etc.
            volatile double* Q = C.quad;
            const uint64_t S = C.quad_stride;
            double mxx = 0., mxy = 0., mxz = 0., myy = 0., myz = 0.;
            for (int ch = c + 1; ch < end; ch = C.meta[ch].y) {
                const volatile double4* q = (const volatile double4*)&C.pos[ch];
                const double dm = q->w;
                const double qx = s_sub(q->x, mx), qy = s_sub(q->y, my), qz = s_sub(q->z, mz);
                const double qr2 = s_add(s_add(s_mul(qx, qx), s_mul(qy, qy)), s_mul(qz, qz));
                mxx = s_add(mxx, s_add(Q[0 * S + ch], s_mul(dm, s_sub(s_mul(s_mul(3., qx), qx), qr2))));
                mxy = s_add(mxy, s_add(Q[1 * S + ch], s_mul(s_mul(s_mul(dm, 3.), qx), qy)));
                mxz = s_add(mxz, s_add(Q[2 * S + ch], s_mul(s_mul(s_mul(dm, 3.), qx), qz)));
                myy = s_add(myy, s_add(Q[3 * S + ch], s_mul(dm, s_sub(s_mul(s_mul(3., qy), qy), qr2))));
                myz = s_add(myz, s_add(Q[4 * S + ch], s_mul(s_mul(s_mul(dm, 3.), qy), qz)));
            }
            Q[0 * S + c] = mxx; Q[1 * S + c] = mxy; Q[2 * S + c] = mxz; Q[3 * S + c] = myy; Q[4 * S + c] = myz;
            Q[5 * S + c] = s_sub(-mxx, myy);            // node->mzz = -node->mxx - node->myy
        }
        c = C.parent[c];
    }
}

__global__ void __launch_bounds__(256) export_kernel(uint64_t n_cells, CellArrays C, rebcu_treecell* out) {
    const uint64_t c = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= n_cells) return;
    const double4 g = C.geo[c], p = C.pos[c];
    const int4 mt = C.meta[c];
    rebcu_treecell t;
    t.x = g.x; t.y = g.y; t.z = g.z; t.w = g.w; t.m = p.w; t.mx = p.x; t.my = p.y; t.mz = p.z;
    t.pt = mt.x; t.skip = mt.y; t.depth = mt.z; t.rootbox = mt.w;
    out[c] = t;
}

// ---- Barnes-Hut walk -----------------------------------------------------------------------------
// 256-bit load of one (mx,my,mz,m) record: one LDG.E.256 instead of two LDG.E.128, which halves the L1
// wavefronts of the divergent per-thread walk (its measured bottleneck).
__device__ __forceinline__ double4 ld_pos256(const double4* p) {
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

struct WalkArgs {
    const double4* pos; const int4* meta; const int2* meta2; uint64_t n_cells;
    const uint32_t* perm; const uint32_t* list; uint64_t n_work;     // work item t -> sorted position (list[t] or t)
    const double* x; const double* y; const double* z;
    double* ax; double* ay; double* az;
    const GhostShifts* ghosts;
    double G, soft2, theta2;
    double w2[W_TABLE];      // squared cell width by depth
    double root_size;
    int windowed;            // G inside the window of the branch-free sqrt/divide (strict_math.cuh)
    const double* quad; uint64_t quad_stride;     // QUADRUPOLE builds: six arrays mxx mxy mxz myy myz mzz, else null
    const double4* rec; const double* m;          // traversal records (mx,my,mz,meta) + masses (records walk), else null
    uint32_t w2_lo;                               // low word of every (normal) squared cell width, see walk_pack_kernel
    int gw_stack;                                 // traversal stack entries the group walk may use (<= GW_STACK)
    unsigned int gw_abort;                        // list entries after which a group gives up and its lanes walk on their own
    // sharded runs with the native exchange: a rank walks the sorted positions [k_begin, k_begin + n_work) -- a compact
    // region of space -- and leaves the accelerations in SORTED order (three arrays of sorted_stride doubles); they are
    // all-gathered and scattered to the index blocks afterwards (tree_gravity)
    uint64_t k_begin;
    double* sorted_out; uint64_t sorted_stride;
};

__device__ __forceinline__ void store_acc(const WalkArgs& a, uint32_t self, uint64_t k, double sx, double sy, double sz) {
    if (a.sorted_out) { a.sorted_out[k] = sx; a.sorted_out[a.sorted_stride + k] = sy; a.sorted_out[2 * a.sorted_stride + k] = sz; }
    else { a.ax[self] = sx; a.ay[self] = sy; a.az[self] = sz; }
}

// MODE 0: strict with the branch-free windowed sqrt/divide (returns the running window key),
//      1: FAST (FMA + rsqrt), 2: strict with the generic __dsqrt_rn/__ddiv_rn.
template <int MODE>
__device__ __forceinline__ unsigned walk_one(const WalkArgs& a, uint32_t self, double px, double py, double pz,
                                             double& sx, double& sy, double& sz) {
    sx = sy = sz = 0.;
    unsigned bad = 0;
    const double negG = -a.G;
    const int ngb = a.ghosts->n;
    const int n_cells = (int)a.n_cells;
    for (int g = 0; g < ngb; g++) {
        // gravity.c:93-97: shifted position = ghost box offset + particle position
        const double gx = s_add(a.ghosts->gb[g].x, px), gy = s_add(a.ghosts->gb[g].y, py), gz = s_add(a.ghosts->gb[g].z, pz);
        int c = 0;
        while (c < n_cells) {
            const double4 q = ld_pos256(a.pos + c);
            const int2 mt = a.meta2[c];               // x: particle index (leaf) or -(depth+1); y: skip
            const double dx = s_sub(gx, q.x), dy = s_sub(gy, q.y), dz = s_sub(gz, q.z);
            const double r2 = s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz));
            if (mt.x < 0) {
                const int depth = -mt.x - 1;
                double w2;
                if (depth < W_TABLE) w2 = a.w2[depth];
                else { double w = a.root_size; for (int d = 0; d < depth; d++) w = s_div(w, 2.); w2 = s_mul(w, w); }
                if (w2 > s_mul(a.theta2, r2)) { c++; continue; }          // tree.c:284: open the cell
            } else if ((uint32_t)mt.x == self) { c = mt.y; continue; }    // tree.c:311
            if (MODE == 1) {
                const double p = fast_m_over_r3(r2 + a.soft2, negG * q.w);
                sx = fma(p, dx, sx); sy = fma(p, dy, sy); sz = fma(p, dz, sz);
            } else if (MODE == 0) {
                const double rs2 = s_add(r2, a.soft2);
                bad = max(bad, strict_window_key(rs2));
                const double r = fsqrt_rn_w(rs2);
                const double p = s_mul(fdiv_rn_w(negG, s_mul(s_mul(r, r), r)), q.w);      // tree.c:292,313
                sx = s_add(sx, s_mul(p, dx)); sy = s_add(sy, s_mul(p, dy)); sz = s_add(sz, s_mul(p, dz));
            } else {
                const double r = s_sqrt(s_add(r2, a.soft2));
                const double p = s_mul(s_div(negG, s_mul(s_mul(r, r), r)), q.w);
                sx = s_add(sx, s_mul(p, dx)); sy = s_add(sy, s_mul(p, dy)); sz = s_add(sz, s_mul(p, dz));
            }
            c = mt.y;
        }
    }
    return bad;
}

__device__ __forceinline__ void walk_generic(const WalkArgs& a, uint32_t self, double px, double py, double pz,
                                          double& sx, double& sy, double& sz) {
    walk_one<2>(a, self, px, py, pz, sx, sy, sz);
}

template <bool FAST>
__global__ void __launch_bounds__(128) walk_kernel(const WalkArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    if (t >= a.n_work) return;
    const uint64_t k = a.list ? a.list[t] : a.k_begin + t;
    const uint32_t self = a.perm[k];
    const double px = a.x[self], py = a.y[self], pz = a.z[self];
    double sx, sy, sz;
    unsigned wkey = STRICT_WINDOW_LIMIT;
    if (FAST || a.windowed) wkey = walk_one<FAST ? 1 : 0>(a, self, px, py, pz, sx, sy, sz);
    if (!FAST && wkey >= STRICT_WINDOW_LIMIT) walk_generic(a, self, px, py, pz, sx, sy, sz);
    store_acc(a, self, k, sx, sy, sz);
}

// ---- shared pieces of the walks ----------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void interact(const WalkArgs& a, double negG, double dx, double dy, double dz, double r2, double m,
                                         double& sx, double& sy, double& sz, unsigned& bad) {
    if (MODE == 1) {
        const double p = fast_m_over_r3(r2 + a.soft2, negG * m);
        sx = fma(p, dx, sx); sy = fma(p, dy, sy); sz = fma(p, dz, sz);
    } else if (MODE == 0) {
        const double rs2 = s_add(r2, a.soft2);
        bad = max(bad, strict_window_key(rs2));
        const double r = fsqrt_rn_w(rs2);
        const double p = s_mul(fdiv_rn_w(negG, s_mul(s_mul(r, r), r)), m);      // tree.c:292,313
        sx = s_add(sx, s_mul(p, dx)); sy = s_add(sy, s_mul(p, dy)); sz = s_add(sz, s_mul(p, dz));
    } else {
        const double r = s_sqrt(s_add(r2, a.soft2));
        const double p = s_mul(s_div(negG, s_mul(s_mul(r, r), r)), m);
        sx = s_add(sx, s_mul(p, dx)); sy = s_add(sy, s_mul(p, dy)); sz = s_add(sz, s_mul(p, dz));
    }
}

// out-of-line copy for the rare branch of the records walk (keeps the hot loop short)
__device__ __noinline__ double cell_w2_rare(double root_size, int depth) {
    double w = root_size;
    for (int d = 0; d < depth; d++) w = s_div(w, 2.);
    return s_mul(w, w);
}

__device__ __forceinline__ double cell_w2(const WalkArgs& a, int depth) {
    if (depth < W_TABLE) return a.w2[depth];
    double w = a.root_size;
    for (int d = 0; d < depth; d++) w = s_div(w, 2.);
    return s_mul(w, w);
}

// ---- walk, default variant ("advance to the next accepted cell, then interact", 32-byte traversal records) ------
// Same interaction list in the same order as walk_one (hence the same bits); what changes is how the warp spends
// its issue slots and its loads.
//  * walk_one evaluates one visited cell per loop trip, so a warp pays the ~35 FP64 instructions of the interaction
//    whenever ANY lane accepts its cell while the lanes that open theirs idle.  Here every lane first runs down its
//    own path (one load, 8 FP64 instructions and a compare per visited cell) until it holds an ACCEPTED cell, the
//    lanes reconverge, and the interaction is evaluated once for all of them (25.4 instead of 23.8 active lanes per
//    instruction, ncu).
//  * The record the traversal needs next is requested BEFORE the interaction, so its latency is covered by the
//    sqrt/divide chain (without this the restructured loop is slower than walk_one: 7.0 vs 5.3 ms).
//  * Every visited cell used to cost two gathers (the 32-byte position record and the 8-byte meta record).  A
//    traversal decision needs the centre of mass and the meta word but not the mass, so the build's (mx,my,mz,m) +
//    (pt|depth, skip) arrays are repacked into (mx,my,mz,meta) + m: ONE 32-byte gather per visited cell; the mass is
//    fetched for accepted cells only and arrives while the sqrt/divide chain runs.
// Measured (B200, disc N=2^20 / 2^22, sheet N=2^20 with 25 ghost boxes): 4.93 / 23.5 / 1.85 ms against 5.28 / 24.8 /
// 1.94 ms for walk_kernel (REBOUND_B200_WALK=v1), repacking included.
// Meta word of a record: low half = tag, high half = skip.  Leaf: tag = particle index (>= 0).  Internal cell: bit 31
// set and, below it, the HIGH word of the cell's squared width w2.  Cell widths are exact halvings of root_size
// (tree.c:100), so every w2 is fl(root_size^2) scaled by a power of four: all of them share their low word (passed as
// a kernel argument) and the walk rebuilds w2 with one logic instruction instead of a table lookup by depth.  A width
// so small that w2 leaves the normal range gets the marker 0x80000000 and the walk recomputes it from the depth.
// Tag of an internal cell at `depth` (see above): bit 31 | high word of its squared width, or the bare marker.
__device__ __forceinline__ unsigned int internal_tag(const WalkArgs& a, int depth) {
    const double w2 = cell_w2(a, depth);
    const unsigned int hi = (unsigned int)__double2hiint(w2), lo = (unsigned int)__double2loint(w2);
    return (lo == a.w2_lo && (hi >> 20) != 0 && hi < 0x7ff00000u) ? (0x80000000u | hi) : 0x80000000u;
}

__global__ void __launch_bounds__(256) walk_pack_kernel(uint64_t c_begin, uint64_t n_cells, const double4* __restrict__ pos, const int2* __restrict__ meta2,
                                                        double4* __restrict__ rec, double* __restrict__ m, WalkArgs a) {
    const uint64_t c = c_begin + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= c_begin + n_cells) return;
    const double4 q = pos[c];
    const int2 mt = meta2[c];
    unsigned int tag = (unsigned int)mt.x;
    if (mt.x < 0) tag = internal_tag(a, -mt.x - 1);
    const long long bits = (long long)(((unsigned long long)(unsigned int)mt.y << 32) | (unsigned long long)tag);
    rec[c] = make_double4(q.x, q.y, q.z, __longlong_as_double(bits));
    m[c] = q.w;
}

template <int MODE>
__device__ __forceinline__ unsigned walk_one_rec(const WalkArgs& a, uint32_t self, double px, double py, double pz,
                                                double& sx, double& sy, double& sz) {
    sx = sy = sz = 0.;
    unsigned bad = 0;
    const double negG = -a.G;
    const int ngb = a.ghosts->n;
    const int n_cells = (int)a.n_cells;
    for (int g = 0; g < ngb; g++) {
        const double gx = s_add(a.ghosts->gb[g].x, px), gy = s_add(a.ghosts->gb[g].y, py), gz = s_add(a.ghosts->gb[g].z, pz);
        int c = 0;
        double4 q = make_double4(0., 0., 0., 0.);
        if (n_cells > 0) q = ld_pos256(a.rec);
        while (true) {
            double dx = 0., dy = 0., dz = 0., r2 = 0., m = 0.;
            bool found = false;
            while (c < n_cells) {
                const long long bits = __double_as_longlong(q.w);
                const int tag = (int)(unsigned int)(unsigned long long)bits, skip = (int)(unsigned int)((unsigned long long)bits >> 32);
                dx = s_sub(gx, q.x); dy = s_sub(gy, q.y); dz = s_sub(gz, q.z);
                r2 = s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz));
                const int hi = tag & 0x7fffffff;
                double w2 = __hiloint2double(hi, (int)a.w2_lo);
                if (tag < 0 && hi == 0) w2 = cell_w2_rare(a.root_size, -a.meta2[c].x - 1);    // beyond the normal range: by depth
                const bool open = tag < 0 && w2 > s_mul(a.theta2, r2);                   // tree.c:284
                const bool acc = tag < 0 ? !open : (uint32_t)tag != self;               // tree.c:311
                const int c0 = c;
                c = open ? c + 1 : skip;
                if (c < n_cells) q = ld_pos256(a.rec + c);
                if (acc) { m = a.m[c0]; found = true; break; }
            }
            if (!found) break;
            interact<MODE>(a, negG, dx, dy, dz, r2, m, sx, sy, sz, bad);
        }
    }
    return bad;
}

template <bool FAST>
__global__ void __launch_bounds__(128) walk_rec_kernel(const WalkArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    if (t >= a.n_work) return;
    const uint64_t k = a.list ? a.list[t] : a.k_begin + t;
    const uint32_t self = a.perm[k];
    const double px = a.x[self], py = a.y[self], pz = a.z[self];
    double sx, sy, sz;
    unsigned wkey = STRICT_WINDOW_LIMIT;
    if (FAST || a.windowed) wkey = walk_one_rec<FAST ? 1 : 0>(a, self, px, py, pz, sx, sy, sz);
    if (!FAST && wkey >= STRICT_WINDOW_LIMIT) walk_generic(a, self, px, py, pz, sx, sy, sz);
    store_acc(a, self, k, sx, sy, sz);
}

// Walk of a QUADRUPOLE build (tree.c:286-304): accepted internal cells add the quadrupole correction, in two
// separate additions per component exactly as the reference does; leaves are monopoles.  Generic IEEE sqrt and
// divide (three divisions per accepted cell), one thread per particle.
template <bool FAST>
__global__ void __launch_bounds__(128) walk_quad_kernel(const WalkArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    if (t >= a.n_work) return;
    const uint64_t k = a.list ? a.list[t] : a.k_begin + t;
    const uint32_t self = a.perm[k];
    const double px = a.x[self], py = a.y[self], pz = a.z[self];
    double sx = 0., sy = 0., sz = 0.;
    const double negG = -a.G;
    const int ngb = a.ghosts->n;
    const int n_cells = (int)a.n_cells;
    const uint64_t S = a.quad_stride;
    for (int g = 0; g < ngb; g++) {
        const double gx = s_add(a.ghosts->gb[g].x, px), gy = s_add(a.ghosts->gb[g].y, py), gz = s_add(a.ghosts->gb[g].z, pz);
        int c = 0;
        while (c < n_cells) {
            const double4 q = ld_pos256(a.pos + c);
            const int2 mt = a.meta2[c];
            const double dx = s_sub(gx, q.x), dy = s_sub(gy, q.y), dz = s_sub(gz, q.z);
            const double r2 = s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz));
            const bool internal = mt.x < 0;
            if (internal) {
                const int depth = -mt.x - 1;
                double w2;
                if (depth < W_TABLE) w2 = a.w2[depth];
                else { double w = a.root_size; for (int d = 0; d < depth; d++) w = s_div(w, 2.); w2 = s_mul(w, w); }
                if (w2 > s_mul(a.theta2, r2)) { c++; continue; }
            } else if ((uint32_t)mt.x == self) { c = mt.y; continue; }
            if (FAST) {
                const double ri = fast_rsqrt(r2 + a.soft2);
                const double ri2 = ri * ri, ri3 = ri2 * ri;
                const double prefact = negG * q.w * ri3;
                if (internal) {
                    const double mxx = a.quad[0 * S + c], mxy = a.quad[1 * S + c], mxz = a.quad[2 * S + c];
                    const double myy = a.quad[3 * S + c], myz = a.quad[4 * S + c], mzz = a.quad[5 * S + c];
                    double qp = a.G * ri3 * ri2;
                    sx += qp * (dx * mxx + dy * mxy + dz * mxz);
                    sy += qp * (dx * mxy + dy * myy + dz * myz);
                    sz += qp * (dx * mxz + dy * myz + dz * mzz);
                    const double mrr = dx * dx * mxx + dy * dy * myy + dz * dz * mzz + 2. * dx * dy * mxy + 2. * dx * dz * mxz + 2. * dy * dz * myz;
                    qp *= -2.5 * ri2 * mrr;
                    sx += (qp + prefact) * dx; sy += (qp + prefact) * dy; sz += (qp + prefact) * dz;
                } else { sx = fma(prefact, dx, sx); sy = fma(prefact, dy, sy); sz = fma(prefact, dz, sz); }
            } else {
                const double r = s_sqrt(s_add(r2, a.soft2));
                const double r3 = s_mul(s_mul(r, r), r);
                const double prefact = s_mul(s_div(negG, r3), q.w);                         // tree.c:292
                if (internal) {
                    const double mxx = a.quad[0 * S + c], mxy = a.quad[1 * S + c], mxz = a.quad[2 * S + c];
                    const double myy = a.quad[3 * S + c], myz = a.quad[4 * S + c], mzz = a.quad[5 * S + c];
                    double qp = s_div(a.G, s_mul(s_mul(r3, r), r));                          // G/(_r*_r*_r*_r*_r), tree.c:294
                    sx = s_add(sx, s_mul(qp, s_add(s_add(s_mul(dx, mxx), s_mul(dy, mxy)), s_mul(dz, mxz))));
                    sy = s_add(sy, s_mul(qp, s_add(s_add(s_mul(dx, mxy), s_mul(dy, myy)), s_mul(dz, myz))));
                    sz = s_add(sz, s_mul(qp, s_add(s_add(s_mul(dx, mxz), s_mul(dy, myz)), s_mul(dz, mzz))));
                    // mrr = dx*dx*mxx + dy*dy*myy + dz*dz*mzz + 2.*dx*dy*mxy + 2.*dx*dz*mxz + 2.*dy*dz*myz   (tree.c:298-299)
                    double mrr = s_mul(s_mul(dx, dx), mxx);
                    mrr = s_add(mrr, s_mul(s_mul(dy, dy), myy));
                    mrr = s_add(mrr, s_mul(s_mul(dz, dz), mzz));
                    mrr = s_add(mrr, s_mul(s_mul(s_mul(2., dx), dy), mxy));
                    mrr = s_add(mrr, s_mul(s_mul(s_mul(2., dx), dz), mxz));
                    mrr = s_add(mrr, s_mul(s_mul(s_mul(2., dy), dz), myz));
                    qp = s_mul(qp, s_mul(s_div(-5.0, s_mul(s_mul(2.0, r), r)), mrr));        // qprefact *= -5.0/(2.0*_r*_r)*mrr
                    const double f = s_add(qp, prefact);
                    sx = s_add(sx, s_mul(f, dx)); sy = s_add(sy, s_mul(f, dy)); sz = s_add(sz, s_mul(f, dz));
                } else {
                    sx = s_add(sx, s_mul(prefact, dx)); sy = s_add(sy, s_mul(prefact, dy)); sz = s_add(sz, s_mul(prefact, dz));
                }
            }
            c = mt.y;
        }
    }
    store_acc(a, self, k, sx, sy, sz);
}

// Warp-cooperative walk.  ncu showed the per-thread walk bound by L1 wavefronts (l1tex data pipe 93 % busy,
// 23 sectors per request): 32 lanes chase 32 different cells.  Here the 32 key-adjacent particles of a warp
// share ONE traversal cursor c: the cell record is fetched with a warp-uniform (broadcast, single wavefront)
// load, every lane still takes its own opening decision and keeps `next` = the first cell it wants to see
// again (c+1 after opening, skip after accepting), and the cursor advances to the minimum over the lanes.
// Each lane therefore visits exactly the cells, in exactly the order, of its own stackless walk: the result
// is bit-identical to walk_kernel, only the memory traffic is shared.
template <int MODE>
__global__ void __launch_bounds__(128) walk_coop_kernel(const WalkArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const bool live = t < a.n_work;
    uint32_t self = 0xffffffffu;
    double px = 0, py = 0, pz = 0;
    uint64_t k = 0;
    if (live) {
        k = a.list ? a.list[t] : a.k_begin + t;
        self = a.perm[k];
        px = a.x[self]; py = a.y[self]; pz = a.z[self];
    }
    double sx = 0., sy = 0., sz = 0.;
    unsigned wkey = 0;
    const double negG = -a.G;
    const int ngb = a.ghosts->n;
    const int n_cells = (int)a.n_cells;
    for (int g = 0; g < ngb; g++) {
        const double gx = s_add(a.ghosts->gb[g].x, px), gy = s_add(a.ghosts->gb[g].y, py), gz = s_add(a.ghosts->gb[g].z, pz);
        int next = live ? 0 : n_cells;
        int c = 0;                                   // warp-uniform cursor = min over lanes of `next`
        while (c < n_cells) {
            const double4 q = a.pos[c];              // uniform address: one wavefront for the whole warp
            const int4 mt = a.meta[c];
            if (next == c) {
                const double dx = s_sub(gx, q.x), dy = s_sub(gy, q.y), dz = s_sub(gz, q.z);
                const double r2 = s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz));
                bool interact = true;
                if (mt.x < 0) {
                    double w2;
                    if (mt.z < W_TABLE) w2 = a.w2[mt.z];
                    else { double w = a.root_size; for (int d = 0; d < mt.z; d++) w = s_div(w, 2.); w2 = s_mul(w, w); }
                    if (w2 > s_mul(a.theta2, r2)) { interact = false; next = c + 1; }     // tree.c:284: open the cell
                } else if ((uint32_t)mt.x == self) interact = false;                     // tree.c:311
                if (interact) {
                    if (MODE == 1) {
                        const double p = fast_m_over_r3(r2 + a.soft2, negG * q.w);
                        sx = fma(p, dx, sx); sy = fma(p, dy, sy); sz = fma(p, dz, sz);
                    } else if (MODE == 0) {
                        const double rs2 = s_add(r2, a.soft2);
                        wkey = max(wkey, strict_window_key(rs2));
                        const double r = fsqrt_rn_w(rs2);
                        const double p = s_mul(fdiv_rn_w(negG, s_mul(s_mul(r, r), r)), q.w);      // tree.c:292,313
                        sx = s_add(sx, s_mul(p, dx)); sy = s_add(sy, s_mul(p, dy)); sz = s_add(sz, s_mul(p, dz));
                    } else {
                        const double r = s_sqrt(s_add(r2, a.soft2));
                        const double p = s_mul(s_div(negG, s_mul(s_mul(r, r), r)), q.w);
                        sx = s_add(sx, s_mul(p, dx)); sy = s_add(sy, s_mul(p, dy)); sz = s_add(sz, s_mul(p, dz));
                    }
                }
                if (next == c) next = mt.y;          // accepted, leaf, or own leaf: continue after the subtree
            }
            c = __reduce_min_sync(0xffffffffu, next);
        }
    }
    if (!live) return;
    if (MODE == 0 && wkey >= STRICT_WINDOW_LIMIT) walk_generic(a, self, px, py, pz, sx, sy, sz);
    store_acc(a, self, k, sx, sy, sz);
}

// ---- FAST group walk: one warp = 32 key-adjacent particles, shared traversal, list-based evaluation -------------------
// north_star's "warp-cooperative Barnes-Hut walk" for REBCU_MODE_FAST.  The per-particle walks above keep every
// particle's own interaction list (bit parity), which leaves the warp divergent: 25 of 32 lanes active and the FP64
// pipe 53 % busy (profiles/r01_walk_rec_ncu.txt).  Here the 32 particles of a warp share ONE traversal and ONE list:
//  * traversal   lane-parallel over a shared-memory stack of sibling ranges (cell, end): each lane pops one range,
//                tests its first cell against the GROUP's bounding box, and pushes the rest of the range (skip, end)
//                and, if the cell has to be opened, its children (cell+1, skip).  32 cells are tested per trip on the
//                pre-order records the build already emits (no child table needed).
//  * criterion   box criterion (first kernels): a cell is accepted iff  w^2 <= theta^2 * d^2  with d the distance from its
//                centre of mass to the nearest point of the group's bounding box: d <= |x_i - com| for every particle i
//                of the group, so the cell would also be accepted by each particle's own test (src/tree.c:284).
//                Shipped (EXACT kernels, see below): where the box cannot decide, every particle is asked -- the group
//                criterion is then exactly as strict as its strictest particle; either way the error against direct
//                summation is at most the reference's.
//  * list        accepted cells and leaves are compacted (ballot + popc) into a shared-memory list (x,y,z,m) with the
//                ghost-box shift already removed; whenever the list cannot take 32 more entries, the warp evaluates
//                every entry for all its particles (broadcast LDS, 16 FP64 instructions per pair, fast_math.cuh; the
//                shipped PAIR kernels give every lane two particles and half the entries, see below).  The own leaf
//                contributes f * 0 = 0 exactly (no per-entry identity test, see gw_evaluate).
// The traversal costs a few instructions per visited cell per warp; the evaluation is pure FP64-pipe work with all
// lanes active.  The sum order differs from the reference's, so this is FAST mode only (tolerance stated in the tests).
// A stack that would overflow (trees deeper than ~GW_STACK levels) sends the warp to the per-particle FAST walk, and so does
// a list that grows past gw_abort entries: 32 key-adjacent particles that straddle a large cell boundary have a bounding
// box many cells wide, against which hardly any cell can be accepted -- a few such groups (lists of 1e5 entries and more)
// took longer than all others together (97 ms instead of ~3 at N = 2^20, profiles/r02_walk_group_v1_ncu.txt).
// Shape of the kernel (template parameters; REBOUND_B200_GW_VARIANT selects one for A/B runs): warps (= groups) per
// CTA, unroll of the evaluation loop, resident CTAs per SM the register allocation aims at, list / stack entries per warp.
constexpr int GW_STACK_MAX = 352;

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// soft2 > 0 always (a vanishing softening is replaced by 1e-200, which no real squared distance notices): the own leaf of a
// lane's particle then contributes f * 0 = 0 exactly, so the list needs no per-entry "is this me" test (the kernel is only
// used without ghost boxes, where a particle's own leaf holds exactly its own position; tree.c:311).
template <int UNROLL>
__device__ __forceinline__ void gw_evaluate(const double4* __restrict__ ent, int n, double px, double py, double pz, double soft2,
                                            double& sx, double& sy, double& sz) {
#pragma unroll UNROLL
    for (int j = 0; j < n; j++) {
        const double4 s = ent[j];
        const double dx = px - s.x, dy = py - s.y, dz = pz - s.z;
        const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, soft2)));
        const double f = fast_m_over_r3(r2, s.w);
        sx = fma(f, dx, sx); sy = fma(f, dy, sy); sz = fma(f, dz, sz);
    }
}

// Paired evaluation (PAIR kernels).  A broadcast LDS.128 writes 512 B of registers however few distinct addresses it reads,
// and the shared-memory pipe returns 128 B per clock: the two loads of an entry cost 8 cycles per warp against 34 cycles
// of FP64 issue -- with four warps per SM partition the two pipes are equally busy (the situation direct_fast_kernel was
// in with one particle per lane).  Here lane L evaluates TWO particles of the group, (L & 15) and (L & 15) + 16, against
// the entries of parity L >> 4: one pair of loads serves two pair terms.  The halves' partial sums meet in gw_pair_finish.
// An odd list is padded with a massless copy of its last entry.  Pair term: fast_m_over_r3 (16 FP64 instructions, full
// precision), or with CHEAP fast_m_over_r3_tree (15; measured and not shipped, see fast_math.cuh).
template <int UNROLL, bool CHEAP>
__device__ __forceinline__ void gw_evaluate_pair(double4* __restrict__ ent, int n, int half, double pxa, double pya, double pza,
                                                 double pxb, double pyb, double pzb, double soft2,
                                                 double& sxa, double& sya, double& sza, double& sxb, double& syb, double& szb) {
    if (n & 1) {
        if ((threadIdx.x & 31) == 0) { double4 e = ent[n - 1]; e.w = 0.; ent[n] = e; }
        __syncwarp();
    }
    const int np = (n + 1) >> 1;
    const double4* __restrict__ mine = ent + half;
#pragma unroll UNROLL
    for (int j = 0; j < np; j++) {
        const double4 s = mine[2 * j];
        const double dxa = pxa - s.x, dya = pya - s.y, dza = pza - s.z;
        const double dxb = pxb - s.x, dyb = pyb - s.y, dzb = pzb - s.z;
        const double r2a = fma(dxa, dxa, fma(dya, dya, fma(dza, dza, soft2)));
        const double r2b = fma(dxb, dxb, fma(dyb, dyb, fma(dzb, dzb, soft2)));
        const double fa = CHEAP ? fast_m_over_r3_tree(r2a, s.w) : fast_m_over_r3(r2a, s.w);
        const double fb = CHEAP ? fast_m_over_r3_tree(r2b, s.w) : fast_m_over_r3(r2b, s.w);
        sxa = fma(fa, dxa, sxa); sya = fma(fa, dya, sya); sza = fma(fa, dza, sza);
        sxb = fma(fb, dxb, sxb); syb = fma(fb, dyb, syb); szb = fma(fb, dzb, szb);
    }
}

// lane L ends up with the complete sum of ITS particle: the a-sums for L < 16, the b-sums for L >= 16
__device__ __forceinline__ double gw_pair_finish(double sa, double sb, int half) {
    sa += __shfl_xor_sync(0xffffffffu, sa, 16);
    sb += __shfl_xor_sync(0xffffffffu, sb, 16);
    return half ? sb : sa;
}

// Exact group criterion (EXACT kernels).  The bounding-box test is conservative: it opens every cell SOME point of the box
// would open, and 32 key-adjacent particles fill their box sparsely (a Z-order run is L-shaped or split as often as not).
// The tightest criterion a shared list allows is "accept iff every particle of the group accepts the cell" (tree.c:284
// for each of them).  It is evaluated only where the box cannot decide: a cell the box accepts is accepted; a cell that
// even the FARTHEST point of the box would open is opened; the cells in between (a third of the internal cells visited)
// are parked in shared memory and every lane tests them against its own particle, the verdicts being OR-ed across the
// warp.  Measured on the CPU model of the walk (disc, N = 2^17, 150 groups): list 1080 entries instead of 1430
// (per-particle lists: 609), 1408 instead of 1916 visited cells, and the longest list 1638 instead of 6778 -- the groups
// that had to give up with the box criterion finish here like all others.
// retry: [0] = number of groups that gave up, [1...] = their first work item; walk_retry_kernel finishes them.
template <int WARPS, int UNROLL, int MINB, int LIST, int STACK, bool PAIR = false, bool EXACT = false, bool CHEAP = true>
__global__ void __launch_bounds__(32 * WARPS, MINB) walk_group_kernel(const WalkArgs a, unsigned long long* __restrict__ stats,
                                                                      unsigned int* __restrict__ retry) {
    __shared__ int2 s_stack[WARPS][STACK];
    __shared__ double4 s_ent[WARPS][LIST + (PAIR ? 1 : 0)];
    __shared__ double4 s_amb[EXACT ? WARPS : 1][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t t0 = ((uint64_t)blockIdx.x * WARPS + w) * 32;
    if (t0 >= a.n_work) return;                              // warp-uniform; no block-wide barrier below
    const uint64_t t = t0 + lane;
    const bool live = t < a.n_work;
    const uint64_t tt = live ? t : a.n_work - 1;             // idle lanes shadow the last particle (keeps the box tight)
    const uint64_t k = a.list ? a.list[tt] : a.k_begin + tt;
    const int self = (int)a.perm[k];
    const double px = a.x[self], py = a.y[self], pz = a.z[self];
    const double lox = warp_min_d(px), hix = warp_max_d(px);
    const double loy = warp_min_d(py), hiy = warp_max_d(py);
    const double loz = warp_min_d(pz), hiz = warp_max_d(pz);
    const double cxg = 0.5 * (lox + hix), cyg = 0.5 * (loy + hiy), czg = 0.5 * (loz + hiz);
    // half extents, rounded up so that the box certainly contains every particle of the group
    const double hxg = fmax(hix - cxg, cxg - lox), hyg = fmax(hiy - cyg, cyg - loy), hzg = fmax(hiz - czg, czg - loz);
    int2* stack = s_stack[w];
    double4* ent = s_ent[w];
    const double soft2 = a.soft2 > 0. ? a.soft2 : 1e-200;
    const unsigned lt = (1u << lane) - 1u;
    const int n_cells = (int)a.n_cells;
    const int ngb = a.ghosts->n;
    const int stack_cap = a.gw_stack < STACK ? a.gw_stack : STACK;
    double sx = 0., sy = 0., sz = 0.;
    // PAIR: this lane's two particles (see gw_evaluate_pair) and their partial sums
    const int half = lane >> 4;
    double pxa = 0., pya = 0., pza = 0., pxb = 0., pyb = 0., pzb = 0., sxb = 0., syb = 0., szb = 0.;
    if (PAIR) {
        pxa = __shfl_sync(0xffffffffu, px, lane & 15); pya = __shfl_sync(0xffffffffu, py, lane & 15); pza = __shfl_sync(0xffffffffu, pz, lane & 15);
        pxb = __shfl_sync(0xffffffffu, px, (lane & 15) + 16); pyb = __shfl_sync(0xffffffffu, py, (lane & 15) + 16); pzb = __shfl_sync(0xffffffffu, pz, (lane & 15) + 16);
    }
    int nl = 0;
    bool overflow = false;
    unsigned int n_ent = 0, n_vis = 0;
    for (int g = 0; g < ngb && !overflow; g++) {
        const double gbx = a.ghosts->gb[g].x, gby = a.ghosts->gb[g].y, gbz = a.ghosts->gb[g].z;
        const double bx = cxg + gbx, by = cyg + gby, bz = czg + gbz;      // centre of the shifted group box (gravity.c:93-97)
        int sp = 0;
        if (n_cells > 0) { if (lane == 0) stack[0] = make_int2(0, n_cells); sp = 1; }
        __syncwarp();
        while (sp > 0) {
            int n = sp < 32 ? sp : 32;
            if (n > stack_cap - sp) n = stack_cap - sp;       // every popped range may push two
            if (n < 1) { overflow = true; break; }
            int c = -1, end = 0;
            if (lane < n) { const int2 e = stack[sp - 1 - lane]; c = e.x; end = e.y; }
            sp -= n;
            __syncwarp();
            bool open = false, amb = false;
            int skip = 0, tag = 0;
            double4 q = make_double4(0., 0., 0., 0.);
            double cm = 0., w2c = 0.;
            if (c >= 0) {
                q = ld_pos256(a.rec + c);
                cm = a.m[c];                                  // requested together with the record: one latency, not two
                const long long bits = __double_as_longlong(q.w);
                tag = (int)(unsigned int)(unsigned long long)bits;
                skip = (int)(unsigned int)((unsigned long long)bits >> 32);
                if (tag < 0) {
                    const int hi = tag & 0x7fffffff;
                    double w2 = __hiloint2double(hi, (int)a.w2_lo);
                    if (hi == 0) w2 = cell_w2_rare(a.root_size, -a.meta2[c].x - 1);
                    const double ux = fabs(q.x - bx), uy = fabs(q.y - by), uz = fabs(q.z - bz);
                    const double ex = fmax(ux - hxg, 0.), ey = fmax(uy - hyg, 0.), ez = fmax(uz - hzg, 0.);
                    const double d2 = fma(ex, ex, fma(ey, ey, ez * ez));
                    open = w2 > a.theta2 * d2;
                    if (EXACT) {
                        const double fx = ux + hxg, fy = uy + hyg, fz = uz + hzg;      // farthest point of the box
                        amb = open && !(w2 > a.theta2 * fma(fx, fx, fma(fy, fy, fz * fz)));
                        w2c = w2;
                    }
                }
            }
            if (EXACT) {
                const unsigned m_amb = __ballot_sync(0xffffffffu, amb);
                if (m_amb) {                                  // warp-uniform
                    double4* parked = s_amb[w];
                    const int mine = __popc(m_amb & lt);
                    if (amb) parked[mine] = make_double4(q.x - gbx, q.y - gby, q.z - gbz, w2c);
                    __syncwarp();
                    const int na = __popc(m_amb);
                    unsigned verdict = 0;                     // bit i: my particle would open parked cell i (tree.c:284)
#pragma unroll 4
                    for (int i = 0; i < na; i++) {
                        const double4 pc = parked[i];
                        const double dx = px - pc.x, dy = py - pc.y, dz = pz - pc.z;
                        const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
                        verdict |= (pc.w > a.theta2 * r2 ? 1u : 0u) << i;
                    }
                    verdict = __reduce_or_sync(0xffffffffu, verdict);
                    if (amb) open = (verdict >> mine) & 1u;
                    __syncwarp();
                }
            }
            const bool acc = c >= 0 && !open;
            const bool sib = c >= 0 && skip < end;
            const unsigned m_sib = __ballot_sync(0xffffffffu, sib), m_open = __ballot_sync(0xffffffffu, open);
            const unsigned m_acc = __ballot_sync(0xffffffffu, acc);
            const int n_sib = __popc(m_sib), n_acc = __popc(m_acc);
            // siblings below, children on top: the next trip continues depth-first (the records it touches are neighbours)
            if (sib) stack[sp + __popc(m_sib & lt)] = make_int2(skip, end);
            if (open) stack[sp + n_sib + __popc(m_open & lt)] = make_int2(c + 1, skip);
            sp += n_sib + __popc(m_open);
            if (nl + n_acc > LIST) {
                __syncwarp();
                if (PAIR) gw_evaluate_pair<UNROLL, CHEAP>(ent, nl, half, pxa, pya, pza, pxb, pyb, pzb, soft2, sx, sy, sz, sxb, syb, szb);
                else gw_evaluate<UNROLL>(ent, nl, px, py, pz, soft2, sx, sy, sz);
                n_ent += nl;
                nl = 0;
                __syncwarp();
                if (n_ent > a.gw_abort) { overflow = true; break; }
            }
            if (acc) {
                const int slot = nl + __popc(m_acc & lt);
                ent[slot] = make_double4(q.x - gbx, q.y - gby, q.z - gbz, cm);
            }
            nl += n_acc;
            n_vis += n;
            __syncwarp();
        }
    }
    if (overflow) {
        // too deep for the stack, or a list far longer than its particles' own lists would be: the group is finished
        // particle by particle (walk_retry_kernel) -- outside this kernel, so that the slow groups do not hold an SM
        if (lane == 0) retry[1 + atomicAdd(&retry[0], 1u)] = (unsigned int)(t0 / 32);
        return;
    }
    __syncwarp();
    if (PAIR) {
        gw_evaluate_pair<UNROLL, CHEAP>(ent, nl, half, pxa, pya, pza, pxb, pyb, pzb, soft2, sx, sy, sz, sxb, syb, szb);
        sx = gw_pair_finish(sx, sxb, half); sy = gw_pair_finish(sy, syb, half); sz = gw_pair_finish(sz, szb, half);
    } else gw_evaluate<UNROLL>(ent, nl, px, py, pz, soft2, sx, sy, sz);
    n_ent += nl;
    const double negG = -a.G;
    if (live) store_acc(a, (uint32_t)self, k, negG * sx, negG * sy, negG * sz);
    if (stats && lane == 0) { atomicAdd(&stats[0], (unsigned long long)n_ent); atomicAdd(&stats[1], (unsigned long long)n_vis); atomicAdd(&stats[2], 1ull); }
}

// The particles of the groups that gave up, one thread each, per-particle criterion, FAST arithmetic.  Launched with a
// grid for the worst case; CTAs beyond the retry count leave at once (no host round trip for the count).
__global__ void __launch_bounds__(128) walk_retry_kernel(const WalkArgs a, const unsigned int* __restrict__ retry) {
    const uint64_t slot = ((uint64_t)blockIdx.x * 128 + threadIdx.x) / 32;
    if (slot >= retry[0]) return;
    const uint64_t t = (uint64_t)retry[1 + slot] * 32 + (threadIdx.x & 31);
    if (t >= a.n_work) return;
    const uint64_t k = a.list ? a.list[t] : a.k_begin + t;
    const uint32_t self = a.perm[k];
    double sx, sy, sz;
    walk_one_rec<1>(a, self, a.x[self], a.y[self], a.z[self], sx, sy, sz);
    store_acc(a, self, k, sx, sy, sz);
}

// Interaction count of the per-particle criterion (what the reference and the STRICT walk evaluate): accepted cells and
// leaves summed over the work items; instrumentation for the roofline figures of bench.py, not part of a step.
__global__ void __launch_bounds__(128) walk_count_kernel(const WalkArgs a, unsigned long long* __restrict__ stats) {
    const uint64_t t = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    unsigned long long n_acc = 0, n_vis = 0;
    if (t < a.n_work) {
        const uint64_t k = a.list ? a.list[t] : a.k_begin + t;
        const uint32_t self = a.perm[k];
        const double px = a.x[self], py = a.y[self], pz = a.z[self];
        const int n_cells = (int)a.n_cells;
        for (int g = 0; g < a.ghosts->n; g++) {
            const double gx = s_add(a.ghosts->gb[g].x, px), gy = s_add(a.ghosts->gb[g].y, py), gz = s_add(a.ghosts->gb[g].z, pz);
            int c = 0;
            while (c < n_cells) {
                const double4 q = ld_pos256(a.rec + c);
                const long long bits = __double_as_longlong(q.w);
                const int tag = (int)(unsigned int)(unsigned long long)bits, skip = (int)(unsigned int)((unsigned long long)bits >> 32);
                const double dx = s_sub(gx, q.x), dy = s_sub(gy, q.y), dz = s_sub(gz, q.z);
                const double r2 = s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz));
                const int hi = tag & 0x7fffffff;
                double w2 = __hiloint2double(hi, (int)a.w2_lo);
                if (tag < 0 && hi == 0) w2 = cell_w2_rare(a.root_size, -a.meta2[c].x - 1);
                const bool open = tag < 0 && w2 > s_mul(a.theta2, r2);
                n_vis++;
                if (tag < 0 ? !open : (uint32_t)tag != self) n_acc++;
                c = open ? c + 1 : skip;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o); n_vis += __shfl_xor_sync(0xffffffffu, n_vis, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&stats[0], n_acc); atomicAdd(&stats[1], n_vis); }
}

__global__ void __launch_bounds__(256) shard_flag_kernel(uint64_t n, const uint32_t* __restrict__ perm, uint32_t b, uint32_t e,
                                                         uint32_t* __restrict__ flag) {
    const uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (k < n) { const uint32_t p = perm[k]; flag[k] = (p >= b && p < e) ? 1u : 0u; }
}
__global__ void __launch_bounds__(256) shard_list_kernel(uint64_t n, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                                         uint32_t* __restrict__ list) {
    const uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (k < n && flag[k]) list[pos[k]] = (uint32_t)k;
}

template <typename T>
int ensure(rebcu_handle* h, T** ptr, size_t count) {
    cudaFree(*ptr);
    *ptr = nullptr;
    CU_TRY(h, cudaMalloc(ptr, count * sizeof(T)));
    return REBCU_OK;
}

}  // namespace

void tree_free(rebcu_handle* h) {
    TreeBuffers& T = h->tree;
    cudaFree(T.keys); cudaFree(T.keys_sorted); cudaFree(T.perm); cudaFree(T.perm_in); cudaFree(T.lcp);
    cudaFree(T.cell_off); cudaFree(T.cell_cnt); cudaFree(T.cells); cudaFree(T.parent); cudaFree(T.ready);
    cudaFree(T.walk_pos); cudaFree(T.walk_geo); cudaFree(T.walk_meta); cudaFree(T.walk_meta2); cudaFree(T.walk_rec); cudaFree(T.walk_m); cudaFree(T.col_rec); cudaFree(T.sort_tmp); cudaFree(T.scan_tmp); cudaFree(T.flags);
    cudaFree(T.shard_list); cudaFree(T.quad); cudaFree(T.acc_sorted);
    cudaFree(T.sh_keys); cudaFree(T.sh_idx); cudaFree(T.sh_hist); cudaFree(T.sh_pstart); cudaFree(T.sh_tab); cudaFree(T.sh_level);
    cudaFree(T.sh_delta); cudaFree(T.sh_top); cudaFree(T.sh_info);
    T = TreeBuffers();
}

static int tree_error(rebcu_handle* h, const int* f) {
    // The reference stops at the first offending particle in index order (tree.c:263-270).
    int best = 0x7fffffff, which = -1;
    for (int k = 0; k < 3; k++) if (f[k] < best) { best = f[k]; which = k; }
    if (which == 0) return rebcu_fail(h, REBCU_ERR_OUTSIDE_BOX, "Particle is outside of simulation box. Cannot add to tree.");
    if (which == 1) return rebcu_fail(h, REBCU_ERR_NONFINITE, "Particle has non-finite coordinates. Cannot add to tree.");
    if (which == 2) return rebcu_fail(h, REBCU_ERR_SAME_COORDINATES, "Cannot add two particles with the same coordinates to the tree.");
    if (f[3] != 0x7fffffff) return rebcu_fail(h, REBCU_ERR_CELL_SIZE_ZERO, "Tree cell has size zero.");
    return REBCU_OK;
}

// Key layout of this build and the per-particle arena; *empty = nothing to build (N == 0).
static int tree_prepare(rebcu_handle* h, const rebcu_config* c, TreeParams& P, bool* empty) {
    TreeBuffers& T = h->tree;
    *empty = false;
    if (c->root_size <= 0.0)
        return rebcu_fail(h, REBCU_ERR_ROOT_SIZE, "Set root_size to a finite value to use a tree based gravity or collision solver.");
    const uint64_t n = h->N;
    T.n_cells = 0;
    if (n == 0) { *empty = true; return REBCU_OK; }
    if (n >= (1ull << 31)) return rebcu_fail(h, REBCU_ERR_ARG, "tree supports N < 2^31 (int indices, tree.h:52)");
    P.clamp = -1; P.bshift = 0; P.bdelta = nullptr;
    P.root_size = c->root_size; P.Nx = c->N_root_x; P.Ny = c->N_root_y; P.Nz = c->N_root_z; P.n = n;
    const uint64_t n_root = (uint64_t)P.Nx * P.Ny * P.Nz;
    P.rbits = 0;
    while ((1ull << P.rbits) < n_root) P.rbits++;
    P.L0 = (63 - P.rbits) / 3;
    if (P.L0 < 1) return rebcu_fail(h, REBCU_ERR_ARG, "too many root boxes");
    // Key prefix: the radix sort costs one pass per 8 key bits, and all 63 path bits are only needed to separate
    // particles closer than 2^-21 of the box.  With log2(N)/2 + 2 levels (a surface distribution fills ~4^level cells)
    // most particles already have a cell of their own; the few that share one form short runs of equal keys, which the
    // exact pairwise descent of tie_kernel / lcp_kernel orders anyway (the same code that handles paths deeper than 21
    // levels).  A run longer than TIE_RUN_MAX (clustered input) aborts the build, which then falls back to full keys
    // for this simulation.  REBOUND_B200_KEY_LEVELS=<n> forces a prefix (tests), =0 switches the prefix off.
    P.prefix = 0;
    {
        static const int forced = [] { const char* e = getenv("REBOUND_B200_KEY_LEVELS"); return e ? atoi(e) : -1; }();
        int want = 0;
        if (forced > 0) want = forced;
        else if (forced < 0 && T.prefix_ok && n >= 65536) {
            want = 2; for (uint64_t q = n; q > 1; q >>= 2) want++;            // ceil-ish log4(N) + 2
            const int over = (P.rbits + 3 * want) % 8;
            if (over >= 1 && over <= 3 && want > 8) want--;                   // one level less saves a whole pass
        }
        if (want >= 1 && want < P.L0 && (forced > 0 ? T.prefix_ok : true)) { P.L0 = want; P.prefix = 1; }
    }

    if (T.cap_n < n) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        const uint64_t cap = h->cap > n ? h->cap : n;
        int err;
        if ((err = ensure(h, &T.keys, cap))) return err;
        if ((err = ensure(h, &T.keys_sorted, cap))) return err;
        if ((err = ensure(h, &T.perm, cap))) return err;
        if ((err = ensure(h, &T.perm_in, cap))) return err;
        if ((err = ensure(h, &T.lcp, cap + 1))) return err;
        if ((err = ensure(h, &T.cell_off, cap + 1))) return err;
        if ((err = ensure(h, &T.cell_cnt, cap + 1))) return err;
        if ((err = ensure(h, &T.shard_list, cap))) return err;
        if (!T.flags) { if ((err = ensure(h, &T.flags, 8))) return err; }
        // radix-sort histogram (256 bins per 2048-element tile) and scan scratch (tile sums)
        const size_t hist_words = prim::radix_hist_words(cap);
        const size_t scan_words = prim::scan_scratch_words(hist_words > cap + 1 ? hist_words : cap + 1);
        cudaFree(T.sort_tmp); cudaFree(T.scan_tmp); T.sort_tmp = T.scan_tmp = nullptr;
        CU_TRY(h, cudaMalloc(&T.sort_tmp, hist_words * sizeof(uint32_t))); T.sort_tmp_bytes = hist_words * sizeof(uint32_t);
        CU_TRY(h, cudaMalloc(&T.scan_tmp, scan_words * sizeof(uint32_t))); T.scan_tmp_bytes = scan_words * sizeof(uint32_t);
        T.cap_n = cap;
    }
    return REBCU_OK;
}

static int tree_cell_capacity(rebcu_handle* h, uint64_t n_cells) {
    TreeBuffers& T = h->tree;
    if (T.cap_cells < n_cells) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        const uint64_t cap = n_cells + n_cells / 8 + 1024;
        int err;
        if ((err = ensure(h, &T.walk_pos, cap))) return err;
        if ((err = ensure(h, &T.walk_geo, cap))) return err;
        if ((err = ensure(h, (int4**)&T.walk_meta, cap))) return err;
        if ((err = ensure(h, &T.parent, cap))) return err;
        if ((err = ensure(h, &T.ready, cap))) return err;
        if ((err = ensure(h, &T.walk_meta2, cap))) return err;
        cudaFree(T.cells); T.cells = nullptr;
        T.cap_cells = cap;
    }
    return REBCU_OK;
}

int tree_build(rebcu_handle* h, const rebcu_config* c) {
    TreeBuffers& T = h->tree;
    TreeParams P;
    bool empty;
    { const int err = tree_prepare(h, c, P, &empty); if (err || empty) return err; }
    const uint64_t n = h->N;
    T.rec_ready = false; T.complete = true;
    const double *x = h->f(F_X), *y = h->f(F_Y), *z = h->f(F_Z), *m = h->f(F_M);
    {
        LaunchScope ls(h, TC_TREEBUILD, 8);
        CU_TRY(h, cudaMemsetAsync(T.flags, 0x7f, 8 * sizeof(int), h->stream));   // 0x7f7f7f7f: "no error"
        key_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(P, x, y, z, T.keys, T.perm_in, T.flags);
        h->launches += prim::radix_sort_pairs(h->stream, T.keys, T.perm_in, T.keys_sorted, T.perm, n, P.rbits + 3 * P.L0,
                                              (uint32_t*)T.sort_tmp, (uint32_t*)T.scan_tmp);
        tie_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(P, T.keys_sorted, T.perm, x, y, z, T.flags);
        lcp_kernel<<<div_up(n + 1, 256), 256, 0, h->stream>>>(P, T.keys_sorted, T.perm, x, y, z, T.lcp, T.flags);
        count_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(n, T.lcp, T.cell_cnt);
        CU_TRY(h, cudaMemsetAsync(T.cell_cnt + n, 0, sizeof(uint32_t), h->stream));
        prim::exclusive_scan_u32(h->stream, T.cell_cnt, T.cell_off, n + 1, (uint32_t*)T.scan_tmp);
    }
    CU_TRY(h, cudaGetLastError());
    // cell count and error flags back to the host
    int* pin = (int*)h->pinned;
    CU_TRY(h, cudaMemcpyAsync(pin, T.flags, 5 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemcpyAsync(pin + 5, T.cell_off + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (P.prefix && pin[4] != 0x7f7f7f7f) {          // a long tie run: this input needs the full keys
        T.prefix_ok = false;
        return tree_build(h, c);
    }
    int f[4];
    for (int k = 0; k < 4; k++) f[k] = (pin[k] == 0x7f7f7f7f) ? 0x7fffffff : pin[k];
    if (f[0] != 0x7fffffff || f[1] != 0x7fffffff || f[2] != 0x7fffffff || f[3] != 0x7fffffff) return tree_error(h, f);
    const uint64_t n_cells = (uint32_t)pin[5];
    { const int err = tree_cell_capacity(h, n_cells); if (err) return err; }
    CellArrays C{T.walk_pos, T.walk_geo, (int4*)T.walk_meta, T.parent, T.ready, T.walk_meta2, nullptr, 0};
    T.has_quad = false;
    if (c->quadrupole) {
        if (T.quad_cap < T.cap_cells) {
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            cudaFree(T.quad); T.quad = nullptr; T.quad_cap = 0;
            CU_TRY(h, cudaMalloc(&T.quad, 6 * T.cap_cells * sizeof(double)));
            T.quad_cap = T.cap_cells;
        }
        CU_TRY(h, cudaMemsetAsync(T.quad, 0, 6 * T.quad_cap * sizeof(double), h->stream));     // leaves: tree.c:148-155
        C.quad = T.quad; C.quad_stride = T.quad_cap;
        T.has_quad = true;
    }
    {
        LaunchScope ls(h, TC_TREEBUILD, 3);
        emit_kernel<<<div_up(n, 128), 128, 0, h->stream>>>(P, T.keys_sorted, T.perm, T.lcp, T.cell_off, x, y, z, m, C);
        adopt_kernel<<<div_up(n_cells, 256), 256, 0, h->stream>>>(0, n_cells, C);
        moment_kernel<<<div_up(n_cells, 256), 256, 0, h->stream>>>(0, n_cells, C);
    }
    CU_TRY(h, cudaGetLastError());
    T.n_cells = n_cells;
    T.built_for_n = (int)n;
    return REBCU_OK;
}

// ---- sharded build (SURVEY 8e stage 2: key-range ownership, per-rank subtrees, replicated top tree) ---------------------
// With W ranks the replicated build above makes every rank sort and emit all N particles.  Here the tree is cut at
// octant level SH_LS below the root cells into "buckets" (4096 per root box):
//   1. every rank computes all keys (positions are replicated after the exchange) and the bucket histogram, hence the
//      same bucket-aligned splitters: rank r owns the buckets [B_r, B_r+1), about N/W particles;
//   2. it compacts, sorts and builds ONLY its buckets: the subtree below every bucket cell (neighbours in different
//      buckets count as sharing SH_LS-1 levels, so no emitted cell spans two buckets) with moments, bottom-up as before;
//   3. the ranks all-gather a small table (cells per bucket, error flags), from which each computes the complete TOP tree
//      -- a cell above bucket level exists iff it holds >= 2 particles -- and with it the global pre-order index of
//      every bucket subtree; the subtrees are emitted directly at those indices, packed into traversal records, and
//      all-gathered (40 B per cell) together with the sorted permutation (4 B per particle);
//   4. every rank then fills in the few hundred top cells: moments from their children in octant order with the
//      reference's expression (tree.c:162-179).
// The result is the same pre-order record array the replicated build + walk_pack produce, bit for bit for everything the
// gravity walks read (centre of mass, mass, skip links, width tags; a lone particle in a bucket is emitted as a leaf at
// bucket depth, which no walk looks at), so sharded STRICT runs stay bit-identical to single-GPU runs.  Host round trips:
// two per build (splitters; cell counts + error flags).  Quadrupole builds, trees deeper than 480 levels and collision
// searches keep the replicated build.
constexpr int SH_LS = 4;
enum { I_NTOT = 0, I_NTOP = 1, I_FLAGS = 2, I_S = 8, I_B = 25, I_P = 42, I_WORDS = 64 };

__global__ void __launch_bounds__(256) bucket_hist_kernel(uint64_t n, const uint64_t* __restrict__ keys, int bshift, uint32_t* __restrict__ hist) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    if (key == ~0ull) return;                          // flagged by key_kernel; the build aborts
    atomicAdd(&hist[key >> bshift], 1u);
}

// One block: exclusive prefix sum of the histogram and the splitters B_r = first bucket whose start is >= r N / W.
__global__ void __launch_bounds__(1024) bucket_split_kernel(uint32_t n_buckets, const uint32_t* __restrict__ hist, int W,
                                                            uint64_t* __restrict__ pstart, uint32_t* __restrict__ info) {
    __shared__ uint64_t part[1024];
    const int t = threadIdx.x;
    const uint32_t chunk = (n_buckets + 1023) / 1024;
    const uint32_t b0 = t * chunk, b1 = min(n_buckets, b0 + chunk);
    uint64_t sum = 0;
    for (uint32_t b = b0; b < b1; b++) sum += hist[b];
    part[t] = sum;
    __syncthreads();
    if (t == 0) { uint64_t run = 0; for (int k = 0; k < 1024; k++) { const uint64_t v = part[k]; part[k] = run; run += v; } pstart[n_buckets] = run; }
    __syncthreads();
    uint64_t run = part[t];
    for (uint32_t b = b0; b < b1; b++) { pstart[b] = run; run += hist[b]; }
    __syncthreads();
    __threadfence_block();
    if (t <= W) {
        const uint64_t total = pstart[n_buckets];
        const uint64_t want = total * (uint64_t)t / (uint64_t)W;
        uint32_t lo = 0, hi = n_buckets;             // first b with pstart[b] >= want
        while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (pstart[mid] >= want) hi = mid; else lo = mid + 1; }
        uint32_t B = (t == 0) ? 0 : (t == W ? n_buckets : lo);
        info[I_B + t] = B;
        info[I_P + t] = (uint32_t)pstart[B];
    }
}

__global__ void __launch_bounds__(256) bucket_flag_kernel(uint64_t n, const uint64_t* __restrict__ keys, int bshift, uint32_t lo, uint32_t hi,
                                                          uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    const uint64_t b = key >> bshift;
    flag[i] = (key != ~0ull && b >= lo && b < hi) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) bucket_scatter_kernel(uint64_t n, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ flag,
                                                             const uint32_t* __restrict__ pos, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n && flag[i]) { keys_out[pos[i]] = keys[i]; idx_out[pos[i]] = (uint32_t)i; }
}

// tab entry of bucket b owned by rank r sits at b + 8 (r + 1); low word = local cell offset of its first particle,
// high word = local cell offset behind its last particle.
__global__ void __launch_bounds__(256) bucket_bounds_kernel(uint64_t n_loc, const uint64_t* __restrict__ keys, int bshift,
                                                            const uint32_t* __restrict__ off, uint32_t* __restrict__ tab_mine) {
    const uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= n_loc) return;
    const uint64_t b = keys[k] >> bshift;
    if (k == 0 || (keys[k - 1] >> bshift) != b) tab_mine[2 * b] = off[k];
    if (k + 1 == n_loc || (keys[k + 1] >> bshift) != b) tab_mine[2 * b + 1] = off[k + 1];
}

struct TopArgs {
    uint32_t n_root, n_buckets; int W;
    const uint32_t* hist; const uint64_t* tab; uint32_t* level; uint64_t level_words;
    long long* delta; int4* top; uint32_t* info; uint32_t top_cap;
};

// One block on every rank (identical inputs, identical outputs): the top tree from the gathered tables.
__global__ void __launch_bounds__(1024) top_layout_kernel(TopArgs A) {
    __shared__ uint32_t s_ntop;
    __shared__ uint32_t lvl_off[SH_LS + 2];
    const int t = threadIdx.x, T = blockDim.x;
    uint32_t* cnt = A.level; uint32_t* cells = A.level + A.level_words; uint32_t* start = A.level + 2 * A.level_words;
    if (t == 0) {
        s_ntop = 0;
        uint32_t o = 0, w = A.n_root;
        for (int d = 0; d <= SH_LS; d++) { lvl_off[d] = o; o += w; w *= 8; }
        lvl_off[SH_LS + 1] = o;
        // flags: minimum over the ranks' status words (tab[B_r + 8 r + k], k < 6)
        for (int k = 0; k < 6; k++) {
            uint32_t v = 0x7f7f7f7fu;
            for (int r = 0; r < A.W; r++) { const uint32_t f = (uint32_t)A.tab[A.info[I_B + r] + 8 * r + k]; if (f < v) v = f; }
            A.info[I_FLAGS + k] = v;
        }
    }
    __syncthreads();
    // bucket level: counts from the histogram, cells from the owners' (first, end) offsets
    for (int r = 0; r < A.W; r++) {
        const uint32_t b0 = A.info[I_B + r], b1 = A.info[I_B + r + 1];
        for (uint32_t b = b0 + t; b < b1; b += T) {
            const uint64_t e = A.tab[b + 8 * (r + 1)];
            cnt[lvl_off[SH_LS] + b] = A.hist[b];
            cells[lvl_off[SH_LS] + b] = A.hist[b] ? (uint32_t)(e >> 32) - (uint32_t)e : 0u;
        }
    }
    __syncthreads();
    for (int d = SH_LS - 1; d >= 0; d--) {
        const uint32_t w = lvl_off[d + 1] - lvl_off[d];
        for (uint32_t p = t; p < w; p += T) {
            uint32_t c = 0, ce = 0;
            for (int o = 0; o < 8; o++) { c += cnt[lvl_off[d + 1] + 8 * p + o]; ce += cells[lvl_off[d + 1] + 8 * p + o]; }
            cnt[lvl_off[d] + p] = c;
            cells[lvl_off[d] + p] = ce + (c >= 2 ? 1u : 0u);
        }
        __syncthreads();
    }
    if (t == 0) {
        uint32_t run = 0;
        for (uint32_t rb = 0; rb < A.n_root; rb++) { start[rb] = run; run += cells[rb]; }
        A.info[I_NTOT] = run;
    }
    __syncthreads();
    for (int d = 0; d < SH_LS; d++) {
        const uint32_t w = lvl_off[d + 1] - lvl_off[d];
        for (uint32_t p = t; p < w; p += T) {
            const uint32_t c = cnt[lvl_off[d] + p];
            uint32_t s = start[lvl_off[d] + p];
            if (c >= 2) {
                const uint32_t slot = atomicAdd(&s_ntop, 1u);
                if (slot < A.top_cap) A.top[slot] = make_int4((int)s, d, (int)(s + cells[lvl_off[d] + p]), (int)c);
                s += 1;
            }
            for (int o = 0; o < 8; o++) { start[lvl_off[d + 1] + 8 * p + o] = s; s += cells[lvl_off[d + 1] + 8 * p + o]; }
        }
        __syncthreads();
    }
    // global index of every bucket subtree minus the owner's local offset of its first cell
    for (int r = 0; r < A.W; r++) {
        const uint32_t b0 = A.info[I_B + r], b1 = A.info[I_B + r + 1];
        for (uint32_t b = b0 + t; b < b1; b += T)
            A.delta[b] = (long long)start[lvl_off[SH_LS] + b] - (long long)(uint32_t)A.tab[b + 8 * (r + 1)];
    }
    // spans of the ranks in the global cell array: from the first cell of a rank's first non-empty bucket to the next rank's
    if (t <= A.W) {
        uint32_t S;
        if (t == 0) S = 0;
        else if (t == A.W) S = A.info[I_NTOT];
        else {
            uint32_t b = A.info[I_B + t];
            while (b < A.n_buckets && A.hist[b] == 0) b++;
            S = b < A.n_buckets ? start[lvl_off[SH_LS] + b] : A.info[I_NTOT];
        }
        A.info[I_S + t] = S;
    }
    __syncthreads();
    if (t == 0) A.info[I_NTOP] = s_ntop;
}

// One block on every rank, after the subtrees have been gathered: records of the top cells, deepest level first.
__global__ void __launch_bounds__(1024) top_moment_kernel(const int4* __restrict__ top, uint32_t n_top, double4* __restrict__ rec, double* __restrict__ m, WalkArgs a) {
    for (int d = SH_LS - 1; d >= 0; d--) {
        for (uint32_t k = threadIdx.x; k < n_top; k += blockDim.x) {
            const int4 tc = top[k];
            if (tc.y != d) continue;
            double mm = 0., mx = 0., my = 0., mz = 0.;
            for (int ch = tc.x + 1; ch < tc.z;) {
                const volatile double4* q = (const volatile double4*)&rec[ch];
                const double dx = q->x, dy = q->y, dz = q->z;
                const long long bits = __double_as_longlong(q->w);
                const double dm = ((const volatile double*)m)[ch];
                mx = s_add(mx, s_mul(dx, dm));
                my = s_add(my, s_mul(dy, dm));
                mz = s_add(mz, s_mul(dz, dm));
                mm = s_add(mm, dm);
                ch = (int)(unsigned int)((unsigned long long)bits >> 32);
            }
            if (mm > 0) { mx = s_div(mx, mm); my = s_div(my, mm); mz = s_div(mz, mm); }
            const unsigned int tag = internal_tag(a, d);
            const long long bits = (long long)(((unsigned long long)(unsigned int)tc.z << 32) | (unsigned long long)tag);
            volatile double4* o = (volatile double4*)&rec[tc.x];
            o->x = mx; o->y = my; o->z = mz; o->w = __longlong_as_double(bits);
            ((volatile double*)m)[tc.x] = mm;
        }
        __threadfence_block();
        __syncthreads();
    }
}

static int walk_args_fill(rebcu_handle* h, const rebcu_config* c, WalkArgs& a);

// Returns REBCU_OK with *used = false when this build has to be (or is better) done replicated.
static int tree_build_sharded(rebcu_handle* h, const rebcu_config* c, bool* used) {
    TreeBuffers& T = h->tree;
    *used = false;
    const int W = h->world, me = h->rank;
    if (!h->comm || W <= 1 || c->quadrupole || !T.shard_ok || T.shard_mode == 0) return REBCU_OK;
    if (T.shard_mode == 2 && h->N < (1ull << 18)) return REBCU_OK;
    TreeParams P;
    bool empty;
    { const int err = tree_prepare(h, c, P, &empty); if (err || empty) return err; }
    if (P.L0 < SH_LS + 1) return REBCU_OK;
    const uint64_t n = h->N;
    const uint32_t n_root = (uint32_t)(P.Nx * P.Ny * P.Nz);
    const uint64_t nb64 = (uint64_t)n_root << (3 * SH_LS);
    if (nb64 > (1ull << 22)) return REBCU_OK;          // too many root boxes for the bucket tables
    const uint32_t n_buckets = (uint32_t)nb64;
    P.bshift = 3 * (P.L0 - SH_LS);
    int err;
    if (T.sh_cap_n < T.cap_n) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        if ((err = ensure(h, &T.sh_keys, T.cap_n))) return err;
        if ((err = ensure(h, &T.sh_idx, T.cap_n))) return err;
        T.sh_cap_n = T.cap_n;
    }
    uint64_t level_words = 0;
    { uint64_t w = n_root; for (int d = 0; d <= SH_LS; d++) { level_words += w; w *= 8; } }
    const uint32_t top_cap = (uint32_t)(level_words - n_buckets);
    if (T.sh_buckets_cap < n_buckets) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        if ((err = ensure(h, &T.sh_hist, (size_t)n_buckets))) return err;
        if ((err = ensure(h, &T.sh_pstart, (size_t)n_buckets + 1))) return err;
        if ((err = ensure(h, &T.sh_tab, (size_t)n_buckets + 8 * (REBCU_MAX_RANKS + 1)))) return err;
        if ((err = ensure(h, &T.sh_level, (size_t)(3 * level_words)))) return err;
        if ((err = ensure(h, &T.sh_delta, (size_t)n_buckets))) return err;
        if ((err = ensure(h, &T.sh_top, (size_t)top_cap + 1))) return err;
        if (!T.sh_info) { if ((err = ensure(h, &T.sh_info, (size_t)I_WORDS))) return err; }
        T.sh_buckets_cap = n_buckets;
    }
    T.rec_ready = false; T.complete = false;
    const double *x = h->f(F_X), *y = h->f(F_Y), *z = h->f(F_Z), *m = h->f(F_M);
    uint32_t* pin = (uint32_t*)h->pinned;
    // ---- 1. all keys, bucket histogram, splitters ----
    {
        LaunchScope ls(h, TC_TREEBUILD, 4);
        CU_TRY(h, cudaMemsetAsync(T.flags, 0x7f, 8 * sizeof(int), h->stream));
        CU_TRY(h, cudaMemsetAsync(T.sh_hist, 0, (size_t)n_buckets * sizeof(uint32_t), h->stream));
        CU_TRY(h, cudaMemsetAsync(T.sh_info, 0, I_WORDS * sizeof(uint32_t), h->stream));
        key_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(P, x, y, z, T.keys, T.perm_in, T.flags);
        bucket_hist_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(n, T.keys, P.bshift, T.sh_hist);
        bucket_split_kernel<<<1, 1024, 0, h->stream>>>(n_buckets, T.sh_hist, W, T.sh_pstart, T.sh_info);
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(pin, T.sh_info, I_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    uint64_t Bk[REBCU_MAX_RANKS + 1], Pk[REBCU_MAX_RANKS + 1], tabB[REBCU_MAX_RANKS + 1];
    for (int r = 0; r <= W; r++) { Bk[r] = pin[I_B + r]; Pk[r] = pin[I_P + r]; tabB[r] = Bk[r] + 8ull * r; }
    tabB[W] = Bk[W] + 8ull * W;
    for (int r = 0; r <= W; r++) T.sh_pk[r] = Pk[r];
    const uint64_t n_loc = Pk[me + 1] - Pk[me];
    uint64_t* keys_loc = T.keys_sorted + Pk[me];
    uint32_t* perm_loc = T.perm + Pk[me];
    TreeParams L = P;
    L.n = n_loc; L.clamp = SH_LS - 1;
    uint32_t* tab_mine = (uint32_t*)(T.sh_tab + tabB[me] + 8);       // entry of bucket b: tab_mine[2 (b - B_me)]; indexed with b below
    // ---- 2. this rank's key range: compact, sort, ties, lcp, cell counts ----
    {
        LaunchScope ls(h, TC_TREEBUILD, 10);
        bucket_flag_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(n, T.keys, P.bshift, (uint32_t)Bk[me], (uint32_t)Bk[me + 1], T.cell_cnt);
        prim::exclusive_scan_u32(h->stream, T.cell_cnt, T.cell_off, n, (uint32_t*)T.scan_tmp);
        bucket_scatter_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(n, T.keys, T.cell_cnt, T.cell_off, T.sh_keys, T.sh_idx);
        // status words (error flags so far and those of the kernels below) + this rank's table entries
        CU_TRY(h, cudaMemsetAsync(T.sh_tab + tabB[me], 0, (size_t)(tabB[me + 1] - tabB[me]) * sizeof(uint64_t), h->stream));
        if (n_loc) {
            h->launches += prim::radix_sort_pairs(h->stream, T.sh_keys, T.sh_idx, keys_loc, perm_loc, n_loc, P.rbits + 3 * P.L0,
                                                  (uint32_t*)T.sort_tmp, (uint32_t*)T.scan_tmp);
            tie_kernel<<<div_up(n_loc, 256), 256, 0, h->stream>>>(L, keys_loc, perm_loc, x, y, z, T.flags);
            lcp_kernel<<<div_up(n_loc + 1, 256), 256, 0, h->stream>>>(L, keys_loc, perm_loc, x, y, z, T.lcp, T.flags);
            count_kernel<<<div_up(n_loc, 256), 256, 0, h->stream>>>(n_loc, T.lcp, T.cell_cnt);
        }
        CU_TRY(h, cudaMemsetAsync(T.cell_cnt + n_loc, 0, sizeof(uint32_t), h->stream));
        prim::exclusive_scan_u32(h->stream, T.cell_cnt, T.cell_off, n_loc + 1, (uint32_t*)T.scan_tmp);
        if (n_loc) bucket_bounds_kernel<<<div_up(n_loc, 256), 256, 0, h->stream>>>(n_loc, keys_loc, P.bshift, T.cell_off, tab_mine - 2 * Bk[me]);
        // flags -> the six status words in front of this rank's entries (32-bit values in 64-bit slots)
        CU_TRY(h, cudaMemcpy2DAsync(T.sh_tab + tabB[me], sizeof(uint64_t), T.flags, sizeof(int), sizeof(int), 6, cudaMemcpyDeviceToDevice, h->stream));
    }
    CU_TRY(h, cudaGetLastError());
    // ---- 3. gather the table, lay out the top tree ----
    { void* ptr = T.sh_tab; int bytes = 8; if ((err = comm_gather_ranges(h, &ptr, &bytes, 1, tabB))) return err; }
    TopArgs A;
    A.n_root = n_root; A.n_buckets = n_buckets; A.W = W; A.hist = T.sh_hist; A.tab = T.sh_tab; A.level = T.sh_level;
    A.level_words = level_words; A.delta = T.sh_delta; A.top = T.sh_top; A.info = T.sh_info; A.top_cap = top_cap;
    {
        LaunchScope ls(h, TC_TREEBUILD, 1);
        top_layout_kernel<<<1, 1024, 0, h->stream>>>(A);
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(pin, T.sh_info, I_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    {
        // every rank sees the same flags (minimum over the ranks) and takes the same decision
        const int* fl = (const int*)(pin + I_FLAGS);
        if (P.prefix && fl[4] != 0x7f7f7f7f) { T.prefix_ok = false; return tree_build_sharded(h, c, used); }
        if (fl[5] != 0x7f7f7f7f) { T.shard_ok = false; return REBCU_OK; }          // very deep tree: replicated build
        int f[4];
        for (int k = 0; k < 4; k++) f[k] = (fl[k] == 0x7f7f7f7f) ? 0x7fffffff : fl[k];
        if (f[0] != 0x7fffffff || f[1] != 0x7fffffff || f[2] != 0x7fffffff || f[3] != 0x7fffffff) return tree_error(h, f);
    }
    const uint64_t n_total = pin[I_NTOT];
    const uint32_t n_top = pin[I_NTOP];
    uint64_t Sk[REBCU_MAX_RANKS + 1];
    for (int r = 0; r <= W; r++) Sk[r] = pin[I_S + r];
    if ((err = tree_cell_capacity(h, n_total))) return err;
    if (T.walk_rec_cap < T.cap_cells) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(T.walk_rec); cudaFree(T.walk_m); T.walk_rec = nullptr; T.walk_m = nullptr; T.walk_rec_cap = 0;
        CU_TRY(h, cudaMalloc(&T.walk_rec, T.cap_cells * sizeof(double4)));
        CU_TRY(h, cudaMalloc(&T.walk_m, T.cap_cells * sizeof(double)));
        T.walk_rec_cap = T.cap_cells;
    }
    WalkArgs a;
    walk_args_fill(h, c, a);
    CellArrays C{T.walk_pos, T.walk_geo, (int4*)T.walk_meta, T.parent, T.ready, T.walk_meta2, nullptr, 0};
    T.has_quad = false;
    const uint64_t s0 = Sk[me], sn = Sk[me + 1] - Sk[me];
    // ---- 4. this rank's subtrees at their global indices: emit, moments, traversal records ----
    if (sn) {
        LaunchScope ls(h, TC_TREEBUILD, 7);
        // cells of the span that no subtree covers (top cells) must look inert to the passes below
        CU_TRY(h, cudaMemsetAsync((int4*)T.walk_meta + s0, 0x80, sn * sizeof(int4), h->stream));
        CU_TRY(h, cudaMemsetAsync(T.walk_meta2 + s0, 0, sn * sizeof(int2), h->stream));
        CU_TRY(h, cudaMemsetAsync(T.parent + s0, 0xff, sn * sizeof(int32_t), h->stream));
        CU_TRY(h, cudaMemsetAsync(T.ready + s0, 0, sn * sizeof(uint32_t), h->stream));
        L.bdelta = T.sh_delta;
        if (n_loc) emit_kernel<<<div_up(n_loc, 128), 128, 0, h->stream>>>(L, keys_loc, perm_loc, T.lcp, T.cell_off, x, y, z, m, C);
        adopt_kernel<<<div_up(sn, 256), 256, 0, h->stream>>>(s0, sn, C);
        moment_kernel<<<div_up(sn, 256), 256, 0, h->stream>>>(s0, sn, C);
        walk_pack_kernel<<<div_up(sn, 256), 256, 0, h->stream>>>(s0, sn, T.walk_pos, T.walk_meta2, T.walk_rec, T.walk_m, a);
    }
    CU_TRY(h, cudaGetLastError());
    // ---- 5. gather records, masses and the sorted permutation; fill in the top cells ----
    { void* ptrs[2] = {T.walk_rec, T.walk_m}; int bytes[2] = {32, 8}; if ((err = comm_gather_ranges(h, ptrs, bytes, 2, Sk))) return err; }
    { void* ptr = T.perm; int bytes = 4; if ((err = comm_gather_ranges(h, &ptr, &bytes, 1, Pk))) return err; }
    if (n_top) {
        LaunchScope ls(h, TC_TREEBUILD, 1);
        top_moment_kernel<<<1, 1024, 0, h->stream>>>(T.sh_top, n_top, T.walk_rec, T.walk_m, a);
    }
    CU_TRY(h, cudaGetLastError());
    T.n_cells = n_total;
    T.built_for_n = (int)n;
    T.rec_ready = true;
    T.last_build_cells_local = sn;
    *used = true;
    return REBCU_OK;
}

extern "C" int rebcu_set_sharded_build(rebcu_handle* h, int mode) {
    if (group_active(h)) return group_run(h, [mode](rebcu_handle* s, int) { return rebcu_set_sharded_build(s, mode); });
    if (mode < 0 || mode > 2) return rebcu_fail(h, REBCU_ERR_ARG, "sharded build mode: 0 never, 1 whenever possible, 2 automatic");
    h->tree.shard_mode = mode;
    return REBCU_OK;
}

// Copies the tree out as rebcu_treecell records (parity tests).
int tree_export(rebcu_handle* h) {
    TreeBuffers& T = h->tree;
    if (T.n_cells == 0) return REBCU_OK;
    if (!T.complete) return rebcu_fail(h, REBCU_ERR_ARG, "the last tree was built per rank (sharded); call rebcu_tree_build for a complete cell array");
    if (!T.cells) CU_TRY(h, cudaMalloc(&T.cells, T.cap_cells * sizeof(rebcu_treecell)));
    CellArrays C{T.walk_pos, T.walk_geo, (int4*)T.walk_meta, T.parent, T.ready, T.walk_meta2, nullptr, 0};
    export_kernel<<<div_up(T.n_cells, 256), 256, 0, h->stream>>>(T.n_cells, C, T.cells);
    CU_TRY(h, cudaGetLastError());
    return REBCU_OK;
}

// Sorted positions (indices into perm) of the particles this rank owns, in key order.  Valid until the next build.
int tree_shard_list(rebcu_handle* h, const uint32_t** list, uint64_t* n_work) {
    TreeBuffers& T = h->tree;
    const uint64_t n = h->N;
    uint64_t b, e; engine_shard(h, &b, &e);
    uint32_t* flag = T.cell_cnt; uint32_t* pos = T.cell_off;      // reuse (build is finished)
    {
        LaunchScope ls(h, TC_TREEWALK, 3);
        shard_flag_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(n, T.perm, (uint32_t)b, (uint32_t)e, flag);
        prim::exclusive_scan_u32(h->stream, flag, pos, n, (uint32_t*)T.scan_tmp);
        shard_list_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(n, flag, pos, T.shard_list);
    }
    CU_TRY(h, cudaGetLastError());
    *list = T.shard_list; *n_work = e - b;
    return REBCU_OK;
}

// accelerations in sorted order -> the rank's own index block
__global__ void __launch_bounds__(256) acc_scatter_kernel(uint64_t n, const uint32_t* __restrict__ perm, const double* __restrict__ sorted,
                                                          uint64_t stride, uint64_t b, uint64_t e, double* ax, double* ay, double* az) {
    const uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const uint32_t i = perm[k];
    if (i >= b && i < e) { ax[i] = sorted[k]; ay[i] = sorted[stride + k]; az[i] = sorted[2 * stride + k]; }
}

// Key-range share of a sharded walk: sorted positions [kb[r], kb[r+1]) belong to rank r (the ranges of the sharded
// build, else equal parts).
static void walk_key_ranges(rebcu_handle* h, uint64_t* kb) {
    TreeBuffers& T = h->tree;
    const int W = h->world;
    for (int r = 0; r <= W; r++) kb[r] = (T.rec_ready && !T.complete) ? T.sh_pk[r] : h->N * (uint64_t)r / (uint64_t)W;
}

static int walk_args_fill(rebcu_handle* h, const rebcu_config* c, WalkArgs& a) {
    TreeBuffers& T = h->tree;
    a.pos = T.walk_pos; a.meta = (const int4*)T.walk_meta; a.meta2 = T.walk_meta2; a.n_cells = T.n_cells;
    a.perm = T.perm; a.list = nullptr; a.n_work = h->N;
    a.quad = nullptr; a.quad_stride = 0; a.rec = nullptr; a.m = nullptr; a.w2_lo = 0;
    a.k_begin = 0; a.sorted_out = nullptr; a.sorted_stride = 0;
    a.x = h->f(F_X); a.y = h->f(F_Y); a.z = h->f(F_Z);
    a.ax = h->f(F_AX); a.ay = h->f(F_AY); a.az = h->f(F_AZ);
    a.ghosts = h->ghosts_dev;
    a.G = c->G; a.soft2 = c->softening * c->softening; a.theta2 = c->opening_angle2;
    a.root_size = c->root_size;
    a.windowed = strict_window_ok(c->G) ? 1 : 0;
    double w = c->root_size;
    for (int d = 0; d < W_TABLE; d++) { a.w2[d] = w * w; w = w / 2.; }
    { const double w20 = a.w2[0]; unsigned long long u; memcpy(&u, &w20, 8); a.w2_lo = (uint32_t)u; }
    // REBOUND_B200_GW_STACK=<n> shrinks the group walk's traversal stack (tests of the overflow path)
    static const int gw_stack = [] { const char* e = getenv("REBOUND_B200_GW_STACK"); const int v = e ? atoi(e) : 0;
                                     return (v >= 2 && v <= GW_STACK_MAX) ? v : GW_STACK_MAX; }();
    a.gw_stack = gw_stack;
    // a group whose list passes ~6 typical per-particle lists (60 log2 N entries at theta^2 = 0.25, ~theta^-3) is cheaper
    // to finish particle by particle; REBOUND_B200_GW_ABORT=<entries> overrides
    static const long abort_forced = [] { const char* e = getenv("REBOUND_B200_GW_ABORT"); return e ? atol(e) : 0l; }();
    double est = 360.0 * log2((double)(h->N > 2 ? h->N : 2));
    if (c->opening_angle2 > 0.0 && c->opening_angle2 < 0.25) est *= pow(0.25 / c->opening_angle2, 1.5);
    if (!(c->opening_angle2 > 0.0)) est = 4e9;
    a.gw_abort = abort_forced > 0 ? (unsigned int)abort_forced : (unsigned int)(est < 4096.0 ? 4096.0 : (est > 4e9 ? 4e9 : est));
    return REBCU_OK;
}

// (mx,my,mz | tag, skip) traversal records + masses of the current tree (walk_pack_kernel)
static int walk_records(rebcu_handle* h, WalkArgs& a) {
    TreeBuffers& T = h->tree;
    if (T.rec_ready) { a.rec = T.walk_rec; a.m = T.walk_m; return REBCU_OK; }      // the sharded build leaves them complete
    if (T.walk_rec_cap < T.cap_cells) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(T.walk_rec); cudaFree(T.walk_m); T.walk_rec = nullptr; T.walk_m = nullptr; T.walk_rec_cap = 0;
        CU_TRY(h, cudaMalloc(&T.walk_rec, T.cap_cells * sizeof(double4)));
        CU_TRY(h, cudaMalloc(&T.walk_m, T.cap_cells * sizeof(double)));
        T.walk_rec_cap = T.cap_cells;
    }
    h->launches++;
    a.rec = T.walk_rec; a.m = T.walk_m;
    walk_pack_kernel<<<div_up(T.n_cells, 256), 256, 0, h->stream>>>(0, T.n_cells, T.walk_pos, T.walk_meta2, T.walk_rec, T.walk_m, a);
    T.rec_ready = true;
    return REBCU_OK;
}

int tree_gravity(rebcu_handle* h, rebcu_config* c) {
    // gravity.c:56.  Every rank holds all positions at this point (they were exchanged after the drift), and
    // the replicated tree needs all of them wrapped, so the check covers the full range on every rank.
    int err = boundary_check_full(h, c);
    if (err) return err;
    bool sharded_build = false;
    err = tree_build_sharded(h, c, &sharded_build);       // several ranks: every rank builds the subtrees of its key range
    if (err) return err;
    if (!sharded_build) err = tree_build(h, c);           // gravity.c:63-71
    if (err) return err;
    const uint64_t n = h->N;
    if (n == 0) return REBCU_OK;
    TreeBuffers& T = h->tree;
    GhostShifts g;
    engine_ghost_shifts(c, c->N_ghost_x, c->N_ghost_y, c->N_ghost_z, &g);
    err = engine_upload_ghosts(h, &g);
    if (err) return err;
    WalkArgs a;
    walk_args_fill(h, c, a);
    static const int variant = [] { const char* e = getenv("REBOUND_B200_WALK");
                                    return (e && strcmp(e, "coop") == 0) ? 2 : (e && strcmp(e, "v1") == 0) ? 1 : (e && strcmp(e, "rec") == 0) ? 3 : 0; }();
    uint64_t kb[REBCU_MAX_RANKS + 1];
    const bool by_key = h->world > 1 && h->comm != nullptr;
    if (by_key) {
        // Several ranks with the native exchange: each walks a contiguous range of SORTED positions -- a compact region of
        // space, so that the 32 particles of a group (and of a warp of the per-particle walks) are neighbours; the
        // particles of an index block are scattered all over the box, which lengthened the group lists by a quarter on two
        // ranks already.  The accelerations come out in sorted order, are all-gathered (24 B per particle, like the
        // positions) and every rank picks those of its index block.  Same bits: a particle's sum does not depend on who
        // computes it.
        if (T.acc_sorted_cap < T.cap_n) {
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            cudaFree(T.acc_sorted); T.acc_sorted = nullptr; T.acc_sorted_cap = 0;
            CU_TRY(h, cudaMalloc(&T.acc_sorted, 3 * T.cap_n * sizeof(double)));
            T.acc_sorted_cap = T.cap_n;
        }
        walk_key_ranges(h, kb);
        a.k_begin = kb[h->rank]; a.n_work = kb[h->rank + 1] - kb[h->rank];
        a.sorted_out = T.acc_sorted; a.sorted_stride = T.acc_sorted_cap;
    } else if (h->world > 1) {
        // exchange through the caller's callback: walk the particles of this rank's index block, visited in key order
        if ((err = tree_shard_list(h, &a.list, &a.n_work))) return err;
    }
    if (a.n_work) {
        LaunchScope ls(h, TC_TREEWALK);
        // REBOUND_B200_WALK selects a walk for A/B runs: unset = records walk (walk_rec_kernel; in FAST mode the group
        // walk, walk_group_kernel), rec = records walk also in FAST mode, v1 = one visited cell per trip on the build's
        // arrays (walk_kernel), coop = warp-cooperative with per-lane lists (walk_coop_kernel; FP64-issue bound with
        // 18.7 of 32 lanes active, profiles/r01_walk_coop_ncu.txt).  All but the group walk give identical bits.
        const unsigned int nb = div_up(a.n_work, 128);
        if (T.has_quad) {
            a.quad = T.quad; a.quad_stride = T.quad_cap;
            if (c->mode == REBCU_MODE_FAST) walk_quad_kernel<true><<<nb, 128, 0, h->stream>>>(a);
            else walk_quad_kernel<false><<<nb, 128, 0, h->stream>>>(a);
        } else if (variant == 0 || variant == 3) {
            if ((err = walk_records(h, a))) return err;
            // the group walk pays without ghost boxes; with them every group would traverse the tree once per box (25 times
            // for the shearing sheet: 2.8 ms against 1.8 ms per-particle at N = 2^20), and a particle's ghost images would
            // need a per-entry identity test
            if (c->mode == REBCU_MODE_FAST && variant == 0 && g.n == 1) {
                CU_TRY(h, cudaMemsetAsync(h->counters + 8, 0, 3 * sizeof(unsigned long long), h->stream));
                // groups that give up are listed in the (now unused) cell-count array and finished by walk_retry_kernel
                unsigned int* retry = T.cell_cnt;
                CU_TRY(h, cudaMemsetAsync(retry, 0, sizeof(unsigned int), h->stream));
                static const int gv = [] { const char* e = getenv("REBOUND_B200_GW_VARIANT"); return e ? (e[0] | 32) - 'a' : -1; }();
                const unsigned int ng = div_up(a.n_work, 32);
                h->launches++;
                switch (gv) {
                    case 1:  walk_group_kernel<1, 4, 20, 160, 352><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 2:  walk_group_kernel<1, 4, 32, 96, 224><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 3:  walk_group_kernel<2, 4, 12, 160, 352><<<div_up(ng, 2), 64, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 4:  walk_group_kernel<1, 8, 16, 160, 352><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 5:  walk_group_kernel<4, 4, 8, 96, 224><<<div_up(ng, 4), 128, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 6:  walk_group_kernel<1, 2, 32, 96, 224><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 0:  walk_group_kernel<4, 4, 5, 160, 352><<<div_up(ng, 4), 128, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    // shapes a-g measured at N = 2^20 / 2^22 (profiles/r02_walk_group_shapes.txt): all within 8 % -- the kernel is
                    // bound by its instruction mix, not by occupancy.  h-q (profiles/r02_walk_group_ab.txt): paired evaluation
                    // -7 %, exact group criterion -16 % (2^24) ... -31 % (2^20) and no group gives up any more.  The 15-instruction
                    // pair term (h-p) would take another 6 % but misses the 1e-12 of the theta = 0 tests (the hardware seed is
                    // good to 2^-19.5 only), so the shipped kernel keeps the 16-instruction term of the direct kernels.
                    case 17: walk_group_kernel<1, 8, 16, 160, 352><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 18: walk_group_kernel<4, 4, 4, 160, 352, true, true, false><<<div_up(ng, 4), 128, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 19: walk_group_kernel<2, 8, 8, 160, 352, true, true, false><<<div_up(ng, 2), 64, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 20: walk_group_kernel<2, 2, 8, 160, 352, true, true, false><<<div_up(ng, 2), 64, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 21: walk_group_kernel<2, 4, 8, 224, 352, true, true, false><<<div_up(ng, 2), 64, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    // paired evaluation (two particles per lane, half the shared-memory loads; gw_evaluate_pair)
                    case 7:  walk_group_kernel<1, 4, 16, 160, 352, true><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 8:  walk_group_kernel<1, 2, 16, 160, 352, true><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 9:  walk_group_kernel<1, 4, 12, 160, 352, true><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 10: walk_group_kernel<2, 4, 8, 160, 352, true><<<div_up(ng, 2), 64, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    // exact group criterion (the intersection of the particles' own criteria) where the box cannot decide
                    case 11: walk_group_kernel<1, 8, 16, 160, 352, false, true><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 12: walk_group_kernel<1, 4, 16, 160, 352, true, true><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 13: walk_group_kernel<1, 2, 16, 160, 352, true, true><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 14: walk_group_kernel<2, 4, 8, 160, 352, true, true><<<div_up(ng, 2), 64, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    case 15: walk_group_kernel<1, 4, 12, 160, 352, true, true><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    // the same with the 16-instruction pair term of the direct kernels (full precision)
                    case 16: walk_group_kernel<1, 4, 16, 160, 352, true, true, false><<<ng, 32, 0, h->stream>>>(a, h->counters + 8, retry); break;
                    default: walk_group_kernel<2, 4, 8, 160, 352, true, true, false><<<div_up(ng, 2), 64, 0, h->stream>>>(a, h->counters + 8, retry); break;
                }
                walk_retry_kernel<<<div_up((uint64_t)ng * 32, 128), 128, 0, h->stream>>>(a, retry);
            } else {
                CU_TRY(h, cudaMemsetAsync(h->counters + 8, 0, 3 * sizeof(unsigned long long), h->stream));   // no group walk: its counters read 0
                if (c->mode == REBCU_MODE_FAST) walk_rec_kernel<true><<<nb, 128, 0, h->stream>>>(a);
                else walk_rec_kernel<false><<<nb, 128, 0, h->stream>>>(a);
            }
        } else if (variant == 1) {
            if (c->mode == REBCU_MODE_FAST) walk_kernel<true><<<nb, 128, 0, h->stream>>>(a);
            else walk_kernel<false><<<nb, 128, 0, h->stream>>>(a);
        } else {
            if (c->mode == REBCU_MODE_FAST) walk_coop_kernel<1><<<nb, 128, 0, h->stream>>>(a);
            else if (a.windowed) walk_coop_kernel<0><<<nb, 128, 0, h->stream>>>(a);
            else walk_coop_kernel<2><<<nb, 128, 0, h->stream>>>(a);
        }
    }
    CU_TRY(h, cudaGetLastError());
    if (by_key) {
        void* ptrs[3] = {T.acc_sorted, T.acc_sorted + T.acc_sorted_cap, T.acc_sorted + 2 * T.acc_sorted_cap};
        int bytes[3] = {8, 8, 8};
        if ((err = comm_gather_ranges(h, ptrs, bytes, 3, kb))) return err;
        uint64_t b, e; engine_shard(h, &b, &e);
        LaunchScope ls(h, TC_TREEWALK);
        acc_scatter_kernel<<<div_up(n, 256), 256, 0, h->stream>>>(n, T.perm, T.acc_sorted, T.acc_sorted_cap, b, e, a.ax, a.ay, a.az);
        CU_TRY(h, cudaGetLastError());
    }
    return REBCU_OK;
}

// Work counters of the tree walk on the CURRENT tree (call right after a TREE force evaluation with the same cfg):
//   out[0] interactions of the per-particle criterion (accepted cells + leaves, summed over this rank's particles and
//          ghost boxes: what the reference and the STRICT walk evaluate), out[1] cells those walks visit,
//   out[2] list entries summed over the groups of the last FAST group walk (each is evaluated by 32 lanes),
//   out[3] cells its traversals tested, out[4] number of groups, out[5] cells in the tree.
extern "C" int rebcu_tree_walk_stats(rebcu_handle* h, const rebcu_config* c, uint64_t* out6) {
    TreeBuffers& T = h->tree;
    for (int k = 0; k < 6; k++) out6[k] = 0;
    if (!h->resident || T.n_cells == 0 || T.built_for_n != (int)h->N) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no current tree: call after a TREE force evaluation");
    CU_TRY(h, cudaSetDevice(h->device));
    WalkArgs a;
    walk_args_fill(h, c, a);
    int err;
    if (h->world > 1 && h->comm) {
        uint64_t kb[REBCU_MAX_RANKS + 1];
        walk_key_ranges(h, kb);
        a.k_begin = kb[h->rank]; a.n_work = kb[h->rank + 1] - kb[h->rank];
    } else if (h->world > 1 && (err = tree_shard_list(h, &a.list, &a.n_work))) return err;
    if ((err = walk_records(h, a))) return err;
    unsigned long long host[8] = {0};
    CU_TRY(h, cudaMemcpyAsync(host + 2, h->counters + 8, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemsetAsync(h->counters + 12, 0, 2 * sizeof(unsigned long long), h->stream));
    if (a.n_work) walk_count_kernel<<<div_up(a.n_work, 128), 128, 0, h->stream>>>(a, h->counters + 12);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(host, h->counters + 12, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 5; k++) out6[k] = host[k];
    out6[5] = T.n_cells;
    return REBCU_OK;
}
