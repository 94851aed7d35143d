// collision.cu -- the search part of reb_collision_search (src/collision.c:49-331).
//
// DIRECT (collision.c:64-124): all ordered pairs (i, j != i) in every ghost box of the innermost ring;
//   a pair is reported when the spheres overlap (r2 <= (r_i+r_j)^2) and approach (dv.dx <= 0).
// TREE (collision.c:197-269, helper :422-503): per projectile, ghost box, root box: depth-first descent
//   pruned with  r2 < (r_i + r_2nd + 0.866 w)^2  on the GEOMETRIC cell centre; same leaf test.
//
// Both are two-pass (count -> exclusive scan -> fill) so that the list comes out in the order the
// reference's serial build produces it -- DIRECT: ghost box, projectile, target; TREE: projectile,
// ghost box, root box, pre-order -- because the list order feeds the rand_r shuffle and the
// order-dependent resolve loop (collision.c:336-404), which stay on the host.
// All predicates use strictly rounded arithmetic in the reference's expression order, so the pair set is
// bit-exact.  `ri` is written for TREE only (the reference leaves it uninitialised in DIRECT).
// r->map / r->N_map / r->N_targets subsets (collision.c:53-58; set by MERCURIUS/TRACE around encounter steps):
//   rebcu_set_collision_subset.  DIRECT maps projectile and target SLOTS through the map and skips equal slots,
//   LINE maps both sides and ignores N_targets, the tree modes only cut the projectile loop (as the reference does).
// Bound: DIRECT FP64 pipe (N^2 predicates); TREE L2/HBM latency on the cell arrays.
#include "engine.cuh"
#include "primitives.cuh"
#include <math.h>
#include <vector>

namespace {

struct ColSoa { const double *x, *y, *z, *vx, *vy, *vz, *r; };

// one LDG.E.256 per 32-byte cell record (half the L1 wavefronts of two 128-bit loads in a divergent walk)
__device__ __forceinline__ double4 ld256(const double4* p) {
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

ColSoa col_soa(const rebcu_handle* h) {
    return ColSoa{h->f(F_X), h->f(F_Y), h->f(F_Z), h->f(F_VX), h->f(F_VY), h->f(F_VZ), h->f(F_R)};
}

// collision.c:95-106 / :457-469
__device__ __forceinline__ bool hit(const rebcu_vec6d& s, double r1, double x2, double y2, double z2, double r2p,
                                    const ColSoa& P, uint32_t j) {
    const double dx = s_sub(s.x, x2), dy = s_sub(s.y, y2), dz = s_sub(s.z, z2);
    const double sr = s_add(r1, r2p);
    const double r2 = s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz));
    if (r2 > s_mul(sr, sr)) return false;
    const double dvx = s_sub(s.vx, P.vx[j]), dvy = s_sub(s.vy, P.vy[j]), dvz = s_sub(s.vz, P.vz[j]);
    if (s_add(s_add(s_mul(dvx, dx), s_mul(dvy, dy)), s_mul(dvz, dz)) > 0) return false;
    return true;
}

// LINE / LINETREE test: the straight-line trajectories over the last step come closer than r1 + r2
// (collision.c:155-177, :513-535).  MIN is the reference's macro ((a) > (b) ? (b) : (a)), collision.c:43.
__device__ __forceinline__ bool hit_line(const rebcu_vec6d& s, double r1, double x2, double y2, double z2, double r2p,
                                         const ColSoa& P, uint32_t j, double dt) {
    const double dx1 = s_sub(s.x, x2), dy1 = s_sub(s.y, y2), dz1 = s_sub(s.z, z2);
    const double r1sq = s_add(s_add(s_mul(dx1, dx1), s_mul(dy1, dy1)), s_mul(dz1, dz1));
    const double dvx = s_sub(s.vx, P.vx[j]), dvy = s_sub(s.vy, P.vy[j]), dvz = s_sub(s.vz, P.vz[j]);
    const double dx2 = s_sub(dx1, s_mul(dt, dvx)), dy2 = s_sub(dy1, s_mul(dt, dvy)), dz2 = s_sub(dz1, s_mul(dt, dvz));
    const double r2sq = s_add(s_add(s_mul(dx2, dx2), s_mul(dy2, dy2)), s_mul(dz2, dz2));
    const double tc = s_div(s_add(s_add(s_mul(dx1, dvx), s_mul(dy1, dvy)), s_mul(dz1, dvz)),
                            s_add(s_add(s_mul(dvx, dvx), s_mul(dvy, dvy)), s_mul(dvz, dvz)));
    double rmin2 = (r1sq > r2sq) ? r2sq : r1sq;
    const double frac = s_div(tc, dt);
    if (frac >= 0. && frac <= 1.) {
        const double dx3 = s_sub(dx1, s_mul(tc, dvx)), dy3 = s_sub(dy1, s_mul(tc, dvy)), dz3 = s_sub(dz1, s_mul(tc, dvz));
        const double r3sq = s_add(s_add(s_mul(dx3, dx3), s_mul(dy3, dy3)), s_mul(dz3, dz3));
        rmin2 = (rmin2 > r3sq) ? r3sq : rmin2;
    }
    const double rsum = s_add(r1, r2p);
    return !(rmin2 > s_mul(rsum, rsum));
}

__device__ __forceinline__ rebcu_vec6d shifted(const rebcu_vec6d& gb, const ColSoa& P, uint32_t i) {
    rebcu_vec6d s;
    s.x = s_add(gb.x, P.x[i]); s.y = s_add(gb.y, P.y[i]); s.z = s_add(gb.z, P.z[i]);
    s.vx = s_add(gb.vx, P.vx[i]); s.vy = s_add(gb.vy, P.vy[i]); s.vz = s_add(gb.vz, P.vz[i]);
    return s;
}

__device__ __forceinline__ void emit(rebcu_collision* out, uint64_t at, uint32_t p1, uint32_t p2, const rebcu_vec6d& gb, uint64_t ri) {
    rebcu_collision c;
    c.p1 = p1; c.p2 = p2; c.gb = gb; c.ri = ri;
    out[at] = c;
}

// ---- DIRECT ------------------------------------------------------------------------------------
// grid.y = ghost box, thread = projectile, targets tiled through shared memory.
// LINE: pairs j > i only, straight-line test over the last step (collision.c:152-189).
// A subset search (collision.c:53-58) runs over SLOTS: projectile slot i and target slot j stand for the particles
// map[i], map[j] (the slot itself without a map); only equal slots are skipped (:92), so a map that lists a
// particle twice reports it against itself, as the reference does.
struct DirectColArgs {
    const uint32_t* map;      // nullptr: slot == particle
    uint32_t n_targ;          // target slots [0, n_targ)   (LINE: [0, n_proj), collision.c:153)
    uint32_t ib, nloc;        // this rank's block of projectile slots
};

template <bool FILL, bool LINE>
__global__ void __launch_bounds__(128) direct_collision_kernel(ColSoa P, DirectColArgs a, const GhostShifts* ghosts,
                                                               uint32_t* __restrict__ count, const uint32_t* __restrict__ off,
                                                               rebcu_collision* __restrict__ out, double dt_last_done) {
    __shared__ double4 tile[128];
    __shared__ uint32_t tile_p[128];
    const uint32_t g = blockIdx.y;
    const uint32_t il = blockIdx.x * 128 + threadIdx.x;
    const uint32_t i = a.ib + il;                              // projectile slot
    const bool valid = il < a.nloc;
    const uint32_t n = a.n_targ;
    const rebcu_vec6d gb = ghosts->gb[g];
    rebcu_vec6d s = gb;
    double r1 = 0;
    uint32_t ip = 0;
    if (valid) { ip = a.map ? a.map[i] : i; s = shifted(gb, P, ip); r1 = P.r[ip]; }
    uint32_t found = 0;
    const uint64_t base = (FILL && valid) ? off[(uint64_t)g * a.nloc + il] : 0;
    for (uint32_t t0 = 0; t0 < n; t0 += 128) {
        __syncthreads();
        const uint32_t j0 = t0 + threadIdx.x;
        if (j0 < n) {
            const uint32_t jp = a.map ? a.map[j0] : j0;
            tile[threadIdx.x] = make_double4(P.x[jp], P.y[jp], P.z[jp], P.r[jp]);
            tile_p[threadIdx.x] = jp;
        }
        __syncthreads();
        const int jn = min(128u, n - t0);
        if (valid) {
            for (int jj = 0; jj < jn; jj++) {
                const uint32_t j = t0 + jj;
                if (LINE ? (j <= i) : (j == i)) continue;
                const double4 q = tile[jj];
                const uint32_t jp = tile_p[jj];
                if (LINE ? hit_line(s, r1, q.x, q.y, q.z, q.w, P, jp, dt_last_done) : hit(s, r1, q.x, q.y, q.z, q.w, P, jp)) {
                    if (FILL) emit(out, base + found, ip, jp, gb, 0);
                    found++;
                }
            }
        }
    }
    if (!FILL && valid) count[(uint64_t)g * a.nloc + il] = found;
}

// ---- TREE ---------------------------------------------------------------------------------------
constexpr int COL_W_TABLE = 64;

struct TreeColArgs {
    const double4* rec;          // one 32-byte record per cell: everything a visit needs (see col_pack_kernel)
    const int4* meta; uint32_t n_cells;      // meta[c].w = root box, read for reported hits only
    double w[COL_W_TABLE];       // cell width by depth (tree.c:100: exact halvings of root_size)
    double root_size;
    const uint32_t* perm; uint32_t n;
    const uint32_t* list; uint32_t n_work;       // sharded: work item t -> sorted position list[t] (else t)
    uint32_t n_proj;                             // only particles [0, n_proj) are projectiles (collision.c:229,:286 with r->map set)
    const GhostShifts* ghosts;
    double r2nd;          // TREE: radius of the second largest particle; LINETREE: maxdrift = dt_last_done*sqrt(max v^2)
    double dt_last_done;
};

// The search visits a cell for its geometric centre and width (internal) or for its particle (leaf), plus the meta
// word (leaf flag / depth, skip).  The build keeps these in three arrays (32-byte centre-of-mass record, 32-byte
// geometry record, 16-byte meta record), which costs two dependent gathers per visited cell; the width is an exact
// halving of root_size per level, so a table indexed by depth replaces it and ONE 32-byte record per cell is enough:
//   leaf      (x, y, z of the particle | pt, skip)            internal  (cx, cy, cz | -(depth+1), skip)
__global__ void __launch_bounds__(256) col_pack_kernel(uint32_t n_cells, const double4* __restrict__ pos, const double4* __restrict__ geo,
                                                       const int4* __restrict__ meta, double4* __restrict__ rec) {
    const uint32_t c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n_cells) return;
    const int4 mt = meta[c];                 // pt, skip, depth, rootbox
    const double4 q = (mt.x >= 0) ? pos[c] : geo[c];
    const int tag = (mt.x >= 0) ? mt.x : -(mt.z + 1);
    const long long bits = (long long)(((unsigned long long)(unsigned int)mt.y << 32) | (unsigned long long)(unsigned int)tag);
    rec[c] = make_double4(q.x, q.y, q.z, __longlong_as_double(bits));
}

__device__ __forceinline__ double col_cell_w(const TreeColArgs& a, int depth) {
    if (depth < COL_W_TABLE) return a.w[depth];
    double w = a.root_size;
    for (int d = 0; d < depth; d++) w = s_div(w, 2.);
    return w;
}

// One traversal per projectile.  Pass 0 counts the hits and parks the first COL_SLOTS of them (target, ghost
// box, root box packed in 8 bytes) in a per-projectile slot row; after the scan, tree_slots_kernel turns the
// parked hits into list entries without touching the tree again.  Only projectiles with more than COL_SLOTS
// hits (rare) walk a second time (pass 1) to write their entries directly.
constexpr int COL_SLOTS = 6;

template <int PASS, bool LINE>
__global__ void __launch_bounds__(128) tree_collision_kernel(ColSoa P, TreeColArgs a, uint32_t* __restrict__ count,
                                                             const uint32_t* __restrict__ off, rebcu_collision* __restrict__ out,
                                                             uint64_t* __restrict__ slots) {
    const uint32_t t = blockIdx.x * 128 + threadIdx.x;
    if (t >= a.n_work) return;
    const uint32_t i = a.perm[a.list ? a.list[t] : t];     // key order => neighbouring lanes walk neighbouring paths
    if (i >= a.n_proj) { if (PASS == 0) count[i] = 0; return; }
    if (PASS == 1 && count[i] <= COL_SLOTS) return;
    const double r1 = P.r[i];
    double reach;
    if (LINE) {   // p1_r_plus_dtv + maxdrift (collision.c:302, 557)
        const double vx = P.vx[i], vy = P.vy[i], vz = P.vz[i];
        const double v = s_sqrt(s_add(s_add(s_mul(vx, vx), s_mul(vy, vy)), s_mul(vz, vz)));
        reach = s_add(s_add(r1, s_mul(a.dt_last_done, v)), a.r2nd);
    } else reach = s_add(r1, a.r2nd);                 // collision.c:492
    uint32_t found = 0;
    const uint64_t base = (PASS == 1) ? off[i] : 0;
    const int ngb = a.ghosts->n;
    for (int g = 0; g < ngb; g++) {
        const rebcu_vec6d gb = a.ghosts->gb[g];
        const rebcu_vec6d s = shifted(gb, P, i);
        uint32_t c = 0;
        double4 q = ld256(a.rec);                                    // n_cells >= 1: the tree holds this projectile
        while (c < a.n_cells) {
            const long long bits = __double_as_longlong(q.w);
            const int tag = (int)(unsigned int)(unsigned long long)bits;
            const uint32_t skip = (uint32_t)((unsigned long long)bits >> 32);
            const uint32_t c0 = c;
            const double4 q0 = q;
            if (tag >= 0) {
                c = skip;
                if (c < a.n_cells) q = ld256(a.rec + c);             // next record in flight during the overlap test
                if ((uint32_t)tag != i) {
                    if (LINE ? hit_line(s, r1, q0.x, q0.y, q0.z, P.r[tag], P, (uint32_t)tag, a.dt_last_done)
                             : hit(s, r1, q0.x, q0.y, q0.z, P.r[tag], P, (uint32_t)tag)) {
                        const int rootbox = a.meta[c0].w;
                        if (PASS == 1) emit(out, base + found, i, (uint32_t)tag, gb, (uint64_t)rootbox);
                        else if (found < COL_SLOTS)
                            slots[(uint64_t)i * COL_SLOTS + found] = (uint64_t)(uint32_t)tag | ((uint64_t)g << 32) | ((uint64_t)rootbox << 40);
                        found++;
                    }
                }
            } else {
                const double dx = s_sub(s.x, q0.x), dy = s_sub(s.y, q0.y), dz = s_sub(s.z, q0.z);
                const double r2 = s_add(s_add(s_mul(dx, dx), s_mul(dy, dy)), s_mul(dz, dz));
                const double rp = s_add(reach, s_mul(0.86602540378443, col_cell_w(a, -tag - 1)));       // collision.c:492
                c = (r2 < s_mul(rp, rp)) ? c + 1 : skip;
                if (c < a.n_cells) q = ld256(a.rec + c);
            }
        }
    }
    if (PASS == 0) count[i] = found;
}

// list entries of the projectiles whose hits all fit their slot row; flags[0] is set if some projectile overflowed
__global__ void __launch_bounds__(256) tree_slots_kernel(uint32_t n, const uint32_t* __restrict__ count, const uint32_t* __restrict__ off,
                                                         const uint64_t* __restrict__ slots, const GhostShifts* ghosts,
                                                         rebcu_collision* __restrict__ out) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint32_t cnt = count[i];
    if (cnt == 0 || cnt > COL_SLOTS) return;
    const uint64_t base = off[i];
    for (uint32_t q = 0; q < cnt; q++) {
        const uint64_t e = slots[(uint64_t)i * COL_SLOTS + q];
        emit(out, base + q, i, (uint32_t)e, ghosts->gb[(e >> 32) & 0xff], e >> 40);
    }
}

__global__ void __launch_bounds__(256) overflow_flag_kernel(uint32_t n, const uint32_t* __restrict__ count, unsigned long long* flag) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const bool over = i < n && count[i] > COL_SLOTS;
    if (__any_sync(0xffffffffu, over) && (threadIdx.x & 31) == 0) atomicAdd(flag, 1ull);
}

// Radius of the second largest particle (reb_simulation_two_largest_particles, simulation.c:718-799) and
// max over particles of vx^2+vy^2+vz^2 (collision.c:273-277): only the VALUES are used, so the reductions are order
// independent.  Two stages: RED_BLOCKS blocks leave their partial results in `part`, one block combines them.
constexpr int RED_BLOCKS = 128;

// two largest of the multiset {a1 >= a2} U {b1 >= b2}
__device__ __forceinline__ void top2_merge(double& a1, double& a2, double b1, double b2) {
    const double m1 = a1 > b1 ? a1 : b1;
    const double lo = a1 > b1 ? b1 : a1;
    const double hi2 = a1 > b1 ? a2 : b2;
    a1 = m1; a2 = lo > hi2 ? lo : hi2;
}

__device__ __forceinline__ void top2_block(double& l1, double& l2, double* s1, double* s2) {
    for (int o = 16; o > 0; o >>= 1) top2_merge(l1, l2, __shfl_down_sync(0xffffffffu, l1, o), __shfl_down_sync(0xffffffffu, l2, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s1[w] = l1; s2[w] = l2; }
    __syncthreads();
    if (w == 0) {
        l1 = (l < (int)(blockDim.x >> 5)) ? s1[l] : -1.0; l2 = (l < (int)(blockDim.x >> 5)) ? s2[l] : -1.0;
        for (int o = 16; o > 0; o >>= 1) top2_merge(l1, l2, __shfl_down_sync(0xffffffffu, l1, o), __shfl_down_sync(0xffffffffu, l2, o));
    }
}

__global__ void __launch_bounds__(256) second_largest_kernel(const double* __restrict__ r, uint32_t n, double* __restrict__ part) {
    __shared__ double s1[8], s2[8];
    double l1 = -1.0, l2 = -1.0;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const double v = r[i];
        if (v > l1) { l2 = l1; l1 = v; } else if (v > l2) l2 = v;
    }
    top2_block(l1, l2, s1, s2);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = l1; part[2 * blockIdx.x + 1] = l2; }
}

__global__ void __launch_bounds__(RED_BLOCKS) second_largest_final_kernel(const double* __restrict__ part, int n_part, uint32_t n, double* out) {
    __shared__ double s1[RED_BLOCKS / 32], s2[RED_BLOCKS / 32];
    double l1 = -1.0, l2 = -1.0;
    if ((int)threadIdx.x < n_part) { l1 = part[2 * threadIdx.x]; l2 = part[2 * threadIdx.x + 1]; }
    top2_block(l1, l2, s1, s2);
    if (threadIdx.x == 0) out[0] = (n >= 2) ? l2 : 0.0;       // collision.c:222-225: no second particle, radius 0
}

__global__ void __launch_bounds__(256) vmax2_kernel(ColSoa P, uint32_t n, double* __restrict__ part) {
    __shared__ double sm[8];
    double m = 0.;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const double vx = P.vx[i], vy = P.vy[i], vz = P.vz[i];
        const double v2 = s_add(s_add(s_mul(vx, vx), s_mul(vy, vy)), s_mul(vz, vz));
        m = (m > v2) ? m : v2;                                // MAX(vmax2, v2), collision.c:44
    }
    for (int o = 16; o > 0; o >>= 1) { const double b = __shfl_down_sync(0xffffffffu, m, o); m = (m > b) ? m : b; }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int w = 1; w < 8; w++) m = (m > sm[w]) ? m : sm[w]; part[blockIdx.x] = m; }
}

__global__ void __launch_bounds__(RED_BLOCKS) vmax2_final_kernel(const double* __restrict__ part, int n_part, double* out) {
    __shared__ double sm[RED_BLOCKS];
    sm[threadIdx.x] = ((int)threadIdx.x < n_part) ? part[threadIdx.x] : 0.;
    __syncthreads();
    for (int w = RED_BLOCKS / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { const double a = sm[threadIdx.x], b = sm[threadIdx.x + w]; sm[threadIdx.x] = (a > b) ? a : b; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

int ensure_lists(rebcu_handle* h, uint64_t n_counts) {
    if (h->col_cap_n < n_counts + 1) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->col_count); cudaFree(h->col_off); cudaFree(h->col_scan_tmp);
        h->col_count = nullptr; h->col_off = nullptr; h->col_scan_tmp = nullptr;
        const uint64_t cap = n_counts + n_counts / 8 + 1024;
        CU_TRY(h, cudaMalloc(&h->col_count, cap * sizeof(uint32_t)));
        CU_TRY(h, cudaMalloc(&h->col_off, cap * sizeof(uint32_t)));
        const size_t tb = prim::scan_scratch_words(cap) * sizeof(uint32_t);
        CU_TRY(h, cudaMalloc(&h->col_scan_tmp, tb));
        h->col_scan_tmp_bytes = tb;
        h->col_cap_n = cap;
    }
    return REBCU_OK;
}

// scan the counts (n_counts entries + a zero sentinel) and fetch the total
int scan_counts(rebcu_handle* h, uint64_t n_counts, uint64_t* total) {
    CU_TRY(h, cudaMemsetAsync(h->col_count + n_counts, 0, sizeof(uint32_t), h->stream));
    prim::exclusive_scan_u32(h->stream, h->col_count, (uint32_t*)h->col_off, n_counts + 1, (uint32_t*)h->col_scan_tmp);
    uint32_t* pin = (uint32_t*)h->pinned;
    CU_TRY(h, cudaMemcpyAsync(pin, (uint32_t*)h->col_off + n_counts, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    *total = pin[0];
    if (*total > h->col_cap) {
        cudaFree(h->col_list); h->col_list = nullptr;
        const uint64_t cap = *total + *total / 4 + 1024;
        CU_TRY(h, cudaMalloc(&h->col_list, cap * sizeof(rebcu_collision)));
        h->col_cap = cap;
    }
    return REBCU_OK;
}

}  // namespace

int collision_search(rebcu_handle* h, const rebcu_config* c) {
    h->col_n = 0; h->col_seg_n = 0; h->col_seg_stride = 0;
    const uint64_t n = h->N;
    if (c->collision == REBCU_COLLISION_NONE || n == 0) return REBCU_OK;
    const bool direct = c->collision == REBCU_COLLISION_DIRECT || c->collision == REBCU_COLLISION_LINE;
    const bool line = c->collision == REBCU_COLLISION_LINE || c->collision == REBCU_COLLISION_LINETREE;
    if (!direct && c->collision != REBCU_COLLISION_TREE && c->collision != REBCU_COLLISION_LINETREE)
        return rebcu_fail(h, REBCU_ERR_ARG, "Collision routine not implemented.");
    // Sharded: the overlap tests read the target's velocity, which only its owner has kept current.
    if (h->world > 1) { const int xerr = engine_exchange(h, REBCU_EXCHANGE_POSITIONS | REBCU_EXCHANGE_VELOCITIES); if (xerr) return xerr; }
    if (n >= (1ull << 31)) return rebcu_fail(h, REBCU_ERR_ARG, "collision search supports N < 2^31");
    // r->map / r->N_map / r->N_targets (collision.c:53-58)
    const uint32_t* map = h->col_map_on ? h->col_map : nullptr;
    const uint64_t n_proj = h->col_map_on ? h->col_map_n : n;
    const uint64_t n_targ = h->col_targets != REBCU_SIZE_MAX ? h->col_targets : n_proj;
    if (h->col_map_on && h->col_map_n && h->col_map_max >= n) return rebcu_fail(h, REBCU_ERR_ARG, "collision subset: map entry >= N");
    if (n_targ > n_proj) return rebcu_fail(h, REBCU_ERR_ARG, "collision subset: N_targets exceeds the number of projectiles");
    if (!direct && n_proj > n) return rebcu_fail(h, REBCU_ERR_ARG, "collision subset: N_map exceeds N in a tree search");
    // this rank's block of projectile slots (the i-block of the force kernels when there is no subset)
    const uint64_t n_slots = direct ? n_proj : n;
    const uint64_t ib = n_slots * (uint64_t)h->rank / (uint64_t)h->world, ie = n_slots * (uint64_t)(h->rank + 1) / (uint64_t)h->world;
    const uint64_t nloc = ie - ib;
    // only the innermost ring of ghost boxes (collision.c:67-69, 214-216)
    GhostShifts g;
    engine_ghost_shifts(c, c->N_ghost_x > 1 ? 1 : c->N_ghost_x, c->N_ghost_y > 1 ? 1 : c->N_ghost_y,
                        c->N_ghost_z > 1 ? 1 : c->N_ghost_z, &g);
    int err;
    ColSoa P = col_soa(h);
    uint64_t total = 0;
    if (direct) {
        if ((err = engine_upload_ghosts(h, &g))) return err;
        const uint64_t n_counts = (uint64_t)g.n * nloc;
        h->col_seg_n = g.n; h->col_seg_stride = nloc;
        if (nloc == 0) return REBCU_OK;
        if ((err = ensure_lists(h, n_counts))) return err;
        dim3 grid(div_up(nloc, 128), g.n);
        const DirectColArgs da{map, (uint32_t)(line ? n_proj : n_targ), (uint32_t)ib, (uint32_t)nloc};
        {
            LaunchScope ls(h, TC_COLLISION, 2);
            if (line) direct_collision_kernel<false, true><<<grid, 128, 0, h->stream>>>(P, da, h->ghosts_dev, h->col_count, nullptr, nullptr, c->dt_last_done);
            else direct_collision_kernel<false, false><<<grid, 128, 0, h->stream>>>(P, da, h->ghosts_dev, h->col_count, nullptr, nullptr, 0.);
        }
        CU_TRY(h, cudaGetLastError());
        if ((err = scan_counts(h, n_counts, &total))) return err;
        if (total) {
            LaunchScope ls(h, TC_COLLISION);
            if (line) direct_collision_kernel<true, true><<<grid, 128, 0, h->stream>>>(P, da, h->ghosts_dev, nullptr, (const uint32_t*)h->col_off, h->col_list, c->dt_last_done);
            else direct_collision_kernel<true, false><<<grid, 128, 0, h->stream>>>(P, da, h->ghosts_dev, nullptr, (const uint32_t*)h->col_off, h->col_list, 0.);
        }
    } else {
        if ((err = tree_build(h, c))) return err;                       // collision.c:200
        if ((err = engine_upload_ghosts(h, &g))) return err;
        if ((err = ensure_lists(h, n))) return err;
        TreeBuffers& T = h->tree;
        TreeColArgs a;
        if (T.col_rec_cap < T.cap_cells) {
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            cudaFree(T.col_rec); T.col_rec = nullptr; T.col_rec_cap = 0;
            CU_TRY(h, cudaMalloc(&T.col_rec, T.cap_cells * sizeof(double4)));
            T.col_rec_cap = T.cap_cells;
        }
        {
            LaunchScope ls(h, TC_COLLISION);
            col_pack_kernel<<<div_up(T.n_cells, 256), 256, 0, h->stream>>>((uint32_t)T.n_cells, T.walk_pos, T.walk_geo, (const int4*)T.walk_meta, T.col_rec);
        }
        a.rec = T.col_rec; a.meta = (const int4*)T.walk_meta; a.n_cells = (uint32_t)T.n_cells;
        a.root_size = c->root_size;
        { double w = c->root_size; for (int d = 0; d < COL_W_TABLE; d++) { a.w[d] = w; w = w / 2.; } }
        a.perm = T.perm; a.n = (uint32_t)n; a.ghosts = h->ghosts_dev;
        a.list = nullptr; a.n_work = (uint32_t)n; a.n_proj = (uint32_t)n_proj;
        h->col_seg_n = 1; h->col_seg_stride = n;
        if (h->world > 1) {
            uint64_t nw = 0;
            if ((err = tree_shard_list(h, &a.list, &nw))) return err;
            a.n_work = (uint32_t)nw;
            CU_TRY(h, cudaMemsetAsync(h->col_count, 0, n * sizeof(uint32_t), h->stream));   // the other ranks' projectiles
        }
        {
            LaunchScope ls(h, TC_COLLISION, 2);
            double* part = h->scratch_big;        // 2 x RED_BLOCKS doubles of the 3584 (the fused test-particle step is not running)
            if (line) {                                                                           // collision.c:273-277
                const int nb = (int)min((unsigned int)RED_BLOCKS, max(1u, div_up(n_proj, 256)));
                vmax2_kernel<<<nb, 256, 0, h->stream>>>(P, (uint32_t)n_proj, part);
                vmax2_final_kernel<<<1, RED_BLOCKS, 0, h->stream>>>(part, nb, h->scratch);
            } else {
                const int nb = (int)min((unsigned int)RED_BLOCKS, max(1u, div_up(n, 256)));
                second_largest_kernel<<<nb, 256, 0, h->stream>>>(P.r, (uint32_t)n, part);
                second_largest_final_kernel<<<1, RED_BLOCKS, 0, h->stream>>>(part, nb, (uint32_t)n, h->scratch);
            }
        }
        double* pin = (double*)(h->pinned + 8);
        CU_TRY(h, cudaMemcpyAsync(pin, h->scratch, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        a.r2nd = line ? c->dt_last_done * sqrt(pin[0]) : pin[0];      // collision.c:278 / :222-225
        a.dt_last_done = c->dt_last_done;
        if (h->col_slots_cap < n) {
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            cudaFree(h->col_slots); h->col_slots = nullptr;
            const uint64_t cap = n + n / 8 + 1024;
            CU_TRY(h, cudaMalloc(&h->col_slots, cap * COL_SLOTS * sizeof(uint64_t)));
            h->col_slots_cap = cap;
        }
        CU_TRY(h, cudaMemsetAsync(h->counters + 4, 0, sizeof(unsigned long long), h->stream));
        {
            LaunchScope ls(h, TC_COLLISION, 2);
            if (line) tree_collision_kernel<0, true><<<div_up(max(a.n_work, 1u), 128), 128, 0, h->stream>>>(P, a, h->col_count, nullptr, nullptr, h->col_slots);
            else tree_collision_kernel<0, false><<<div_up(max(a.n_work, 1u), 128), 128, 0, h->stream>>>(P, a, h->col_count, nullptr, nullptr, h->col_slots);
            overflow_flag_kernel<<<div_up(n, 256), 256, 0, h->stream>>>((uint32_t)n, h->col_count, h->counters + 4);
        }
        CU_TRY(h, cudaGetLastError());
        CU_TRY(h, cudaMemcpyAsync(h->pinned + 16, h->counters + 4, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        if ((err = scan_counts(h, n, &total))) return err;          // synchronises the stream
        if (total) {
            LaunchScope ls(h, TC_COLLISION, 2);
            tree_slots_kernel<<<div_up(n, 256), 256, 0, h->stream>>>((uint32_t)n, h->col_count, (const uint32_t*)h->col_off, h->col_slots,
                                                                   h->ghosts_dev, h->col_list);
            if (h->pinned[16]) {     // some projectile has more hits than slots: those walk again and write directly
                if (line) tree_collision_kernel<1, true><<<div_up(max(a.n_work, 1u), 128), 128, 0, h->stream>>>(P, a, h->col_count, (const uint32_t*)h->col_off, h->col_list, nullptr);
                else tree_collision_kernel<1, false><<<div_up(max(a.n_work, 1u), 128), 128, 0, h->stream>>>(P, a, h->col_count, (const uint32_t*)h->col_off, h->col_list, nullptr);
            }
        }
    }
    CU_TRY(h, cudaGetLastError());
    h->col_n = total;
    return REBCU_OK;
}

// r->map / r->N_map / r->N_targets of the next searches (src/rebound.h:257-258,344; collision.c:53-58).
extern "C" int rebcu_set_collision_subset(rebcu_handle* h, const uint64_t* map, uint64_t N_map, uint64_t N_targets) {
    if (group_active(h)) return group_run(h, [=](rebcu_handle* s, int) { return rebcu_set_collision_subset(s, map, N_map, N_targets); });
    CU_TRY(h, cudaSetDevice(h->device));
    h->col_targets = N_targets;
    h->col_map_on = map != nullptr;
    h->col_map_n = map ? N_map : 0;
    h->col_map_max = 0;
    if (!map || N_map == 0) return REBCU_OK;
    if (N_map >= (1ull << 31)) return rebcu_fail(h, REBCU_ERR_ARG, "collision subset supports N_map < 2^31");
    std::vector<uint32_t> m32(N_map);
    for (uint64_t i = 0; i < N_map; i++) {
        if (map[i] >= (1ull << 31)) { h->col_map_on = false; h->col_map_n = 0; return rebcu_fail(h, REBCU_ERR_ARG, "collision subset: map entry >= 2^31"); }
        m32[i] = (uint32_t)map[i];
        if (map[i] > h->col_map_max) h->col_map_max = map[i];
    }
    if (h->col_map_cap < N_map) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->col_map); h->col_map = nullptr; h->col_map_cap = 0;
        const uint64_t cap = N_map + N_map / 4 + 256;
        CU_TRY(h, cudaMalloc(&h->col_map, cap * sizeof(uint32_t)));
        h->col_map_cap = cap;
    }
    // pageable source: the copy is staged before the call returns, so m32 may go out of scope
    CU_TRY(h, cudaMemcpyAsync(h->col_map, m32.data(), N_map * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return REBCU_OK;
}

extern "C" int rebcu_collisions_segments(rebcu_handle* h, uint64_t* counts, uint64_t cap, uint64_t* n_segments) {
    *n_segments = h->col_seg_n;
    if (h->col_seg_n == 0) return REBCU_OK;
    if (cap < h->col_seg_n) return rebcu_fail(h, REBCU_ERR_CAPACITY, "segment buffer too small");
    if (h->col_seg_n == 1) { counts[0] = h->col_n; return REBCU_OK; }
    if (h->col_n == 0) { for (uint64_t s = 0; s < h->col_seg_n; s++) counts[s] = 0; return REBCU_OK; }
    // DIRECT / LINE: entry offsets are the scanned counts, ghost box major: segment s starts at off[s * stride]
    CU_TRY(h, cudaSetDevice(h->device));
    uint32_t* pin = (uint32_t*)h->pinned;       // 32 words of 8 bytes >= 28 offsets
    for (uint64_t s = 0; s <= h->col_seg_n; s++)
        CU_TRY(h, cudaMemcpyAsync(pin + s, (const uint32_t*)h->col_off + s * h->col_seg_stride, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    for (uint64_t s = 0; s < h->col_seg_n; s++) counts[s] = pin[s + 1] - pin[s];
    return REBCU_OK;
}
