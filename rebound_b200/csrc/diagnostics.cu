// diagnostics.cu -- device-side versions of the O(N) / O(N^2) host loops users and parity checks call between
// steps, so that a resident simulation does not have to be downloaded for them:
//   reb_simulation_energy            src/tools.c:108-162   (kinetic over N_interact, potential over i<N_active, j>i)
//   reb_simulation_com               src/tools.c:401-408, 376-399 (mass-weighted means of x, v, a)
//   reb_simulation_angular_momentum  src/tools.c:164-174
//   exit_max_distance / exit_min_distance scans of run_heartbeat  src/simulation.c:242-272  (exact: flags only)
// The reference accumulates each of these in ONE scalar in index order; a parallel sum cannot reproduce that
// rounding sequence, so these are the only entry points of the library that are not bit-identical: they agree
// with the reference to ~1e-13 relative (tests state 1e-12) and are deterministic run to run (fixed partial
// layout, fixed-order final reduction, compensated summation throughout).
// Also here: the DFMA throughput probe bench.py uses as the measured FP64 peak.
#include "engine.cuh"

namespace {

constexpr int DIAG_SUMS = 14;   // e_kin, M, Mx My Mz, Mvx Mvy Mvz, Max May Maz, Lx Ly Lz
constexpr int DIAG_BLOCK = 256;

struct Kahan {
    double s = 0, c = 0;
    __device__ __forceinline__ void add(double v) {
        const double y = __dsub_rn(v, c), t = __dadd_rn(s, y);
        c = __dsub_rn(__dsub_rn(t, s), y); s = t;
    }
};

// block-wide sum in a fixed order (warp shuffle tree, then warp 0 over the warp results)
__device__ double block_sum(double v, double* sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (w == 0) for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;      // valid in thread 0
}

struct DiagSoa { const double *x, *y, *z, *vx, *vy, *vz, *ax, *ay, *az, *m; };

__global__ void __launch_bounds__(DIAG_BLOCK) moments_kernel(DiagSoa P, uint64_t n, uint64_t n_interact, double* __restrict__ partial) {
    __shared__ double sm[32];
    Kahan k[DIAG_SUMS];
    for (uint64_t i = (uint64_t)blockIdx.x * DIAG_BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * DIAG_BLOCK) {
        const double m = P.m[i], x = P.x[i], y = P.y[i], z = P.z[i], vx = P.vx[i], vy = P.vy[i], vz = P.vz[i];
        if (i < n_interact) k[0].add(0.5 * m * (vx * vx + vy * vy + vz * vz));
        k[1].add(m);
        k[2].add(m * x); k[3].add(m * y); k[4].add(m * z);
        k[5].add(m * vx); k[6].add(m * vy); k[7].add(m * vz);
        k[8].add(m * P.ax[i]); k[9].add(m * P.ay[i]); k[10].add(m * P.az[i]);
        k[11].add(m * (y * vz - z * vy)); k[12].add(m * (z * vx - x * vz)); k[13].add(m * (x * vy - y * vx));
    }
#pragma unroll
    for (int q = 0; q < DIAG_SUMS; q++) {
        const double v = block_sum(k[q].s, sm);
        if (threadIdx.x == 0) partial[(uint64_t)blockIdx.x * DIAG_SUMS + q] = v;
    }
}

// Potential energy: block (bx, by) handles targets i in its 128-wide tile against sources j of chunk `by`,
// j > i and j < n_interact only (tools.c:121-130).  Per pair: -G m_i m_j / sqrt(dx^2+dy^2+dz^2), unsoftened.
__global__ void __launch_bounds__(128) potential_kernel(DiagSoa P, uint64_t n_active, uint64_t n_interact, uint64_t j_chunk, double G,
                                                        double* __restrict__ partial) {
    __shared__ double4 tile[128];
    __shared__ double sm[32];
    const uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const bool valid = i < n_active;
    double xi = 0, yi = 0, zi = 0, mi = 0;
    if (valid) { xi = P.x[i]; yi = P.y[i]; zi = P.z[i]; mi = P.m[i]; }
    const uint64_t j_begin = (uint64_t)blockIdx.y * j_chunk;
    const uint64_t j_end = min(n_interact, j_begin + j_chunk);
    const uint64_t i_first = (uint64_t)blockIdx.x * 128;
    Kahan acc;
    // tiles entirely at or below the block's first target hold no pair j > i
    uint64_t t0 = j_begin;
    if (t0 + 128 <= i_first + 1) t0 = j_begin + ((i_first + 1 - j_begin) / 128) * 128;
    for (; t0 < j_end; t0 += 128) {
        __syncthreads();
        const uint64_t j0 = t0 + threadIdx.x;
        tile[threadIdx.x] = (j0 < j_end) ? make_double4(P.x[j0], P.y[j0], P.z[j0], P.m[j0]) : make_double4(0, 0, 0, 0);
        __syncthreads();
        const int jn = (int)min((uint64_t)128, j_end - t0);
        if (valid) {
            for (int jj = 0; jj < jn; jj++) {
                if (t0 + jj <= i) continue;
                const double4 q = tile[jj];
                const double dx = xi - q.x, dy = yi - q.y, dz = zi - q.z;
                acc.add(-(G * q.w * mi / sqrt(dx * dx + dy * dy + dz * dz)));
            }
        }
    }
    const double v = block_sum(acc.s, sm);
    if (threadIdx.x == 0) partial[(uint64_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// out[q] = sum over p of partial[p * stride + q], in index order, compensated
__global__ void __launch_bounds__(32) final_sum_kernel(const double* __restrict__ partial, uint64_t n_partial, int stride, int n_sums, double* out) {
    const int q = threadIdx.x;
    if (q >= n_sums) return;
    Kahan k;
    for (uint64_t p = 0; p < n_partial; p++) k.add(partial[p * stride + q]);
    out[q] = k.s;
}

__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double a, double b) {
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int k = 0; k < iters; k++) {
        r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
        r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
    const double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 12345.678) out[0] = s;       // never true; keeps the chain alive
}

int ensure_diag(rebcu_handle* h, uint64_t words) {
    if (h->diag_cap < words) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->diag_partial); h->diag_partial = nullptr; h->diag_cap = 0;
        CU_TRY(h, cudaMalloc(&h->diag_partial, words * sizeof(double)));
        h->diag_cap = words;
    }
    return REBCU_OK;
}

DiagSoa diag_soa(const rebcu_handle* h) {
    return DiagSoa{h->f(F_X), h->f(F_Y), h->f(F_Z), h->f(F_VX), h->f(F_VY), h->f(F_VZ), h->f(F_AX), h->f(F_AY), h->f(F_AZ), h->f(F_M)};
}

// the 14 moment sums into host memory
int moments(rebcu_handle* h, uint64_t n_interact, double* out14) {
    const uint64_t n = h->N;
    const unsigned int nb = (unsigned int)min((uint64_t)1184, max((uint64_t)1, (uint64_t)div_up(n, DIAG_BLOCK)));
    int err = ensure_diag(h, (uint64_t)nb * DIAG_SUMS + 64);
    if (err) return err;
    {
        LaunchScope ls(h, TC_PACK, 2);
        moments_kernel<<<nb, DIAG_BLOCK, 0, h->stream>>>(diag_soa(h), n, n_interact, h->diag_partial);
        final_sum_kernel<<<1, 32, 0, h->stream>>>(h->diag_partial, nb, DIAG_SUMS, DIAG_SUMS, h->scratch);
    }
    CU_TRY(h, cudaGetLastError());
    double* pin = (double*)h->pinned;      // 32 words
    CU_TRY(h, cudaMemcpyAsync(pin, h->scratch, DIAG_SUMS * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < DIAG_SUMS; q++) out14[q] = pin[q];
    return REBCU_OK;
}

// ---- exit conditions of run_heartbeat (simulation.c:242-272): pure predicates, strictly rounded, exact --------
// flags[0] = some particle has x^2+y^2+z^2 > max2 (:250-253)
__global__ void __launch_bounds__(256) escape_kernel(DiagSoa P, uint64_t n, double max2, unsigned int* __restrict__ flags) {
    bool any = false;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const double x = P.x[i], y = P.y[i], z = P.z[i];
        any |= s_add(s_add(s_mul(x, x), s_mul(y, y)), s_mul(z, z)) > max2;
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) flags[0] = 1u;
}

// flags[1] = some pair j < i has |x_i - x_j|^2 < min2 (:262-269).  Block (bx, by): particles i of tile bx against
// the j tiles by, by + gridDim.y, ... <= bx, staged in shared memory.
__global__ void __launch_bounds__(128) encounter_kernel(DiagSoa P, uint64_t n, double min2, unsigned int* __restrict__ flags) {
    __shared__ double sx[128], sy[128], sz[128];
    const uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const bool valid = i < n;
    const double xi = valid ? P.x[i] : 0., yi = valid ? P.y[i] : 0., zi = valid ? P.z[i] : 0.;
    bool any = false;
    for (uint64_t t = blockIdx.y; t <= blockIdx.x; t += gridDim.y) {
        const uint64_t j0 = t * 128 + threadIdx.x;
        __syncthreads();
        if (j0 < n) { sx[threadIdx.x] = P.x[j0]; sy[threadIdx.x] = P.y[j0]; sz[threadIdx.x] = P.z[j0]; }
        __syncthreads();
        if (!valid) continue;
        const uint64_t jb = t * 128;
        const int jn = (int)min((uint64_t)128, i > jb ? i - jb : 0);        // j < i only
        for (int jj = 0; jj < jn; jj++) {
            const double x = s_sub(xi, sx[jj]), y = s_sub(yi, sy[jj]), z = s_sub(zi, sz[jj]);
            any |= s_add(s_add(s_mul(x, x), s_mul(y, y)), s_mul(z, z)) < min2;
        }
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) flags[1] = 1u;
}

}  // namespace

extern "C" {

// The exit checks run_heartbeat makes after every step (src/simulation.c:242-272).
int rebcu_exit_check(rebcu_handle* h, double exit_max_distance, double exit_min_distance, int* escape, int* encounter) {
    if (group_active(h)) {
        // every rank scans all particles (positions are exchanged inside); the leader's verdict is returned
        int esc[REBCU_MAX_RANKS] = {0}, enc[REBCU_MAX_RANKS] = {0};
        int* pe = esc; int* pn = enc;
        const int err = group_run(h, [=](rebcu_handle* s, int r) { return rebcu_exit_check(s, exit_max_distance, exit_min_distance, &pe[r], &pn[r]); });
        *escape = esc[0]; *encounter = enc[0];
        return err;
    }
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    *escape = 0; *encounter = 0;
    const uint64_t n = h->N;
    // the reference tests the distances for truth (`if (r->exit_max_distance)`): zero switches a check off
    if (n == 0 || (!exit_max_distance && !exit_min_distance)) return REBCU_OK;
    if (h->world > 1) { const int xerr = engine_exchange(h, REBCU_EXCHANGE_POSITIONS); if (xerr) return xerr; }     // every rank scans all particles
    unsigned int* flags = (unsigned int*)(h->counters + 12);
    CU_TRY(h, cudaMemsetAsync(flags, 0, 2 * sizeof(unsigned int), h->stream));
    if (exit_max_distance) {
        LaunchScope ls(h, TC_BOUNDARY);
        escape_kernel<<<min(div_up(n, 256), 148u * 8u), 256, 0, h->stream>>>(diag_soa(h), n, exit_max_distance * exit_max_distance, flags);
    }
    if (exit_min_distance && n > 1) {
        LaunchScope ls(h, TC_BOUNDARY);
        const unsigned int gx = div_up(n, 128);
        const unsigned int gy = gx < 592u ? max(1u, min(gx, 592u / gx)) : 1u;
        encounter_kernel<<<dim3(gx, gy), 128, 0, h->stream>>>(diag_soa(h), n, exit_min_distance * exit_min_distance, flags);
    }
    CU_TRY(h, cudaGetLastError());
    unsigned int* pin = (unsigned int*)(h->pinned + 20);
    CU_TRY(h, cudaMemcpyAsync(pin, flags, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    *escape = pin[0] != 0; *encounter = pin[1] != 0;
    return REBCU_OK;
}

int rebcu_energy(rebcu_handle* h, const rebcu_config* cfg, double* out3) {
    GROUP_UNSUPPORTED(h, "rebcu_energy");
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    out3[0] = out3[1] = out3[2] = 0.;
    const uint64_t n = h->N;
    if (n == 0) return REBCU_OK;
    const uint64_t n_active = (cfg->N_active == REBCU_SIZE_MAX) ? n : min(cfg->N_active, n);
    const uint64_t n_interact = (cfg->testparticle_type == 0) ? n_active : n;       // tools.c:119
    double m14[DIAG_SUMS];
    int err = moments(h, n_interact, m14);
    if (err) return err;
    double e_pot = 0.;
    if (n_active > 0 && n_interact > 1) {
        const unsigned int gx = (unsigned int)div_up(n_active, 128);
        // enough j-chunks to fill the GPU when there are few targets (N_active << N with testparticle_type 1)
        uint64_t gy = max((uint64_t)1, (uint64_t)(4 * 148) / gx);
        gy = min(gy, (uint64_t)div_up(n_interact, 128));
        uint64_t j_chunk = (uint64_t)div_up(div_up(n_interact, gy), 128) * 128;
        gy = div_up(n_interact, j_chunk);
        if ((err = ensure_diag(h, (uint64_t)gx * gy + 64))) return err;
        {
            LaunchScope ls(h, TC_DIRECT, 2);
            potential_kernel<<<dim3(gx, (unsigned int)gy), 128, 0, h->stream>>>(diag_soa(h), n_active, n_interact, j_chunk, cfg->G, h->diag_partial);
            final_sum_kernel<<<1, 32, 0, h->stream>>>(h->diag_partial, (uint64_t)gx * gy, 1, 1, h->scratch);
        }
        CU_TRY(h, cudaGetLastError());
        double* pin = (double*)h->pinned;
        CU_TRY(h, cudaMemcpyAsync(pin, h->scratch, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        e_pot = pin[0];
    }
    out3[0] = m14[0]; out3[1] = e_pot; out3[2] = m14[0] + e_pot;
    return REBCU_OK;
}

int rebcu_com(rebcu_handle* h, double* out10) {
    GROUP_UNSUPPORTED(h, "rebcu_com");
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    for (int q = 0; q < 10; q++) out10[q] = 0.;
    if (h->N == 0) return REBCU_OK;
    double m14[DIAG_SUMS];
    int err = moments(h, 0, m14);
    if (err) return err;
    const double M = m14[1];
    out10[0] = M;
    for (int q = 0; q < 9; q++) out10[1 + q] = (M > 0.) ? m14[2 + q] / M : m14[2 + q];      // tools.c:387-398
    return REBCU_OK;
}

int rebcu_angular_momentum(rebcu_handle* h, double* out3) {
    GROUP_UNSUPPORTED(h, "rebcu_angular_momentum");
    if (!h->resident) return rebcu_fail(h, REBCU_ERR_NOT_RESIDENT, "no resident particles");
    CU_TRY(h, cudaSetDevice(h->device));
    out3[0] = out3[1] = out3[2] = 0.;
    if (h->N == 0) return REBCU_OK;
    double m14[DIAG_SUMS];
    int err = moments(h, 0, m14);
    if (err) return err;
    out3[0] = m14[11]; out3[1] = m14[12]; out3[2] = m14[13];
    return REBCU_OK;
}

// Sustained DFMA rate of this device in TFLOP/s (2 flop per DFMA): 8 independent chains per thread, 8 resident
// 256-thread CTAs per SM, timed with CUDA events on the handle's stream.
int rebcu_measure_fp64_peak(rebcu_handle* h, double* tflops) {
    CU_TRY(h, cudaSetDevice(h->device));
    int sms = 0;
    CU_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const int iters = 1 << 15, blocks = sms * 8;
    cudaEvent_t a, b;
    CU_TRY(h, cudaEventCreate(&a)); CU_TRY(h, cudaEventCreate(&b));
    dfma_probe_kernel<<<blocks, 256, 0, h->stream>>>(h->scratch, 1 << 10, 1.0000001, 1e-9);     // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CU_TRY(h, cudaEventRecord(a, h->stream));
        dfma_probe_kernel<<<blocks, 256, 0, h->stream>>>(h->scratch, iters, 1.0000001, 1e-9);
        CU_TRY(h, cudaEventRecord(b, h->stream));
        CU_TRY(h, cudaEventSynchronize(b));
        float ms = 0;
        CU_TRY(h, cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    h->launches += 4;
    cudaEventDestroy(a); cudaEventDestroy(b);
    CU_TRY(h, cudaGetLastError());
    *tflops = (double)blocks * 256.0 * 8.0 * iters * 2.0 / (best * 1e-3) / 1e12;
    return REBCU_OK;
}

}  // extern "C"
