// primitives.cuh -- hand-written device-wide exclusive scan and stable LSD radix sort (64-bit keys with
// 32-bit payloads) used by the octree build, the collision list assembly and the open-boundary compaction.
//
// Scan:  three launches -- per-tile sums, scan of the tile sums (one CTA), per-tile scan + offset.
//        HBM traffic 12 B per element (read, read, write).
// Sort:  8-bit digits, least significant first; per pass
//          radix_hist_kernel     per-tile digit histograms, stored digit-major  hist[digit][tile]
//          exclusive scan        over the flattened histogram = global start of every (digit, tile) bucket
//          radix_scatter_kernel  re-reads the tile, ranks every element stably inside its tile
//                                (warp match_any + per-warp running counters + cross-warp prefix), reorders the
//                                tile by digit in shared memory, then writes each digit run to its bucket, so the
//                                global stores are coalesced runs instead of a 256-way scatter.
//        Stable, so equal keys keep their index order (the duplicate detection in tree.cu relies on it).
//        HBM traffic per pass 8 B (histogram) + 12 B + 12 B per element.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prim {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 2048

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    return v;
}

// CTA-wide exclusive scan of one value per thread (SCAN_THREADS threads); returns the exclusive prefix and
// writes the CTA total to *total.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t s = (lane < SCAN_THREADS / 32) ? warp_sums[lane] : 0;
        const uint32_t si = warp_incl_scan(s, lane);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1) block_total = si;
    }
    __syncthreads();
    const uint32_t r = warp_sums[w] + incl - v;
    *total = block_total;
    __syncthreads();
    return r;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ sums) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x; if (i < n) s += in[i]; }
    uint32_t total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// exclusive scan of m values by one CTA (in place)
static __global__ void __launch_bounds__(SCAN_THREADS) scan_small_kernel(uint32_t* __restrict__ v, uint64_t m) {
    uint32_t carry = 0;
    for (uint64_t base = 0; base < m; base += SCAN_THREADS) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t x = (i < m) ? v[i] : 0;
        uint32_t total;
        const uint32_t ex = block_excl_scan(x, &total);
        if (i < m) v[i] = carry + ex;
        carry += total;
    }
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n,
                                                                  const uint32_t* __restrict__ tile_off) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;   // blocked: 8 consecutive items per thread
    uint32_t x[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { x[k] = (base + k < n) ? in[base + k] : 0; s += x[k]; }
    uint32_t total;
    uint32_t ex = block_excl_scan(s, &total) + tile_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += x[k]; }
}

// scratch needed (uint32 words) for a scan of n elements: tile sums of the level, recursively
inline size_t scan_scratch_words(uint64_t n) {
    size_t words = 0;
    uint64_t m = (n + SCAN_TILE - 1) / SCAN_TILE;
    words += m;
    return words + 64;
}

// out[i] = sum_{j<i} in[i]  (n < 2^32 * ...; sums must fit 32 bits).  in == out is allowed.
inline void exclusive_scan_u32(cudaStream_t s, const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch) {
    if (n == 0) return;
    const uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, n, scratch);
    scan_small_kernel<<<1, SCAN_THREADS, 0, s>>>(scratch, tiles);
    scan_apply_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, out, n, scratch);
}

// ---- radix sort ---------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                         // elements per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;      // 2048 elements per CTA (33 KB of static shared memory)
constexpr int RS_BINS = 256;

static __global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t mask,
                                                                uint32_t* __restrict__ hist, uint32_t n_tiles) {
    __shared__ uint32_t bins[RS_BINS];
    bins[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&bins[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = bins[threadIdx.x];
}

static __global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                   uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                   uint64_t n, int shift, uint32_t mask, const uint32_t* __restrict__ bucket_start,
                                                                   uint32_t n_tiles) {
    __shared__ uint32_t warp_cnt[RS_WARPS][RS_BINS];     // running per-warp digit counters, then per-warp exclusive offsets
    __shared__ uint32_t digit_start[RS_BINS];            // start of each digit run inside the reordered tile
    __shared__ uint32_t digit_global[RS_BINS];           // global start of this tile's bucket for each digit
    __shared__ uint64_t s_keys[RS_TILE];
    __shared__ uint32_t s_vals[RS_TILE];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RS_WARPS * RS_BINS; k += RS_THREADS) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();
    const uint64_t tile_base = (uint64_t)blockIdx.x * RS_TILE;
    const uint64_t warp_base = tile_base + (uint64_t)w * (RS_ITEMS * 32);    // each warp owns a contiguous slice
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS];
    uint32_t rank[RS_ITEMS];       // stable rank of the element among the warp's elements with the same digit
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint64_t i = warp_base + (uint64_t)k * 32 + lane;
        const bool ok = i < n;
        key[k] = ok ? keys_in[i] : ~0ull;
        val[k] = ok ? vals_in[i] : 0u;
        const uint32_t d = ok ? ((uint32_t)(key[k] >> shift) & mask) : 0x100u;   // 0x100: padding, matches nothing real
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (ok) {
            const int leader = __ffs(peers) - 1;
            if (lane == leader) { base = warp_cnt[w][d]; warp_cnt[w][d] = base + __popc(peers); }
            base = __shfl_sync(peers, base, leader);
        }
        rank[k] = base + before;
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps, and the digit's total in this tile
    uint32_t tot = 0;
    {
        const int d = threadIdx.x;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { const uint32_t c = warp_cnt[ww][d]; warp_cnt[ww][d] = tot; tot += c; }
        digit_global[d] = bucket_start[(uint64_t)d * n_tiles + blockIdx.x];
    }
    uint32_t tile_total;
    const uint32_t dstart = block_excl_scan(tot, &tile_total);
    digit_start[threadIdx.x] = dstart;
    __syncthreads();
    // reorder the tile by digit in shared memory (stable)
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint64_t i = warp_base + (uint64_t)k * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[k] >> shift) & mask;
            const uint32_t pos = digit_start[d] + warp_cnt[w][d] + rank[k];
            s_keys[pos] = key[k];
            s_vals[pos] = val[k];
        }
    }
    __syncthreads();
    // write the digit runs to their global buckets: consecutive threads -> consecutive addresses inside a run
    for (uint32_t pos = threadIdx.x; pos < tile_total; pos += RS_THREADS) {
        const uint64_t kk = s_keys[pos];
        const uint32_t d = (uint32_t)(kk >> shift) & mask;
        const uint64_t dst = (uint64_t)digit_global[d] + (pos - digit_start[d]);
        keys_out[dst] = kk;
        vals_out[dst] = s_vals[pos];
    }
}

struct RadixScratch {
    uint32_t* hist = nullptr;      // RS_BINS * n_tiles
    uint32_t* scan_tmp = nullptr;
    uint64_t* keys_alt = nullptr;  // ping-pong buffers
    uint32_t* vals_alt = nullptr;
};

inline size_t radix_hist_words(uint64_t n) { return (size_t)RS_BINS * ((n + RS_TILE - 1) / RS_TILE); }

// Sorts (keys, vals) by the low `bits` bits of the key, stable.  The result ends up in (keys_out, vals_out);
// the inputs are used as ping-pong space and are clobbered.  Returns the number of kernels launched.
inline int radix_sort_pairs(cudaStream_t s, uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out,
                            uint64_t n, int bits, uint32_t* hist, uint32_t* scan_tmp) {
    if (n == 0) return 0;
    const int passes = (bits + 7) / 8;
    const uint32_t n_tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint64_t* ka = keys_in; uint32_t* va = vals_in; uint64_t* kb = keys_out; uint32_t* vb = vals_out;
    // with an even number of passes the data would end in the input buffers: start by copying once so that
    // the final pass always lands in (keys_out, vals_out)
    if (passes % 2 == 0) {
        cudaMemcpyAsync(keys_out, keys_in, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(vals_out, vals_in, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
        ka = keys_out; va = vals_out; kb = keys_in; vb = vals_in;
    }
    int launches = 0;
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        const int rem = bits - shift;                               // the last digit may be narrower than 8 bits
        const uint32_t mask = rem >= 8 ? 0xffu : ((1u << rem) - 1u);
        radix_hist_kernel<<<n_tiles, RS_THREADS, 0, s>>>(ka, n, shift, mask, hist, n_tiles);
        exclusive_scan_u32(s, hist, hist, (uint64_t)RS_BINS * n_tiles, scan_tmp);
        radix_scatter_kernel<<<n_tiles, RS_THREADS, 0, s>>>(ka, va, kb, vb, n, shift, mask, hist, n_tiles);
        launches += 5;
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    return launches;
}

}  // namespace prim
