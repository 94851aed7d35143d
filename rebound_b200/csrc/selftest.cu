// selftest.cu -- on-device check that the branch-free fast-range sqrt / divide of strict_math.cuh agree
// bit for bit with the IEEE-correct generic __dsqrt_rn / __ddiv_rn wherever they do not raise their
// out-of-range flag (and that the flag is raised rarely on ordinary operands).
#include "engine.cuh"
#include "strict_math.cuh"
#include "primitives.cuh"

namespace {

__device__ __forceinline__ uint64_t mix(uint64_t x) {          // splitmix64
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// Operand families: 0 uniform mantissa, moderate exponents (what pair distances look like); 1 any bit
// pattern (exercises the flag: negatives, denormals, inf, nan); 2 exact squares +- 1 ulp (rounding ties).
__device__ double operand(uint64_t h, int family) {
    if (family == 1) return __longlong_as_double((long long)h);
    const uint64_t mant = h & 0x000fffffffffffffull;
    const int e = 1023 + (int)((h >> 52) % 121) - 60;
    double v = __longlong_as_double((long long)(((uint64_t)e << 52) | mant));
    if (family == 2) {
        const double s = (double)(uint32_t)(h >> 20) * 1.0000000001;
        v = __longlong_as_double(__double_as_longlong(s * s) + (long long)(h % 3) - 1);
    }
    return v;
}

__global__ void math_selftest_kernel(uint64_t n, uint64_t seed, unsigned long long* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t h1 = mix(seed + 2 * i), h2 = mix(seed + 2 * i + 1);
    const int family = (int)(i % 3);
    const double x = operand(h1, family), y = operand(h2, family == 2 ? 0 : family);
    unsigned bs = 0, bd = 0;
    const double s_fast = fsqrt_rn(x, bs), s_ref = __dsqrt_rn(x);
    const double d_fast = fdiv_rn(x, y, bd), d_ref = __ddiv_rn(x, y);
    if (!bs && __double_as_longlong(s_fast) != __double_as_longlong(s_ref)) atomicAdd(&out[0], 1ull);
    if (!bd && __double_as_longlong(d_fast) != __double_as_longlong(d_ref)) atomicAdd(&out[1], 1ull);
    if (family != 1) { if (bs) atomicAdd(&out[2], 1ull); if (bd) atomicAdd(&out[3], 1ull); }
}

}  // namespace

// Sorts caller-provided (key, value) pairs by the low `bits` key bits with the hand-written radix sort of
// primitives.cuh (the tree build's sorter) so that tests can compare it with a host stable sort.
extern "C" int rebcu_selftest_sort(rebcu_handle* h, uint64_t* keys, uint32_t* vals, uint64_t n, int bits) {
    CU_TRY(h, cudaSetDevice(h->device));
    if (n == 0) return REBCU_OK;
    uint64_t *k0 = nullptr, *k1 = nullptr; uint32_t *v0 = nullptr, *v1 = nullptr, *hist = nullptr, *tmp = nullptr;
    const size_t hw = prim::radix_hist_words(n), sw = prim::scan_scratch_words(hw);
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc(&k0, n * 8)) || (e = cudaMalloc(&k1, n * 8)) || (e = cudaMalloc(&v0, n * 4)) || (e = cudaMalloc(&v1, n * 4)) ||
        (e = cudaMalloc(&hist, hw * 4)) || (e = cudaMalloc(&tmp, sw * 4))) {
        cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(hist); cudaFree(tmp);
        return rebcu_cuda_fail(h, e, "selftest_sort allocation");
    }
    cudaMemcpyAsync(k0, keys, n * 8, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(v0, vals, n * 4, cudaMemcpyHostToDevice, h->stream);
    h->launches += prim::radix_sort_pairs(h->stream, k0, v0, k1, v1, n, bits, hist, tmp);
    cudaMemcpyAsync(keys, k1, n * 8, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(vals, v1, n * 4, cudaMemcpyDeviceToHost, h->stream);
    e = cudaStreamSynchronize(h->stream);
    cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(hist); cudaFree(tmp);
    if (e != cudaSuccess) return rebcu_cuda_fail(h, e, "selftest_sort");
    return REBCU_OK;
}

// Exclusive scan of caller-provided counts with the hand-written scan of primitives.cuh.
extern "C" int rebcu_selftest_scan(rebcu_handle* h, uint32_t* values, uint64_t n) {
    CU_TRY(h, cudaSetDevice(h->device));
    if (n == 0) return REBCU_OK;
    uint32_t *d = nullptr, *tmp = nullptr;
    CU_TRY(h, cudaMalloc(&d, n * 4));
    cudaError_t e = cudaMalloc(&tmp, prim::scan_scratch_words(n) * 4);
    if (e != cudaSuccess) { cudaFree(d); return rebcu_cuda_fail(h, e, "selftest_scan allocation"); }
    cudaMemcpyAsync(d, values, n * 4, cudaMemcpyHostToDevice, h->stream);
    prim::exclusive_scan_u32(h->stream, d, d, n, tmp);
    h->launches += 3;
    cudaMemcpyAsync(values, d, n * 4, cudaMemcpyDeviceToHost, h->stream);
    e = cudaStreamSynchronize(h->stream);
    cudaFree(d); cudaFree(tmp);
    if (e != cudaSuccess) return rebcu_cuda_fail(h, e, "selftest_scan");
    return REBCU_OK;
}

extern "C" int rebcu_selftest_math(rebcu_handle* h, uint64_t n_samples, uint64_t seed, uint64_t* result4) {
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaMemsetAsync(h->counters, 0, 4 * sizeof(unsigned long long), h->stream));
    {
        LaunchScope ls(h, TC_DIRECT);
        math_selftest_kernel<<<div_up(n_samples, 256), 256, 0, h->stream>>>(n_samples, seed, h->counters);
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(h->pinned, h->counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 4; k++) result4[k] = h->pinned[k];
    return REBCU_OK;
}
