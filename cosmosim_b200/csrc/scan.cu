// Device-wide exclusive scan of int32 arrays (three launches: chunk sums, top-level scan of
// the chunk sums by one CTA, chunk-local scan + offset).  Used by the uniform-grid broad
// phase (cell counts -> cell starts) and by the CSR build of the flagged pairs (row counts ->
// row starts).  The element count may live on the device (n_ptr) so no host sync is needed.
#include "cosmo_device.cuh"
#include "cosmo_internal.h"

namespace cosmo {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;  // 4096

__device__ __forceinline__ long long scan_len(const int64_t *n_ptr, long long n_max, long long n_add) {
    long long n = n_ptr ? *n_ptr + n_add : n_max;
    return n < n_max ? (n < 0 ? 0 : n) : n_max;
}

__device__ __forceinline__ int block_exclusive_scan_256(int v, int *total) {
    __shared__ int s_w[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        const int c = s_w[x];
        before += x < w ? c : 0;
        tot += c;
    }
    *total = tot;
    return before + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const int *in, int *blk, const int64_t *n_ptr, long long n_max, long long n_add) {
    const long long n = scan_len(n_ptr, n_max, n_add);
    const long long c0 = (long long)blockIdx.x * SCAN_CHUNK;
    if (c0 >= n) return;
    int sum = 0;
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        const long long k = c0 + q * SCAN_THREADS + threadIdx.x;
        if (k < n) sum += in[k];
    }
    int tot;
    block_exclusive_scan_256(sum, &tot);
    if (threadIdx.x == 0) blk[blockIdx.x] = tot;
}

// one CTA: exclusive scan of the chunk sums in place; total -> blk[COSMO_SCAN_BLOCKS]
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_top(int *blk, const int64_t *n_ptr, long long n_max, long long n_add) {
    const long long n = scan_len(n_ptr, n_max, n_add);
    const int nblk = (int)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
    int running = 0;
    for (int b0 = 0; b0 < nblk; b0 += SCAN_THREADS) {
        const int b = b0 + threadIdx.x;
        const int v = b < nblk ? blk[b] : 0;
        int tot;
        const int ex = block_exclusive_scan_256(v, &tot);
        if (b < nblk) blk[b] = running + ex;
        running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) blk[COSMO_SCAN_BLOCKS] = running;
}

// out[k] = sum(in[0..k)); out[n] = total.  out may alias in.
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_down(const int *in, int *out, const int *blk, const int64_t *n_ptr, long long n_max, long long n_add) {
    const long long n = scan_len(n_ptr, n_max, n_add);
    const long long c0 = (long long)blockIdx.x * SCAN_CHUNK;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = blk[COSMO_SCAN_BLOCKS];
    if (c0 >= n) return;
    int v[SCAN_ITEMS];
    int sum = 0;
    const long long k0 = c0 + (long long)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        v[q] = k0 + q < n ? in[k0 + q] : 0;
        sum += v[q];
    }
    int tot;
    int ex = block_exclusive_scan_256(sum, &tot) + blk[blockIdx.x];
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        if (k0 + q < n) out[k0 + q] = ex;
        ex += v[q];
    }
}

int launch_scan_exclusive(const int *in, int *out, int *blk, const int64_t *n_ptr, long long n_max,
                          long long n_add, cudaStream_t s) {
    if (n_max <= 0) return COSMO_OK;
    long long chunks = (n_max + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (chunks > COSMO_SCAN_BLOCKS) {
        set_error("scan of %lld items exceeds COSMO_SCAN_BLOCKS", n_max);
        return COSMO_E_ARG;
    }
    k_scan_reduce<<<(int)chunks, SCAN_THREADS, 0, s>>>(in, blk, n_ptr, n_max, n_add);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_scan_reduce"));
    k_scan_top<<<1, SCAN_THREADS, 0, s>>>(blk, n_ptr, n_max, n_add);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_scan_top"));
    k_scan_down<<<(int)chunks, SCAN_THREADS, 0, s>>>(in, out, blk, n_ptr, n_max, n_add);
    return COSMO_LAUNCH_CHECK("k_scan_down");
}

}  // namespace cosmo
