// Pipe micro-benchmarks: measure the FMA-pipe / FP64-pipe / MUFU peaks that the roofline of
// the all-pairs kernels is quoted against (SURVEY.md section 8d asks for measured, not
// derived, denominators).  Each thread runs 8 independent dependency chains so latency is
// covered with a handful of warps per scheduler.
#include "cosmo_device.cuh"
#include "cosmo_internal.h"

namespace cosmo {

template <int KIND>
__global__ void __launch_bounds__(1024) k_microbench(int iters, double *out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float f[8];
    double d[8];
    uint64_t u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        f[k] = 1.0f + 1e-3f * (float)(k + (t & 7));
        d[k] = 1.0 + 1e-3 * (double)(k + (t & 7));
        u[k] = pack2(f[k], f[k] * 0.5f);
    }
    const float fa = 0.9999f, fb = 1e-4f;
    const double da = 0.9999, db = 1e-4;
    const uint64_t ua = pack2(fa, fa), ub = pack2(fb, fb);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (KIND == 0) f[k] = fmaf(f[k], fa, fb);
                if (KIND == 1) u[k] = fma2(u[k], ua, ub);
                if (KIND == 2) d[k] = __fma_rn(d[k], da, db);
                if (KIND == 3) f[k] = rsqrt_approx_f32(f[k]);
                if (KIND == 4) d[k] = __dadd_rn(d[k], db);
                if (KIND == 5) d[k] = __dmul_rn(d[k], da);
                if (KIND == 6) f[k] = __fmul_rn(f[k], fa);
                if (KIND == 7) f[k] = __fadd_rn(f[k], fb);
                if (KIND == 8) d[k] = rsqrt_approx_f64(d[k]);
            }
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float a, b;
        unpack2(u[k], a, b);
        acc += (double)f[k] + d[k] + (double)a + (double)b;
    }
    out[t] = acc;
}

int launch_microbench(int kind, int blocks, int threads, int iters, double *out,
                      int64_t *ops_per_thread, cudaStream_t s) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0 || !out) {
        set_error("cosmo_microbench: bad launch shape");
        return COSMO_E_ARG;
    }
    if (ops_per_thread) *ops_per_thread = (int64_t)iters * 16 * 8;
    switch (kind) {
        case 0: k_microbench<0><<<blocks, threads, 0, s>>>(iters, out); break;
        case 1: k_microbench<1><<<blocks, threads, 0, s>>>(iters, out); break;
        case 2: k_microbench<2><<<blocks, threads, 0, s>>>(iters, out); break;
        case 3: k_microbench<3><<<blocks, threads, 0, s>>>(iters, out); break;
        case 4: k_microbench<4><<<blocks, threads, 0, s>>>(iters, out); break;
        case 5: k_microbench<5><<<blocks, threads, 0, s>>>(iters, out); break;
        case 6: k_microbench<6><<<blocks, threads, 0, s>>>(iters, out); break;
        case 7: k_microbench<7><<<blocks, threads, 0, s>>>(iters, out); break;
        case 8: k_microbench<8><<<blocks, threads, 0, s>>>(iters, out); break;
        default: set_error("cosmo_microbench: unknown kind %d", kind); return COSMO_E_ARG;
    }
    return COSMO_LAUNCH_CHECK("k_microbench");
}

}  // namespace cosmo
