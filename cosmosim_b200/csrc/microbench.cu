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

// Operand-traffic probes of the packed FP32 pipe (what bounds k_accel_f32_sym once everything else is
// out of the way): an FFMA2 reads three 64-bit register pairs.
//   KIND 9   fma2 with three live, changing operands (no operand reuse possible)
//   KIND 10  fma2(v[k], const, u[k]): two changing operands, one constant
//   KIND 11  the pair chain of the symmetric kernel on registers only (2 FADD2, 6 FFMA2, 4 FMUL2, 2 MUFU
//            per two pairs; no shared memory, no shuffles, no barriers): the ceiling of the instruction mix
//   KIND 12  the same chain with the two MUFU replaced by moves of the argument
//   KIND 13 / 14  DFMA with three changing operands / two changing operands and a constant
template <int KIND>
__global__ void __launch_bounds__(256, 2) k_microbench2(int iters, double *out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t u[8], v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float f = 1.0f + 1e-3f * (float)(k + (t & 7));
        u[k] = pack2(f, f * 0.5f);
        v[k] = pack2(0.5f * f, 0.25f * f);
    }
    const uint64_t ua = pack2(0.9999f, 0.9999f);
    if (KIND == 13 || KIND == 14) {
        double d[8], e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { d[k] = 1.0 + 1e-3 * (double)(k + (t & 7)); e[k] = 0.5 * d[k]; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (KIND == 13) d[k] = __fma_rn(e[(k + r) & 7], d[(k + 3) & 7], d[k]);
                    if (KIND == 14) d[k] = __fma_rn(e[k], 0.9999, d[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = (uint64_t)__double_as_longlong(d[k]) >> 3;
    } else if (KIND == 9 || KIND == 10) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (KIND == 9) u[k] = fma2(v[(k + r) & 7], u[(k + 3) & 7], u[k]);
                    if (KIND == 10) u[k] = fma2(v[k], ua, u[k]);
                }
            }
        }
    } else {
        // 4 row planets (x, y, -m as scalars; packed sums) against 4 column planets per step, like one
        // rotation step of the kernel; the column values change every step so nothing is hoisted
        float xi[4], yi[4], nmi[4];
        uint64_t rx2[4], ry2[4], tx2[2], ty2[2], xj2[2], yj2[2], mj2[2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            xi[k] = 0.1f * (float)(k + (t & 3)); yi[k] = 0.2f * (float)(k + (t & 5)); nmi[k] = -1.0f - (float)k;
            rx2[k] = 0ull; ry2[k] = 0ull;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) { tx2[h] = 0ull; ty2[h] = 0ull; xj2[h] = u[h]; yj2[h] = u[2 + h]; mj2[h] = v[h]; }
        const uint64_t eps2 = pack2(1e-9f, 1e-9f), step = pack2(1e-3f, 2e-3f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint64_t xik, yik, nmik;
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(xik) : "f"(xi[k]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(yik) : "f"(yi[k]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(nmik) : "f"(nmi[k]));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t dx = sub2(xj2[h], xik);
                    const uint64_t dy = sub2(yj2[h], yik);
                    const uint64_t r2 = fma2(dy, dy, fma2(dx, dx, eps2));
                    float r2a, r2b;
                    unpack2(r2, r2a, r2b);
                    const uint64_t rinv = KIND == 11 ? pack2(rsqrt_approx_f32(r2a), rsqrt_approx_f32(r2b)) : r2;
                    const uint64_t rinv3 = mul2(mul2(rinv, rinv), rinv);
                    const uint64_t sj = mul2(mj2[h], rinv3);
                    const uint64_t nsi = mul2(rinv3, nmik);
                    rx2[k] = fma2(dx, sj, rx2[k]);
                    ry2[k] = fma2(dy, sj, ry2[k]);
                    tx2[h] = fma2(dx, nsi, tx2[h]);
                    ty2[h] = fma2(dy, nsi, ty2[h]);
                }
            }
            // "next column chunk": cheap register-only change of the column operands (1 FADD2-class each,
            // 6 per 96 -- counted as overhead, not as work)
#pragma unroll
            for (int h = 0; h < 2; ++h) { xj2[h] = add2(xj2[h], step); yj2[h] = add2(yj2[h], step); mj2[h] = add2(mj2[h], step); }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { u[k] = add2(rx2[k], ry2[k]); }
        u[4] = add2(tx2[0], ty2[0]); u[5] = add2(tx2[1], ty2[1]);
    }
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float a, b;
        unpack2(u[k], a, b);
        acc += (double)a + (double)b;
    }
    out[t] = acc;
}

int launch_microbench(int kind, int blocks, int threads, int iters, double *out,
                      int64_t *ops_per_thread, cudaStream_t s) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0 || !out) {
        set_error("cosmo_microbench: bad launch shape");
        return COSMO_E_ARG;
    }
    if (ops_per_thread) *ops_per_thread = (int64_t)iters * ((kind == 11 || kind == 12) ? 96 : 16 * 8);   // packed FMA-pipe instructions
    switch (kind) {
        case 9: k_microbench2<9><<<blocks, threads, 0, s>>>(iters, out); break;
        case 10: k_microbench2<10><<<blocks, threads, 0, s>>>(iters, out); break;
        case 11: k_microbench2<11><<<blocks, threads, 0, s>>>(iters, out); break;
        case 12: k_microbench2<12><<<blocks, threads, 0, s>>>(iters, out); break;
        case 13: k_microbench2<13><<<blocks, threads, 0, s>>>(iters, out); break;
        case 14: k_microbench2<14><<<blocks, threads, 0, s>>>(iters, out); break;
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
