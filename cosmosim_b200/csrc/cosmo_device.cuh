// Device-side helpers: sm_100a PTX wrappers (mbarrier, bulk async copy, packed f32x2
// math, MUFU approximations) and the exact scalar arithmetic the merge resolver needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cosmo {

typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------
// shared-memory address / mbarrier / bulk async copy (TMA engine, 1-D: SASS UBLKCP)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// global -> shared bulk copy completing on an mbarrier; bytes % 16 == 0, both 16-B aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Spin on the phase parity.  Bounded: a lost transaction traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// ------------------------------------------------------------------------------------
// packed FP32x2 math (Blackwell FFMA2 / FMUL2 / FADD2) on 64-bit register pairs
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ float rsqrt_approx_f32(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// MUFU.RSQ64H: ~20-bit seed from the high word of a double
__device__ __forceinline__ double rsqrt_approx_f64(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// ------------------------------------------------------------------------------------
// exact scalar helpers of the merge resolver (Python numeric semantics, planet.py:74-91)
// ------------------------------------------------------------------------------------
// float(int) with round-half-even, like CPython's int.__float__
__host__ __device__ inline double u128_to_double(u128 v) {
    if (v == 0) return 0.0;
    uint64_t hi = (uint64_t)(v >> 64), lo = (uint64_t)v;
    int top;
#ifdef __CUDA_ARCH__
    top = hi ? 127 - __clzll((long long)hi) : 63 - __clzll((long long)lo);
#else
    top = hi ? 127 - __builtin_clzll(hi) : 63 - __builtin_clzll(lo);
#endif
    if (top <= 52) return (double)lo;
    int shift = top - 52;
    uint64_t mant = (uint64_t)(v >> shift);
    u128 rem = v & ((((u128)1) << shift) - 1);
    u128 half = ((u128)1) << (shift - 1);
    if (rem > half || (rem == half && (mant & 1))) ++mant;
    // mant in [2^52, 2^53]; scale by 2^shift exactly through the exponent field
    double m = (double)mant;
    uint64_t bits;
    memcpy(&bits, &m, 8);
    bits += ((uint64_t)shift) << 52;
    memcpy(&m, &bits, 8);
    return m;
}

// floor(f) as u128 for 0 <= f < 2^128; *has_frac = (f != floor(f))
__host__ __device__ inline u128 double_floor_u128(double f, bool *has_frac) {
    uint64_t bits;
    memcpy(&bits, &f, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1075;  // f = mant * 2^e
    uint64_t mant = (bits & 0xfffffffffffffULL) | (((bits >> 52) & 0x7ff) ? (1ULL << 52) : 0);
    if (e >= 0) {
        *has_frac = false;
        return ((u128)mant) << e;
    }
    if (e <= -64) {
        *has_frac = mant != 0;
        return 0;
    }
    uint64_t ip = mant >> (-e);
    *has_frac = (mant & ((1ULL << (-e)) - 1)) != 0;
    return (u128)ip;
}

// Python's exact `a >= b` for (int|float) x (int|float) non-negative masses
__host__ __device__ inline bool mass_ge(bool a_int, u128 ai, double af, bool b_int, u128 bi, double bf) {
    const double two128 = 3.402823669209384634633746074317682e38;
    if (a_int && b_int) return ai >= bi;
    if (!a_int && !b_int) return af >= bf;
    if (a_int) {  // int >= float
        if (bf != bf) return false;
        if (bf < 0.0) return true;
        if (bf >= two128) return false;
        bool frac;
        u128 fi = double_floor_u128(bf, &frac);
        if (ai != fi) return ai > fi;
        return !frac;
    }
    // float >= int
    if (af != af) return false;
    if (af < 0.0) return false;
    if (af >= two128) return true;
    bool frac;
    u128 fi = double_floor_u128(af, &frac);
    if (fi != bi) return fi > bi;
    return true;
}

// pow(x, fl(1/3)) for finite x > 0, rounded like a correctly-rounded libm pow
// (planet.py:82 computes the merged radius as ((3 V)/(4 pi)) ** (1/3), i.e. C pow with the
// exponent 0.333333333333333314829616256247...; that differs from cbrt by several ulps).
//   x^c = cbrt(x) * exp(-delta ln x),   delta = 1/3 - fl(1/3) = 2^-54 / 3
// y0 = cbrt(x) to a few ulps; one Newton correction from the exactly-evaluated residual
// x - y0^3 plus the first-order -y0*delta*ln(x) term are added in a single rounding.
__host__ __device__ inline double pow_third(double x) {
    const double delta = 1.850371707708594e-17;  // 2^-54 / 3
    double y0 = cbrt(x);
    double a = y0 * y0;
    double ae = fma(y0, y0, -a);
    double b = a * y0;
    double be = fma(a, y0, -b);
    double r = ((x - b) - be) - ae * y0;
    double corr = r / (3.0 * a) - y0 * (delta * log(x));
    return y0 + corr;
}

}  // namespace cosmo
