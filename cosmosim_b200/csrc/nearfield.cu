// FP64 near field of the FP32 fast path (stage A).
//
// The FP32 all-pairs kernels round every coordinate to 24 bits: at 1 AU that is a grid of 9 km, and
// a neighbour 10^7..10^8 m away -- touching, about to merge -- then has a separation known to only
// 3..4 digits.  Such pairs dominate the error of the fast path (1e-2..1e-1 of a planet's acceleration)
// while being a vanishing fraction of the N^2 pairs.  This pass finds them with the periodic grid of
// the broad phase (grid.cuh) and, for every ordered pair closer than
//     R_c(i, j) = 2^-9 max(|x_i|, |y_i|, |x_j|, |y_j|)      (2^14 FP32 ulps of the coordinates)
// replaces the FP32 pair term by the FP64 direct-difference term m_j (p_j - p_i) / d^3
// (cosmosim/util/blas.py:31 without the expanded d^2; no softening, the where= mask for d == 0):
//     corr_i = sum_near [ exact term  -  the very term the FP32 kernel added ]
// The FP32 term is REPLAYED with the kernel's own instruction sequence (sub, fma, fma, rsqrt.approx,
// mul, mul, mul -- accel.cu / accel_sym.cu use the same order), so the subtraction removes exactly
// what was added, up to the rounding of the FP32 running sums.  Pairs beyond R_c keep a relative
// error of a few ulp / d <= ~5e-4 in their own term (measured: tests/test_gpu_parity.py).
//
// Cost: one more binning of the planets (on the start-of-frame positions, with LAST frame's grid:
// those positions are last frame's new positions) and ~16 candidate tests per planet: O(N), ~0.1 %
// of the all-pairs pass at N = 2^20.  Sharded runs deal the cell-sorted chunks round-robin (symmetric
// kernels: the corrections ride the all-reduce of the partial sums) or own a row range (one-sided).
#include <math.h>

#include "grid.cuh"

namespace cosmo {

int launch_grid_build_old2(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cosmo_grid *gd,
                           int mode, double reuse_floor, cudaStream_t s);
int launch_grid_build_old3(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cosmo_grid *gd,
                           int mode, double reuse_floor, cudaStream_t s);

#define COSMO_F32_TINY_NF 8.673617379884035e-19f /* 2^-60, the softening of the FP32 kernels */
constexpr double NEAR_REL = 1.0 / 512.0;          // R_c / max |coordinate|
constexpr int NF_CHUNK = 1024;

// first slot of the first cell starting at or after slot k * NF_CHUNK (cell boundaries are the same
// on every rank, the order inside a cell is not)
__global__ void __launch_bounds__(128) k_nf_ranges(cosmo_scratch sc, const cosmo_grid *gd, int *bounds) {
    const int ncell = (int)gd->ncell;
    const int total = sc.cell_start[ncell];
    const int nchunk = (total + NF_CHUNK - 1) / NF_CHUNK;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= nchunk; k += gridDim.x * blockDim.x) {
        const long long target = (long long)k * NF_CHUNK;
        int bound = total;
        if (target < total) {
            int lo = 0, hi = ncell;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sc.cell_start[mid] >= target) hi = mid; else lo = mid + 1;
            }
            bound = sc.cell_start[lo];
        }
        bounds[k] = bound;
    }
}

// correction of the ordered pair (i <- j): exact FP64 term minus the term the FP32 kernel added,
// replayed instruction for instruction
__device__ __forceinline__ void near_term(const cosmo_state &st, const cosmo_scratch &sc, const double2 pi, float xis,
                                          float yis, int j, double unscale, double &cx, double &cy) {
    const double2 pj = reinterpret_cast<const double2 *>(st.pos)[j];
    const double dx = pj.x - pi.x, dy = pj.y - pi.y;
    const double r2 = dx * dx + dy * dy;
    const float dxs = __fsub_rn(sc.pk_x[j], xis), dys = __fsub_rn(sc.pk_y[j], yis);
    const float r2s = __fmaf_rn(dys, dys, __fmaf_rn(dxs, dxs, COSMO_F32_TINY_NF));
    const float rinv = rsqrt_approx_f32(r2s);
    const float sj = __fmul_rn(sc.pk_m[j], __fmul_rn(__fmul_rn(rinv, rinv), rinv));
    const double sjd = (double)sj * unscale;
    // d == 0 keeps the undivided (zero) difference like the reference's where= mask
    const double w = r2 > 0.0 ? st.mass[j] / (r2 * sqrt(r2)) : 0.0;
    cx += dx * w - (double)dxs * sjd;
    cy += dy * w - (double)dys * sjd;
}

// The order of the planets inside a cell comes from atomics; the corrections are therefore added
// in ascending partner index -- a fixed order, so the result is reproducible run to run and rank
// to rank.  Near partners are few (usually none): up to NEAR_KEEP of them are kept sorted in
// registers; a row with more falls back to repeated "next larger partner" scans.
constexpr int NEAR_KEEP = 6;

__device__ __forceinline__ void near_row(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_grid *gd, int i,
                                         double unscale) {
    const double2 *pos = reinterpret_cast<const double2 *>(st.pos);
    const double2 pi = pos[i];
    const float xis = sc.pk_x[i], yis = sc.pk_y[i];
    const double ai = fmax(fabs(pi.x), fabs(pi.y));
    int s0[6], s1[6];
    cand_ranges(sc, sc.cell_of[i], gd->c, s0, s1);
    auto is_near = [&](int j) {
        const double2 pj = pos[j];
        const double dx = pj.x - pi.x, dy = pj.y - pi.y;
        const double rc = NEAR_REL * fmax(ai, fmax(fabs(pj.x), fabs(pj.y)));
        return j != i && dx * dx + dy * dy < rc * rc;
    };
    int keep[NEAR_KEEP], cnt = 0;
#pragma unroll
    for (int d = 0; d < 6; ++d)
        for (int s = s0[d]; s < s1[d]; ++s) {
            const int j = sc.cell_items[s];
            if (!is_near(j)) continue;
            if (cnt < NEAR_KEEP) {          // sorted insert
                int q = cnt;
#pragma unroll
                for (int t = NEAR_KEEP - 1; t > 0; --t)
                    if (t <= q && keep[t - 1] > j) { keep[t] = keep[t - 1]; q = t - 1; }
                keep[q] = j;
            }
            ++cnt;
        }
    double cx = 0.0, cy = 0.0;
    if (cnt <= NEAR_KEEP) {
#pragma unroll
        for (int t = 0; t < NEAR_KEEP; ++t)
            if (t < cnt) near_term(st, sc, pi, xis, yis, keep[t], unscale, cx, cy);
    } else {
        for (int prev = -1;;) {             // a crowd: partners in ascending order, one scan each
            int best = 0x7fffffff;
            for (int d = 0; d < 6; ++d)
                for (int s = s0[d]; s < s1[d]; ++s) {
                    const int j = sc.cell_items[s];
                    if (j > prev && j < best && is_near(j)) best = j;
                }
            if (best == 0x7fffffff) break;
            near_term(st, sc, pi, xis, yis, best, unscale, cx, cy);
            prev = best;
        }
    }
    reinterpret_cast<double2 *>(sc.corr)[i] = make_double2(cx, cy);
}

// own_mode 0: rows [i_begin, i_end) belong to this rank (one-sided kernels integrate them locally);
// own_mode 1: the cell-sorted chunks are dealt round-robin to the ranks (symmetric kernels: every
// rank's partial sums are all-reduced, any partition of the rows will do).
__global__ void __launch_bounds__(128)
k_nearfield(cosmo_state st, cosmo_scratch sc, const cosmo_grid *gd, const int *bounds, int own_mode, int i_begin,
            int i_end, int shard_index, int shard_count, double unscale) {
    const int total = sc.cell_start[(int)gd->ncell];
    if (own_mode == 0) {
        for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += gridDim.x * blockDim.x) {
            const int i = sc.cell_items[slot];
            if (i >= i_begin && i < i_end) near_row(st, sc, gd, i, unscale);
        }
        return;
    }
    const int nchunk = (total + NF_CHUNK - 1) / NF_CHUNK;
    for (int chunk = shard_index + blockIdx.x * shard_count; chunk < nchunk; chunk += gridDim.x * shard_count) {
        const int s_lo = bounds[chunk], s_hi = bounds[chunk + 1];
        for (int slot = s_lo + threadIdx.x; slot < s_hi; slot += blockDim.x)
            near_row(st, sc, gd, sc.cell_items[slot], unscale);
    }
}

// ---------------------------------------------------------------------------------------
// The same pass for the 3-D step (state3d.cu): positions, FP32 stage and corrections are planes
// [3][cap] / [4][cap]; the planets are binned by their (x, y) projection (objects close in 3-D are at
// least as close in projection) and the cutoff is tested on the 3-D distance.  The FP32 term replayed
// is the one of k3_accel_f32 / k3_accel_f32_sym.
// ---------------------------------------------------------------------------------------
struct Near3 {
    const double *pos;    // [3][cap] start-of-step positions
    const float *pk;      // [4][cap] scaled x, y, z, mass
    const double *mass;   // [cap]
    double *corr;         // [3][cap]
    size_t cap;
};

__device__ __forceinline__ void near_row3(const cosmo_scratch &sc, const Near3 &a, const cosmo_grid *gd, int i,
                                          double unscale) {
    const size_t cap = a.cap;
    const double xi = a.pos[i], yi = a.pos[cap + i], zi = a.pos[2 * cap + i];
    const float xis = a.pk[i], yis = a.pk[cap + i], zis = a.pk[2 * cap + i];
    const double ai = fmax(fabs(xi), fmax(fabs(yi), fabs(zi)));
    int s0[6], s1[6];
    cand_ranges(sc, sc.cell_of[i], gd->c, s0, s1);
    auto is_near = [&](int j) {
        const double xj = a.pos[j], yj = a.pos[cap + j], zj = a.pos[2 * cap + j];
        const double dx = xj - xi, dy = yj - yi, dz = zj - zi;
        const double rc = NEAR_REL * fmax(ai, fmax(fabs(xj), fmax(fabs(yj), fabs(zj))));
        return j != i && dx * dx + dy * dy + dz * dz < rc * rc;
    };
    double c[3] = {0.0, 0.0, 0.0};
    auto term = [&](int j) {
        const double dx = a.pos[j] - xi, dy = a.pos[cap + j] - yi, dz = a.pos[2 * cap + j] - zi;
        const double r2 = dx * dx + dy * dy + dz * dz;
        const float dxs = __fsub_rn(a.pk[j], xis), dys = __fsub_rn(a.pk[cap + j], yis), dzs = __fsub_rn(a.pk[2 * cap + j], zis);
        const float r2s = __fmaf_rn(dzs, dzs, __fmaf_rn(dys, dys, __fmaf_rn(dxs, dxs, COSMO_F32_TINY_NF)));
        const float rinv = rsqrt_approx_f32(r2s);
        const float sj = __fmul_rn(a.pk[3 * cap + j], __fmul_rn(__fmul_rn(rinv, rinv), rinv));
        const double sjd = (double)sj * unscale;
        const double w = r2 > 0.0 ? a.mass[j] / (r2 * sqrt(r2)) : 0.0;
        c[0] += dx * w - (double)dxs * sjd;
        c[1] += dy * w - (double)dys * sjd;
        c[2] += dz * w - (double)dzs * sjd;
    };
    for (int prev = -1;;) {                 // partners in ascending index order (fixed summation order)
        int best = 0x7fffffff;
        for (int d = 0; d < 6; ++d)
            for (int s = s0[d]; s < s1[d]; ++s) {
                const int j = sc.cell_items[s];
                if (j > prev && j < best && is_near(j)) best = j;
            }
        if (best == 0x7fffffff) break;
        term(best);
        prev = best;
    }
    a.corr[i] = c[0]; a.corr[cap + i] = c[1]; a.corr[2 * cap + i] = c[2];
}

__global__ void __launch_bounds__(128)
k_nearfield3(cosmo_scratch sc, Near3 a, const cosmo_grid *gd, const int *bounds, int own_mode, int i_begin, int i_end,
             int shard_index, int shard_count, double unscale) {
    const int total = sc.cell_start[(int)gd->ncell];
    if (own_mode == 0) {
        for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += gridDim.x * blockDim.x) {
            const int i = sc.cell_items[slot];
            if (i >= i_begin && i < i_end) near_row3(sc, a, gd, i, unscale);
        }
        return;
    }
    const int nchunk = (total + NF_CHUNK - 1) / NF_CHUNK;
    for (int chunk = shard_index + blockIdx.x * shard_count; chunk < nchunk; chunk += gridDim.x * shard_count) {
        const int s_lo = bounds[chunk], s_hi = bounds[chunk + 1];
        for (int slot = s_lo + threadIdx.x; slot < s_hi; slot += blockDim.x) near_row3(sc, a, gd, sc.cell_items[slot], unscale);
    }
}

// st2 / sc2: shadow structs of the 3-D step (pos = planes, the cell arrays, chunk_alive, grid_dev, tune_hist,
// rinf, status); p2 carries n_upper, shard, near_mode, exponents.  sym: symmetric kernel in use.
int launch_nearfield3(const cosmo_state &st2, const cosmo_scratch &sc2, const cosmo_step_params &p2, const float *pk,
                      const double *mass, double *corr3, int sym, cudaStream_t s) {
    if (p2.n_upper <= 0) return COSMO_OK;
    if (!corr3 || !sc2.grid_dev || !sc2.chunk_alive || !pk) {
        set_error("3-D near field: scratch lacks corr / grid_dev / chunk bounds");
        return COSMO_E_ARG;
    }
    cosmo_grid *gd = sc2.grid_dev + 1;
    const size_t cap = (size_t)st2.cap;
    for (int d = 0; d < 3; ++d)
        COSMO_TRY(check_cuda(cudaMemsetAsync(corr3 + d * cap, 0, (size_t)p2.n_upper * sizeof(double), s), "corr clear"));
    COSMO_TRY(launch_grid_build_old3(st2, sc2, p2, gd, p2.near_mode == 2 ? 2 : 3, 0.5 * NEAR_REL, s));
    const int sh_n = p2.shard_count > 1 ? p2.shard_count : 1;
    const int own_mode = (sym && sh_n > 1) ? 1 : 0;
    const int nchunk = (p2.n_upper + NF_CHUNK - 1) / NF_CHUNK + 1;
    const double unscale = scalbn(1.0, p2.mass_exp - 2 * p2.pos_exp);
    if (own_mode) {
        k_nf_ranges<<<(nchunk + 127) / 128, 128, 0, s>>>(sc2, gd, sc2.chunk_alive);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_nf_ranges"));
    }
    int blocks = own_mode ? (nchunk + sh_n - 1) / sh_n : (p2.n_upper + 127) / 128;
    if (blocks < 1) blocks = 1;
    if (!own_mode && blocks > 148 * 16) blocks = 148 * 16;
    const int ib = (sh_n > 1 && !sym) ? p2.i_begin : 0, ie = (sh_n > 1 && !sym) ? p2.i_end : p2.n_upper;
    Near3 a = {st2.pos, pk, mass, corr3, cap};
    k_nearfield3<<<blocks, 128, 0, s>>>(sc2, a, gd, sc2.chunk_alive, own_mode, ib, ie, p2.shard_index, sh_n, unscale);
    return COSMO_LAUNCH_CHECK("k_nearfield3");
}

int launch_nearfield(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cudaStream_t s) {
    if (p.n_upper <= 0) return COSMO_OK;
    if (!sc.corr || !sc.grid_dev || !sc.chunk_alive) {
        set_error("near field: scratch lacks corr / grid_dev / chunk_alive");
        return COSMO_E_ARG;
    }
    cosmo_grid *gd = sc.grid_dev + 1;        // second slot: stage B's own grid (slot 0) is left alone
    COSMO_TRY(check_cuda(cudaMemsetAsync(sc.corr, 0, (size_t)p.n_upper * 2 * sizeof(double), s), "corr clear"));
    // near_mode 1: last frame's stage-B grid, cells no smaller than the near-field cutoff of the bulk
    // (2^-10 of half the bulk box); near_mode 2: no such grid yet -> tune one on the start positions
    COSMO_TRY(launch_grid_build_old2(st, sc, p, gd, p.near_mode == 2 ? 2 : 3, 0.5 * NEAR_REL, s));
    const int sym = (p.opts & (COSMO_OPT_F32_SYM)) ? 1 : 0;
    const int sh_n = p.shard_count > 1 ? p.shard_count : 1;
    const int own_mode = (sym && sh_n > 1) ? 1 : 0;
    const int nchunk = (p.n_upper + NF_CHUNK - 1) / NF_CHUNK + 1;
    const double unscale = scalbn(1.0, p.mass_exp - 2 * p.pos_exp);
    if (own_mode) {
        k_nf_ranges<<<(nchunk + 127) / 128, 128, 0, s>>>(sc, gd, sc.chunk_alive);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_nf_ranges"));
    }
    int blocks = own_mode ? (nchunk + sh_n - 1) / sh_n : (p.n_upper + 127) / 128;
    if (blocks < 1) blocks = 1;
    if (!own_mode && blocks > 148 * 16) blocks = 148 * 16;
    // one-sided single GPU / symmetric single GPU: every row (i_begin = 0, i_end = n)
    const int ib = (sh_n > 1 && !sym) ? p.i_begin : 0, ie = (sh_n > 1 && !sym) ? p.i_end : p.n_upper;
    k_nearfield<<<blocks, 128, 0, s>>>(st, sc, gd, sc.chunk_alive, own_mode, ib, ie, p.shard_index, sh_n, unscale);
    return COSMO_LAUNCH_CHECK("k_nearfield");
}

}  // namespace cosmo
