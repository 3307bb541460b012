// Stage B: collision-pair detection on the NEW positions.  Replaces pairwise_distances +
// collision_matrix (reference cosmosim/core/universe_old.py:180,192-193).
//
// Two modes, same exact predicate (pred.cuh), same output (unsorted list of flagged ordered
// pairs; stage C orders it):
//   all-pairs   the reference's N x N test, tiled like stage A; used for small N;
//   grid        uniform-grid broad phase: a counting sort of the planets into square cells no
//               smaller than the largest possible flagged separation of two non-giant planets,
//               so every flagged pair of non-giants sits in adjacent cells; "giants" (radius
//               above giant_radius: the star) are tested against everybody.  The candidate
//               set is a superset; every candidate goes through the exact predicate, so the
//               flagged set is identical to the all-pairs one.
#include <math.h>

#include "grid.cuh"
#include "pred.cuh"

namespace cosmo {

// All-pairs detection: same tiling as stage A (targets in registers, sources through a
// bulk-copy/mbarrier ring).  7 FP64-pipe instructions per ordered pair for the candidate
// filter d2 <= (r'_i + r'_j)^2 with radii inflated by 2^-30 (a superset of the flagged
// pairs); the rare candidates go through the exact predicate with the true radii.
template <int BLOCK, int IB, int TJ>
__global__ void __launch_bounds__(BLOCK)
k_detect_allpairs(cosmo_state st, cosmo_scratch sc, int i_begin, int i_end, int jsplit) {
    __shared__ __align__(16) double2 s_xy[2][TJ];
    __shared__ __align__(16) double s_xx[2][TJ];
    __shared__ __align__(16) double s_r[2][TJ];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int base = i_begin + blockIdx.x * (BLOCK * IB);
    if (base >= i_hi) return;
    const int tid = threadIdx.x;
    const double2 *pn = reinterpret_cast<const double2 *>(sc.pos_new);

    double qx[IB], qy[IB], xxi[IB], ri[IB];
    int ii[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int i = base + k * BLOCK + tid;
        ii[k] = i < i_hi ? i : -1;
        int ic = min(i, i_hi - 1);
        double2 p = pn[ic];
        qx[k] = -2.0 * p.x; qy[k] = -2.0 * p.y;
        xxi[k] = sc.xx_new[ic];
        ri[k] = sc.rinf[ic];
    }

    const int tiles = (n + TJ - 1) / TJ;
    const int t0 = (int)(((long long)blockIdx.y * tiles) / jsplit);
    const int t1 = (int)(((long long)(blockIdx.y + 1) * tiles) / jsplit);
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t kTileBytes = TJ * (sizeof(double2) + 2 * sizeof(double));
    auto issue = [&](int t, int b) {
        const size_t j0 = (size_t)t * TJ;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kTileBytes);
        bulk_g2s(&s_xy[b][0], pn + j0, TJ * sizeof(double2), &s_bar[b]);
        bulk_g2s(&s_xx[b][0], sc.xx_new + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_r[b][0], sc.rinf + j0, TJ * sizeof(double), &s_bar[b]);
    };
    if (tid == 0 && t0 < t1) issue(t0, 0);

    for (int t = t0; t < t1; ++t) {
        const int it = t - t0;
        const int b = it & 1;
        if (tid == 0 && t + 1 < t1) issue(t + 1, b ^ 1);
        mbar_wait(&s_bar[b], (it >> 1) & 1);
        const int j0 = t * TJ;
        const int cnt = min(TJ, n - j0);
#pragma unroll 4
        for (int jj = 0; jj < cnt; ++jj) {
            const double2 pj = s_xy[b][jj];
            const double xxj = s_xx[b][jj];
            const double rj = s_r[b][jj];
#pragma unroll
            for (int k = 0; k < IB; ++k) {
                double d2 = pred_d2(qx[k], qy[k], xxi[k], pj.x, pj.y, xxj);
                double rs = ri[k] + rj;
                if (d2 <= rs * rs) {
                    const int i = ii[k], j = j0 + jj;
                    if (i >= 0 && i != j && pred_exact(d2, st.radius[i], st.radius[j]))
                        emit_pair(st, sc, i, j);
                }
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------
// Uniform-grid broad phase
// ---------------------------------------------------------------------------------------
// The grid of the current frame lives on the device (scratch.grid_dev, cosmo_grid): it is either
// copied from the step parameters (detect_mode 1, k_grid_set) or TUNED ON THE DEVICE from the
// frame's own positions and radii (detect_mode 2, k_grid_hist + k_grid_tune) -- the host never
// reads planet data while stepping.  Only two properties matter for correctness, and both are
// enforced here whatever the tuning says: the cell edge is at least the largest separation at which
// two non-giants can be flagged, and every planet above the giant radius is brute-forced.


// Tuning scratch (scratch.tune_hist, int32 words, all zero between frames):
//   [0, 16384)           pass 1: x coordinates of a strided sample of the planets
//                        pass 2: 128 x 128 occupancy of the bulk box (all planets)
//   [16384, 32768)       pass 1: y coordinates
//   [32768, 36864)       pass 1: inflated radii of ALL planets
//   [36864, 36928)       32 doubles handed from k_grid_tune_box to k_grid_tune
// Coordinates are binned by the top 14 bits of an order-preserving key of their FP32 value
// (sign, exponent, 5 mantissa bits: 3 % relative resolution anywhere, no range to choose);
// radii by exponent + 4 mantissa bits (6 %).
constexpr int TUNE_XY_BINS = 16384, TUNE_R_BINS = 4096, TUNE_OCC = 128, TUNE_XY_SHIFT = 18;
constexpr int TUNE_SAMPLE = 131072;       // coordinates sampled for the quantile box
constexpr int TUNE_AUX = 2 * TUNE_XY_BINS + TUNE_R_BINS;   // word offset of the doubles
static_assert(TUNE_AUX + 64 == COSMO_TUNE_HIST_LEN, "header and kernels disagree on tune_hist");
static_assert(TUNE_OCC * TUNE_OCC <= TUNE_XY_BINS, "occupancy map must fit the x histogram");

__device__ __forceinline__ uint32_t tune_key(double v) {
    const float f = (float)fmin(fmax(v, -3.0e38), 3.0e38);
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ double tune_key_value(uint32_t key) {
    const uint32_t u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    return (double)__uint_as_float(u);
}
__device__ __forceinline__ int tune_rbin(double r) {
    const float f = (float)fmin(fmax(r, 0.0), 3.0e38);
    return (int)((__float_as_uint(f) >> 19) & (TUNE_R_BINS - 1));
}
__device__ __forceinline__ double tune_rbin_upper(int b) {
    return (double)__uint_as_float(((uint32_t)b << 19) | 0x7ffffu);
}

// pass 1: max |pos_new|^2 of the frame and, with TUNE, the 1-D tuning histograms
template <int SRC, bool TUNE>
__global__ void __launch_bounds__(256) k_grid_hist(cosmo_state st, cosmo_scratch sc) {
    __shared__ int s_r[TUNE ? TUNE_R_BINS : 1];
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    if (TUNE) {
        for (int q = threadIdx.x; q < TUNE_R_BINS; q += blockDim.x) s_r[q] = 0;
        __syncthreads();
    }
    const int stride = max(1, n / TUNE_SAMPLE);
    double m = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        m = fmax(m, xx_of<SRC>(st, sc, k));
        if (TUNE) {
            atomicAdd(&s_r[tune_rbin(sc.rinf[k])], 1);
            if (k % stride == 0) {
                const double2 p = xy_of<SRC>(st, sc, k);
                atomicAdd(&sc.tune_hist[tune_key(p.x) >> TUNE_XY_SHIFT], 1);
                atomicAdd(&sc.tune_hist[TUNE_XY_BINS + (tune_key(p.y) >> TUNE_XY_SHIFT)], 1);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0)  // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_PMAX2]),
                  (unsigned long long)__double_as_longlong(m));
    if (TUNE) {
        __syncthreads();
        for (int q = threadIdx.x; q < TUNE_R_BINS; q += blockDim.x)
            if (s_r[q]) atomicAdd(&sc.tune_hist[2 * TUNE_XY_BINS + q], s_r[q]);
    }
}

// Planets farther from the origin than 2^24 giant radii are brute-forced like giants: the rounding
// slack of the reference's expanded d^2 grows with |p|^2, and one planet flung out to 10^4 AU must
// not inflate every cell of the grid (grid_hmin then uses this bound instead of the true maximum).
__device__ __forceinline__ double grid_far2(double rg) {
    const double f = 16777216.0 * rg;
    return f * f;
}

// Smallest admissible cell edge.  The reference compares the COMPUTED distance (expanded form,
// absolute error E in d^2 of a few ulps of |p|^2) with r_i + r_j, so two non-giants (inflated
// radius <= rg, |p|^2 <= far2) can be flagged while truly sqrt(4 rg^2 + E) apart; the cell edge is
// never smaller than that (plus slack for the rounding of the cell coordinate itself).
__device__ __forceinline__ double grid_hmin(const cosmo_state &st, double rg) {
    const double pmax2 = fmin(__longlong_as_double(st.status[COSMO_ST_PMAX2]), grid_far2(rg));
    const double e = 64.0 * 1.1102230246251565e-16 * pmax2;
    return sqrt(4.0 * rg * rg + e) * (1.0 + 1e-6);
}

// detect_mode 1: the caller's grid (cell_size, origin of the linear box, cells per side, giant radius)
__global__ void k_grid_set(cosmo_state st, cosmo_scratch sc, cosmo_grid *gd, double h, double ox, double oy, double rg, int c) {
    cosmo_grid g;
    const double half_in = 0.5 * (double)c * h;    // the centre is fixed by the caller's (h, origin)
    g.h = fmax(h, grid_hmin(st, rg));
    g.cx = ox + half_in; g.cy = oy + half_in; g.rg = rg; g.far2 = grid_far2(rg);
    g.ncell = (int64_t)c * c; g.c = c; g.w = 0; g.source = 1; g.reserved = 0;
    g.box = (double)c * h; g.h_tuned = h; g.rho = 0.0; g.cand_total = 0; g.n_counted = 0;
    *gd = g;
}

// exclusive prefix sum over the 1024 threads of the CTA; *total = sum of all
__device__ __forceinline__ int block_scan_1024(int v, int *s_w, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = s_w[lane], y = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, y, o);
            if (lane >= o) y += t;
        }
        s_w[32 + lane] = y - x;                 // exclusive warp offsets
        if (lane == 31) s_w[64] = y;
    }
    __syncthreads();
    *total = s_w[64];
    return s_w[32 + w] + inc - v;
}

// detect_mode 2, step 1 (one CTA): the BULK box and the radius order statistics.
//   box   1.5 x the 10 % .. 90 % quantile range of the coordinates (NOT min/max, not even 2 % .. 98 %:
//         in a violent scene a few per cent of the planets are flung far out within a few frames;
//         the periodic grid has no border, they alias into the bulk's slots a few at a time)
//   big   (g+1)-th largest inflated radius (upper bin edge) for g = 0, 1, 2, 4, ..., 1024
// Results go to the aux doubles; the 1-D histograms return to zero.
__global__ void __launch_bounds__(1024) k_grid_tune_box(cosmo_state st, cosmo_scratch sc) {
    __shared__ int s_w[65];
    __shared__ double s_q[4];          // x lo, x hi, y lo, y hi
    __shared__ double s_big[12];
    const int tid = threadIdx.x;
    if (tid < 4) s_q[tid] = 0.0;
    if (tid < 12) s_big[tid] = -1.0;
    // ---- coordinate quantiles: thread t owns bins [16 t, 16 t + 16) of each axis (four 16-byte loads)
    for (int axis = 0; axis < 2; ++axis) {
        int4 *hist4 = reinterpret_cast<int4 *>(sc.tune_hist + axis * TUNE_XY_BINS) + tid * 4;
        int v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int4 t4 = hist4[q];
            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
            hist4[q] = make_int4(0, 0, 0, 0);
        }
        int loc = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) loc += v[q];
        int total;
        const int before = block_scan_1024(loc, s_w, &total);
        const long long r_lo = (long long)(0.10 * total), r_hi = min((long long)(0.90 * total), (long long)total - 1);
        if (total > 0) {
            int run = before;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const uint32_t bin = (uint32_t)(tid * 16 + q);
                if (run <= r_lo && r_lo < run + v[q]) s_q[2 * axis] = tune_key_value(bin << TUNE_XY_SHIFT);
                if (run <= r_hi && r_hi < run + v[q])
                    s_q[2 * axis + 1] = tune_key_value((bin << TUNE_XY_SHIFT) | ((1u << TUNE_XY_SHIFT) - 1u));
                run += v[q];
            }
        }
        __syncthreads();
    }
    // ---- radius order statistics from the top: thread t owns bins [4 t, 4 t + 4)
    {
        int4 *hist4 = reinterpret_cast<int4 *>(sc.tune_hist + 2 * TUNE_XY_BINS) + tid;
        const int4 t4 = *hist4;
        *hist4 = make_int4(0, 0, 0, 0);
        const int v[4] = {t4.x, t4.y, t4.z, t4.w};
        const int loc = v[0] + v[1] + v[2] + v[3];
        int total;
        const int before = block_scan_1024(loc, s_w, &total);
        int above = total - before - loc;               // planets in bins above this thread's
#pragma unroll
        for (int q = 3; q >= 0; --q) {
            const int cnt = v[q];
            if (cnt) {
#pragma unroll
                for (int gi = 0; gi < 12; ++gi) {
                    const int g = gi == 0 ? 0 : 1 << (gi - 1);
                    if (above < g + 1 && g + 1 <= above + cnt) s_big[gi] = tune_rbin_upper(tid * 4 + q);
                }
            }
            above += cnt;
        }
    }
    __syncthreads();
    double *aux = reinterpret_cast<double *>(sc.tune_hist + TUNE_AUX);
    if (tid < 12) aux[4 + tid] = s_big[tid];
    if (tid != 0) return;
    double box = fmax(s_q[1] - s_q[0], s_q[3] - s_q[2]) * 1.5;
    const double pmax = sqrt(__longlong_as_double(st.status[COSMO_ST_PMAX2]));
    if (!(box > 0.0) || !isfinite(box)) box = fmax(2.0 * pmax, 1e-300);
    double cx = 0.5 * (s_q[0] + s_q[1]), cy = 0.5 * (s_q[2] + s_q[3]);
    if (!isfinite(cx) || !isfinite(cy)) { cx = 0.0; cy = 0.0; }
    aux[0] = cx; aux[1] = cy; aux[2] = box;
}

// detect_mode 2, step 2: 128 x 128 occupancy of the bulk box
template <int SRC>
__global__ void __launch_bounds__(256) k_grid_occupancy(cosmo_state st, cosmo_scratch sc) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const double *aux = reinterpret_cast<const double *>(sc.tune_hist + TUNE_AUX);
    const double cx = aux[0], cy = aux[1], inv = (double)TUNE_OCC / aux[2];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double2 p = xy_of<SRC>(st, sc, k);
        const double u = (p.x - cx) * inv + 0.5 * TUNE_OCC, v = (p.y - cy) * inv + 0.5 * TUNE_OCC;
        if (u >= 0.0 && u < (double)TUNE_OCC && v >= 0.0 && v < (double)TUNE_OCC)
            atomicAdd(&sc.tune_hist[(int)v * TUNE_OCC + (int)u], 1);
    }
}

// detect_mode 2, step 3 (one CTA): the frame's grid.
//   rho   density a planet sees on average (a plain n / area is wrong by orders of magnitude for a
//         disk, a ring or a cluster): sum c (c - 1) / (sum c * cell area) over the occupancy map, or --
//         finer than that map can resolve -- the density MEASURED by the previous frame's grid,
//         candidates / (9 h^2) per planet (k_grid_density; integer sums, identical on every rank)
//   rg    giant radius from a cost model: g giants cost g*n exact tests, everybody else
//         9 cells * rho * (2 rg)^2 candidates; g in {0, 1, 2, 4, ..., 1024}
//   h     cell edge for ~16 candidates per planet, never below 2 rg (nor below grid_hmin), never so
//         small that the linear part of the grid could not cover the bulk box
// and returns the occupancy map to zero.
__global__ void __launch_bounds__(1024) k_grid_tune(cosmo_state st, cosmo_scratch sc, cosmo_grid *gd) {
    __shared__ double s_a[32], s_b[32];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    double a = 0.0, b = 0.0;
    for (int q = tid; q < TUNE_OCC * TUNE_OCC; q += 1024) {
        const double c = (double)sc.tune_hist[q];
        a += c * (c - 1.0);
        b += c;
        sc.tune_hist[q] = 0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { s_a[wp] = a; s_b[wp] = b; }
    __syncthreads();
    if (tid != 0) return;
    a = 0.0; b = 0.0;
    for (int q = 0; q < 32; ++q) { a += s_a[q]; b += s_b[q]; }
    double *aux = reinterpret_cast<double *>(sc.tune_hist + TUNE_AUX);
    cosmo_grid g;
    g.cx = aux[0]; g.cy = aux[1];
    const double box = aux[2];
    const double dn = (double)max(n, 1);
    const double cell_area = (box / TUNE_OCC) * (box / TUNE_OCC);
    double rho = b > 0.0 ? a / (b * cell_area) : 0.0;
    rho = fmax(rho, dn / (box * box) * 1e-3);              // nearly empty map: fall back towards the mean density
    const cosmo_grid prev = *gd;
    if (prev.source == 2 && prev.n_counted > 0 && prev.cand_total > 0 && prev.h > 0.0 && isfinite(prev.h))
        rho = fmax(rho, (double)prev.cand_total / ((double)prev.n_counted * 9.0 * prev.h * prev.h));
    double best = -1.0, rg = 0.0;
    for (int gi = 0; gi < 12; ++gi) {
        const int gg = gi == 0 ? 0 : 1 << (gi - 1);
        const double big = aux[4 + gi];
        if (gg > n - 1 || big < 0.0) break;
        const double cost = gg * dn + dn * 9.0 * rho * (2.0 * big) * (2.0 * big);
        if (best < 0.0 || cost < best) { best = cost; rg = big; }
    }
    int cmax = (int)floor(sqrt((double)sc.grid_cells));
    while ((long long)cmax * cmax > sc.grid_cells) --cmax;
    if (cmax < 3) cmax = 3;
    const double span = fmax(box, 4.0 * rg) * 1.1;    // one period holds the bulk: bulk cells do not alias each other
    double h = fmax(fmax(2.0 * rg * 1.001, span / cmax), sqrt(16.0 / (9.0 * rho)));   // ~16 candidates per planet
    int c = (int)ceil(span / h);
    c = max(3, min(cmax, c));
    rg = fmax(rg, h / (2.0 * 1.001));     // cells wider than 2 rg: admit larger planets for free
    g.h_tuned = h;
    g.h = fmax(h, grid_hmin(st, rg));
    g.rg = rg; g.far2 = grid_far2(rg); g.rho = rho;
    g.c = c; g.w = 0; g.ncell = (int64_t)g.c * g.c; g.source = 2; g.reserved = 0; g.box = box;
    g.cand_total = 0; g.n_counted = 0;
    *gd = g;
}

template <int SRC>
__global__ void __launch_bounds__(256)
k_grid_count(cosmo_state st, cosmo_scratch sc, const cosmo_grid *gd, int all_in_cells) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const cosmo_grid g = *gd;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        // all_in_cells (near-field pass): every planet is binned, there is no brute-forced set
        if (!all_in_cells && (sc.rinf[k] > g.rg || xx_of<SRC>(st, sc, k) > g.far2)) {
            unsigned long long idx =
                atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_GIANTS]), 1ull);
            if ((long long)idx < sc.giant_cap)
                sc.giants[idx] = k;
            else
                atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                         (unsigned long long)COSMO_ERR_GIANT_OVERFLOW);
            sc.cell_of[k] = -1;
        } else {
            const double2 p = xy_of<SRC>(st, sc, k);
            const int cell = cell_coord(p.y, g.cy, g.h, g.c) * g.c + cell_coord(p.x, g.cx, g.h, g.c);
            sc.cell_of[k] = cell;
            atomicAdd(&sc.cell_count[cell], 1);
        }
    }
}

// counting-sort scatter; counts return to zero, so cell_count needs no per-frame memset
__global__ void __launch_bounds__(256) k_grid_fill(cosmo_state st, cosmo_scratch sc) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int cell = sc.cell_of[k];
        if (cell >= 0) {
            const int slot = sc.cell_start[cell] + atomicSub(&sc.cell_count[cell], 1) - 1;
            sc.cell_items[slot] = k;
        }
    }
}

// Candidates the grid produces, summed over ALL non-giant planets (every rank runs this over all
// rows, so the sums -- integers -- are identical everywhere): feedback for the next frame's tuner.
__global__ void __launch_bounds__(256) k_grid_density(cosmo_state st, cosmo_scratch sc, cosmo_grid *gd) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int c = gd->c;
    unsigned long long cand = 0, cnt = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int cell = sc.cell_of[k];
        if (cell < 0) continue;
        int s0[6], s1[6];
        cand += (unsigned long long)cand_ranges(sc, cell, c, s0, s1);
        ++cnt;
    }
    for (int o = 16; o > 0; o >>= 1) {
        cand += __shfl_xor_sync(0xffffffffu, cand, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(reinterpret_cast<unsigned long long *>(&gd->cand_total), cand);
        atomicAdd(reinterpret_cast<unsigned long long *>(&gd->n_counted), cnt);
    }
}

struct GridRow {
    double qx, qy, qz, xxi, ri;
};

template <int SRC>
__device__ __forceinline__ GridRow load_row(const cosmo_state &st, const cosmo_scratch &sc, int i) {
    GridRow r;
    const double2 p = xy_of<SRC>(st, sc, i);
    r.qx = -2.0 * p.x; r.qy = -2.0 * p.y;
    r.qz = (SRC == SRC_NEW3) ? -2.0 * sc.pos_new[2 * (size_t)st.cap + i] : 0.0;
    r.xxi = sc.xx_new[i]; r.ri = sc.rinf[i];
    return r;
}

template <int SRC>
__device__ __forceinline__ void test_and_emit(const cosmo_state &st, const cosmo_scratch &sc, int i, const GridRow &r,
                                              int j) {
    const double2 pj = xy_of<SRC>(st, sc, j);
    double d2;
    if (SRC == SRC_NEW3) {
        const double m2dot = __fma_rn(r.qz, sc.pos_new[2 * (size_t)st.cap + j], __fma_rn(r.qy, pj.y, __dmul_rn(r.qx, pj.x)));
        d2 = __dadd_rn(__dadd_rn(m2dot, r.xxi), sc.xx_new[j]);
    } else {
        d2 = pred_d2(r.qx, r.qy, r.xxi, pj.x, pj.y, sc.xx_new[j]);
    }
    const double rs = r.ri + sc.rinf[j];
    if (!(d2 <= rs * rs)) return;
    const bool hit = (SRC == SRC_NEW3) ? sqrt(fmax(d2, 0.0)) <= __dadd_rn(st.radius[i], st.radius[j])
                        : pred_exact(d2, st.radius[i], st.radius[j]);
    if (hit) emit_pair(st, sc, i, j);
}

// The cell-sorted order is cut into chunks of GRID_CHUNK slots whose boundaries are snapped to
// CELL boundaries (the order of the planets inside a cell comes from atomics and differs from rank
// to rank, the cell populations do not), and the chunks are dealt round-robin to the ranks: the
// crowded inner cells of a disk -- most of the candidate work -- are spread over all of them.
// bounds[k] = first slot of the first cell starting at or after slot k * GRID_CHUNK.
constexpr int GRID_CHUNK = 1024;
constexpr int GRID_HEAVY = 128;   // candidates above which a row goes to the warp-per-row pass

__global__ void __launch_bounds__(128) k_grid_ranges(cosmo_state st, cosmo_scratch sc, const cosmo_grid *gd) {
    const int ncell = (int)gd->ncell;
    const int total = sc.cell_start[ncell];
    const int nchunk = (total + GRID_CHUNK - 1) / GRID_CHUNK;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= nchunk; k += gridDim.x * blockDim.x) {
        const long long target = (long long)k * GRID_CHUNK;
        int bound = total;
        if (target < total) {
            int lo = 0, hi = ncell;                    // smallest cell with cell_start[cell] >= target
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sc.cell_start[mid] >= target) hi = mid; else lo = mid + 1;
            }
            bound = sc.cell_start[lo];
        }
        sc.chunk_alive[k] = bound;                     // scratch shared with stage C's compaction (later)
    }
}

// One thread per non-giant target row, taken in CELL-SORTED order: the lanes of a warp then
// work on planets of the same or neighbouring cells, so their candidate loops have similar trip
// counts (the disk's surface density varies 20x between its inner and outer edge) and their
// gathers hit the same lines.  Every row is produced by exactly one rank.
template <int SRC>
__global__ void __launch_bounds__(128)
k_grid_detect(cosmo_state st, cosmo_scratch sc, const cosmo_grid *gd, int shard_index, int shard_count) {
    const int c = gd->c;
    const int total = sc.cell_start[c * c];
    const int nchunk = (total + GRID_CHUNK - 1) / GRID_CHUNK;
    unsigned cand_total = 0;
  for (int chunk = shard_index + blockIdx.x * shard_count; chunk < nchunk; chunk += gridDim.x * shard_count) {
   const int s_lo = sc.chunk_alive[chunk], s_hi = sc.chunk_alive[chunk + 1];
   // blockIdx.y splits the chunk's slots: a normal chunk is one slot per thread, an oversized one
   // (a crowded cell is never split across chunks) is still shared by gridDim.y blocks
   for (int slot = s_lo + blockIdx.y * blockDim.x + threadIdx.x; slot < s_hi; slot += gridDim.y * blockDim.x) {
    const int i = sc.cell_items[slot];
    const int cell = sc.cell_of[i];
    const GridRow row = load_row<SRC>(st, sc, i);
    // candidate ranges: the three cells of a grid row are contiguous in the sorted item array
    int s0[6], s1[6];
    const long long cand = cand_ranges(sc, cell, c, s0, s1);
    const int ng = (int)min((long long)st.status[COSMO_ST_N_GIANTS], (long long)sc.giant_cap);
    cand_total += (unsigned)cand;
    if (cand + ng > GRID_HEAVY) {
        // a crowded neighbourhood (dense inner disk, a tight cluster): one thread walking thousands
        // of candidates would be the critical path of the whole frame -> warp-per-row pass
        const unsigned long long h =
            atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_HEAVY]), 1ull);
        sc.label[h] = i;       // `label` is free until stage C
        continue;
    }
#pragma unroll
    for (int d = 0; d < 6; ++d)
        for (int s = s0[d]; s < s1[d]; ++s) {
            const int j = sc.cell_items[s];
            if (j != i) test_and_emit<SRC>(st, sc, i, row, j);
        }
    for (int g = 0; g < ng; ++g) test_and_emit<SRC>(st, sc, i, row, sc.giants[g]);
   }
  }
    // diagnostic only (all lanes are back together after the loop)
    const unsigned tot = __reduce_add_sync(0xffffffffu, cand_total);
    if ((threadIdx.x & 31) == 0 && tot)
        atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_CAND]), (unsigned long long)tot);
}

// Warp-per-row pass for the rows k_grid_detect set aside: the lanes stride over the candidates.
template <int SRC>
__global__ void __launch_bounds__(256) k_grid_detect_heavy(cosmo_state st, cosmo_scratch sc, const cosmo_grid *gd) {
    const int c = gd->c;
    const int nh = (int)st.status[COSMO_ST_N_HEAVY];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    const int ng = (int)min((long long)st.status[COSMO_ST_N_GIANTS], (long long)sc.giant_cap);
    for (int h = warp; h < nh; h += nwarp) {
        const int i = sc.label[h];
        const int cell = sc.cell_of[i];
        const GridRow row = load_row<SRC>(st, sc, i);
        int s0[6], s1[6];
        cand_ranges(sc, cell, c, s0, s1);
        for (int d = 0; d < 6; ++d)
            for (int s = s0[d] + lane; s < s1[d]; s += 32) {
                const int j = sc.cell_items[s];
                if (j != i) test_and_emit<SRC>(st, sc, i, row, j);
            }
        for (int g = lane; g < ng; g += 32) test_and_emit<SRC>(st, sc, i, row, sc.giants[g]);
    }
}

// rows of the giants: giant g against every planet (blockIdx.y strides over the giants)
template <int SRC>
__global__ void __launch_bounds__(256)
k_giant_rows(cosmo_state st, cosmo_scratch sc, int i_begin, int i_end) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int ng = (int)min((long long)st.status[COSMO_ST_N_GIANTS], (long long)sc.giant_cap);
    for (int gi = blockIdx.y; gi < ng; gi += gridDim.y) {
        const int g = sc.giants[gi];
        if (g < i_begin || g >= i_hi) continue;
        const GridRow row = load_row<SRC>(st, sc, g);
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
            if (j != g) test_and_emit<SRC>(st, sc, g, row, j);
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
constexpr int DET_BLOCK = 128, DET_IB = 2, DET_TJ = 256;
constexpr int DETS_BLOCK = 64, DETS_IB = 1, DETS_TJ = 64;

int detect_residency(int *rows_per_cta, int *tile) {
    int nb = 0;
    *rows_per_cta = DET_BLOCK * DET_IB; *tile = DET_TJ;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_detect_allpairs<DET_BLOCK, DET_IB, DET_TJ>, DET_BLOCK, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nb;
}

// near-field pass: last frame's stage-B grid serves this frame's start positions (they ARE last
// frame's new positions, up to merges); its cells only have to be wider than the near-field cutoff
__global__ void k_grid_reuse(const cosmo_grid *from, cosmo_grid *to, double h_floor_rel) {
    cosmo_grid g = *from;
    g.h = fmax(g.h, h_floor_rel * g.box);
    g.cand_total = 0; g.n_counted = 0; g.source = 3;
    *to = g;
}

// Binning of the planets into the periodic grid *gd: (grid from the parameters | tuned on the device
// | last frame's) -> cell populations -> exclusive scan -> cell-sorted item array.
template <int SRC>
static int grid_build(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cosmo_grid *gd,
                      int mode, bool all_in_cells, double reuse_floor, cudaStream_t s) {
    if (!sc.cell_count || !sc.cell_of || !sc.cell_items || !gd || sc.grid_cells < 9 || (!all_in_cells && !sc.giants)) {
        set_error("grid: scratch lacks the cell arrays / grid_dev");
        return COSMO_E_ARG;
    }
    int nb = (p.n_upper + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > 148 * 8) nb = 148 * 8;
    if (mode == 2) {
        // grid tuned on the device from this frame's positions and radii: no host round trip
        if (!sc.tune_hist) {
            set_error("grid: device-side tuning needs scratch.tune_hist");
            return COSMO_E_ARG;
        }
        k_grid_hist<SRC, true><<<nb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_hist"));
        k_grid_tune_box<<<1, 1024, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_tune_box"));
        k_grid_occupancy<SRC><<<nb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_occupancy"));
        k_grid_tune<<<1, 1024, 0, s>>>(st, sc, gd);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_tune"));
    } else if (mode == 1) {
        const int c = p.cells_per_side;
        if (c < 3 || (long long)c * c > sc.grid_cells) {
            set_error("grid: cells_per_side=%d must be >= 3 and fit grid_cells=%d", c, sc.grid_cells);
            return COSMO_E_ARG;
        }
        k_grid_hist<SRC, false><<<nb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_hist"));
        k_grid_set<<<1, 1, 0, s>>>(st, sc, gd, p.cell_size, p.grid_origin_x, p.grid_origin_y, p.giant_radius, c);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_set"));
    } else {
        k_grid_reuse<<<1, 1, 0, s>>>(sc.grid_dev, gd, reuse_floor);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_reuse"));
    }
    k_grid_count<SRC><<<nb, 256, 0, s>>>(st, sc, gd, all_in_cells ? 1 : 0);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_count"));
    // the cell count c*c lives on the device (gd->ncell); blocks beyond it exit at once
    const long long scan_max = mode == 1 ? (long long)p.cells_per_side * p.cells_per_side : (long long)sc.grid_cells;
    COSMO_TRY(launch_scan_exclusive(sc.cell_count, sc.cell_start, sc.scan_blk, &gd->ncell, scan_max, 0, s));
    k_grid_fill<<<nb, 256, 0, s>>>(st, sc);
    return COSMO_LAUNCH_CHECK("k_grid_fill");
}

int launch_grid_build_old2(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cosmo_grid *gd,
                           int mode, double reuse_floor, cudaStream_t s) {
    return grid_build<SRC_OLD2>(st, sc, p, gd, mode, true, reuse_floor, s);
}

int launch_grid_build_old3(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cosmo_grid *gd,
                           int mode, double reuse_floor, cudaStream_t s) {
    return grid_build<SRC_OLD3>(st, sc, p, gd, mode, true, reuse_floor, s);
}

template <int SRC>
static int launch_grid(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cudaStream_t s) {
    const int rows = p.i_end - p.i_begin;
    cosmo_grid *gd = sc.grid_dev;
    COSMO_TRY(grid_build<SRC>(st, sc, p, gd, p.detect_mode == 2 ? 2 : 1, false, 0.0, s));
    int nb = (p.n_upper + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > 148 * 8) nb = 148 * 8;
    if (p.detect_mode == 2) {
        k_grid_density<<<nb, 256, 0, s>>>(st, sc, gd);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_density"));
    }
    const int sh_n = p.shard_count > 0 ? p.shard_count : 1;
    const int sh_i = p.shard_count > 0 ? p.shard_index : 0;
    const int nchunk = (p.n_upper + GRID_CHUNK - 1) / GRID_CHUNK + 1;
    k_grid_ranges<<<(nchunk + 127) / 128, 128, 0, s>>>(st, sc, gd);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_ranges"));
    const int blocks = (nchunk + sh_n - 1) / sh_n;
    k_grid_detect<SRC><<<dim3(blocks > 0 ? blocks : 1, GRID_CHUNK / 128), 128, 0, s>>>(st, sc, gd, sh_i, sh_n);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_detect"));
    k_grid_detect_heavy<SRC><<<296, 256, 0, s>>>(st, sc, gd);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_detect_heavy"));
    if (rows > 0) {
        k_giant_rows<SRC><<<dim3(nb > 296 ? 296 : nb, 8), 256, 0, s>>>(st, sc, p.i_begin, p.i_end);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_giant_rows"));
    }
    return COSMO_OK;
}

int launch_detect_grid(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cudaStream_t s,
                       bool planes3) {
    return planes3 ? launch_grid<SRC_NEW3>(st, sc, p, s) : launch_grid<SRC_NEW2>(st, sc, p, s);
}

int launch_detect(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                  cudaStream_t s) {
    if (!(p.opts & COSMO_OPT_COLLISIONS)) return COSMO_OK;
    const int rows = p.i_end - p.i_begin;
    if (rows <= 0 && p.detect_mode == 0) return COSMO_OK;
    if (p.detect_mode != 0) return launch_detect_grid(st, sc, p, s, false);
    int jsplit = p.jsplit_detect < 1 ? 1 : p.jsplit_detect;
    if (p.small_tiles) {
        const int per = DETS_BLOCK * DETS_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        k_detect_allpairs<DETS_BLOCK, DETS_IB, DETS_TJ><<<grid, DETS_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit);
    } else {
        const int per = DET_BLOCK * DET_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        k_detect_allpairs<DET_BLOCK, DET_IB, DET_TJ><<<grid, DET_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit);
    }
    return COSMO_LAUNCH_CHECK("k_detect_allpairs");
}

}  // namespace cosmo
