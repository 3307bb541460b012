/*
 * cosmo_oracle.c -- CPU restatement of the reference's per-timestep physics update.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this library; the product (cosmosim_b200/)
 * never imports, links or calls it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4), so the pin is the reference itself: oracle/gen_golden.py runs the
 * unmodified reference in the build container and commits its outputs under tests/golden/;
 * tests/test_oracle_golden.py checks every function below against those fixtures, and
 * tests/test_oracle_vs_reference.py re-checks against the live reference whenever
 * /root/reference is present.
 *
 * What is restated (reference file:line):
 *   oracle_accel       cosmosim/util/blas.py:4-41 (acc_blas), called at
 *                      cosmosim/core/universe_old.py:174.  The arithmetic lives in OpenBLAS
 *                      (scipy.linalg.blas zhpr/dspr2/zhpmv; unpinned, scipy 1.18.1 /
 *                      OpenBLAS 0.3.31.dev here); its published algorithm is restated
 *                      elementwise with the exact rounding sequence of the d^2 chain.
 *   oracle_integrate   universe_old.py:177-178.
 *   oracle_flags       universe_old.py:180,192-193 + sklearn.metrics.pairwise_distances
 *                      (sklearn 1.9.0 euclidean_distances: row_norms via einsum, dot via
 *                      OpenBLAS dgemm K=2) -- restated elementwise.
 *   oracle_frame       universe_old.py:345-350 (update_planet_lists, interact,
 *                      destroy_escaped_planets) with the merge loop :196-210 and
 *                      Planet.absorb/destroy cosmosim/core/planet.py:74-91.
 *
 * Build: gcc -O3 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every product/sum below that the reference rounds
 * separately must be rounded separately here; the only fused operation is the explicit
 * fma() that mirrors the dgemm micro-kernel.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ---------------------------------------------------------------------------------------
 * State: one entry per planet in INSERTION order (the order of Universe.planets,
 * universe_old.py:28,42).  Masses follow Python semantics: the example scenes draw masses
 * with random.randint, i.e. arbitrary-precision ints (examples/star_with_disk.py:39,
 * universe_old.py:53) while the star's is a float; `mass += other.mass` (planet.py:80) is
 * exact for int+int and float64 otherwise.  mass_f64 always holds float(mass).
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int64_t n;
    double *pos;          /* [n][2]; undefined for dead planets (position=None, planet.py:90) */
    double *vel;          /* [n][2] */
    double *mass_f64;     /* [n]  float(mass), round-to-nearest-even                           */
    uint64_t *mass_hi;    /* [n]  high/low words of the integer mass when mass_is_int          */
    uint64_t *mass_lo;
    uint8_t *mass_is_int; /* [n] */
    double *radius;       /* [n] */
    double *volume;       /* [n] */
    uint8_t *alive;       /* [n] */
    uint8_t *immobile;    /* [n] */
    int64_t *eaten;       /* [n] planets_eaten */
} oracle_state;

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* float(int) for a 128-bit unsigned int, round-half-even like CPython's int.__float__ */
static double u128_to_double(u128 v) {
    if (v == 0) return 0.0;
    int top = 127;
    while (!((v >> top) & 1)) --top;          /* index of the leading one */
    if (top <= 52) return (double)(uint64_t)v;
    int shift = top - 52;
    uint64_t mant = (uint64_t)(v >> shift);   /* 53 significant bits */
    u128 rem = v & ((((u128)1) << shift) - 1);
    u128 half = ((u128)1) << (shift - 1);
    if (rem > half || (rem == half && (mant & 1))) ++mant;
    return ldexp((double)mant, shift);        /* mant may be 2^53: still exact */
}

/* ---------------------------------------------------------------------------------------
 * A1: acc_blas restated (blas.py:9-41).  pos [n][2], mass [n] (already float64), out [n][2].
 *
 *   per dim k:  R_k(i,j) = fl(fl(q_ik * p_jk) - 2),  q_ik = -2 p_ik   (zhpr, blas.py:11;
 *               real part of  -2 (p_i - 1j)(p_j + 1j))
 *   c(i,j)    = fl(fl(R_x + R_y) + 6)                              (blas.py:23; 3rd row is 0)
 *   d2(i,j)   = fl(fl(c(i,j) - h_hi) - h_lo),  h_k = c(k,k)/2, hi=max(i,j), lo=min(i,j)
 *                                                                   (dspr2 = two axpy per
 *                                                                   column, blas.py:27)
 *   w         = fl(d2 * fl(sqrt(d2)));  t_k = fl(2 fl(p_jk - p_ik) / w) if d2 != 0 else
 *               2 fl(p_jk - p_ik)                                   (blas.py:31, where= mask)
 *   a_ik      = G * sum_j fl(t_k * (m_j / 2))                       (zhpmv alpha=.5, blas.py:40-41)
 * The summation order inside zhpmv is OpenBLAS-internal and not reproducible.  The oracle adds
 * the (bit-exactly reproduced) pair terms with a compensated (Neumaier) sum, i.e. it returns the
 * correctly rounded sum of the reference's terms to within an ulp or two; the real acc_blas
 * agrees with it to <= 2e-14 on the fixture scenes (tests pin this).  In crowded scenes, rows
 * whose large neighbour terms cancel (sum of |terms| / |sum| up to ~500 at N = 65536) carry a
 * summation-order noise of ~1e-13 in ANY plain summation order, OpenBLAS' included.
 * ------------------------------------------------------------------------------------- */
static inline double chain_c(double qix, double qiy, double pjx, double pjy) {
    double rx = qix * pjx;
    rx = rx - 2.0;
    double ry = qiy * pjy;
    ry = ry - 2.0;
    double c = rx + ry;
    return c + 6.0;
}

static inline void accel_rows_impl(int64_t n, const double *pos, const double *mass, double G,
                                   int64_t i_begin, int64_t i_end, double *out, const int compensated) {
    double *h = (double *)malloc((size_t)n * sizeof(double));
    double *mh = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k) {
        double x = pos[2 * k], y = pos[2 * k + 1];
        h[k] = 0.5 * chain_c(-2.0 * x, -2.0 * y, x, y);
        mh[k] = 0.5 * mass[k];
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = i_begin; i < i_end; ++i) {
        const double pix = pos[2 * i], piy = pos[2 * i + 1];
        const double qix = -2.0 * pix, qiy = -2.0 * piy;
        const double hi_i = h[i];
        double sx = 0.0, sy = 0.0, cx = 0.0, cy = 0.0;   /* Neumaier-compensated sums */
        for (int64_t j = 0; j < n; ++j) {
            const double pjx = pos[2 * j], pjy = pos[2 * j + 1];
            double c = chain_c(qix, qiy, pjx, pjy);
            double d2;
            if (j >= i) { d2 = c - h[j]; d2 = d2 - hi_i; }
            else        { d2 = c - hi_i; d2 = d2 - h[j]; }
            double ix = 2.0 * (pjx - pix), iy = 2.0 * (pjy - piy);
            if (d2 != 0.0) {
                double w = d2 * sqrt(d2);
                ix = ix / w;
                iy = iy / w;
            }
            if (compensated) {
                double tx = ix * mh[j], ty = iy * mh[j];
                double nx = sx + tx, ny = sy + ty;
                cx += (fabs(sx) >= fabs(tx)) ? ((sx - nx) + tx) : ((tx - nx) + sx);
                cy += (fabs(sy) >= fabs(ty)) ? ((sy - ny) + ty) : ((ty - ny) + sy);
                sx = nx;
                sy = ny;
            } else {
                sx += ix * mh[j];
                sy += iy * mh[j];
            }
        }
        sx = sx + cx;
        sy = sy + cy;
        out[2 * (i - i_begin)] = G * sx;
        out[2 * (i - i_begin) + 1] = G * sy;
    }
    free(h);
    free(mh);
}

/* parity checks: compensated sum of the reference's pair terms */
void oracle_accel_rows(int64_t n, const double *pos, const double *mass, double G,
                       int64_t i_begin, int64_t i_end, double *out /* [(i_end-i_begin)][2] */) {
    accel_rows_impl(n, pos, mass, G, i_begin, i_end, out, 1);
}

/* CPU baseline timing (bench.py): same terms, plain ascending-j accumulation -- what a CPU port
   written for speed would do (the compensated sum costs ~25 %) */
void oracle_accel_rows_plain(int64_t n, const double *pos, const double *mass, double G,
                             int64_t i_begin, int64_t i_end, double *out) {
    accel_rows_impl(n, pos, mass, G, i_begin, i_end, out, 0);
}

void oracle_accel(int64_t n, const double *pos, const double *mass, double G, double *out) {
    oracle_accel_rows(n, pos, mass, G, 0, n, out);
}

/* the d^2 matrix entry alone (used by the tests that pin the chain bit-for-bit) */
void oracle_d2_rows(int64_t n, const double *pos, int64_t i_begin, int64_t i_end, double *d2_out) {
    double *h = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k) {
        double x = pos[2 * k], y = pos[2 * k + 1];
        h[k] = 0.5 * chain_c(-2.0 * x, -2.0 * y, x, y);
    }
    for (int64_t i = i_begin; i < i_end; ++i) {
        const double qix = -2.0 * pos[2 * i], qiy = -2.0 * pos[2 * i + 1];
        for (int64_t j = 0; j < n; ++j) {
            double c = chain_c(qix, qiy, pos[2 * j], pos[2 * j + 1]);
            double d2;
            if (j >= i) { d2 = c - h[j]; d2 = d2 - h[i]; }
            else        { d2 = c - h[i]; d2 = d2 - h[j]; }
            d2_out[(i - i_begin) * n + j] = d2;
        }
    }
    free(h);
}

/* A2: universe_old.py:177-178 -- numpy elementwise, two roundings per update. */
void oracle_integrate(int64_t n, const double *p0, const double *v0, const double *a, double dt,
                      double *p, double *v) {
    for (int64_t k = 0; k < 2 * n; ++k) {
        double adt = a[k] * dt;
        double vn = v0[k] + adt;
        double vdt = vn * dt;
        v[k] = vn;
        p[k] = p0[k] + vdt;
    }
}

/* ---------------------------------------------------------------------------------------
 * A3: collision predicate on the NEW positions (universe_old.py:180,192-193).
 *   XX_i  = fl(fl(x_i^2) + fl(y_i^2))                    sklearn row_norms (einsum)
 *   dot   = fma(y_i, y_j, fl(x_i * x_j))                 OpenBLAS dgemm, K = 2
 *   d2    = fl(fl(-2 dot + XX_i) + XX_j)                 distances += XX; distances += YY
 *   d     = sqrt(max(d2, 0)); diagonal forced to 0       np.maximum / fill_diagonal
 *   flag  = max(d, 1) <= fl(r_i + r_j)                   np.clip(d,1,None) <= add.outer
 * Known limit of "bit-exact" here: the dot's contraction is the one of OpenBLAS' main dgemm
 * micro-kernel (verified bitwise on whole N x N matrices for N mod 8 < 4, incl. the fixture
 * sizes 1000/1001).  For N mod 8 >= 4 OpenBLAS handles a 4-wide block of tail rows/columns with
 * a kernel that does not fuse the K = 2 dot; those entries differ by a few ulps of |p|^2
 * (<= 3e-12 relative in d).  The flag can only differ if d is within that distance of r_i + r_j.
 * Emits the flagged ordered pairs (i,j), i != j, row-major ascending -- exactly the set and
 * order in which the merge loop's `if c and i != j` fires (universe_old.py:203-204).
 * Returns the number of flagged pairs (may exceed cap; only the first cap are stored).
 * ------------------------------------------------------------------------------------- */
static inline int flag_one(double xi, double yi, double xxi, double ri,
                           double xj, double yj, double xxj, double rj) {
    double dot = fma(yi, yj, xi * xj);
    double d2 = -2.0 * dot + xxi;
    d2 = d2 + xxj;
    if (d2 < 0.0) d2 = 0.0;
    double d = sqrt(d2);
    if (d < 1.0) d = 1.0;
    return d <= ri + rj;
}

int64_t oracle_flags_rows(int64_t n, const double *p, const double *radius,
                          int64_t i_begin, int64_t i_end, int64_t *pairs, int64_t cap) {
    double *xx = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k) {
        double x2 = p[2 * k] * p[2 * k];
        double y2 = p[2 * k + 1] * p[2 * k + 1];
        xx[k] = x2 + y2;
    }
    int64_t rows = i_end - i_begin;
    int64_t *row_cnt = (int64_t *)calloc((size_t)rows + 1, sizeof(int64_t));
    /* pass 1: count per row (parallel); pass 2: fill in row order */
#pragma omp parallel for schedule(static)
    for (int64_t i = i_begin; i < i_end; ++i) {
        double xi = p[2 * i], yi = p[2 * i + 1], xxi = xx[i], ri = radius[i];
        int64_t c = 0;
        for (int64_t j = 0; j < n; ++j)
            if (j != i && flag_one(xi, yi, xxi, ri, p[2 * j], p[2 * j + 1], xx[j], radius[j])) ++c;
        row_cnt[i - i_begin] = c;
    }
    int64_t total = 0;
    for (int64_t r = 0; r < rows; ++r) {
        int64_t c = row_cnt[r];
        if (c) {
            int64_t i = i_begin + r;
            double xi = p[2 * i], yi = p[2 * i + 1], xxi = xx[i], ri = radius[i];
            for (int64_t j = 0; j < n; ++j)
                if (j != i && flag_one(xi, yi, xxi, ri, p[2 * j], p[2 * j + 1], xx[j], radius[j])) {
                    if (total < cap) { pairs[2 * total] = i; pairs[2 * total + 1] = j; }
                    ++total;
                }
        }
    }
    free(row_cnt);
    free(xx);
    return total;
}

int64_t oracle_flags(int64_t n, const double *p, const double *radius, int64_t *pairs, int64_t cap) {
    return oracle_flags_rows(n, p, radius, 0, n, pairs, cap);
}

/* the full clipped distance d(i,j) for the tests that pin A3 bit-for-bit */
void oracle_dist_rows(int64_t n, const double *p, int64_t i_begin, int64_t i_end, double *d_out) {
    for (int64_t i = i_begin; i < i_end; ++i) {
        double xi = p[2 * i], yi = p[2 * i + 1];
        double xxi = xi * xi; { double y2 = yi * yi; xxi = xxi + y2; }
        for (int64_t j = 0; j < n; ++j) {
            double xj = p[2 * j], yj = p[2 * j + 1];
            double xxj = xj * xj; { double y2 = yj * yj; xxj = xxj + y2; }
            double dot = fma(yi, yj, xi * xj);
            double d2 = -2.0 * dot + xxi;
            d2 = d2 + xxj;
            if (d2 < 0.0) d2 = 0.0;
            double d = (i == j) ? 0.0 : sqrt(d2);
            d_out[(i - i_begin) * n + j] = d;
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * A7: energies and angular momentum (universe_old.py:180-189,210) on the new positions p,
 * new velocities v and the pre-merge masses m (float(mass)):
 *   d = clip(dist, 1, None); d_inv = 1/d; V = -G*m*d_inv (column j scaled by m_j), diag 0
 *   U = m * sum(V, axis=1); K = 0.5*m*|v|^2; total = sum(K+U); L = sum(m * cross(p, v))
 * numpy sums pairwise (float masses) or left-to-right (object dtype, Python int masses); the
 * order is not restated -- agreement is to rounding (tests: 1e-12).
 * ------------------------------------------------------------------------------------- */
void oracle_energies(int64_t n, const double *p, const double *v, const double *m, double G,
                     double *energy_out /* [n] */, double *totals_out /* [2] */) {
    double *xx = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k) {
        double x2 = p[2 * k] * p[2 * k];
        double y2 = p[2 * k + 1] * p[2 * k + 1];
        xx[k] = x2 + y2;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double xi = p[2 * i], yi = p[2 * i + 1], sum = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            double dot = fma(yi, p[2 * j + 1], xi * p[2 * j]);
            double d2 = -2.0 * dot + xx[i];
            d2 = d2 + xx[j];
            if (d2 < 0.0) d2 = 0.0;
            double d = sqrt(d2);
            if (d < 1.0) d = 1.0;
            double vij = -G * m[j];
            vij = vij * (1.0 / d);
            sum += vij;
        }
        double vx = v[2 * i], vy = v[2 * i + 1];
        double vmag = sqrt(vx * vx + vy * vy);
        double K = 0.5 * m[i] * (vmag * vmag);
        energy_out[i] = K + m[i] * sum;
    }
    double tot = 0.0, L = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        tot += energy_out[i];
        L += m[i] * (p[2 * i] * v[2 * i + 1] - p[2 * i + 1] * v[2 * i]);
    }
    totals_out[0] = tot;
    totals_out[1] = L;
    free(xx);
}

/* ---------------------------------------------------------------------------------------
 * A4/A5: ordered merge loop (universe_old.py:196-210) + Planet.absorb/destroy
 * (planet.py:74-91), Python numeric semantics.
 * ------------------------------------------------------------------------------------- */
static inline u128 mass_int(const oracle_state *s, int64_t k) {
    return (((u128)s->mass_hi[k]) << 64) | (u128)s->mass_lo[k];
}

/* Python's exact `a >= b` between int and float operands */
static int int_ge_float(u128 a, double f) {          /* a >= f */
    if (isnan(f)) return 0;
    if (f < 0.0) return 1;
    if (f >= 3.402823669209384634633746074317682e38) return 0;  /* 2^128 */
    double fl = floor(f);
    u128 fi = (u128)fl;                                /* exact: fl < 2^128, 53-bit mantissa */
    if (a != fi) return a > fi;
    return !(f > fl);                                  /* a == floor(f): a >= f iff no fraction */
}

static int float_ge_int(double f, u128 b) {          /* f >= b */
    if (isnan(f)) return 0;
    if (f < 0.0) return 0;
    if (f >= 3.402823669209384634633746074317682e38) return 1;
    double fl = floor(f);
    u128 fi = (u128)fl;
    if (fi != b) return fi > b;
    return 1;                                          /* floor(f) == b  =>  f >= b */
}

static int mass_ge(const oracle_state *s, int64_t a, int64_t b) {  /* planet.mass >= other.mass */
    if (s->mass_is_int[a] && s->mass_is_int[b]) return mass_int(s, a) >= mass_int(s, b);
    if (s->mass_is_int[a]) return int_ge_float(mass_int(s, a), s->mass_f64[b]);
    if (s->mass_is_int[b]) return float_ge_int(s->mass_f64[a], mass_int(s, b));
    return s->mass_f64[a] >= s->mass_f64[b];
}

/* returns 1 if a merge happened (both alive) */
static int absorb(oracle_state *s, int64_t a /* self */, int64_t o /* planet */) {
    if (!(s->alive[a] && s->alive[o])) return 0;                       /* planet.py:75 */
    double ma = s->mass_f64[a], mo = s->mass_f64[o];
    double msum;
    int both_int = s->mass_is_int[a] && s->mass_is_int[o];
    u128 isum = 0;
    if (both_int) { isum = mass_int(s, a) + mass_int(s, o); msum = u128_to_double(isum); }
    else          { msum = ma + mo; }                                  /* float(int) + float */
    if (!s->immobile[a]) {                                             /* planet.py:76-79 */
        for (int c = 0; c < 2; ++c) {
            double t1 = ma * s->pos[2 * a + c];
            double t2 = mo * s->pos[2 * o + c];
            double num = t1 + t2;
            s->pos[2 * a + c] = num / msum;
            double u1 = ma * s->vel[2 * a + c];
            double u2 = mo * s->vel[2 * o + c];
            double nv = u1 + u2;
            s->vel[2 * a + c] = nv / msum;
        }
    }
    if (both_int) {                                                    /* planet.py:80 */
        s->mass_hi[a] = (uint64_t)(isum >> 64);
        s->mass_lo[a] = (uint64_t)isum;
    } else {
        s->mass_is_int[a] = 0;
    }
    s->mass_f64[a] = msum;
    s->volume[a] = s->volume[a] + s->volume[o];                        /* planet.py:81 */
    {
        double v3 = 3.0 * s->volume[a];
        double den = 4.0 * M_PI;
        s->radius[a] = pow(v3 / den, 1.0 / 3.0);                       /* planet.py:82 */
    }
    s->alive[o] = 0;                                                   /* planet.py:86-91 */
    s->vel[2 * o] = 0.0;
    s->vel[2 * o + 1] = 0.0;
    s->eaten[a] += 1 + s->eaten[o];                                    /* planet.py:84 */
    return 1;
}

/*
 * One frame of Universe.simulate() minus UI (universe_old.py:345-350):
 *   update_planet_lists -> interact -> destroy_escaped_planets (if bounded).
 * Outputs (all optional, sized for the active count na <= n, returned):
 *   active_idx [na]      insertion index of each active planet
 *   acc, p_new, v_new    [na][2] per-step tensors in active order
 *   flag_pairs [cap][2]  flagged ordered pairs (active indices), *n_flags = total count
 *   events     [cap][2]  merge stream (absorber, absorbed) as INSERTION indices, in
 *                        visitation order; *n_events = count
 */
int64_t oracle_frame(oracle_state *s, double G, double dt, int collisions, int bounded, double r_max,
                     int64_t *active_idx_out, double *acc_out, double *pnew_out, double *vnew_out,
                     int64_t *flag_pairs_out, int64_t *n_flags_out, int64_t *events_out,
                     int64_t *n_events_out, int64_t cap) {
    int64_t n = s->n, na = 0;
    int64_t *act = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    uint8_t *escaped = (uint8_t *)calloc((size_t)n, 1);
    for (int64_t k = 0; k < n; ++k)
        if (s->alive[k]) {
            act[na++] = k;
            double x = s->pos[2 * k], y = s->pos[2 * k + 1];
            double x2 = x * x, y2 = y * y;
            escaped[k] = sqrt(x2 + y2) > r_max;                        /* universe_old.py:75-76 */
        }
    if (n_flags_out) *n_flags_out = 0;
    if (n_events_out) *n_events_out = 0;
    if (na == 0) { free(act); free(escaped); return 0; }

    double *m = (double *)malloc((size_t)na * sizeof(double));
    double *rad = (double *)malloc((size_t)na * sizeof(double));
    double *p0 = (double *)malloc((size_t)na * 2 * sizeof(double));
    double *v0 = (double *)malloc((size_t)na * 2 * sizeof(double));
    double *a = (double *)malloc((size_t)na * 2 * sizeof(double));
    double *p = (double *)malloc((size_t)na * 2 * sizeof(double));
    double *v = (double *)malloc((size_t)na * 2 * sizeof(double));
    for (int64_t i = 0; i < na; ++i) {                                  /* universe_old.py:170-172,192 */
        int64_t k = act[i];
        m[i] = s->mass_f64[k];
        rad[i] = s->radius[k];
        p0[2 * i] = s->pos[2 * k]; p0[2 * i + 1] = s->pos[2 * k + 1];
        v0[2 * i] = s->vel[2 * k]; v0[2 * i + 1] = s->vel[2 * k + 1];
    }
    oracle_accel(na, p0, m, G, a);
    oracle_integrate(na, p0, v0, a, dt, p, v);

    int64_t nflags = 0;
    int64_t *pairs = NULL;
    if (collisions) {
        int64_t pcap = 1024;
        pairs = (int64_t *)malloc((size_t)pcap * 2 * sizeof(int64_t));
        nflags = oracle_flags(na, p, rad, pairs, pcap);
        if (nflags > pcap) {
            pcap = nflags;
            pairs = (int64_t *)realloc(pairs, (size_t)pcap * 2 * sizeof(int64_t));
            nflags = oracle_flags(na, p, rad, pairs, pcap);
        }
    }

    int64_t nev = 0, cursor = 0;
    for (int64_t i = 0; i < na; ++i) {                                  /* universe_old.py:196-210 */
        int64_t P = act[i];
        if (s->alive[P]) {
            if (!s->immobile[P]) {
                s->vel[2 * P] = v[2 * i]; s->vel[2 * P + 1] = v[2 * i + 1];
                s->pos[2 * P] = p[2 * i]; s->pos[2 * P + 1] = p[2 * i + 1];
            }
            while (cursor < nflags && pairs[2 * cursor] < i) ++cursor;
            while (cursor < nflags && pairs[2 * cursor] == i) {
                int64_t O = act[pairs[2 * cursor + 1]];
                int64_t win, lose;
                if (mass_ge(s, P, O)) { win = P; lose = O; } else { win = O; lose = P; }
                if (absorb(s, win, lose)) {
                    if (events_out && nev < cap) { events_out[2 * nev] = win; events_out[2 * nev + 1] = lose; }
                    ++nev;
                }
                ++cursor;
            }
        }
    }
    if (bounded)                                                        /* universe_old.py:349-350,78-80 */
        for (int64_t i = 0; i < na; ++i) {
            int64_t k = act[i];
            if (escaped[k]) { s->alive[k] = 0; s->vel[2 * k] = 0.0; s->vel[2 * k + 1] = 0.0; }
        }

    if (active_idx_out) memcpy(active_idx_out, act, (size_t)na * sizeof(int64_t));
    if (acc_out) memcpy(acc_out, a, (size_t)na * 2 * sizeof(double));
    if (pnew_out) memcpy(pnew_out, p, (size_t)na * 2 * sizeof(double));
    if (vnew_out) memcpy(vnew_out, v, (size_t)na * 2 * sizeof(double));
    if (flag_pairs_out) {
        int64_t c = nflags < cap ? nflags : cap;
        if (c > 0) memcpy(flag_pairs_out, pairs, (size_t)c * 2 * sizeof(int64_t));
    }
    if (n_flags_out) *n_flags_out = nflags;
    if (n_events_out) *n_events_out = nev;

    free(pairs); free(m); free(rad); free(p0); free(v0); free(a); free(p); free(v);
    free(act); free(escaped);
    return na;
}
