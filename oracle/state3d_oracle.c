/*
 * state3d_oracle.c -- CPU restatement of the reference's NEW-API step State.interact()
 * (cosmosim/core/universe.py:94-129) with Object.absorb/destroy (universe.py:39-54).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this library; the product (cosmosim_b200/)
 * never imports, links or calls it.
 *
 * Parity status: PINNED against the reference run in the build container
 * (oracle/gen_golden3d.py -> tests/golden/state3d_*.npz; tests/test_state3d_oracle.py also
 * re-checks against the live reference when /root/reference is present).  The reference's
 * State.interact passes force_all_finite= to sklearn (universe.py:107), which sklearn >= 1.8
 * rejects; the harness drops that keyword and nothing else (oracle/ref_harness3d.py).
 *
 * What is restated (reference file:line):
 *   oracle3_accel_rows   cosmosim/util/blas.py:4-41 with 3-D input (call site universe.py:101):
 *                        the chain of cosmo_oracle.c with a third dimension,
 *                        c = fl(fl(fl(R_x + R_y) + R_z) + 6)  (np.sum over the 3 rows, blas.py:23)
 *   oracle3_flags_rows   universe.py:107-108 + sklearn euclidean_distances on (N,3) float64:
 *                          XX   = fl(fl(fl(x^2) + fl(z^2)) + fl(y^2))     (einsum, 3 columns)
 *                          dot  = fma(z,z', fma(y,y', fl(x x')))          (OpenBLAS dgemm, K = 3)
 *                          d2   = fl(fl(-2 dot + XX_i) + XX_j); d = sqrt(max(d2,0)), diag 0
 *                          flag = d <= fl(r_i + r_j)                      (no clip to 1 here)
 *                        both pinned bitwise on whole matrices (same OpenBLAS tail-kernel caveat
 *                        as cosmo_oracle.c: N mod 8 >= 4 touches a few tail rows/columns)
 *   oracle3_frame        universe.py:94-129: radii from get_radius() (universe.py:33-37), the
 *                        integrator :103-104, the ordered absorb loop :111-125 (the row object
 *                        absorbs whenever `obj.mass >= other.mass`, with NO check that the other
 *                        still exists -- a destroyed object has mass 0.0 and position 0, so it is
 *                        absorbed again, which re-rounds the absorber's position/velocity/density
 *                        and turns an int mass into a float), removal of the absorbed objects
 *                        :126-128.
 * Python numeric semantics are restated exactly: masses and densities are ints (arbitrary
 * precision in Python; here up to 128 bits, larger values set an error) or floats; int (+,*) int
 * is exact, int / int is the correctly rounded true division, mixed operations convert the int
 * with float() first; numpy converts a Python int operand with float() as well.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------- */
/* Python numbers                                                                          */
/* ------------------------------------------------------------------------------------- */
typedef struct { int is_int; u128 i; double f; } pynum;

static int g_overflow;   /* sticky: an int result exceeded 128 bits */

static double u128_to_double3(u128 v) {
    if (v == 0) return 0.0;
    int top = 127;
    while (!((v >> top) & 1)) --top;
    if (top <= 52) return (double)(uint64_t)v;
    int shift = top - 52;
    uint64_t mant = (uint64_t)(v >> shift);
    u128 rem = v & ((((u128)1) << shift) - 1);
    u128 half = ((u128)1) << (shift - 1);
    if (rem > half || (rem == half && (mant & 1))) ++mant;
    return ldexp((double)mant, shift);
}

static pynum pn_int(u128 v) { pynum r; r.is_int = 1; r.i = v; r.f = 0.0; return r; }
static pynum pn_float(double v) { pynum r; r.is_int = 0; r.i = 0; r.f = v; return r; }
static double pn_to_double(pynum a) { return a.is_int ? u128_to_double3(a.i) : a.f; }

static pynum pn_add(pynum a, pynum b) {
    if (a.is_int && b.is_int) {
        u128 s = a.i + b.i;
        if (s < a.i) g_overflow = 1;
        return pn_int(s);
    }
    return pn_float(pn_to_double(a) + pn_to_double(b));
}

static int bitlen(u128 v) { int n = 0; while (v) { ++n; v >>= 1; } return n; }

static pynum pn_mul(pynum a, pynum b) {
    if (a.is_int && b.is_int) {
        if (a.i && b.i && bitlen(a.i) + bitlen(b.i) > 128) g_overflow = 1;
        return pn_int(a.i * b.i);
    }
    return pn_float(pn_to_double(a) * pn_to_double(b));
}

/* correctly rounded a / b for ints (CPython long_true_divide), b > 0 */
static double u128_truediv(u128 a, u128 b) {
    if (a == 0) return 0.0;
    /* long division producing the quotient's leading 55+ bits and a sticky bit:
       scale so that q = floor(a * 2^s / b) has between 55 and 56 significant bits */
    int s = 55 - (bitlen(a) - bitlen(b));          /* q has 55 or 56 bits */
    /* shift-subtract division of a*2^s (s may be negative) with a 129-bit safe remainder */
    u128 q = 0, rem = 0;
    int total_bits = bitlen(a) + (s > 0 ? s : 0);
    int drop = s < 0 ? -s : 0;                      /* low bits of a that only feed the sticky */
    int sticky = 0;
    for (int k = total_bits - 1; k >= drop; --k) {
        int bit;
        if (s > 0) bit = (k >= s) ? (int)((a >> (k - s)) & 1) : 0;
        else bit = (int)((a >> k) & 1);
        int carry = (int)(rem >> 127);
        rem = (rem << 1) | (u128)bit;
        q <<= 1;
        if (carry || rem >= b) { rem -= b; q |= 1; }
    }
    if (rem) sticky = 1;
    if (drop) { if (a & ((((u128)1) << drop) - 1)) sticky = 1; }
    /* q = floor(a * 2^s / b) exactly (for s<0: floor((a >> drop) / b), remainder folded in sticky) */
    int qb = bitlen(q);                             /* 55 or 56 */
    int sh = qb - 53;
    uint64_t mant = (uint64_t)(q >> sh);
    u128 low = q & ((((u128)1) << sh) - 1);
    u128 half = ((u128)1) << (sh - 1);
    if (low > half || (low == half && (sticky || (mant & 1)))) ++mant;
    return ldexp((double)mant, sh - s);
}

static pynum pn_truediv(pynum a, pynum b) {
    if (a.is_int && b.is_int) return pn_float(u128_truediv(a.i, b.i));
    return pn_float(pn_to_double(a) / pn_to_double(b));
}

static u128 floor_u128(double f, int *frac) {
    double fl = floor(f);
    *frac = (fl != f);
    /* fl < 2^128 guaranteed by the callers */
    int e;
    double m = frexp(fl, &e);                     /* fl = m * 2^e, 0.5 <= m < 1 */
    if (fl == 0.0) return 0;
    uint64_t mant = (uint64_t)ldexp(m, 53);       /* 53-bit integer */
    int sh = e - 53;
    return sh >= 0 ? ((u128)mant) << sh : (u128)(mant >> (-sh));
}

/* Python's exact a >= b */
static int pn_ge(pynum a, pynum b) {
    const double two128 = 3.402823669209384634633746074317682e38;
    if (a.is_int && b.is_int) return a.i >= b.i;
    if (!a.is_int && !b.is_int) return a.f >= b.f;
    int frac;
    if (a.is_int) {
        if (b.f != b.f) return 0;
        if (b.f < 0.0) return 1;
        if (b.f >= two128) return 0;
        u128 fi = floor_u128(b.f, &frac);
        if (a.i != fi) return a.i > fi;
        return !frac;
    }
    if (a.f != a.f) return 0;
    if (a.f < 0.0) return 0;
    if (a.f >= two128) return 1;
    u128 fi = floor_u128(a.f, &frac);
    if (fi != b.i) return fi > b.i;
    return 1;
}

/* exported for the tests that pin the helpers against Python itself */
double oracle3_truediv(uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo) {
    return u128_truediv(((u128)ahi << 64) | alo, ((u128)bhi << 64) | blo);
}

/* ------------------------------------------------------------------------------------- */
/* A1 in three dimensions                                                                  */
/* ------------------------------------------------------------------------------------- */
static inline double chain3_c(double qx, double qy, double qz, double px, double py, double pz) {
    double rx = qx * px; rx = rx - 2.0;
    double ry = qy * py; ry = ry - 2.0;
    double rz = qz * pz; rz = rz - 2.0;
    double c = rx + ry;
    c = c + rz;
    return c + 6.0;
}

static void accel3_rows_impl(int64_t n, const double *pos, const double *mass, double G,
                             int64_t i_begin, int64_t i_end, double *out, const int compensated) {
    double *h = (double *)malloc((size_t)n * sizeof(double));
    double *mh = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k) {
        double x = pos[3 * k], y = pos[3 * k + 1], z = pos[3 * k + 2];
        h[k] = 0.5 * chain3_c(-2.0 * x, -2.0 * y, -2.0 * z, x, y, z);
        mh[k] = 0.5 * mass[k];
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = i_begin; i < i_end; ++i) {
        const double pi[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        const double qx = -2.0 * pi[0], qy = -2.0 * pi[1], qz = -2.0 * pi[2];
        const double hi_i = h[i];
        double s[3] = {0, 0, 0}, comp[3] = {0, 0, 0};
        for (int64_t j = 0; j < n; ++j) {
            const double *pj = pos + 3 * j;
            double c = chain3_c(qx, qy, qz, pj[0], pj[1], pj[2]);
            double d2;
            if (j >= i) { d2 = c - h[j]; d2 = d2 - hi_i; }
            else        { d2 = c - hi_i; d2 = d2 - h[j]; }
            double t[3];
            for (int k = 0; k < 3; ++k) { double df = pj[k] - pi[k]; t[k] = 2.0 * df; }
            if (d2 != 0.0) {
                double w = d2 * sqrt(d2);
                for (int k = 0; k < 3; ++k) t[k] = t[k] / w;
            }
            for (int k = 0; k < 3; ++k) {
                double term = t[k] * mh[j];
                if (compensated) {
                    double ns = s[k] + term;
                    comp[k] += (fabs(s[k]) >= fabs(term)) ? ((s[k] - ns) + term) : ((term - ns) + s[k]);
                    s[k] = ns;
                } else {
                    s[k] += term;
                }
            }
        }
        for (int k = 0; k < 3; ++k) {
            double v = s[k] + comp[k];
            out[3 * (i - i_begin) + k] = G * v;
        }
    }
    free(h);
    free(mh);
}

void oracle3_accel_rows(int64_t n, const double *pos, const double *mass, double G, int64_t i_begin,
                        int64_t i_end, double *out) {
    accel3_rows_impl(n, pos, mass, G, i_begin, i_end, out, 1);
}

void oracle3_accel_rows_plain(int64_t n, const double *pos, const double *mass, double G, int64_t i_begin,
                              int64_t i_end, double *out) {
    accel3_rows_impl(n, pos, mass, G, i_begin, i_end, out, 0);
}

void oracle3_d2_rows(int64_t n, const double *pos, int64_t i_begin, int64_t i_end, double *d2_out) {
    double *h = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k) {
        double x = pos[3 * k], y = pos[3 * k + 1], z = pos[3 * k + 2];
        h[k] = 0.5 * chain3_c(-2.0 * x, -2.0 * y, -2.0 * z, x, y, z);
    }
    for (int64_t i = i_begin; i < i_end; ++i)
        for (int64_t j = 0; j < n; ++j) {
            double c = chain3_c(-2.0 * pos[3 * i], -2.0 * pos[3 * i + 1], -2.0 * pos[3 * i + 2], pos[3 * j],
                                pos[3 * j + 1], pos[3 * j + 2]);
            double d2;
            if (j >= i) { d2 = c - h[j]; d2 = d2 - h[i]; }
            else        { d2 = c - h[i]; d2 = d2 - h[j]; }
            d2_out[(i - i_begin) * n + j] = d2;
        }
    free(h);
}

/* universe.py:103-104 */
void oracle3_integrate(int64_t n, const double *p0, const double *v0, const double *a, double dt, double *p,
                       double *v) {
    for (int64_t k = 0; k < 3 * n; ++k) {
        double adt = a[k] * dt;
        double vn = v0[k] + adt;
        double vdt = vn * dt;
        v[k] = vn;
        p[k] = p0[k] + vdt;
    }
}

/* ------------------------------------------------------------------------------------- */
/* collision predicate (universe.py:107-108)                                               */
/* ------------------------------------------------------------------------------------- */
static inline double row_norm3(const double *p) {
    double x2 = p[0] * p[0], y2 = p[1] * p[1], z2 = p[2] * p[2];
    double s = x2 + z2;
    return s + y2;
}

static inline double dist3(const double *pi, double xxi, const double *pj, double xxj) {
    double dot = fma(pi[2], pj[2], fma(pi[1], pj[1], pi[0] * pj[0]));
    double d2 = -2.0 * dot + xxi;
    d2 = d2 + xxj;
    if (d2 < 0.0) d2 = 0.0;
    return sqrt(d2);
}

void oracle3_dist_rows(int64_t n, const double *p, int64_t i_begin, int64_t i_end, double *d_out) {
    for (int64_t i = i_begin; i < i_end; ++i) {
        double xxi = row_norm3(p + 3 * i);
        for (int64_t j = 0; j < n; ++j)
            d_out[(i - i_begin) * n + j] = (i == j) ? 0.0 : dist3(p + 3 * i, xxi, p + 3 * j, row_norm3(p + 3 * j));
    }
}

int64_t oracle3_flags_rows(int64_t n, const double *p, const double *radius, int64_t i_begin, int64_t i_end,
                           int64_t *pairs, int64_t cap) {
    double *xx = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k) xx[k] = row_norm3(p + 3 * k);
    int64_t rows = i_end - i_begin;
    int64_t *row_cnt = (int64_t *)calloc((size_t)rows + 1, sizeof(int64_t));
#pragma omp parallel for schedule(static)
    for (int64_t i = i_begin; i < i_end; ++i) {
        int64_t c = 0;
        for (int64_t j = 0; j < n; ++j)
            if (j != i && dist3(p + 3 * i, xx[i], p + 3 * j, xx[j]) <= radius[i] + radius[j]) ++c;
        row_cnt[i - i_begin] = c;
    }
    int64_t total = 0;
    for (int64_t r = 0; r < rows; ++r) {
        if (!row_cnt[r]) continue;
        int64_t i = i_begin + r;
        for (int64_t j = 0; j < n; ++j)
            if (j != i && dist3(p + 3 * i, xx[i], p + 3 * j, xx[j]) <= radius[i] + radius[j]) {
                if (total < cap) { pairs[2 * total] = i; pairs[2 * total + 1] = j; }
                ++total;
            }
    }
    free(row_cnt);
    free(xx);
    return total;
}

/* ------------------------------------------------------------------------------------- */
/* the frame                                                                               */
/* ------------------------------------------------------------------------------------- */
typedef struct {
    int64_t n;            /* objects in State.objects (all exist at the start of a step) */
    double *pos;          /* [n][3] */
    double *vel;          /* [n][3] */
    double *mass_f64;     /* [n] float(mass) */
    uint64_t *mass_hi;    /* [n] integer mass when mass_is_int */
    uint64_t *mass_lo;
    uint8_t *mass_is_int;
    double *dens_f64;     /* [n] float(density); an exact integer < 2^53 when dens_is_int */
    uint8_t *dens_is_int;
    int64_t *id;          /* [n] insertion index */
} oracle3_state;

static pynum load_mass(const oracle3_state *s, int64_t k) {
    if (s->mass_is_int[k]) return pn_int(((u128)s->mass_hi[k] << 64) | s->mass_lo[k]);
    return pn_float(s->mass_f64[k]);
}
static pynum load_dens(const oracle3_state *s, int64_t k) {
    if (s->dens_is_int[k]) return pn_int((u128)(uint64_t)s->dens_f64[k]);
    return pn_float(s->dens_f64[k]);
}

/* Object.get_radius(): ((3*(mass/density))/(4*math.pi))**(1/3)  (universe.py:33-37) */
static double radius_of(pynum mass, pynum dens) {
    double vol = pn_to_double(pn_truediv(mass, dens));
    double x = 3.0 * vol;
    x = x / (4.0 * 3.141592653589793);
    return pow(x, 1.0 / 3.0);
}

void oracle3_radii(const oracle3_state *s, double *radius_out) {
    for (int64_t k = 0; k < s->n; ++k) radius_out[k] = radius_of(load_mass(s, k), load_dens(s, k));
}

/*
 * One State.interact(collisions) (universe.py:94-129).  Mutates *s in place (absorbed objects are
 * removed and the arrays compacted, s->n shrinks).  Outputs (may be NULL): acc_out/p_new_out/
 * v_new_out [n0][3], radius_out [n0] (pre-merge radii), pairs_out [pair_cap][2] flagged ordered
 * pairs as slot indices, n_pairs_out; events_out [event_cap][4] = (absorber id, absorbed id,
 * absorbed-still-existed, 0) in call order of Object.absorb.  Returns the number of absorb calls,
 * or -1 if an integer exceeded 128 bits.
 */
int64_t oracle3_frame(oracle3_state *s, double dt, double G, int collisions, double *acc_out,
                      double *p_new_out, double *v_new_out, double *radius_out, int64_t *pairs_out,
                      int64_t pair_cap, int64_t *n_pairs_out, int64_t *events_out, int64_t event_cap) {
    const int64_t n = s->n;
    g_overflow = 0;
    double *acc = (double *)malloc((size_t)n * 3 * sizeof(double));
    double *p = (double *)malloc((size_t)n * 3 * sizeof(double));
    double *v = (double *)malloc((size_t)n * 3 * sizeof(double));
    double *r = (double *)malloc((size_t)n * sizeof(double));
    oracle3_radii(s, r);
    oracle3_accel_rows(n, s->pos, s->mass_f64, G, 0, n, acc);
    oracle3_integrate(n, s->pos, s->vel, acc, dt, p, v);
    if (acc_out) memcpy(acc_out, acc, (size_t)n * 3 * sizeof(double));
    if (p_new_out) memcpy(p_new_out, p, (size_t)n * 3 * sizeof(double));
    if (v_new_out) memcpy(v_new_out, v, (size_t)n * 3 * sizeof(double));
    if (radius_out) memcpy(radius_out, r, (size_t)n * sizeof(double));

    int64_t n_pairs = 0, cap = 0;
    int64_t *pairs = NULL;
    if (collisions) {
        cap = 4096;
        pairs = (int64_t *)malloc((size_t)cap * 2 * sizeof(int64_t));
        n_pairs = oracle3_flags_rows(n, p, r, 0, n, pairs, cap);
        if (n_pairs > cap) {
            cap = n_pairs;
            pairs = (int64_t *)realloc(pairs, (size_t)cap * 2 * sizeof(int64_t));
            n_pairs = oracle3_flags_rows(n, p, r, 0, n, pairs, cap);
        }
    }
    if (n_pairs_out) *n_pairs_out = n_pairs;
    if (pairs_out)
        for (int64_t k = 0; k < n_pairs && k < pair_cap; ++k) {
            pairs_out[2 * k] = pairs[2 * k];
            pairs_out[2 * k + 1] = pairs[2 * k + 1];
        }

    /* live Python-side state of every object during the loop */
    pynum *mass = (pynum *)malloc((size_t)n * sizeof(pynum));
    pynum *dens = (pynum *)malloc((size_t)n * sizeof(pynum));
    uint8_t *exists = (uint8_t *)malloc((size_t)n);
    for (int64_t k = 0; k < n; ++k) { mass[k] = load_mass(s, k); dens[k] = load_dens(s, k); exists[k] = 1; }
    double *cp = s->pos, *cv = s->vel;     /* objects keep their OLD rows until visited */

    int64_t n_events = 0, q = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!exists[i]) {
            while (q < n_pairs && pairs[2 * q] == i) ++q;
            continue;
        }
        for (int k = 0; k < 3; ++k) { cv[3 * i + k] = v[3 * i + k]; cp[3 * i + k] = p[3 * i + k]; }
        for (; q < n_pairs && pairs[2 * q] == i; ++q) {
            int64_t j = pairs[2 * q + 1];
            if (!pn_ge(mass[i], mass[j])) continue;
            /* Object.absorb (universe.py:39-46) */
            if (events_out && n_events < event_cap) {
                events_out[4 * n_events] = s->id[i];
                events_out[4 * n_events + 1] = s->id[j];
                events_out[4 * n_events + 2] = exists[j];
                events_out[4 * n_events + 3] = 0;
            }
            ++n_events;
            pynum m_total = pn_add(mass[i], mass[j]);
            double mi = pn_to_double(mass[i]), mj = pn_to_double(mass[j]), mt = pn_to_double(m_total);
            for (int k = 0; k < 3; ++k) {
                double a = mi * cp[3 * i + k], b = mj * cp[3 * j + k];
                double sum = a + b;
                cp[3 * i + k] = sum / mt;
                a = mi * cv[3 * i + k]; b = mj * cv[3 * j + k];
                sum = a + b;
                cv[3 * i + k] = sum / mt;
            }
            dens[i] = pn_truediv(pn_add(pn_mul(mass[i], dens[i]), pn_mul(mass[j], dens[j])), m_total);
            mass[i] = pn_add(mass[i], mass[j]);
            /* obj.destroy() (universe.py:48-54) */
            exists[j] = 0;
            mass[j] = pn_float(0.0);
            for (int k = 0; k < 3; ++k) { cp[3 * j + k] = 0.0; cv[3 * j + k] = 0.0; }
        }
    }

    /* remove the absorbed objects (universe.py:126-128), keep list order */
    int64_t w = 0;
    for (int64_t k = 0; k < n; ++k) {
        if (!exists[k]) continue;
        for (int d = 0; d < 3; ++d) { s->pos[3 * w + d] = cp[3 * k + d]; s->vel[3 * w + d] = cv[3 * k + d]; }
        s->mass_is_int[w] = (uint8_t)mass[k].is_int;
        s->mass_hi[w] = mass[k].is_int ? (uint64_t)(mass[k].i >> 64) : 0;
        s->mass_lo[w] = mass[k].is_int ? (uint64_t)mass[k].i : 0;
        s->mass_f64[w] = pn_to_double(mass[k]);
        s->dens_is_int[w] = (uint8_t)dens[k].is_int;
        s->dens_f64[w] = pn_to_double(dens[k]);
        s->id[w] = s->id[k];
        ++w;
    }
    s->n = w;
    free(acc); free(p); free(v); free(r); free(pairs); free(mass); free(dens); free(exists);
    return g_overflow ? -1 : n_events;
}
