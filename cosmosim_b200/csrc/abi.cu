// extern "C" surface of libcosmosim_b200.so (see include/cosmosim_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "cosmo_device.cuh"
#include "cosmo_internal.h"

namespace cosmo {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return COSMO_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return COSMO_E_CUDA;
}

static long long g_launches = 0;

int check_launch(const char *what) {
    ++g_launches;
    return check_cuda(cudaGetLastError(), what);
}

static int check_args(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p) {
    if (!st || !sc || !p) {
        set_error("null state/scratch/params");
        return COSMO_E_ARG;
    }
    if (st->cap <= 0 || st->cap % COSMO_PAD != 0) {
        set_error("cap=%d must be a positive multiple of %d", st->cap, COSMO_PAD);
        return COSMO_E_ARG;
    }
    if (p->n_upper < 0 || p->n_upper > st->cap || p->i_begin < 0 || p->i_end > p->n_upper ||
        p->i_begin > p->i_end) {
        set_error("bad ranges: n_upper=%d cap=%d rows=[%d,%d)", p->n_upper, st->cap, p->i_begin, p->i_end);
        return COSMO_E_ARG;
    }
    if (p->jsplit > sc->jsplit_max) {
        set_error("jsplit=%d exceeds jsplit_max=%d", p->jsplit, sc->jsplit_max);
        return COSMO_E_ARG;
    }
    return COSMO_OK;
}

// cost model for the source-range split: waves x tiles per CTA
static int suggest_jsplit(int rows, int n, int rows_per_cta, int tile, int slots, int jmax) {
    if (rows <= 0 || n <= 0) return 1;
    const long long iblocks = (rows + rows_per_cta - 1) / rows_per_cta;
    const long long tiles = (n + tile - 1) / tile;
    long long best_cost = -1;
    int best = 1;
    for (int js = 1; js <= jmax && js <= tiles; ++js) {
        const long long waves = (iblocks * js + slots - 1) / slots;
        const long long per = (tiles + js - 1) / js;
        const long long cost = waves * per * 1024 + js;  // mild preference for fewer splits
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = js;
        }
    }
    return best;
}

}  // namespace cosmo

using namespace cosmo;

extern "C" {

int cosmo_abi_version(void) { return COSMO_ABI_VERSION; }

long long cosmo_launch_count(void) { return g_launches; }

const char *cosmo_last_error(void) { return g_err; }

int cosmo_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    COSMO_TRY(check_cuda(cudaGetDevice(&dev), "cudaGetDevice"));
    cudaDeviceProp prop;
    COSMO_TRY(check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties"));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (prop.major != 10) {
        set_error("device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
        return COSMO_E_NO_DEVICE;
    }
    return COSMO_OK;
}

int cosmo_suggest_jsplit(int32_t rows, int32_t n, int kernel, int32_t jsplit_max) {
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    }
    if (jsplit_max < 1) jsplit_max = 1;
    int rows_per_cta = 256, tile = 256, resident = 0;
    if (kernel == 3 || kernel == 5) resident = detect_residency(&rows_per_cta, &tile);
    else resident = accel_residency(kernel < 0 || kernel > 4 ? 0 : kernel, &rows_per_cta, &tile);
    if (kernel == 4 || kernel == 5) {  // small-N shapes: 64 rows x 64-source tiles per CTA
        rows_per_cta = 64;
        tile = 64;
        resident = 16;
    }
    if (resident <= 0) resident = kernel == 1 || kernel == 2 ? 3 : 8;  // no device: documented defaults
    return suggest_jsplit(rows, n, rows_per_cta, tile, sms * resident, jsplit_max);
}

int cosmo_begin_frame(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
                      void *stream) {
    COSMO_TRY(check_args(st, sc, p));
    return launch_prep(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo_accel_integrate(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
                          void *stream) {
    COSMO_TRY(check_args(st, sc, p));
    return launch_accel(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo_energies(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p, void *stream) {
    COSMO_TRY(check_args(st, sc, p));
    return launch_energies(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo_integrate(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p, void *stream) {
    COSMO_TRY(check_args(st, sc, p));
    return launch_sym_integrate(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo_detect_pairs(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
                       void *stream) {
    COSMO_TRY(check_args(st, sc, p));
    return launch_detect(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo_resolve_commit(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
                         void *stream) {
    COSMO_TRY(check_args(st, sc, p));
    return launch_resolve_commit(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo_step(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p, void *stream) {
    COSMO_TRY(check_args(st, sc, p));
    cudaStream_t s = (cudaStream_t)stream;
    COSMO_TRY(launch_prep(*st, *sc, *p, s));
    COSMO_TRY(launch_accel(*st, *sc, *p, s));
    if (p->opts & COSMO_OPT_ENERGY) COSMO_TRY(launch_energies(*st, *sc, *p, s));
    COSMO_TRY(launch_detect(*st, *sc, *p, s));
    return launch_resolve_commit(*st, *sc, *p, s);
}

int cosmo_append_pairs(const cosmo_state *st, const cosmo_scratch *sc, const uint64_t *src,
                       const int64_t *count_ptr, int64_t max_count, void *stream) {
    if (!st || !sc || !src || !count_ptr) {
        set_error("cosmo_append_pairs: null argument");
        return COSMO_E_ARG;
    }
    return launch_append_pairs(*st, *sc, src, count_ptr, max_count, (cudaStream_t)stream);
}

int cosmo_seal_pair_row(const cosmo_state *st, int64_t *row, void *stream) {
    if (!st || !st->status || !row) {
        set_error("cosmo_seal_pair_row: null argument");
        return COSMO_E_ARG;
    }
    return launch_seal_pair_row(st->status, row, (cudaStream_t)stream);
}

int cosmo_append_pair_rows(const cosmo_state *st, const cosmo_scratch *sc, const int64_t *rows, int32_t n_rows,
                           int64_t max_count, void *stream) {
    if (!st || !sc || !rows || n_rows <= 0 || max_count <= 0) {
        set_error("cosmo_append_pair_rows: bad argument");
        return COSMO_E_ARG;
    }
    return launch_append_pair_rows(*st, *sc, rows, n_rows, max_count, (cudaStream_t)stream);
}

// ---- function-level seam: acc_blas replacement on raw arrays --------------------------
namespace {
struct AccelWs {
    cosmo_state st;
    cosmo_scratch sc;
    int cap;
    int64_t bytes;
};

constexpr int kAccelJsplitMax = 16;

AccelWs carve(int32_t n, char *base) {
    AccelWs w;
    memset(&w, 0, sizeof(w));
    const int cap = (n + COSMO_PAD - 1) / COSMO_PAD * COSMO_PAD;
    w.cap = cap;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return p;
    };
    w.st.cap = cap;
    w.st.n_total = cap;
    w.st.status = (int64_t *)take(COSMO_STATUS_LEN * 8);
    w.sc.tile_ctr = (int32_t *)take((size_t)(cap / 64 + 4) * 4);
    w.st.pos = (double *)take((size_t)cap * 16);
    w.st.vel = (double *)take((size_t)cap * 16);
    w.st.mass = (double *)take((size_t)cap * 8);
    w.st.radius = (double *)take((size_t)cap * 8);
    w.sc.h = (double *)take((size_t)cap * 8);
    w.sc.rinf = (double *)take((size_t)cap * 8);
    w.sc.pk_x = (float *)take((size_t)cap * 4);
    w.sc.pk_y = (float *)take((size_t)cap * 4);
    w.sc.pk_m = (float *)take((size_t)cap * 4);
    w.sc.pos_new = (double *)take((size_t)cap * 16);
    w.sc.vel_new = (double *)take((size_t)cap * 16);
    w.sc.xx_new = (double *)take((size_t)cap * 8);
    w.sc.acc = (double *)take((size_t)cap * 16);
    w.sc.partial = (double *)take((size_t)cap * 16 * kAccelJsplitMax);
    w.sc.jsplit_max = kAccelJsplitMax;
    w.sc.alive = (uint8_t *)take((size_t)cap);
    w.sc.escaped = (uint8_t *)take((size_t)cap);
    w.bytes = (int64_t)off;
    return w;
}
}  // namespace

int64_t cosmo_accel_workspace_bytes(int32_t n) {
    if (n <= 0) return 0;
    return carve(n, nullptr).bytes;
}

int cosmo_accel(const double *pos, const double *mass, int32_t n, double G, int fp32, int32_t pos_exp,
                int32_t mass_exp, double *acc_out, void *workspace, void *stream) {
    if (!pos || !mass || !acc_out || !workspace || n <= 0) {
        set_error("cosmo_accel: bad argument");
        return COSMO_E_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    AccelWs w = carve(n, (char *)workspace);
    const int64_t status0[COSMO_STATUS_LEN] = {n};
    // status, padded inputs
    COSMO_TRY(check_cuda(cudaMemcpyAsync(w.st.status, status0, sizeof(status0), cudaMemcpyHostToDevice, s),
                         "status upload"));
    COSMO_TRY(check_cuda(cudaMemcpyAsync(w.st.pos, pos, (size_t)n * 16, cudaMemcpyDeviceToDevice, s), "pos copy"));
    COSMO_TRY(check_cuda(cudaMemcpyAsync(w.st.mass, mass, (size_t)n * 8, cudaMemcpyDeviceToDevice, s), "mass copy"));
    COSMO_TRY(check_cuda(cudaMemsetAsync(w.st.vel, 0, (size_t)w.cap * 16, s), "vel clear"));
    COSMO_TRY(check_cuda(cudaMemsetAsync(w.st.radius, 0, (size_t)w.cap * 8, s), "radius clear"));
    cosmo_step_params p;
    memset(&p, 0, sizeof(p));
    p.G = G;
    p.dt = 0.0;
    p.r_max = 1e300;
    p.opts = COSMO_OPT_STORE_ACC | (fp32 ? COSMO_OPT_FP32 : 0u);
    p.n_upper = n;
    p.i_begin = 0;
    p.i_end = n;
    p.pos_exp = pos_exp;
    p.mass_exp = mass_exp;
    p.jsplit = cosmo_suggest_jsplit(n, n, fp32 ? 1 : 0, kAccelJsplitMax);
    COSMO_TRY(launch_prep(w.st, w.sc, p, s));
    COSMO_TRY(launch_accel(w.st, w.sc, p, s));
    return check_cuda(cudaMemcpyAsync(acc_out, w.sc.acc, (size_t)n * 16, cudaMemcpyDeviceToDevice, s), "acc copy");
}

int cosmo_accel_host(const double *pos_host, const double *mass_host, int32_t n, double G, int fp32,
                     int32_t pos_exp, int32_t mass_exp, double *acc_out_host, void *dev_workspace,
                     int64_t dev_workspace_bytes, void *stream) {
    if (!pos_host || !mass_host || !acc_out_host || !dev_workspace || n <= 0) {
        set_error("cosmo_accel_host: bad argument");
        return COSMO_E_ARG;
    }
    const int cap = (n + COSMO_PAD - 1) / COSMO_PAD * COSMO_PAD;
    const int64_t need = cosmo_accel_workspace_bytes(n) + (int64_t)cap * 40;
    if (dev_workspace_bytes < need) {
        set_error("cosmo_accel_host: workspace %lld < %lld bytes", (long long)dev_workspace_bytes, (long long)need);
        return COSMO_E_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char *base = (char *)dev_workspace;
    double *d_pos = (double *)base;
    double *d_mass = (double *)(base + (size_t)cap * 16);
    double *d_acc = (double *)(base + (size_t)cap * 24);
    void *ws = base + (size_t)cap * 40;
    COSMO_TRY(check_cuda(cudaMemcpyAsync(d_pos, pos_host, (size_t)n * 16, cudaMemcpyHostToDevice, s), "pos H2D"));
    COSMO_TRY(check_cuda(cudaMemcpyAsync(d_mass, mass_host, (size_t)n * 8, cudaMemcpyHostToDevice, s), "mass H2D"));
    COSMO_TRY(cosmo_accel(d_pos, d_mass, n, G, fp32, pos_exp, mass_exp, d_acc, ws, stream));
    COSMO_TRY(check_cuda(cudaMemcpyAsync(acc_out_host, d_acc, (size_t)n * 16, cudaMemcpyDeviceToHost, s), "acc D2H"));
    return check_cuda(cudaStreamSynchronize(s), "cosmo_accel_host sync");
}

int cosmo_microbench(int kind, int blocks, int threads, int iters, double *out, int64_t *ops_per_thread,
                     void *stream) {
    return launch_microbench(kind, blocks, threads, iters, out, ops_per_thread, (cudaStream_t)stream);
}

int cosmo_sym_tuned_available(void) { return sym_tuned_available(); }

// host-callable copies of the resolver's exact scalar helpers (CPU tests pin them against
// Python's own int/float semantics and libm pow)
double cosmo_host_pow_third(double x) { return pow_third(x); }

double cosmo_host_u128_to_double(uint64_t hi, uint64_t lo) {
    return u128_to_double((((u128)hi) << 64) | lo);
}

int cosmo_host_mass_ge(int a_int, uint64_t a_hi, uint64_t a_lo, double a_f, int b_int, uint64_t b_hi,
                       uint64_t b_lo, double b_f) {
    return mass_ge(a_int != 0, (((u128)a_hi) << 64) | a_lo, a_f, b_int != 0, (((u128)b_hi) << 64) | b_lo, b_f)
               ? 1
               : 0;
}

}  // extern "C"
