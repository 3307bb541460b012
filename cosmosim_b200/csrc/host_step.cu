// Whole frames for callers that are not Python: one device allocation carved into every state and
// scratch array (cosmo_workspace_*), default launch parameters (cosmo_params_default), and the
// host-buffer frame cosmo_step_host -- what the reference's `self.interact()` call
// (cosmosim/core/universe_old.py:347) binds to when the planets live in host memory.
#include <limits.h>
#include <string.h>

#include "cosmo_device.cuh"
#include "cosmo_internal.h"

namespace cosmo {

namespace {

constexpr int kJsplitMax = 128;
constexpr int kGiantCap = 1 << 16;

struct Sizes {
    int cap;
    long long n_total, pair_cap, event_cap, grid_cells, colbuf_rows;
};

Sizes sizes_for(int n_total, uint32_t opts) {
    Sizes z;
    z.n_total = n_total < 1 ? 1 : n_total;
    z.cap = (int)((z.n_total + COSMO_PAD - 1) / COSMO_PAD * COSMO_PAD);
    z.pair_cap = 4LL * z.cap > (1 << 16) ? 4LL * z.cap : (1 << 16);
    z.event_cap = 2 * z.n_total > (1 << 16) ? 2 * z.n_total : (1 << 16);
    z.grid_cells = 8LL * z.cap > (1 << 16) ? 8LL * z.cap : (1 << 16);
    if (z.grid_cells > (long long)COSMO_SCAN_BLOCKS * 4096 - 1) z.grid_cells = (long long)COSMO_SCAN_BLOCKS * 4096 - 1;
    z.colbuf_rows = 0;
    if (opts & COSMO_OPT_F32_SYM) z.colbuf_rows = 2 * ((z.cap / 1024 + 3) / 4 + 1);   // double2 per super row of four row blocks
    if (opts & COSMO_OPT_F64_SYM) z.colbuf_rows = 2LL * (z.cap / 512) > z.colbuf_rows ? 2LL * (z.cap / 512) : z.colbuf_rows;
    return z;
}

// walks the layout; with base == nullptr it only measures
struct Carver {
    char *base;
    size_t off = 0;
    template <typename T>
    T *take(long long count) {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += ((size_t)count * sizeof(T) + 255) / 256 * 256;
        return p;
    }
};

void carve(int n_total, uint32_t opts, char *base, cosmo_state *st, cosmo_scratch *sc, size_t *bytes) {
    const Sizes z = sizes_for(n_total, opts);
    const long long cap = z.cap, nt = z.n_total;
    cosmo_state s;
    cosmo_scratch c;
    memset(&s, 0, sizeof(s));
    memset(&c, 0, sizeof(c));
    Carver w{base};
    s.cap = z.cap; s.n_total = (int32_t)nt; s.event_cap = z.event_cap;
    s.status = w.take<int64_t>(COSMO_STATUS_LEN);
    s.pos = w.take<double>(2 * cap); s.vel = w.take<double>(2 * cap); s.mass = w.take<double>(cap);
    s.mass_hi = w.take<uint64_t>(cap); s.mass_lo = w.take<uint64_t>(cap);
    s.radius = w.take<double>(cap); s.volume = w.take<double>(cap); s.eaten = w.take<int64_t>(cap);
    s.id = w.take<int32_t>(cap); s.flags = w.take<uint8_t>(cap);
    s.energy = w.take<double>(cap); s.totals = w.take<double>(4);
    s.rec_mass = w.take<double>(nt); s.rec_mass_hi = w.take<uint64_t>(nt); s.rec_mass_lo = w.take<uint64_t>(nt);
    s.rec_radius = w.take<double>(nt); s.rec_volume = w.take<double>(nt); s.rec_eaten = w.take<int64_t>(nt);
    s.rec_death = w.take<int32_t>(nt); s.rec_flags = w.take<uint8_t>(nt);
    s.events = w.take<int64_t>(3 * z.event_cap);
    c.h = w.take<double>(cap); c.pk_x = w.take<float>(cap); c.pk_y = w.take<float>(cap); c.pk_m = w.take<float>(cap);
    c.acc = w.take<double>(2 * cap);
    c.pos_new = w.take<double>(2 * cap); c.vel_new = w.take<double>(2 * cap); c.xx_new = w.take<double>(cap);
    c.rinf = w.take<double>(cap);
    c.partial = w.take<double>((long long)kJsplitMax * 2 * cap); c.jsplit_max = kJsplitMax;
    c.tile_ctr = w.take<int32_t>(cap / 64 + 4);
    c.alive = w.take<uint8_t>(cap); c.escaped = w.take<uint8_t>(cap);
    c.pairs = w.take<uint64_t>(z.pair_cap); c.pairs_sorted = w.take<uint64_t>(z.pair_cap); c.pair_cap = z.pair_cap;
    c.cur_pos = w.take<double>(2 * cap); c.cur_vel = w.take<double>(2 * cap);
    c.mod_row = w.take<int32_t>(cap); c.mod_step = w.take<int32_t>(cap);
    c.chunk_alive = w.take<int32_t>(cap / 1024 + 2);
    c.row_count = w.take<int32_t>(cap + 1); c.row_start = w.take<int32_t>(cap + 1);
    c.pmin = w.take<int32_t>(cap); c.pmax = w.take<int32_t>(cap);
    c.scan_blk = w.take<int32_t>(COSMO_SCAN_BLOCKS + 1);
    c.pair_flag = w.take<int32_t>(z.pair_cap); c.cplx_list = w.take<int32_t>(z.pair_cap);
    c.scan_a = w.take<int32_t>(z.pair_cap + 1); c.ev_tmp = w.take<int64_t>(2 * z.pair_cap);
    c.label = w.take<int32_t>(cap); c.comp_items = w.take<int32_t>(z.pair_cap); c.comp_root = w.take<int32_t>(z.pair_cap);
    c.t_pos = w.take<double>(2 * cap); c.t_vel = w.take<double>(2 * cap); c.t_mass = w.take<double>(cap);
    c.t_mass_hi = w.take<uint64_t>(cap); c.t_mass_lo = w.take<uint64_t>(cap);
    c.t_radius = w.take<double>(cap); c.t_volume = w.take<double>(cap); c.t_eaten = w.take<int64_t>(cap);
    c.t_id = w.take<int32_t>(cap); c.t_flags = w.take<uint8_t>(cap); c.t_energy = w.take<double>(cap);
    c.cell_of = w.take<int32_t>(cap); c.cell_items = w.take<int32_t>(cap);
    c.cell_count = w.take<int32_t>(z.grid_cells + 1); c.cell_start = w.take<int32_t>(z.grid_cells + 1);
    c.giants = w.take<int32_t>(kGiantCap);
    c.grid_cells = (int32_t)z.grid_cells; c.giant_cap = kGiantCap;
    c.grid_dev = w.take<cosmo_grid>(2);
    c.tune_hist = w.take<int32_t>(COSMO_TUNE_HIST_LEN);
    c.corr = w.take<double>(2 * cap);
    c.colbuf_rows = z.colbuf_rows;
    c.colbuf = z.colbuf_rows ? w.take<float>(z.colbuf_rows * 2 * cap) : nullptr;
    if (st) *st = s;
    if (sc) *sc = c;
    if (bytes) *bytes = w.off;
}

__global__ void k_fill_i32(int32_t *p, long long n, int32_t v) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        p[k] = v;
}

}  // namespace

}  // namespace cosmo

using namespace cosmo;

extern "C" {

int64_t cosmo_workspace_bytes(int32_t n_total, uint32_t opts) {
    size_t bytes = 0;
    carve(n_total, opts, nullptr, nullptr, nullptr, &bytes);
    return (int64_t)bytes;
}

int cosmo_workspace_init(void *dev_base, int64_t dev_bytes, int32_t n_total, uint32_t opts, cosmo_state *st,
                         cosmo_scratch *sc, void *stream) {
    if (!dev_base || !st || !sc || n_total < 0 || ((uintptr_t)dev_base & 255)) {
        set_error("cosmo_workspace_init: bad argument (the base must be 256-byte aligned)");
        return COSMO_E_ARG;
    }
    const int64_t need = cosmo_workspace_bytes(n_total, opts);
    if (dev_bytes < need) {
        set_error("cosmo_workspace_init: %lld bytes given, %lld needed", (long long)dev_bytes, (long long)need);
        return COSMO_E_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    size_t bytes = 0;
    carve(n_total, opts, (char *)dev_base, st, sc, &bytes);
    // everything starts zeroed except the four arrays whose idle value is not zero; colbuf (the
    // bulk of a symmetric-kernel workspace) is write-before-read and left alone
    const size_t zero_bytes = sc->colbuf ? (size_t)((char *)sc->colbuf - (char *)dev_base) : bytes;
    COSMO_TRY(check_cuda(cudaMemsetAsync(dev_base, 0, zero_bytes, s), "workspace clear"));
    const long long cap = st->cap;
    k_fill_i32<<<64, 256, 0, s>>>(st->rec_death, st->n_total, -1);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_fill_i32(rec_death)"));
    k_fill_i32<<<64, 256, 0, s>>>(sc->mod_step, cap, -1);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_fill_i32(mod_step)"));
    k_fill_i32<<<64, 256, 0, s>>>(sc->pmin, cap, INT_MAX);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_fill_i32(pmin)"));
    k_fill_i32<<<64, 256, 0, s>>>(sc->pmax, cap, -1);
    return COSMO_LAUNCH_CHECK("k_fill_i32(pmax)");
}

int cosmo_params_default(int32_t n, double G, double dt, double r_max, uint32_t opts, double pos_max, double mass_max,
                         const cosmo_scratch *sc, cosmo_step_params *p) {
    if (!p || n < 0) {
        set_error("cosmo_params_default: bad argument");
        return COSMO_E_ARG;
    }
    memset(p, 0, sizeof(*p));
    p->G = G; p->dt = dt; p->r_max = r_max;
    p->n_upper = n; p->i_begin = 0; p->i_end = n;
    p->shard_index = 0; p->shard_count = 1;
    const bool fp32 = (opts & COSMO_OPT_FP32) != 0;
    // symmetric kernels (each unordered pair once) from 32768 (FP32) / 16384 (FP64) planets, if the
    // workspace has the column-partial buffer for them
    opts &= ~(COSMO_OPT_F32_SYM | COSMO_OPT_F64_SYM | COSMO_OPT_DEFER_INTEGRATE | COSMO_OPT_PEER_EXCHANGE);
    const long long nblk32 = (n + 1023) / 1024, nblk64 = (n + 511) / 512;
    bool sym = false;
    if (fp32 && n >= 32768 && sc && sc->colbuf && sc->colbuf_rows >= 2 * ((nblk32 + 3) / 4)) { opts |= COSMO_OPT_F32_SYM; sym = true; }
    if (!fp32 && n >= 16384 && sc && sc->colbuf && sc->colbuf_rows >= 2 * nblk64) { opts |= COSMO_OPT_F64_SYM; sym = true; }
    p->opts = opts;
    {   // FP32 path: exact power-of-two rescaling, max |coordinate| -> [32, 64), max mass -> [2^19, 2^20)
        int e = 0;
        if (pos_max > 0 && isfinite(pos_max)) { frexp(pos_max, &e); p->pos_exp = e - 6; }
        if (mass_max > 0 && isfinite(mass_max)) { frexp(mass_max, &e); p->mass_exp = e - 20; }
    }
    if (fp32) p->near_mode = 2;   // callers that keep stepping may switch to 1 once a grid-mode stage B has run
    p->small_tiles = n <= 4096 ? 1 : 0;
    const bool collisions = (opts & COSMO_OPT_COLLISIONS) != 0;
    const bool grid = collisions && n >= 8192 && sc && sc->cell_count && sc->grid_dev && sc->tune_hist;
    p->detect_mode = grid ? 2 : 0;
    p->resolve_mode = grid ? 1 : 0;
    p->jsplit = p->jsplit_detect = 1;
    const int jmax = sc ? sc->jsplit_max : 1;
    if (sym) {
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        // ~16 CTAs per resident slot; FP32: as many row blocks per CTA as that allows (4 or 2 -- a CTA works
        // on two row blocks at once, so 1 would idle half of its warps)
        const long long nblk = fp32 ? nblk32 : nblk64;
        long long want = nblk > 1 ? (nblk * nblk) / (2LL * 16 * 2 * sms) : 1;
        const long long lo = (nblk + jmax - 1) / jmax;
        if (fp32) {
            p->sym_rows = want >= 16 ? 4 : 2;
            want /= p->sym_rows;
        }
        long long chunk = want < lo ? lo : want;
        p->sym_chunk = (int32_t)(chunk < 1 ? 1 : (chunk > 32 ? 32 : chunk));   // the launcher raises it if partial rows run out
    } else if (n > 0) {
        p->jsplit = cosmo_suggest_jsplit(n, n, fp32 ? 1 : (p->small_tiles ? 4 : 0), jmax);
    }
    if (collisions && !grid && n > 0) p->jsplit_detect = cosmo_suggest_jsplit(n, n, p->small_tiles ? 5 : 3, jmax);
    return COSMO_OK;
}

int cosmo_step_host(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
                    const cosmo_host_frame *f, void *stream) {
    if (!st || !sc || !p || !f || !f->status_out || f->n < 0 || f->n > st->cap || p->n_upper != f->n ||
        (f->n > 0 && (!f->pos || !f->vel || !f->mass || !f->radius || !f->volume || !f->id || !f->flags))) {
        set_error("cosmo_step_host: bad argument (n=%d cap=%d n_upper=%d)", f ? f->n : -1, st ? st->cap : -1,
                  p ? p->n_upper : -1);
        return COSMO_E_ARG;
    }
    if (p->shard_count > 1) {
        set_error("cosmo_step_host drives one device; sharded frames are sequenced by the caller");
        return COSMO_E_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)f->n;
    // ---- H2D: the n active planets and a fresh status block
#define COSMO_H2D(dst, src, bytes, what)                                                                      \
    if ((src) && (bytes))                                                                                     \
        COSMO_TRY(check_cuda(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, s), what " H2D"))
    COSMO_H2D(st->pos, f->pos, n * 16, "pos");
    COSMO_H2D(st->vel, f->vel, n * 16, "vel");
    COSMO_H2D(st->mass, f->mass, n * 8, "mass");
    COSMO_H2D(st->radius, f->radius, n * 8, "radius");
    COSMO_H2D(st->volume, f->volume, n * 8, "volume");
    COSMO_H2D(st->id, f->id, n * 4, "id");
    COSMO_H2D(st->flags, f->flags, n, "flags");
    if (f->mass_hi && f->mass_lo) {
        COSMO_H2D(st->mass_hi, f->mass_hi, n * 8, "mass_hi");
        COSMO_H2D(st->mass_lo, f->mass_lo, n * 8, "mass_lo");
    } else {
        COSMO_TRY(check_cuda(cudaMemsetAsync(st->mass_hi, 0, n * 8, s), "mass_hi clear"));
        COSMO_TRY(check_cuda(cudaMemsetAsync(st->mass_lo, 0, n * 8, s), "mass_lo clear"));
    }
    if (f->eaten) COSMO_H2D(st->eaten, f->eaten, n * 8, "eaten");
    else COSMO_TRY(check_cuda(cudaMemsetAsync(st->eaten, 0, n * 8, s), "eaten clear"));
#undef COSMO_H2D
    // the status words travel in status_out (the caller's pinned buffer, overwritten below)
    memset(f->status_out, 0, COSMO_STATUS_LEN * sizeof(int64_t));
    f->status_out[COSMO_ST_N_ACTIVE] = f->n;
    f->status_out[COSMO_ST_STEP] = f->frame;
    COSMO_TRY(check_cuda(cudaMemcpyAsync(st->status, f->status_out, COSMO_STATUS_LEN * sizeof(int64_t),
                                         cudaMemcpyHostToDevice, s), "status H2D"));
    // ---- the frame
    if (f->n > 0) COSMO_TRY(cosmo_step(st, sc, p, stream));
    // ---- D2H: status first (it says how many survivors and events there are), then exactly those
    COSMO_TRY(check_cuda(cudaMemcpyAsync(f->status_out, st->status, COSMO_STATUS_LEN * sizeof(int64_t),
                                         cudaMemcpyDeviceToHost, s), "status D2H"));
    COSMO_TRY(check_cuda(cudaStreamSynchronize(s), "cosmo_step_host sync (status)"));
    const size_t m = (size_t)f->status_out[COSMO_ST_N_ACTIVE];
#define COSMO_D2H(dst, src, bytes, what)                                                                      \
    if ((dst) && (bytes))                                                                                     \
        COSMO_TRY(check_cuda(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, s), what " D2H"))
    COSMO_D2H(f->pos_out, st->pos, m * 16, "pos");
    COSMO_D2H(f->vel_out, st->vel, m * 16, "vel");
    COSMO_D2H(f->mass_out, st->mass, m * 8, "mass");
    COSMO_D2H(f->mass_hi_out, st->mass_hi, m * 8, "mass_hi");
    COSMO_D2H(f->mass_lo_out, st->mass_lo, m * 8, "mass_lo");
    COSMO_D2H(f->radius_out, st->radius, m * 8, "radius");
    COSMO_D2H(f->volume_out, st->volume, m * 8, "volume");
    COSMO_D2H(f->eaten_out, st->eaten, m * 8, "eaten");
    COSMO_D2H(f->id_out, st->id, m * 4, "id");
    COSMO_D2H(f->flags_out, st->flags, m, "flags");
    long long ne = f->status_out[COSMO_ST_N_EVENTS];
    if (ne > st->event_cap) ne = st->event_cap;
    if (ne > f->events_cap) ne = f->events_cap;
    COSMO_D2H(f->events_out, st->events, (size_t)(ne > 0 ? ne : 0) * 24, "events");
#undef COSMO_D2H
    return check_cuda(cudaStreamSynchronize(s), "cosmo_step_host sync");
}

}  // extern "C"
