// Stage B (collision-pair detection) and stage C (ordered merge resolver, commit, stable
// compaction).  Replaces pairwise_distances + collision_matrix (reference
// cosmosim/core/universe_old.py:180,192-193), the Python merge loop (:196-210),
// Planet.absorb/destroy (cosmosim/core/planet.py:74-91) and destroy_escaped_planets
// (universe_old.py:78-80,349-350).
#include <math.h>

#include "cosmo_device.cuh"
#include "cosmo_internal.h"
#include "pred.cuh"

namespace cosmo {

// ---------------------------------------------------------------------------------------
// Pair list utilities
// ---------------------------------------------------------------------------------------
__global__ void k_append_pairs(cosmo_state st, cosmo_scratch sc, const uint64_t *src,
                               const int64_t *count_ptr, long long max_count) {
    long long cnt = *count_ptr;
    if (cnt > max_count) {
        cnt = max_count;
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                     (unsigned long long)COSMO_ERR_PAIR_OVERFLOW);
    }
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < cnt;
         k += (long long)gridDim.x * blockDim.x) {
        uint64_t key = src[k];
        emit_pair(st, sc, (int)(key >> 32), (int)(key & 0xffffffffu));
    }
}

// sharded frame: stage B wrote this rank's pairs directly behind row[0]; publish the count there
// and reset the counter that cosmo_append_pair_rows re-fills from all ranks' rows
__global__ void k_seal_pair_row(int64_t *status, int64_t *row) {
    row[0] = status[COSMO_ST_N_PAIRS];
    status[COSMO_ST_N_PAIRS] = 0;
}

// all ranks' lists at once: row r = (count | pairs[max_count]); blockIdx.y = row
__global__ void k_append_pair_rows(cosmo_state st, cosmo_scratch sc, const int64_t *rows, long long max_count) {
    const int64_t *row = rows + (size_t)blockIdx.y * (size_t)(max_count + 1);
    long long cnt = row[0];
    if (cnt > max_count) {
        cnt = max_count;
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                     (unsigned long long)COSMO_ERR_PAIR_OVERFLOW);
    }
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < cnt; k += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = (uint64_t)row[1 + k];
        emit_pair(st, sc, (int)(key >> 32), (int)(key & 0xffffffffu));
    }
}

// Rank sort of the (unique) keys i<<32|j: row-major order = the order in which the
// reference's double loop meets the flags.  O(M^2) compares spread over the grid; M is
// normally a handful.
__global__ void __launch_bounds__(256) k_sort_pairs(cosmo_state st, cosmo_scratch sc) {
    __shared__ uint64_t s_keys[256];
    long long M = st.status[COSMO_ST_N_PAIRS];
    if (M > sc.pair_cap) M = sc.pair_cap;
    if (M == 0) return;
    const long long span = (long long)gridDim.x * blockDim.x;
    const long long rounds = (M + span - 1) / span;
    for (long long r = 0; r < rounds; ++r) {
        const long long k = r * span + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        const uint64_t key = k < M ? sc.pairs[k] : ~0ull;
        long long rank = 0;
        for (long long b0 = 0; b0 < M; b0 += 256) {
            __syncthreads();
            s_keys[threadIdx.x] = b0 + threadIdx.x < M ? sc.pairs[b0 + threadIdx.x] : ~0ull;
            __syncthreads();
            const int lim = (int)min(256ll, M - b0);
            for (int q = 0; q < lim; ++q) rank += s_keys[q] < key;
        }
        if (k < M) sc.pairs_sorted[rank] = key;
    }
}

// ---------------------------------------------------------------------------------------
// Ordered merge resolver.  Single thread: the reference's loop is inherently sequential
// (live masses, stale positions) and the pair list is tiny.
//
// The reference walks i = 0..n-1 ("row"), first storing p_new[i], v_new[i] into planet i if
// it is alive and mobile, then absorbing along the flagged j of that row.  Equivalent
// lazy evaluation for the planets that occur in flagged pairs -- value of planet k when
// row i is being processed:
//     immobile k           : the stored (old) state, never modified
//     k <= i (visited)     : merged value if k was modified at a row >= k, else p_new[k]
//     k >  i (not visited) : merged value if modified at all (stale otherwise: old state)
// A modification made before a mobile planet's own row is discarded when the row reaches
// it (universe_old.py:199-201 overwrite) -- only mass/volume/radius/planets_eaten persist.
// ---------------------------------------------------------------------------------------
// Everything the replay of a pair may need from one planet, fetched with independent loads up
// front: the replay threads are latency-bound (one thread per cluster), so one round trip to L2
// per pair instead of a dozen dependent ones is what matters here.
struct BodyRec {
    uint8_t alive, flags;
    double mass, volume;
    uint64_t mhi, mlo;
    long long eaten;
    int id, mod_step, mod_row;
    double2 cur_p, cur_v, new_p, new_v, old_p, old_v;
};

__device__ __forceinline__ BodyRec load_body(const cosmo_state &st, const cosmo_scratch &sc, int k) {
    BodyRec b;
    b.alive = sc.alive[k];
    b.flags = st.flags[k];
    b.mass = st.mass[k];
    b.volume = st.volume[k];
    b.mhi = st.mass_hi[k];
    b.mlo = st.mass_lo[k];
    b.eaten = st.eaten[k];
    b.id = st.id[k];
    b.mod_step = sc.mod_step[k];
    b.mod_row = sc.mod_row[k];
    b.cur_p = reinterpret_cast<const double2 *>(sc.cur_pos)[k];
    b.cur_v = reinterpret_cast<const double2 *>(sc.cur_vel)[k];
    b.new_p = reinterpret_cast<const double2 *>(sc.pos_new)[k];
    b.new_v = reinterpret_cast<const double2 *>(sc.vel_new)[k];
    b.old_p = reinterpret_cast<const double2 *>(st.pos)[k];
    b.old_v = reinterpret_cast<const double2 *>(st.vel)[k];
    return b;
}

struct PV {
    double2 p, v;
};

// Value of planet k while row `row` of the reference's loop is being processed (see above).
__device__ __forceinline__ PV current_value(const BodyRec &b, int k, int row, int stamp) {
    PV out;
    if (b.flags & COSMO_F_IMMOBILE) {
        out.p = b.old_p;
        out.v = b.old_v;
        return out;
    }
    const bool mod = b.mod_step == stamp;
    const bool use_cur = (k <= row) ? (mod && b.mod_row >= k) : mod;
    if (use_cur) {
        out.p = b.cur_p;
        out.v = b.cur_v;
    } else if (k <= row) {
        out.p = b.new_p;
        out.v = b.new_v;
    } else {
        out.p = b.old_p;
        out.v = b.old_v;
    }
    return out;
}

__device__ __forceinline__ double merge1(double ma, double a, double mo, double o, double msum) {
    // ((m_s * x_s) + (m_o * x_o)) / (m_s + m_o), numpy elementwise, no contraction
    return __ddiv_rn(__dadd_rn(__dmul_rn(ma, a), __dmul_rn(mo, o)), msum);
}

// Replays ONE flagged pair (row i, partner j) exactly as planet.py:74-85 / universe_old.py:204-209
// would at that point of the double loop.  Returns false when absorb()'s guard fails (one of
// the two already dead).  On success writes the event ids to ev[0..1].
__device__ bool replay_pair(const cosmo_state &st, const cosmo_scratch &sc, int i, int j, int stamp,
                            int64_t *ev) {
    const BodyRec bi = load_body(st, sc, i), bj = load_body(st, sc, j);
    if (!bi.alive || !bj.alive) return false;  // absorb() guard, planet.py:75
    const bool i_int = bi.flags & COSMO_F_MASS_INT, j_int = bj.flags & COSMO_F_MASS_INT;
    const u128 mi = (((u128)bi.mhi) << 64) | bi.mlo;
    const u128 mj = (((u128)bj.mhi) << 64) | bj.mlo;
    // universe_old.py:206: `planet.mass >= other_planet.mass` -- the row planet wins ties
    const bool i_wins = mass_ge(i_int, mi, bi.mass, j_int, mj, bj.mass);
    const int W = i_wins ? i : j, L = i_wins ? j : i;
    const BodyRec &bw = i_wins ? bi : bj, &bl = i_wins ? bj : bi;
    const double mW = bw.mass, mL = bl.mass;
    double msum;
    const bool both_int = i_int && j_int;
    u128 isum = 0;
    if (both_int) {
        isum = mi + mj;
        msum = u128_to_double(isum);
    } else {
        msum = __dadd_rn(mW, mL);
    }
    if (!(bw.flags & COSMO_F_IMMOBILE)) {  // planet.py:76-79
        const PV w = current_value(bw, W, i, stamp);
        const PV l = current_value(bl, L, i, stamp);
        reinterpret_cast<double2 *>(sc.cur_pos)[W] =
            make_double2(merge1(mW, w.p.x, mL, l.p.x, msum), merge1(mW, w.p.y, mL, l.p.y, msum));
        reinterpret_cast<double2 *>(sc.cur_vel)[W] =
            make_double2(merge1(mW, w.v.x, mL, l.v.x, msum), merge1(mW, w.v.y, mL, l.v.y, msum));
        sc.mod_row[W] = i;
        sc.mod_step[W] = stamp;
    }
    st.mass[W] = msum;  // planet.py:80
    if (both_int) {
        st.mass_hi[W] = (uint64_t)(isum >> 64);
        st.mass_lo[W] = (uint64_t)isum;
    } else {
        st.flags[W] = bw.flags & ~COSMO_F_MASS_INT;
    }
    const double vol = __dadd_rn(bw.volume, bl.volume);  // planet.py:81
    st.volume[W] = vol;
    st.radius[W] = pow_third(__ddiv_rn(__dmul_rn(3.0, vol), 4.0 * 3.141592653589793));  // :82
    st.eaten[W] = bw.eaten + 1 + bl.eaten;  // planet.py:84
    sc.alive[L] = 0;                        // planet.py:83,86-91
    ev[0] = bw.id;
    ev[1] = bl.id;
    return true;
}

// resolve_mode 0: one thread walks the rank-sorted pair list (few pairs, few launches)
__device__ void resolve_sequential(const cosmo_state &st, const cosmo_scratch &sc) {
    long long M = st.status[COSMO_ST_N_PAIRS];
    if (M > sc.pair_cap) M = sc.pair_cap;
    if (M == 0) return;
    const int stamp = (int)st.status[COSMO_ST_STEP];
    long long nev = st.status[COSMO_ST_N_EVENTS];
    for (long long q = 0; q < M; ++q) {
        const uint64_t key = sc.pairs_sorted[q];
        int64_t ev[2];
        if (!replay_pair(st, sc, (int)(key >> 32), (int)(key & 0xffffffffu), stamp, ev)) continue;
        if (nev < st.event_cap) {
            st.events[3 * nev + 0] = stamp;
            st.events[3 * nev + 1] = ev[0];
            st.events[3 * nev + 2] = ev[1];
        } else {
            st.status[COSMO_ST_ERROR] |= COSMO_ERR_EVENT_OVERFLOW;
        }
        ++nev;
    }
    st.status[COSMO_ST_N_EVENTS] = nev;
    st.status[COSMO_ST_N_COMPLEX] = M;
}

__global__ void k_resolve(cosmo_state st, cosmo_scratch sc) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    resolve_sequential(st, sc);
}

// ---------------------------------------------------------------------------------------
// resolve_mode 1: row-major (CSR) ordering of the flagged pairs by counting sort, then
//   * "simple" pairs -- both planets have each other as their ONLY partner in the whole
//     flagged set -- cannot interact with any other pair, so the sequential loop's outcome
//     for them is independent of everything else: one thread each, in parallel;
//   * everything else (chains, triples) is split into connected components, each replayed
//     by one thread in sorted order;
//   * merge events land in a slot per sorted pair and are appended to the log in sorted
//     order, which is the visitation order of the reference's double loop.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long pair_count(const cosmo_state &st, const cosmo_scratch &sc) {
    long long M = st.status[COSMO_ST_N_PAIRS];
    return M > sc.pair_cap ? sc.pair_cap : M;
}

__global__ void __launch_bounds__(256) k_csr_count(cosmo_state st, cosmo_scratch sc) {
    const long long M = pair_count(st, sc);
    if (blockIdx.x == 0 && threadIdx.x == 0) st.status[COSMO_ST_N_HEAVY] = 0;   // re-used as the long-row counter
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs[q];
        const int i = (int)(key >> 32), j = (int)(key & 0xffffffffu);
        atomicAdd(&sc.row_count[i], 1);
        atomicMin(&sc.pmin[i], j); atomicMax(&sc.pmax[i], j);
        atomicMin(&sc.pmin[j], i); atomicMax(&sc.pmax[j], i);
    }
}

__global__ void __launch_bounds__(256) k_csr_fill(cosmo_state st, cosmo_scratch sc) {
    const long long M = pair_count(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs[q];
        const int i = (int)(key >> 32);
        const int slot = sc.row_start[i] + atomicSub(&sc.row_count[i], 1) - 1;  // counts return to 0
        sc.pairs_sorted[slot] = key;
    }
}

// order every row by partner index: short rows by the owner of the row's first slot (insertion
// sort), long rows -- the star of a crowded disk is flagged with hundreds of planets -- by a rank
// sort spread over a whole CTA (k_csr_sort_long_rows)
constexpr int ROW_LONG = 48;

__global__ void __launch_bounds__(256) k_csr_sort_long_rows(cosmo_state st, cosmo_scratch sc) {
    const int nlong = (int)st.status[COSMO_ST_N_HEAVY];           // rows listed by k_csr_sort_rows
    uint64_t *tmp = reinterpret_cast<uint64_t *>(sc.ev_tmp);      // free until the replay kernels
    for (int r = blockIdx.x; r < nlong; r += gridDim.x) {
        const int i = sc.cplx_list[r];                            // cplx_list is free until k_scatter_bit(2)
        const int s0 = sc.row_start[i], s1 = sc.row_start[i + 1];
        for (int a = s0 + threadIdx.x; a < s1; a += blockDim.x) {
            const uint64_t key = sc.pairs_sorted[a];
            int rank = 0;
            for (int b = s0; b < s1; ++b) rank += sc.pairs_sorted[b] < key;
            tmp[s0 + rank] = key;
        }
        __syncthreads();
        for (int a = s0 + threadIdx.x; a < s1; a += blockDim.x) sc.pairs_sorted[a] = tmp[a];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_csr_sort_rows(cosmo_state st, cosmo_scratch sc) {
    const long long M = pair_count(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(sc.pairs_sorted[q] >> 32);
        const int s0 = sc.row_start[i], s1 = sc.row_start[i + 1];
        sc.pair_flag[q] = 0;
        if (q != s0 || s1 - s0 < 2) continue;
        if (s1 - s0 > ROW_LONG) {      // handed to k_csr_sort_long_rows (N_HEAVY is free after stage B)
            const unsigned long long r =
                atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_HEAVY]), 1ull);
            sc.cplx_list[r] = i;
            continue;
        }
        for (int a = s0 + 1; a < s1; ++a) {
            const uint64_t key = sc.pairs_sorted[a];
            int b = a - 1;
            while (b >= s0 && sc.pairs_sorted[b] > key) {
                sc.pairs_sorted[b + 1] = sc.pairs_sorted[b];
                --b;
            }
            sc.pairs_sorted[b + 1] = key;
        }
    }
}

__global__ void __launch_bounds__(256) k_resolve_simple(cosmo_state st, cosmo_scratch sc) {
    const long long M = pair_count(st, sc);
    const int stamp = (int)st.status[COSMO_ST_STEP];
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs_sorted[q];
        const int i = (int)(key >> 32), j = (int)(key & 0xffffffffu);
        const bool simple = sc.pmin[i] == j && sc.pmax[i] == j && sc.pmin[j] == i && sc.pmax[j] == i;
        if (!simple) {
            sc.pair_flag[q] = 2;
            continue;
        }
        // of the two orientations only the one the loop meets first can do anything
        const bool first = i < j || sc.row_start[j + 1] == sc.row_start[j];
        if (first && replay_pair(st, sc, i, j, stamp, &sc.ev_tmp[2 * q])) sc.pair_flag[q] = 1;
    }
}

// Stable (order-preserving) compaction of the sorted-pair indices carrying a flag bit, multi-CTA:
// 0/1 marks -> device-wide exclusive scan (scan.cu) -> scatter.  bit 2: indices of the clustered
// pairs -> cplx_list; bit 1: the merge events -> appended to the log in sorted-pair order, which
// is the visitation order of the reference's double loop.
__global__ void __launch_bounds__(256) k_mark_bit(cosmo_state st, cosmo_scratch sc, int bit) {
    const long long M = pair_count(st, sc);
    if (bit == 1 && blockIdx.x == 0 && threadIdx.x == 0) st.status[COSMO_ST_EV_BASE] = st.status[COSMO_ST_N_EVENTS];
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x)
        sc.scan_a[q] = (sc.pair_flag[q] & bit) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_scatter_bit(cosmo_state st, cosmo_scratch sc, int bit) {
    const long long M = pair_count(st, sc);
    const int stamp = (int)st.status[COSMO_ST_STEP];
    const long long base = st.status[COSMO_ST_EV_BASE];
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        if (!(sc.pair_flag[q] & bit)) continue;
        const long long pos = sc.scan_a[q];
        if (bit == 2) {
            sc.cplx_list[pos] = (int)q;
        } else if (base + pos < st.event_cap) {
            st.events[3 * (base + pos) + 0] = stamp;
            st.events[3 * (base + pos) + 1] = sc.ev_tmp[2 * q];
            st.events[3 * (base + pos) + 2] = sc.ev_tmp[2 * q + 1];
        } else {
            atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                     (unsigned long long)COSMO_ERR_EVENT_OVERFLOW);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const long long total = M > 0 ? sc.scan_a[M] : 0;
        if (bit == 2) st.status[COSMO_ST_N_COMPLEX] = total;
        else st.status[COSMO_ST_N_EVENTS] = base + total;
    }
}

// Pairs that are not simple form clusters (chains, triples, in a crowded disk also larger blobs).
// Clusters are disjoint sets of planets, so they are independent of each other; inside a cluster
// the reference's order matters.  Connected components by lock-free union-find over the cluster
// edges (hook the larger root under the smaller with atomicCAS, path halving on find): three
// multi-CTA kernels, no iteration to a fixed point; the representative is the smallest planet
// index of the component.  A counting sort then groups the pairs by component, a parallel rank
// sort orders every group by sorted-pair index (= visitation order), and one thread per component
// replays its pairs.
__device__ __forceinline__ int uf_find(int *parent, int v) {
    int p = *(volatile int *)&parent[v];
    while (p != v) {
        const int gp = *(volatile int *)&parent[p];
        if (gp != p) parent[v] = gp;     // path halving (benign race: only ever moves towards the root)
        v = p;
        p = gp;
    }
    return v;
}

__global__ void __launch_bounds__(256) k_uf_init(cosmo_state st, cosmo_scratch sc) {
    const long long C = st.status[COSMO_ST_N_COMPLEX];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs_sorted[sc.cplx_list[c]];
        const int i = (int)(key >> 32), j = (int)(key & 0xffffffffu);
        sc.label[i] = i;
        sc.label[j] = j;
    }
}

__global__ void __launch_bounds__(256) k_uf_union(cosmo_state st, cosmo_scratch sc) {
    const long long C = st.status[COSMO_ST_N_COMPLEX];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs_sorted[sc.cplx_list[c]];
        int a = uf_find(sc.label, (int)(key >> 32)), b = uf_find(sc.label, (int)(key & 0xffffffffu));
        while (a != b) {
            const int hi = max(a, b), lo = min(a, b);
            const int old = atomicCAS(&sc.label[hi], hi, lo);    // hook root `hi` under `lo`
            if (old == hi) break;
            a = uf_find(sc.label, old);                           // someone else moved it: retry from the new roots
            b = uf_find(sc.label, lo);
        }
    }
}

// flatten (label = root = smallest index of the component) and count the pairs per component
__global__ void __launch_bounds__(256) k_uf_count(cosmo_state st, cosmo_scratch sc) {
    const long long C = st.status[COSMO_ST_N_COMPLEX];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs_sorted[sc.cplx_list[c]];
        const int root = uf_find(sc.label, (int)(key >> 32));
        sc.comp_root[c] = root;
        atomicAdd(&sc.row_count[root], 1);
    }
}

__global__ void __launch_bounds__(256) k_comp_fill(cosmo_state st, cosmo_scratch sc) {
    const long long C = st.status[COSMO_ST_N_COMPLEX];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x) {
        const int root = sc.comp_root[c];
        const int slot = sc.row_start[root] + atomicSub(&sc.row_count[root], 1) - 1;
        sc.comp_items[slot] = sc.cplx_list[c];
    }
}

// order every component's pairs by sorted-pair index: rank of a pair = number of smaller indices in
// its component (all pairs in parallel; components are small to moderate)
__global__ void __launch_bounds__(256) k_comp_rank(cosmo_state st, cosmo_scratch sc) {
    const long long C = st.status[COSMO_ST_N_COMPLEX];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x) {
        const int q = sc.cplx_list[c];
        const int root = sc.comp_root[c];
        const int s0 = sc.row_start[root], s1 = sc.row_start[root + 1];
        int rank = 0;
        for (int s = s0; s < s1; ++s) rank += sc.comp_items[s] < q;
        sc.scan_a[s0 + rank] = q;          // scan_a is free between the two ordered compactions
    }
}

// one thread per component root: replay the component's pairs in visitation order
__global__ void __launch_bounds__(128) k_comp_resolve(cosmo_state st, cosmo_scratch sc) {
    if (st.status[COSMO_ST_N_COMPLEX] == 0) return;
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int s0 = sc.row_start[v], s1 = sc.row_start[v + 1];
    if (s1 <= s0) return;
    const int stamp = (int)st.status[COSMO_ST_STEP];
    for (int a = s0; a < s1; ++a) {
        const int q = sc.scan_a[a];
        const uint64_t key = sc.pairs_sorted[q];
        if (replay_pair(st, sc, (int)(key >> 32), (int)(key & 0xffffffffu), stamp, &sc.ev_tmp[2 * q]))
            sc.pair_flag[q] |= 1;
    }
}

// re-arm the self-resetting partner arrays
__global__ void __launch_bounds__(256) k_csr_cleanup(cosmo_state st, cosmo_scratch sc) {
    const long long M = pair_count(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs[q];
        const int i = (int)(key >> 32), j = (int)(key & 0xffffffffu);
        sc.pmin[i] = 0x7fffffff; sc.pmax[i] = -1;
        sc.pmin[j] = 0x7fffffff; sc.pmax[j] = -1;
    }
}

// ---------------------------------------------------------------------------------------
// Commit + stable compaction.  1024 slots per CTA (256 threads x 4 consecutive... slots
// are taken strided so loads coalesce).
// ---------------------------------------------------------------------------------------
constexpr int CHUNK = 1024;

__global__ void __launch_bounds__(256) k_commit(cosmo_state st, cosmo_scratch sc, int bounded) {
    __shared__ int s_cnt[8];
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int c0 = blockIdx.x * CHUNK;
    if (c0 >= n) return;
    const int stamp = (int)st.status[COSMO_ST_STEP];
    double2 *pos = reinterpret_cast<double2 *>(st.pos);
    double2 *vel = reinterpret_cast<double2 *>(st.vel);
    int survivors = 0, escaped_now = 0;
    for (int q = 0; q < CHUNK / 256; ++q) {
        const int k = c0 + q * 256 + threadIdx.x;
        if (k >= n) break;
        const bool alive = sc.alive[k] != 0;
        const uint8_t fl = st.flags[k];
        if (alive && !(fl & COSMO_F_IMMOBILE)) {
            const bool use_cur = sc.mod_step[k] == stamp && sc.mod_row[k] >= k;
            pos[k] = use_cur ? reinterpret_cast<const double2 *>(sc.cur_pos)[k]
                             : reinterpret_cast<const double2 *>(sc.pos_new)[k];
            vel[k] = use_cur ? reinterpret_cast<const double2 *>(sc.cur_vel)[k]
                             : reinterpret_cast<const double2 *>(sc.vel_new)[k];
        }
        const bool esc = bounded && sc.escaped[k];  // flagged BEFORE interact, universe_old.py:345-350
        const bool live = alive && !esc;
        if (alive && esc) ++escaped_now;
        if (!live) {
            const int id = st.id[k];
            st.rec_mass[id] = st.mass[k];
            st.rec_mass_hi[id] = st.mass_hi[k];
            st.rec_mass_lo[id] = st.mass_lo[k];
            st.rec_radius[id] = st.radius[k];
            st.rec_volume[id] = st.volume[k];
            st.rec_eaten[id] = st.eaten[k];
            st.rec_flags[id] = fl;
            st.rec_death[id] = stamp;
        }
        sc.alive[k] = live ? 1 : 0;
        survivors += live;
    }
    // block reduction of the survivor / escape counts
    int packed = survivors | (escaped_now << 16);
    for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = packed;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += s_cnt[w];
        const int surv = tot & 0xffff, esc = tot >> 16;
        sc.chunk_alive[blockIdx.x] = surv;
        const int len = min(CHUNK, n - c0);
        if (len != surv)
            atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_DEAD]),
                      (unsigned long long)(len - surv));
        if (esc)
            atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_ESCAPED]),
                      (unsigned long long)esc);
    }
}

// survivors of chunk c move to [sum(chunk_alive[0..c)) + rank within chunk): stable.
__global__ void __launch_bounds__(256) k_compact_scatter(cosmo_state st, cosmo_scratch sc) {
    if (st.status[COSMO_ST_N_DEAD] == 0) return;
    __shared__ int s_red[8];
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int c0 = blockIdx.x * CHUNK;
    if (c0 >= n) return;
    // exclusive prefix of the chunk counts (redundant per CTA; chunk counts are few)
    int part = 0;
    for (int c = threadIdx.x; c < (int)blockIdx.x; c += 256) part += sc.chunk_alive[c];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        s_base = t;
    }
    __syncthreads();
    int running = s_base;
    const double2 *pos = reinterpret_cast<const double2 *>(st.pos);
    const double2 *vel = reinterpret_cast<const double2 *>(st.vel);
    for (int q = 0; q < CHUNK / 256; ++q) {
        const int k = c0 + q * 256 + threadIdx.x;
        const int live = (k < n && sc.alive[k]) ? 1 : 0;
        // block-wide exclusive scan of `live` over 256 threads
        const unsigned ballot = __ballot_sync(0xffffffffu, live);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const int in_warp = __popc(ballot & ((1u << lane) - 1));
        __syncthreads();
        if (lane == 0) s_warp[w] = __popc(ballot);
        __syncthreads();
        int before = 0, total = 0;
        for (int x = 0; x < 8; ++x) {
            const int c = s_warp[x];
            before += x < w ? c : 0;
            total += c;
        }
        if (live) {
            const int d = running + before + in_warp;
            reinterpret_cast<double2 *>(sc.t_pos)[d] = pos[k];
            reinterpret_cast<double2 *>(sc.t_vel)[d] = vel[k];
            sc.t_mass[d] = st.mass[k];
            sc.t_mass_hi[d] = st.mass_hi[k];
            sc.t_mass_lo[d] = st.mass_lo[k];
            sc.t_radius[d] = st.radius[k];
            sc.t_volume[d] = st.volume[k];
            sc.t_eaten[d] = st.eaten[k];
            sc.t_id[d] = st.id[k];
            sc.t_flags[d] = st.flags[k];
            if (st.energy) sc.t_energy[d] = st.energy[k];
        }
        running += total;
    }
}

__global__ void __launch_bounds__(256) k_compact_copyback(cosmo_state st, cosmo_scratch sc) {
    const long long dead = st.status[COSMO_ST_N_DEAD];
    if (dead == 0) return;
    const int n_new = (int)(st.status[COSMO_ST_N_ACTIVE] - dead);
    double2 *pos = reinterpret_cast<double2 *>(st.pos);
    double2 *vel = reinterpret_cast<double2 *>(st.vel);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_new; k += gridDim.x * blockDim.x) {
        pos[k] = reinterpret_cast<const double2 *>(sc.t_pos)[k];
        vel[k] = reinterpret_cast<const double2 *>(sc.t_vel)[k];
        st.mass[k] = sc.t_mass[k];
        st.mass_hi[k] = sc.t_mass_hi[k];
        st.mass_lo[k] = sc.t_mass_lo[k];
        st.radius[k] = sc.t_radius[k];
        st.volume[k] = sc.t_volume[k];
        st.eaten[k] = sc.t_eaten[k];
        st.id[k] = sc.t_id[k];
        st.flags[k] = sc.t_flags[k];
        if (st.energy) st.energy[k] = sc.t_energy[k];
    }
}

__global__ void k_end_frame(cosmo_state st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        st.status[COSMO_ST_N_ACTIVE] -= st.status[COSMO_ST_N_DEAD];
        st.status[COSMO_ST_STEP] += 1;  // universe_old.py:361
    }
}


// ---------------------------------------------------------------------------------------
// Stage C for small scenes (n_upper <= TAIL_N, resolve_mode 0) in ONE launch of ONE CTA: rank sort
// of the pairs, the sequential replay, commit (universe_old.py:199-201,349-350), stable compaction
// held in registers, frame counter.  At the size of the reference's example scenes a frame is
// launch-bound; this replaces six launches.
// ---------------------------------------------------------------------------------------
constexpr int TAIL_N = 2048;
constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_PER = TAIL_N / TAIL_THREADS;

__global__ void __launch_bounds__(TAIL_THREADS) k_tail_small(cosmo_state st, cosmo_scratch sc, int collisions,
                                                             int bounded) {
    __shared__ uint64_t s_keys[TAIL_THREADS];
    __shared__ int s_warp[TAIL_THREADS / 32];
    __shared__ int s_esc[TAIL_THREADS / 32];
    __shared__ int s_total, s_escaped;
    const int tid = threadIdx.x;
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int stamp = (int)st.status[COSMO_ST_STEP];
    if (collisions) {
        long long M = st.status[COSMO_ST_N_PAIRS];
        if (M > sc.pair_cap) M = sc.pair_cap;
        for (long long k0 = 0; k0 < M; k0 += TAIL_THREADS) {   // rank sort of the unique keys
            const long long k = k0 + tid;
            const uint64_t key = k < M ? sc.pairs[k] : ~0ull;
            long long rank = 0;
            for (long long b0 = 0; b0 < M; b0 += TAIL_THREADS) {
                __syncthreads();
                s_keys[tid] = b0 + tid < M ? sc.pairs[b0 + tid] : ~0ull;
                __syncthreads();
                const int lim = (int)min((long long)TAIL_THREADS, M - b0);
                for (int q = 0; q < lim; ++q) rank += s_keys[q] < key;
            }
            if (k < M) sc.pairs_sorted[rank] = key;
        }
        __threadfence_block();
        __syncthreads();
        if (tid == 0) resolve_sequential(st, sc);
        __threadfence_block();
        __syncthreads();
    }
    // commit (k_commit) + gather of the survivors into registers: TAIL_PER consecutive slots per thread
    double2 ps[TAIL_PER], vs[TAIL_PER];
    double ms[TAIL_PER], rd[TAIL_PER], vo[TAIL_PER], en[TAIL_PER];
    uint64_t mh[TAIL_PER], ml[TAIL_PER];
    long long ea[TAIL_PER];
    int idv[TAIL_PER];
    uint8_t flv[TAIL_PER];
    int live[TAIL_PER], mine = 0, esc_mine = 0;
#pragma unroll
    for (int q = 0; q < TAIL_PER; ++q) {
        const int k = tid * TAIL_PER + q;
        live[q] = 0;
        if (k >= n) continue;
        const bool alive = sc.alive[k] != 0;
        const uint8_t fl = st.flags[k];
        const bool esc = bounded && sc.escaped[k];   // flagged BEFORE interact, universe_old.py:345-350
        if (alive && esc) ++esc_mine;
        if (alive && !esc) {
            live[q] = 1;
            if (!(fl & COSMO_F_IMMOBILE)) {
                const bool use_cur = sc.mod_step[k] == stamp && sc.mod_row[k] >= k;
                ps[q] = use_cur ? reinterpret_cast<const double2 *>(sc.cur_pos)[k]
                                : reinterpret_cast<const double2 *>(sc.pos_new)[k];
                vs[q] = use_cur ? reinterpret_cast<const double2 *>(sc.cur_vel)[k]
                                : reinterpret_cast<const double2 *>(sc.vel_new)[k];
            } else {
                ps[q] = reinterpret_cast<const double2 *>(st.pos)[k];
                vs[q] = reinterpret_cast<const double2 *>(st.vel)[k];
            }
            ms[q] = st.mass[k]; mh[q] = st.mass_hi[k]; ml[q] = st.mass_lo[k];
            rd[q] = st.radius[k]; vo[q] = st.volume[k]; ea[q] = st.eaten[k]; idv[q] = st.id[k]; flv[q] = fl;
            en[q] = st.energy ? st.energy[k] : 0.0;
        } else {
            const int id = st.id[k];                 // Planet keeps mass / radius after destroy()
            st.rec_mass[id] = st.mass[k];
            st.rec_mass_hi[id] = st.mass_hi[k];
            st.rec_mass_lo[id] = st.mass_lo[k];
            st.rec_radius[id] = st.radius[k];
            st.rec_volume[id] = st.volume[k];
            st.rec_eaten[id] = st.eaten[k];
            st.rec_flags[id] = fl;
            st.rec_death[id] = stamp;
        }
        mine += live[q];
    }
    // block-wide exclusive scan of `mine`, block sum of the escape count
    const int lane = tid & 31, w = tid >> 5;
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int esc_warp = __reduce_add_sync(0xffffffffu, esc_mine);
    if (lane == 31) s_warp[w] = incl;
    if (lane == 0) s_esc[w] = esc_warp;
    __syncthreads();
    if (w == 0) {
        int v = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        s_warp[lane] = v;
        const int e = __reduce_add_sync(0xffffffffu, s_esc[lane]);
        if (lane == 31) { s_total = v; s_escaped = e; }
    }
    __syncthreads();   // every source slot has been read into registers above: the in-place scatter is safe
    int dst = incl - mine + (w ? s_warp[w - 1] : 0);
#pragma unroll
    for (int q = 0; q < TAIL_PER; ++q) {
        if (!live[q]) continue;
        reinterpret_cast<double2 *>(st.pos)[dst] = ps[q];
        reinterpret_cast<double2 *>(st.vel)[dst] = vs[q];
        st.mass[dst] = ms[q]; st.mass_hi[dst] = mh[q]; st.mass_lo[dst] = ml[q];
        st.radius[dst] = rd[q]; st.volume[dst] = vo[q]; st.eaten[dst] = ea[q]; st.id[dst] = idv[q];
        st.flags[dst] = flv[q];
        if (st.energy) st.energy[dst] = en[q];
        ++dst;
    }
    if (tid == 0) {
        st.status[COSMO_ST_N_DEAD] = n - s_total;
        st.status[COSMO_ST_N_ESCAPED] += s_escaped;
        st.status[COSMO_ST_N_ACTIVE] = s_total;
        st.status[COSMO_ST_STEP] += 1;  // universe_old.py:361
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
int launch_append_pairs(const cosmo_state &st, const cosmo_scratch &sc, const uint64_t *src,
                        const int64_t *count_ptr, int64_t max_count, cudaStream_t s) {
    k_append_pairs<<<32, 256, 0, s>>>(st, sc, src, count_ptr, (long long)max_count);
    return COSMO_LAUNCH_CHECK("k_append_pairs");
}

int launch_seal_pair_row(int64_t *status, int64_t *row, cudaStream_t s) {
    k_seal_pair_row<<<1, 1, 0, s>>>(status, row);
    return COSMO_LAUNCH_CHECK("k_seal_pair_row");
}

int launch_append_pair_rows(const cosmo_state &st, const cosmo_scratch &sc, const int64_t *rows, int n_rows,
                            long long max_count, cudaStream_t s) {
    k_append_pair_rows<<<dim3(32, n_rows), 256, 0, s>>>(st, sc, rows, max_count);
    return COSMO_LAUNCH_CHECK("k_append_pair_rows");
}

int launch_resolve_commit(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                          cudaStream_t s) {
    if (p.n_upper <= TAIL_N && (p.resolve_mode == 0 || !(p.opts & COSMO_OPT_COLLISIONS))) {
        k_tail_small<<<1, TAIL_THREADS, 0, s>>>(st, sc, (p.opts & COSMO_OPT_COLLISIONS) ? 1 : 0,
                                                (p.opts & COSMO_OPT_BOUNDED) ? 1 : 0);
        return COSMO_LAUNCH_CHECK("k_tail_small");
    }
    const int chunks = (p.n_upper + CHUNK - 1) / CHUNK > 0 ? (p.n_upper + CHUNK - 1) / CHUNK : 1;
    if ((p.opts & COSMO_OPT_COLLISIONS) && p.resolve_mode == 0) {
        k_sort_pairs<<<64, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_sort_pairs"));
        k_resolve<<<1, 32, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_resolve"));
    } else if (p.opts & COSMO_OPT_COLLISIONS) {
        const int pb = 296;
        k_csr_count<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_csr_count"));
        COSMO_TRY(launch_scan_exclusive(sc.row_count, sc.row_start, sc.scan_blk, &st.status[COSMO_ST_N_ACTIVE],
                                        p.n_upper, 0, s));
        k_csr_fill<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_csr_fill"));
        k_csr_sort_rows<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_csr_sort_rows"));
        k_csr_sort_long_rows<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_csr_sort_long_rows"));
        k_resolve_simple<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_resolve_simple"));
        k_mark_bit<<<pb, 256, 0, s>>>(st, sc, 2);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_mark_bit(2)"));
        COSMO_TRY(launch_scan_exclusive(sc.scan_a, sc.scan_a, sc.scan_blk, &st.status[COSMO_ST_N_PAIRS],
                                        sc.pair_cap, 0, s));
        k_scatter_bit<<<pb, 256, 0, s>>>(st, sc, 2);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_scatter_bit(2)"));
        k_uf_init<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_uf_init"));
        k_uf_union<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_uf_union"));
        k_uf_count<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_uf_count"));
        COSMO_TRY(launch_scan_exclusive(sc.row_count, sc.row_start, sc.scan_blk, &st.status[COSMO_ST_N_ACTIVE],
                                        p.n_upper, 0, s));
        k_comp_fill<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_comp_fill"));
        k_comp_rank<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_comp_rank"));
        k_comp_resolve<<<(p.n_upper + 127) / 128 > 0 ? (p.n_upper + 127) / 128 : 1, 128, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_comp_resolve"));
        k_mark_bit<<<pb, 256, 0, s>>>(st, sc, 1);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_mark_bit(1)"));
        COSMO_TRY(launch_scan_exclusive(sc.scan_a, sc.scan_a, sc.scan_blk, &st.status[COSMO_ST_N_PAIRS],
                                        sc.pair_cap, 0, s));
        k_scatter_bit<<<pb, 256, 0, s>>>(st, sc, 1);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_scatter_bit(1)"));
        k_csr_cleanup<<<pb, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_csr_cleanup"));
    }
    k_commit<<<chunks, 256, 0, s>>>(st, sc, (p.opts & COSMO_OPT_BOUNDED) ? 1 : 0);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_commit"));
    k_compact_scatter<<<chunks, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_compact_scatter"));
    int cb = chunks * 4;
    if (cb > 148 * 8) cb = 148 * 8;
    k_compact_copyback<<<cb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_compact_copyback"));
    k_end_frame<<<1, 32, 0, s>>>(st);
    return COSMO_LAUNCH_CHECK("k_end_frame");
}

}  // namespace cosmo
