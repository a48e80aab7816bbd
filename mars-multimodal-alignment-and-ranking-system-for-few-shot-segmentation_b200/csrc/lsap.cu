// Exact rectangular linear-sum assignment on the device: the solver behind Matcher's
// bidirectional patch matching, scipy.optimize.linear_sum_assignment(S, maximize=True) at
// matcher/Matcher.py:449-450 (forward: fg support rows x query patches) and :471-472 (reverse).
// Shortest augmenting paths (Jonker-Volgenant) with dense reduced costs in float64: one CTA per problem,
// block-wide arg-min over the sinks, parallel relaxation, sequential augmentation.  The smaller side is
// always the set of sources, like scipy's rectangular solver every source gets assigned.
#include <algorithm>

#include "common.cuh"

namespace marsb200 {

constexpr int LSAP_THREADS = 512;
constexpr double LSAP_INF = 1e300;

// per source: u, dsrc (f64), id / assigned sink / reached list (i32); per sink: v, dist (f64), id / predecessor /
// assigned source (u16) and a scanned flag: 23 bytes, so that the 5-shot reverse matching (<= 1369 x 6845) fits
__host__ __device__ inline size_t lsap_smem_bytes(int t_cap, int m_cap) {
    return (size_t)t_cap * (2 * 8 + 3 * 4) + (size_t)m_cap * (2 * 8 + 3 * 2 + 1) + 64;
}
constexpr unsigned short LSAP_NONE = 0xffff;

__device__ inline void lsap_argmin(const double* dist, const unsigned char* scanned, int M, double& best, int& best_j,
                                   double* s_val, int* s_idx) {
    double v = LSAP_INF;
    int j = 0x7fffffff;
    for (int t = threadIdx.x; t < M; t += LSAP_THREADS) {
        const double k = scanned[t] ? LSAP_INF : dist[t];
        if (k < v) {
            v = k;
            j = t;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oj = __shfl_xor_sync(0xffffffffu, j, o);
        if (ov < v || (ov == v && oj < j)) {
            v = ov;
            j = oj;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_val[threadIdx.x >> 5] = v;
        s_idx[threadIdx.x >> 5] = j;
    }
    __syncthreads();
    best = s_val[0];
    best_j = s_idx[0];
#pragma unroll
    for (int w = 1; w < LSAP_THREADS / 32; ++w) {
        if (s_val[w] < best || (s_val[w] == best && s_idx[w] < best_j)) {
            best = s_val[w];
            best_j = s_idx[w];
        }
    }
}

// sim [E, R, C]; row_sel [E, R] / col_sel [E, C] choose the participating rows / columns (null = all).
// row_to_col [E, R]: assigned column of each selected row or -1; objective [E]: sum of the assigned similarities.
__global__ void __launch_bounds__(LSAP_THREADS, 1) lsap_kernel(const float* __restrict__ sim, const uint8_t* __restrict__ row_sel,
                                                             const uint8_t* __restrict__ col_sel, int R, int Ccols,
                                                             int maximize, int t_cap, int m_cap,
                                                             int32_t* __restrict__ row_to_col,
                                                             double* __restrict__ objective, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char lsap_smem_raw[];
    double* u = reinterpret_cast<double*>(lsap_smem_raw);  // [t_cap]
    double* dsrc = u + t_cap;                              // [t_cap]
    double* v = dsrc + t_cap;                              // [m_cap]
    double* dist = v + m_cap;                              // [m_cap]
    int* src_id = reinterpret_cast<int*>(dist + m_cap);    // [t_cap] row (or column) index of source i
    int* src_sink = src_id + t_cap;                        // [t_cap] sink assigned to source i
    int* list = src_sink + t_cap;                          // [t_cap] reached sources
    unsigned short* sink_id = reinterpret_cast<unsigned short*>(list + t_cap);  // [m_cap]
    unsigned short* pred_src = sink_id + m_cap;            // [m_cap]
    unsigned short* sink_src = pred_src + m_cap;           // [m_cap] source assigned to sink j, or LSAP_NONE
    unsigned char* scanned = reinterpret_cast<unsigned char*>(sink_src + m_cap);  // [m_cap]
    __shared__ double s_val[2][LSAP_THREADS / 32];
    __shared__ int s_idx[2][LSAP_THREADS / 32];
    __shared__ int s_nr, s_nc, s_nreached;
    const int tid = threadIdx.x;
    const int64_t e = blockIdx.x;
    const float* S = sim + e * (int64_t)R * Ccols;
    int32_t* out = row_to_col + e * R;

    // selected rows / columns (ascending); the lists land in the sink arrays first and are swapped below if needed
    for (int r = tid; r < R; r += LSAP_THREADS) out[r] = -1;
    if (tid == 0) {
        int nr = 0, nc = 0;
        for (int r = 0; r < R; ++r) nr += (!row_sel || row_sel[e * R + r]) ? 1 : 0;
        for (int c = 0; c < Ccols; ++c) nc += (!col_sel || col_sel[e * Ccols + c]) ? 1 : 0;
        s_nr = nr;
        s_nc = nc;
    }
    __syncthreads();
    const int nr = s_nr, nc = s_nc;
    const bool rows_are_sources = nr <= nc;
    const int T = rows_are_sources ? nr : nc, M = rows_are_sources ? nc : nr;
    if (T == 0) {
        if (tid == 0) objective[e] = 0.0;
        return;
    }
    if (T > t_cap || M > m_cap) {
        if (tid == 0) {
            objective[e] = nan("");
            atomicMax(status, max(T, M));
        }
        return;
    }
    if (tid == 0) {
        int a = 0, b = 0;
        for (int r = 0; r < R; ++r)
            if (!row_sel || row_sel[e * R + r]) {
                if (rows_are_sources) src_id[a++] = r;
                else sink_id[a++] = (unsigned short)r;
            }
        for (int c = 0; c < Ccols; ++c)
            if (!col_sel || col_sel[e * Ccols + c]) {
                if (rows_are_sources) sink_id[b++] = (unsigned short)c;
                else src_id[b++] = c;
            }
    }
    __syncthreads();
    const double sign = maximize ? -1.0 : 1.0;
    auto cost = [&](int i, int j) -> double {
        const int r = rows_are_sources ? src_id[i] : sink_id[j];
        const int c = rows_are_sources ? sink_id[j] : src_id[i];
        return sign * (double)S[(int64_t)r * Ccols + c];
    };

    // duals: sinks may stay unassigned (inequality), so their dual starts at 0 and only decreases; the sources are
    // all assigned (equality), so u_i = min_j c_ij makes every reduced cost non-negative
    for (int i = tid >> 5; i < T; i += LSAP_THREADS / 32) {
        double mn = LSAP_INF;
        for (int j = tid & 31; j < M; j += 32) mn = fmin(mn, cost(i, j));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if ((tid & 31) == 0) {
            u[i] = mn;
            src_sink[i] = -1;
        }
    }
    for (int j = tid; j < M; j += LSAP_THREADS) {
        v[j] = 0.0;
        sink_src[j] = LSAP_NONE;
    }
    __syncthreads();

    for (int r = 0; r < T; ++r) {
        const double ur = u[r];
        for (int j = tid; j < M; j += LSAP_THREADS) {
            const double d = cost(r, j) - ur - v[j];
            dist[j] = d;
            scanned[j] = 0;
            pred_src[j] = (unsigned short)r;
        }
        if (tid == 0) {
            dsrc[r] = 0.0;
            list[0] = r;
            s_nreached = 1;
        }
        __syncthreads();
        double D;
        int jstar;
        for (int step = 0;; ++step) {
            lsap_argmin(dist, scanned, M, D, jstar, s_val[step & 1], s_idx[step & 1]);
            const int i = sink_src[jstar] == LSAP_NONE ? -1 : (int)sink_src[jstar];  // only changes in the augmentation
            if (tid == 0) scanned[jstar] = 1;
            if (i < 0) break;
            if (tid == 0) {
                dsrc[i] = D;
                list[s_nreached++] = i;
            }
            const double base = D - u[i];
            __syncthreads();  // scanned[jstar] visible before the relaxation reads it
            for (int j = tid; j < M; j += LSAP_THREADS) {
                if (!scanned[j]) {
                    const double nd = base + cost(i, j) - v[j];
                    if (nd < dist[j]) {
                        dist[j] = nd;
                        pred_src[j] = (unsigned short)i;
                    }
                }
            }
            __syncthreads();
        }
        __syncthreads();
        const int nreached = s_nreached;
        for (int k = tid; k < nreached; k += LSAP_THREADS) {
            const int i = list[k];
            u[i] += D - dsrc[i];
        }
        for (int j = tid; j < M; j += LSAP_THREADS)
            if (scanned[j]) v[j] -= D - dist[j];
        __syncthreads();
        if (tid == 0) {  // flip the assignments along the path back to r
            int j = jstar;
            while (true) {
                const int i = pred_src[j];
                const int prev = src_sink[i];
                sink_src[j] = (unsigned short)i;
                src_sink[i] = j;
                if (i == r) break;
                j = prev;
            }
        }
        __syncthreads();
    }

    double acc = 0.0;
    for (int i = tid; i < T; i += LSAP_THREADS) {
        const int j = src_sink[i];
        const int r = rows_are_sources ? src_id[i] : sink_id[j];
        const int c = rows_are_sources ? sink_id[j] : src_id[i];
        out[r] = c;
        acc += (double)S[(int64_t)r * Ccols + c];
    }
    acc = warp_sum(acc);
    __syncthreads();
    if ((tid & 31) == 0) s_val[0][tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < LSAP_THREADS / 32; ++w) total += s_val[0][w];
        objective[e] = total;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Near-square problems (the 5-shot forward matching, 1374 fg support rows x 1369 query patches): the per-source
// Dijkstra above needs ~n searches of up to n steps each and runs at scipy's speed.  The transport solver's schedule
// (emd.cu) is the better fit: a PHASE grows one shortest-path forest from ALL unmatched sources and augments along the
// tree path of every free sink it settles; a WAVE settles every unscanned sink at the current minimum distance at
// once.  With unit supplies the flow is a matching (two index arrays instead of flow lists) and a tree path is dead as
// soon as one of its sources lies on a path augmented earlier in the phase (`used`).  Rows are the sources and columns
// the sinks (a source relaxes a contiguous row of `sim`); the smaller side is padded with zero-cost dummy nodes to a
// square problem, so both sides are equality-constrained and every dual may move freely - which nodes end up on the
// dummies is exactly the rectangular optimum.  The dispatcher only takes this path when the padding is small.
#ifdef MARSB200_LSQ_PROFILE
#define LQ_TIC() long long lq_last = clock64(), lq0 = 0, lq1 = 0, lq2 = 0, lq3 = 0, lq4 = 0, lq5 = 0; int lq_ph = 0, lq_wv = 0
#define LQ_LAP(x) do { const long long t = clock64(); x += t - lq_last; lq_last = t; } while (0)
#define LQ_PRINT() do { if (threadIdx.x == 0 && blockIdx.x == 0) printf("lsq n=%d phases %d waves %d | init %lld argmin %lld settle %lld augment %lld relax %lld dual %lld\n", n, lq_ph, lq_wv, lq0, lq1, lq2, lq3, lq4, lq5); } while (0)
#else
#define LQ_TIC() do { } while (0)
#define LQ_LAP(x) do { } while (0)
#define LQ_PRINT() do { } while (0)
#endif
constexpr int LSQ_THREADS = 512;  // measured at 1374 x 1369: 256 threads 76 ms, 512 threads 51 ms, 1024 threads 58 ms
constexpr unsigned short LSQ_NONE = 0xffff;

constexpr int LSQ_SLOTS = 4;  // sinks per thread, kept in registers: n <= LSQ_SLOTS * LSQ_THREADS = 2048

__host__ __device__ inline size_t lsq_smem_bytes(int n_cap) { return (size_t)n_cap * (3 * 8 + 7 * 2 + 2) + 64; }

// A thread owns the sinks tid, tid + 512, ... and keeps their distance, dual, column offset and scanned flag in REGISTERS for
// the whole solve (the arg-min, the settle pass and the relaxation touch shared memory only for what other threads read:
// the predecessor of a sink, the matching, the sources' duals).  One CTA per SM by design: no register cap.
__global__ void __launch_bounds__(LSQ_THREADS, 1) lsap_square_kernel(const float* __restrict__ sim, const uint8_t* __restrict__ row_sel,
                                                                      const uint8_t* __restrict__ col_sel, int R, int Ccols,
                                                                      int maximize, int n_cap, int32_t* __restrict__ row_to_col,
                                                                      double* __restrict__ objective, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char lsq_smem_raw[];
    double* u = reinterpret_cast<double*>(lsq_smem_raw);  // [n] source duals
    double* dsrc = u + n_cap;                              // [n] distance at which a source was reached
    double* v0 = dsrc + n_cap;                             // [n] sink duals after the column reduction (read by the row reduction only)
    unsigned short* row_id = reinterpret_cast<unsigned short*>(v0 + n_cap);  // [n] row of source i (dummy beyond nr)
    unsigned short* col_id = row_id + n_cap;               // [n] column of sink j (dummy beyond nc)
    unsigned short* src_sink = col_id + n_cap;             // [n] sink matched to source i, LSQ_NONE = unmatched
    unsigned short* sink_src = src_sink + n_cap;           // [n] source matched to sink j, LSQ_NONE = free
    unsigned short* pred_src = sink_src + n_cap;           // [n] source that gave sink j its distance
    unsigned short* newlist = pred_src + n_cap;            // [n] sources reached in the current wave (roots at a phase start)
    unsigned short* freelist = newlist + n_cap;            // [n] free sinks settled in the current wave
    unsigned char* reached = reinterpret_cast<unsigned char*>(freelist + n_cap);  // [n]
    unsigned char* used = reached + n_cap;                 // [n] source lies on a path augmented in this phase
    __shared__ double s_red[LSQ_THREADS / 32];
    __shared__ int s_cnt[LSQ_THREADS / 32];
    __shared__ int s_nnew, s_nfree, s_left, s_roots_left, s_free_unscanned;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t e = blockIdx.x;
    const float* S = sim + e * (int64_t)R * Ccols;
    int32_t* out = row_to_col + e * R;

    // ---- selected rows / columns in ascending order (ordered block compaction)
    for (int r = tid; r < R; r += LSQ_THREADS) out[r] = -1;
    auto compact = [&](int n, const uint8_t* sel, unsigned short* ids) -> int {
        int base = 0;
        for (int k0 = 0; k0 < n; k0 += LSQ_THREADS) {
            const int k = k0 + tid;
            const bool p = k < n && (!sel || sel[k]);
            const unsigned bal = __ballot_sync(0xffffffffu, p);
            if (lane == 0) s_cnt[warp] = __popc(bal);
            __syncthreads();
            int off = base, total = base;
            for (int w = 0; w < LSQ_THREADS / 32; ++w) {
                if (w < warp) off += s_cnt[w];
                total += s_cnt[w];
            }
            if (p) {
                const int pos = off + __popc(bal & ((1u << lane) - 1u));
                if (pos < n_cap) ids[pos] = (unsigned short)k;
            }
            base = total;
            __syncthreads();
        }
        return base;
    };
    const int nr = compact(R, row_sel ? row_sel + e * R : nullptr, row_id);
    const int nc = compact(Ccols, col_sel ? col_sel + e * Ccols : nullptr, col_id);
    const int n = max(nr, nc);
    if (nr == 0 || nc == 0) {
        if (tid == 0) objective[e] = 0.0;
        return;
    }
    if (n > n_cap || n > LSQ_SLOTS * LSQ_THREADS) {
        if (tid == 0) {
            objective[e] = nan("");
            atomicMax(status, n);
        }
        return;
    }
    const float sign = maximize ? -1.f : 1.f;

    // ---- the thread's own sinks: column pointer (null on a dummy sink), dual, distance, flags
    const float* colp[LSQ_SLOTS];
    double rv[LSQ_SLOTS], rd[LSQ_SLOTS];
    bool exist[LSQ_SLOTS], scanned[LSQ_SLOTS];
#pragma unroll
    for (int t = 0; t < LSQ_SLOTS; ++t) {
        const int j = tid + t * LSQ_THREADS;
        exist[t] = j < n;
        colp[t] = (exist[t] && j < nc) ? S + col_id[j] : nullptr;
        scanned[t] = false;
        rd[t] = 0.0;
        // column reduction: v_j = min_i c_ij (0 on / through a dummy)
        double mn = nr < n ? 0.0 : 1e300;
        if (colp[t] != nullptr) {
            int i = 0;
            for (; i + 4 <= nr; i += 4) {
                const float c0 = colp[t][(int64_t)row_id[i] * Ccols], c1 = colp[t][(int64_t)row_id[i + 1] * Ccols];
                const float c2 = colp[t][(int64_t)row_id[i + 2] * Ccols], c3 = colp[t][(int64_t)row_id[i + 3] * Ccols];
                mn = fmin(fmin(mn, (double)(sign * c0)), fmin((double)(sign * c1), fmin((double)(sign * c2), (double)(sign * c3))));
            }
            for (; i < nr; ++i) mn = fmin(mn, (double)(sign * colp[t][(int64_t)row_id[i] * Ccols]));
        } else {
            mn = 0.0;
        }
        rv[t] = mn;
        if (exist[t]) {
            v0[j] = mn;
            sink_src[j] = LSQ_NONE;
        }
    }
    __syncthreads();
    // row reduction: u_i = min_j (c_ij - v_j) (a warp per source, contiguous row): every row and column has a tight arc
    for (int i = warp; i < n; i += LSQ_THREADS / 32) {
        double mn = 1e300;
        if (i < nr) {
            const float* row = S + (int64_t)row_id[i] * Ccols;
            for (int j = lane; j < n; j += 32) mn = fmin(mn, (j < nc ? (double)(sign * row[col_id[j]]) : 0.0) - v0[j]);
        } else {
            for (int j = lane; j < n; j += 32) mn = fmin(mn, -v0[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (lane == 0) {
            u[i] = mn;
            src_sink[i] = LSQ_NONE;
        }
    }
    if (tid == 0) s_left = n;
    __syncthreads();

    // relax the thread's unscanned sinks against `cnt` listed sources reached at distance dsrc[.]
    auto relax = [&](const unsigned short* list, int cnt) {
        bool live[LSQ_SLOTS];
        int best_i[LSQ_SLOTS];
        bool any = false;
#pragma unroll
        for (int t = 0; t < LSQ_SLOTS; ++t) {
            live[t] = exist[t] && !scanned[t];
            best_i[t] = -1;
            any |= live[t];
        }
        if (!any) return;
        int k = 0;
        for (; k + 2 <= cnt; k += 2) {  // two sources x four sinks of independent gathers in flight
            const int ia = list[k], ib = list[k + 1];
            const double base_a = dsrc[ia] - u[ia], base_b = dsrc[ib] - u[ib];
            const int64_t ra = ia < nr ? (int64_t)row_id[ia] * Ccols : -1, rb = ib < nr ? (int64_t)row_id[ib] * Ccols : -1;
            float ca[LSQ_SLOTS], cb[LSQ_SLOTS];
#pragma unroll
            for (int t = 0; t < LSQ_SLOTS; ++t) {
                ca[t] = (live[t] && colp[t] != nullptr && ra >= 0) ? sign * colp[t][ra] : 0.f;
                cb[t] = (live[t] && colp[t] != nullptr && rb >= 0) ? sign * colp[t][rb] : 0.f;
            }
#pragma unroll
            for (int t = 0; t < LSQ_SLOTS; ++t) {
                const double da = (base_a + (double)ca[t]) - rv[t];
                if (live[t] && da < rd[t]) {
                    rd[t] = da;
                    best_i[t] = ia;
                }
                const double db = (base_b + (double)cb[t]) - rv[t];
                if (live[t] && db < rd[t]) {
                    rd[t] = db;
                    best_i[t] = ib;
                }
            }
        }
        for (; k < cnt; ++k) {
            const int i = list[k];
            const double base = dsrc[i] - u[i];
            const int64_t roff = i < nr ? (int64_t)row_id[i] * Ccols : -1;
            float cv[LSQ_SLOTS];
#pragma unroll
            for (int t = 0; t < LSQ_SLOTS; ++t) cv[t] = (live[t] && colp[t] != nullptr && roff >= 0) ? sign * colp[t][roff] : 0.f;
#pragma unroll
            for (int t = 0; t < LSQ_SLOTS; ++t) {
                const double d = (base + (double)cv[t]) - rv[t];
                if (live[t] && d < rd[t]) {
                    rd[t] = d;
                    best_i[t] = i;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < LSQ_SLOTS; ++t)
            if (best_i[t] >= 0) pred_src[tid + t * LSQ_THREADS] = (unsigned short)best_i[t];
    };

    LQ_TIC();
    while (s_left > 0) {  // uniform: shared state only changes between barriers
        // ---- phase start: every unmatched source is a root at distance 0
#ifdef MARSB200_LSQ_PROFILE
        lq_ph++;
#endif
        if (tid == 0) {
            s_nnew = 0;
            s_nfree = 0;
            s_free_unscanned = 0;
        }
        __syncthreads();
        int free_local = 0;
        for (int i = tid; i < n; i += LSQ_THREADS) {
            const bool root = src_sink[i] == LSQ_NONE;
            reached[i] = root ? 1 : 0;
            used[i] = 0;
            dsrc[i] = 0.0;
            if (root) newlist[atomicAdd(&s_nnew, 1)] = (unsigned short)i;
            free_local += sink_src[i] == LSQ_NONE ? 1 : 0;  // (sinks: same index range)
        }
#pragma unroll
        for (int t = 0; t < LSQ_SLOTS; ++t) {
            scanned[t] = false;
            rd[t] = 1e300;
        }
        free_local = warp_sum(free_local);
        if (lane == 0 && free_local) atomicAdd(&s_free_unscanned, free_local);
        __syncthreads();
        const int nroots = s_nnew;
        relax(newlist, nroots);
        __syncthreads();  // everyone has read the root list
        if (tid == 0) {
            s_roots_left = nroots;
            s_nnew = 0;
        }
        __syncthreads();

        LQ_LAP(lq0);
        double D = 0.0;
        while (true) {
#ifdef MARSB200_LSQ_PROFILE
            lq_wv++;
#endif
            // ---- wave: settle every unscanned sink at the minimum distance
            double mn = 1e300;
#pragma unroll
            for (int t = 0; t < LSQ_SLOTS; ++t)
                if (exist[t] && !scanned[t]) mn = fmin(mn, rd[t]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            if (lane == 0) s_red[warp] = mn;
            __syncthreads();
            mn = s_red[lane & (LSQ_THREADS / 32 - 1)];
#pragma unroll
            for (int o = LSQ_THREADS / 64; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            const double dmin = mn;
            LQ_LAP(lq1);
            if (dmin >= 1e300) break;  // every sink is scanned
            D = dmin;
#pragma unroll
            for (int t = 0; t < LSQ_SLOTS; ++t)
                if (exist[t] && !scanned[t] && rd[t] == dmin) {
                    scanned[t] = true;
                    const int j = tid + t * LSQ_THREADS;
                    const unsigned short i = sink_src[j];
                    if (i == LSQ_NONE) {
                        freelist[atomicAdd(&s_nfree, 1)] = (unsigned short)j;
                    } else {  // the matched source becomes reachable through the tight backward arc
                        reached[i] = 1;
                        dsrc[i] = dmin;
                        newlist[atomicAdd(&s_nnew, 1)] = i;
                    }
                }
            __syncthreads();
            LQ_LAP(lq2);
            const int nfree = s_nfree;
            if (nfree > 0) {
                if (tid == 0) {
                    // ---- augment along the tree path of every settled free sink (sequential: the paths must be
                    // vertex-disjoint); a path through a source used earlier in this phase, or ending in a spent root, is dead
                    s_free_unscanned -= nfree;
                    for (int b = 0; b < nfree; ++b) {
                        const int j = freelist[b];
                        int i = pred_src[j];
                        bool ok = true;
                        while (true) {
                            if (used[i]) {
                                ok = false;
                                break;
                            }
                            const unsigned short m = src_sink[i];
                            if (m == LSQ_NONE) break;  // an unmatched, unused source: the root
                            i = pred_src[m];
                        }
                        if (!ok) continue;
                        int jc = j;
                        i = pred_src[jc];
                        while (true) {
                            const unsigned short prev = src_sink[i];
                            src_sink[i] = (unsigned short)jc;
                            sink_src[jc] = (unsigned short)i;
                            used[i] = 1;
                            if (prev == LSQ_NONE) break;
                            jc = prev;
                            i = pred_src[jc];
                        }
                        s_left -= 1;
                        s_roots_left -= 1;
                    }
                    s_nfree = 0;
                }
                __syncthreads();
                // the phase ends when everything is matched, no root is left to augment from, or no unscanned sink is free
                if (s_left <= 0 || s_roots_left <= 0 || s_free_unscanned <= 0) break;
            }
            LQ_LAP(lq3);
            const int nnew = s_nnew;
            if (nnew > 0) relax(newlist, nnew);
            __syncthreads();
            if (tid == 0) s_nnew = 0;
            LQ_LAP(lq4);
        }
        // ---- dual update: keeps every matched arc tight and all reduced costs non-negative
        for (int i = tid; i < n; i += LSQ_THREADS)
            if (reached[i]) u[i] += D - dsrc[i];
#pragma unroll
        for (int t = 0; t < LSQ_SLOTS; ++t)
            if (scanned[t]) rv[t] -= D - rd[t];
        __syncthreads();
        LQ_LAP(lq5);
    }
    LQ_PRINT();

    double acc = 0.0;
    for (int i = tid; i < nr; i += LSQ_THREADS) {
        const int j = src_sink[i];
        if (j < nc) {
            out[row_id[i]] = col_id[j];
            acc += (double)S[(int64_t)row_id[i] * Ccols + col_id[j]];
        }
    }
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < LSQ_THREADS / 32; ++w) total += s_red[w];
        objective[e] = total;
    }
}

}  // namespace marsb200

using namespace marsb200;

extern "C" int marsb200_lsap(const float* sim, const uint8_t* row_sel, const uint8_t* col_sel, int E, int R, int C,
                             int maximize, int t_cap, int m_cap, int32_t* row_to_col, double* objective, int32_t* status,
                             void* stream) {
    MARS_REQUIRE(sim && row_to_col && objective && status, "null pointer");
    MARS_REQUIRE(E > 0 && R > 0 && C > 0 && R <= 65534 && C <= 65534, "shape (R, C <= 65534)");
    if (t_cap <= 0 || m_cap <= 0) {  // no bound on the selected counts given: size for the whole matrix
        t_cap = R < C ? R : C;
        m_cap = R < C ? C : R;
    }
    cudaStream_t s = as_stream(stream);
    MARS_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    // near-square problems take the multi-source phase solver on the zero-padded square (see lsap_square_kernel)
    const bool near_square = m_cap >= 64 && m_cap - t_cap <= std::max(8, m_cap / 16) && m_cap <= LSQ_SLOTS * LSQ_THREADS &&
                             lsq_smem_bytes(m_cap) <= 220 * 1024;
    if (near_square) {
        const size_t smem = lsq_smem_bytes(m_cap);
        MARS_CUDA_OK(cudaFuncSetAttribute(lsap_square_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lsap_square_kernel<<<E, LSQ_THREADS, smem, s>>>(sim, row_sel, col_sel, R, C, maximize, m_cap, row_to_col, objective, status);
        MARS_LAUNCH_OK();
        return MARSB200_OK;
    }
    const size_t smem = lsap_smem_bytes(t_cap, m_cap);
    MARS_REQUIRE(smem <= 220 * 1024, "problem too large for the shared-memory state (28*min + 23*max bytes <= 220 KB)");
    MARS_CUDA_OK(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsap_kernel<<<E, LSAP_THREADS, smem, s>>>(sim, row_sel, col_sel, R, C, maximize, t_cap, m_cap, row_to_col, objective,
                                               status);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}
