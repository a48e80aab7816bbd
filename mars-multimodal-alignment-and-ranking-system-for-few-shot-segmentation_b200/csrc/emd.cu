// Exact earth mover's distance between the pooled support mask and every pooled proposal
// (SURVEY.md row A7 / 8f-1): the transportation LP the reference solves with POT's ot.emd2
// (FilteringMergingModule.py:142-169, matcher/Matcher.py:1187-1194), uniform marginals 1/T and 1/M,
// cost = C[fg support rows][proposal patches] in float64.
//
// One CTA per (episode, proposal), LPs handed out largest first from a device-side queue.  Flows are kept in
// integer units of 1/(T*M) (supplies M, demands T), so the primal solution is exact.  The solver is the
// primal-dual (Hungarian-type) method for the transportation problem on dense reduced costs, organised for a
// CTA instead of a scalar core:
//   * a PHASE grows a shortest-path forest from ALL sources that still have supply (multi-source Dijkstra over
//     the sinks) and augments along the tree path of every sink with open demand it settles - a few hundred
//     phases instead of a Dijkstra per augmentation;
//   * a WAVE settles every unscanned sink whose tentative distance EQUALS the current minimum at once.  Tree
//     arcs of earlier phases have reduced cost exactly 0, so ties are the rule: 5-8x fewer sequential steps than
//     one sink per step, with no tolerance involved (only exactly equal keys are merged);
//   * the flow is a set of per-sink linked lists in shared memory (a basic solution has at most T + M - 1
//     positive arcs), so settling a saturated sink, the capacity walk and the augmentation never touch global
//     memory; the only global traffic is the gather of cost entries c(i, j) = C[rows[i]][cols[j]] from the
//     episode's L2-resident cost matrix when new sources relax the unscanned sinks.
// The optimal value is unique, so the result matches any exact LP solver up to float64 rounding.
#include <algorithm>

#include "common.cuh"

namespace marsb200 {

#ifdef MARSB200_EMD_PROFILE
__device__ long long g_emd_prof[16];
#define EP_TIC() const long long ep_t0 = clock64()
#define EP_LAP(slot) do { const long long ep_t1 = clock64(); ep_acc[slot] += ep_t1 - ep_last; ep_last = ep_t1; } while (0)
#define EP_COUNT(slot) do { ep_acc[slot] += 1; } while (0)
#else
#define EP_LAP(slot) do { } while (0)
#define EP_COUNT(slot) do { } while (0)
#endif

#ifndef MARSB200_EMD_THREADS
#define MARSB200_EMD_THREADS 256
#endif
constexpr int EMD_THREADS = MARSB200_EMD_THREADS;
constexpr int EMD_MAX_CTAS = 1024 / EMD_THREADS;  // resident CTAs per SM the register file allows at 64 registers per thread
constexpr int EMD_WARPS = EMD_THREADS / 32;
constexpr double EMD_INF = 1e300;
constexpr int EMD_GLOBAL_CTAS = 148;  // CTAs (and state slabs) of the global-state launch

__host__ __device__ inline int emd_pool_nodes(int t_cap, int n_cap) { return 2 * (t_cap + n_cap) + 64; }

struct EmdSmem {
    double* u;       // [t_cap] source duals
    double* dsrc;    // [t_cap] distance at which a source was reached in this phase
    double* v;       // [n_cap] sink duals
    double* dist;    // [n_cap] tentative / final sink distances
    int* reached;    // [t_cap] 0/1 (int: claimed with atomicExch)
    int* soff;       // [t_cap] offset of source i in the cost matrix (row * N, or the patch column when transposed)
    int* koff;       // [n_cap] offset of sink j (patch column, or row * N when transposed): c(i, j) = C[soff[i] + koff[j]]
    short* supply;   // [t_cap]
    short* capflow;  // [t_cap] flow on the tree arc (pred_sink[i] -> i) = what a path through i can take back
    short* pred_sink;  // [t_cap] sink through which source i was reached, -1 for a root
    short* newlist;  // [t_cap] sources reached in the current wave (also the root list while a phase starts)
    short* demand;   // [n_cap]
    short* pred_src;  // [n_cap] source that gave sink j its distance
    short* batch;    // [n_cap] sinks settled in the current wave
    short* head;     // [n_cap] first flow node of sink j, -1 = none
    short* node_src;   // [pool] flow nodes: source, flow, next node of the same sink (or next free node)
    short* node_flow;  // [pool]
    short* node_next;  // [pool]
    unsigned char* scanned;  // [n_cap]
};

__host__ __device__ inline size_t emd_smem_bytes(int t_cap, int n_cap) {
    return (size_t)t_cap * (2 * 8 + 2 * 4 + 4 * 2) + (size_t)n_cap * (2 * 8 + 4 + 4 * 2 + 1) + (size_t)emd_pool_nodes(t_cap, n_cap) * 6 + 64;
}

__device__ inline EmdSmem emd_carve(unsigned char* base, int t_cap, int n_cap) {
    EmdSmem s;
    const int pool = emd_pool_nodes(t_cap, n_cap);
    double* d = reinterpret_cast<double*>(base);
    s.u = d;
    s.dsrc = s.u + t_cap;
    s.v = s.dsrc + t_cap;
    s.dist = s.v + n_cap;
    s.reached = reinterpret_cast<int*>(s.dist + n_cap);
    s.soff = s.reached + t_cap;
    s.koff = s.soff + t_cap;
    short* h = reinterpret_cast<short*>(s.koff + n_cap);
    s.supply = h;
    s.capflow = s.supply + t_cap;
    s.pred_sink = s.capflow + t_cap;
    s.newlist = s.pred_sink + t_cap;
    s.demand = s.newlist + t_cap;
    s.pred_src = s.demand + n_cap;
    s.batch = s.pred_src + n_cap;
    s.head = s.batch + n_cap;
    s.node_src = s.head + n_cap;
    s.node_flow = s.node_src + pool;
    s.node_next = s.node_flow + pool;
    s.scanned = reinterpret_cast<unsigned char*>(s.node_next + pool);
    return s;
}

// ordered compaction of the indices k in [0, n) with pred(k) into out[0..cap); returns the total count
template <typename Pred>
__device__ inline int block_compact(int n, int cap, short* out, Pred pred, int* s_warp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += EMD_THREADS) {
        const int k = k0 + tid;
        const bool p = k < n && pred(k);
        const unsigned bal = __ballot_sync(0xffffffffu, p);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = base, total = base;
        for (int w = 0; w < EMD_WARPS; ++w) {
            if (w < warp) off += s_warp[w];
            total += s_warp[w];
        }
        if (p) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            if (pos < cap) out[pos] = (short)k;
        }
        base = total;
        __syncthreads();
    }
    return base;
}

// order-preserving map double -> uint64 (handles the odd -1e-17 a rounded reduced cost can produce)
__device__ __forceinline__ unsigned long long dkey(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// block-wide minimum of dist over the unscanned sinks; every thread gets the result
__device__ inline double block_min_key(const EmdSmem& s, int M, unsigned long long* s_val) {
    double v = EMD_INF;
    for (int j = threadIdx.x; j < M; j += EMD_THREADS)
        if (!s.scanned[j]) v = fmin(v, s.dist[j]);
    // warp minimum with two integer reductions (high word, then low word among the lanes that hold the minimum high word)
    const unsigned long long k = dkey(v);
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    if ((threadIdx.x & 31) == 0) s_val[threadIdx.x >> 5] = ((unsigned long long)mhi << 32) | mlo;
    __syncthreads();
    unsigned long long best = s_val[0];
#pragma unroll
    for (int w = 1; w < EMD_WARPS; ++w) best = min(best, s_val[w]);
    return dkey_inv(best);
}

// ---- sizes of every LP and their processing order (largest T*M first: the launch ends with the small ones)
__global__ void __launch_bounds__(256) emd_sizes_kernel(const uint8_t* __restrict__ row_fg, const uint32_t* __restrict__ pooled,
                                                         int P, int64_t m_rows, int npw, int64_t total,
                                                         int64_t* __restrict__ key) {
    const int lane = threadIdx.x & 31;
    const int64_t lp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (lp >= total) return;
    const int64_t e = lp / P;
    int t = 0, m = 0;
    for (int64_t r = lane; r < m_rows; r += 32) t += row_fg[e * m_rows + r] ? 1 : 0;
    for (int w = lane; w < npw; w += 32) m += __popc(pooled[lp * npw + w]);
    t = warp_sum(t);
    m = warp_sum(m);
    if (lane == 0) key[lp] = (int64_t)t * m;
}

// Proposals of one episode with the same pooled bitmap are the same LP (same support rows, same columns): only the
// first one is solved, the others copy its value.  SAM proposals that differ by a few pixels pool to identical
// bitmaps all the time.  One warp per LP compares against the earlier LPs of the episode with the same size key.
__global__ void __launch_bounds__(256) emd_dedupe_kernel(const uint32_t* __restrict__ pooled, int P, int npw, int64_t total,
                                                          const int64_t* __restrict__ key, int32_t* __restrict__ dup_of) {
    const int lane = threadIdx.x & 31;
    const int64_t lp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (lp >= total) return;
    const int64_t first = lp / P * P;
    const int64_t mine = key[lp];
    int32_t found = -1;
    for (int64_t q = first; q < lp && found < 0; ++q) {
        if (key[q] != mine) continue;
        bool same = true;
        for (int w = lane; w < npw; w += 32) same &= pooled[lp * npw + w] == pooled[q * npw + w];
        if (__all_sync(0xffffffffu, same)) found = (int32_t)q;
    }
    if (lane == 0) dup_of[lp] = found;
}

// duplicates sort last and are skipped by the solver: their key becomes negative after every warp has read the keys
__global__ void __launch_bounds__(256) emd_mark_dups_kernel(int64_t total, const int32_t* __restrict__ dup_of,
                                                             int64_t* __restrict__ key) {
    const int64_t lp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lp < total && dup_of[lp] >= 0) key[lp] = -1;
}

__global__ void __launch_bounds__(256) emd_copy_dups_kernel(int64_t total, const int32_t* __restrict__ dup_of,
                                                             double* __restrict__ out) {
    const int64_t lp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lp < total && dup_of[lp] >= 0) out[lp] = out[dup_of[lp]];
}

__global__ void __launch_bounds__(256) emd_rank_kernel(const int64_t* __restrict__ key, int64_t total,
                                                        int32_t* __restrict__ order) {
    __shared__ int64_t s_key[256];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t mine = i < total ? key[i] : 0;
    int64_t rank = 0;
    for (int64_t k0 = 0; k0 < total; k0 += 256) {
        __syncthreads();
        s_key[threadIdx.x] = (k0 + threadIdx.x < total) ? key[k0 + threadIdx.x] : -1;
        __syncthreads();
        const int lim = (int)min((int64_t)256, total - k0);
        for (int k = 0; k < lim; ++k) {
            const int64_t other = s_key[k];
            rank += (other > mine || (other == mine && k0 + k < i)) ? 1 : 0;
        }
    }
    if (i < total) order[rank] = (int32_t)i;
}

// For every sink j (all of them when INIT, the unscanned ones otherwise): best = min over the listed sources i of
// c(i, j) - u[i], c(i, j) = cmat[soff[i] + koff[j]].  INIT starts a phase (dist = best - v, or v = best and dist = 0 in the very first phase);
// otherwise the sink is relaxed with dbase + best - v.  `lanes` (a power of two <= 32) threads share a sink and
// split the source list, so small problems still use the whole CTA.  A thread owns up to EMD_SPT sinks and walks
// the source list four at a time: up to 4 * EMD_SPT independent gathers are in flight per thread.
constexpr int EMD_SPT = 3;
constexpr int EMD_CHUNK = 4;

template <bool INIT>
__device__ inline void scan_sources(const EmdSmem& s, const float* __restrict__ cmat, int M, const short* list,
                                    int n, int lanes, double dbase, bool first_phase) {
    const int tid = threadIdx.x;
    const int sub = tid & (lanes - 1);
    const int groups = EMD_THREADS / lanes;
    // all threads of a group run the same trip count (shuffles below): iterate over the padded sink range
    const int m_pad = (M + groups * EMD_SPT - 1) / (groups * EMD_SPT) * (groups * EMD_SPT);
    for (int j0 = tid / lanes; j0 < m_pad; j0 += groups * EMD_SPT) {
        bool live[EMD_SPT];
        const float* col[EMD_SPT];
        double best[EMD_SPT];
        int best_i[EMD_SPT];
#pragma unroll
        for (int t = 0; t < EMD_SPT; ++t) {
            const int j = j0 + t * groups;
            live[t] = j < M && (INIT || !s.scanned[j]);
            best[t] = EMD_INF;
            best_i[t] = -1;
            col[t] = cmat + (live[t] ? s.koff[j] : 0);
        }
        bool any = false;
#pragma unroll
        for (int t = 0; t < EMD_SPT; ++t) any |= live[t];
        if (any) {
            // full chunks: EMD_CHUNK sources x EMD_SPT sinks of independent gathers in flight
            int k = sub;
            for (; k + (EMD_CHUNK - 1) * lanes < n; k += EMD_CHUNK * lanes) {
                int idx[EMD_CHUNK];
                int off[EMD_CHUNK];
                double ui[EMD_CHUNK];
#pragma unroll
                for (int q = 0; q < EMD_CHUNK; ++q) {
                    idx[q] = list[k + q * lanes];
                    off[q] = s.soff[idx[q]];
                    ui[q] = s.u[idx[q]];
                }
                float cv[EMD_SPT][EMD_CHUNK];
#pragma unroll
                for (int t = 0; t < EMD_SPT; ++t)
#pragma unroll
                    for (int q = 0; q < EMD_CHUNK; ++q) cv[t][q] = live[t] ? col[t][off[q]] : 0.f;
#pragma unroll
                for (int t = 0; t < EMD_SPT; ++t)
#pragma unroll
                    for (int q = 0; q < EMD_CHUNK; ++q) {
                        const double d = (double)cv[t][q] - ui[q];
                        if (d < best[t]) {
                            best[t] = d;
                            best_i[t] = idx[q];
                        }
                    }
            }
            // remainder: one source at a time, still EMD_SPT gathers in flight
            for (; k < n; k += lanes) {
                const int i = list[k];
                const int off = s.soff[i];
                const double ui = s.u[i];
                float cv[EMD_SPT];
#pragma unroll
                for (int t = 0; t < EMD_SPT; ++t) cv[t] = live[t] ? col[t][off] : 0.f;
#pragma unroll
                for (int t = 0; t < EMD_SPT; ++t) {
                    const double d = (double)cv[t] - ui;
                    if (d < best[t]) {
                        best[t] = d;
                        best_i[t] = i;
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < EMD_SPT; ++t) {
            for (int o = lanes >> 1; o > 0; o >>= 1) {  // groups are aligned sub-warps
                const double ob = __shfl_xor_sync(0xffffffffu, best[t], o);
                const int oi = __shfl_xor_sync(0xffffffffu, best_i[t], o);
                if (ob < best[t] || (ob == best[t] && oi >= 0 && (best_i[t] < 0 || oi < best_i[t]))) {
                    best[t] = ob;
                    best_i[t] = oi;
                }
            }
            const int j = j0 + t * groups;
            if (live[t] && sub == 0 && best_i[t] >= 0) {
                if (INIT) {
                    if (first_phase) {  // v_j = min_i c_ij: every sink starts with a tight arc
                        s.v[j] = best[t];
                        s.dist[j] = 0.0;
                    } else {
                        s.dist[j] = best[t] - s.v[j];
                    }
                    s.pred_src[j] = (short)best_i[t];
                } else {
                    const double nd = (dbase + best[t]) - s.v[j];
                    if (nd < s.dist[j]) {
                        s.dist[j] = nd;
                        s.pred_src[j] = (short)best_i[t];
                    }
                }
            }
        }
    }
}

// every source feeding sink j becomes reachable at distance d (tight backward arcs)
__device__ inline void expand_feeders(const EmdSmem& s, int j, double d, int* nnew) {
    for (int n = s.head[j]; n >= 0; n = s.node_next[n]) {
        const int i = s.node_src[n];
        if (atomicExch(&s.reached[i], 1) == 0) {
            s.dsrc[i] = d;
            s.pred_sink[i] = (short)j;
            s.capflow[i] = s.node_flow[n];
            s.newlist[atomicAdd(nnew, 1)] = (short)i;
        }
    }
}

// gstate == nullptr (GLOBAL_STATE false): the fast path, solver state in shared memory, sized by (t_cap, m_cap); a problem that exceeds the caps
// is appended to the overflow list instead of being solved.  GLOBAL_STATE = true: the same solver with its state in a
// per-CTA slab of global memory (L2-resident), sized for ANY problem of the episode shape (t_cap = all support rows,
// m_cap = N); it only walks the overflow list, so it costs one empty launch when nothing overflowed.  Together: no capacity
// limit below the 16-bit index range, like the reference's ot.emd2 (FilteringMergingModule.py:162-166).
__global__ void __launch_bounds__(EMD_THREADS, EMD_MAX_CTAS) emd_kernel(const float* __restrict__ cost, const uint8_t* __restrict__ row_fg,
                                                           const uint32_t* __restrict__ pooled, int P, int64_t m_rows,
                                                           int N, int npw, int t_cap, int m_cap, int total_lps,
                                                           const int32_t* __restrict__ order, int32_t* __restrict__ counter,
                                                           const int32_t* __restrict__ dup_of, double* __restrict__ out,
                                                           int* __restrict__ status, int32_t* __restrict__ ovf_count,
                                                           int32_t* __restrict__ ovf_list, unsigned char* __restrict__ gstate,
                                                           size_t gstate_stride) {
#define EMD_GLOBAL_STATE 0
#include "emd_solver_body.inc"
#undef EMD_GLOBAL_STATE
}

__global__ void __launch_bounds__(EMD_THREADS, 1) emd_global_state_kernel(const float* __restrict__ cost, const uint8_t* __restrict__ row_fg,
                                                           const uint32_t* __restrict__ pooled, int P, int64_t m_rows,
                                                           int N, int npw, int t_cap, int m_cap, int total_lps,
                                                           const int32_t* __restrict__ order, int32_t* __restrict__ counter,
                                                           const int32_t* __restrict__ dup_of, double* __restrict__ out,
                                                           int* __restrict__ status, int32_t* __restrict__ ovf_count,
                                                           int32_t* __restrict__ ovf_list, unsigned char* __restrict__ gstate,
                                                           size_t gstate_stride) {
#define EMD_GLOBAL_STATE 1
#include "emd_solver_body.inc"
#undef EMD_GLOBAL_STATE
}

}  // namespace marsb200

using namespace marsb200;

#ifdef MARSB200_EMD_PROFILE
extern "C" int marsb200_debug_emd_profile(long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_emd_prof, sizeof(long long) * 16);
    if (reset) {
        long long z[16] = {0};
        cudaMemcpyToSymbol(g_emd_prof, z, sizeof(z));
    }
    return 0;
}
#endif

extern "C" {

static int emd_ctas_per_sm(int t_cap, int m_cap) {  // 64 registers x 256 threads: at most 4 by the register file
    const size_t smem = emd_smem_bytes(t_cap, std::max(t_cap, m_cap));
    return (int)std::max<size_t>(1, std::min<size_t>(EMD_MAX_CTAS, (size_t)(227 * 1024) / (smem + 1024)));
}

constexpr size_t EMD_SMEM_LIMIT = 200 * 1024;
constexpr int EMD_INDEX_LIMIT = 32767;  // 16-bit node / row indices

// Largest t <= t_cap whose state (with m_cap sinks) fits the shared-memory fast path.
static int emd_fast_t_cap(int t_cap, int m_cap) {
    int t = std::max(1, t_cap);
    while (t > 1 && (emd_smem_bytes(t, std::max(t, m_cap)) > EMD_SMEM_LIMIT ||
                     emd_pool_nodes(t, std::max(t, m_cap)) > EMD_INDEX_LIMIT))
        t = t - std::max(1, t / 16);
    return t;
}
// Caps of the global-state launch: every problem of the episode shape, as far as 16-bit indices reach.
static int emd_full_t_cap(int64_t m_rows, int N) {
    int t = (int)std::min<int64_t>(m_rows, EMD_INDEX_LIMIT);
    while (t > 1 && emd_pool_nodes(t, std::max(t, N)) > EMD_INDEX_LIMIT) t = t - std::max(1, t / 64);
    return t;
}
static size_t emd_gstate_stride(int64_t m_rows, int N) {
    const int t = emd_full_t_cap(m_rows, N);
    return (emd_smem_bytes(t, std::max(t, N)) + 255) / 256 * 256;
}

// Workspace: size keys, processing order, duplicate links, overflow list, two queue heads + the overflow count, and one
// state slab per SM for the global-state launch (the fast path keeps flows and duals in shared memory).
int64_t marsb200_emd_workspace_bytes(int E, int P, int N, int64_t m_rows, int t_cap, int m_cap) {
    if (E <= 0 || P <= 0 || N <= 0 || m_rows <= 0) return 0;
    (void)t_cap;
    (void)m_cap;
    const int64_t lps = (int64_t)E * P;
    const int64_t head = (lps * 8 + lps * 4 + lps * 4 + lps * 4 + 256 + 255) / 256 * 256;
    return head + (int64_t)EMD_GLOBAL_CTAS * (int64_t)emd_gstate_stride(m_rows, N);
}

int marsb200_emd_scores(const float* cost, const uint8_t* row_fg, const uint32_t* pooled, int E, int P, int64_t m_rows,
                        int N, int t_cap, int m_cap, void* workspace, int64_t workspace_bytes, double* out,
                        int32_t* status, void* stream) {
    MARS_REQUIRE(cost && row_fg && pooled && workspace && out && status, "null pointer");
    MARS_REQUIRE(E > 0 && P > 0 && m_rows > 0 && N > 0 && N <= 32767 && m_rows <= 32767, "shape");
    MARS_REQUIRE((int64_t)E * P < (1ll << 31), "too many problems");
    if (m_cap <= 0 || m_cap > N) m_cap = N;
    if (t_cap <= 0 || t_cap > m_rows) t_cap = (int)m_rows;
    MARS_REQUIRE(workspace_bytes >= marsb200_emd_workspace_bytes(E, P, N, m_rows, t_cap, m_cap), "workspace too small");
    MARS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    MARS_REQUIRE(m_rows * (int64_t)N < (1ll << 31), "cost matrix too large for 32-bit offsets");
    // (t_cap, m_cap) size the shared-memory fast path; what does not fit it is solved by the global-state launch
    t_cap = emd_fast_t_cap(t_cap, m_cap);
    const size_t smem = emd_smem_bytes(t_cap, std::max(t_cap, m_cap));
    cudaStream_t s = as_stream(stream);
    int num_sms = 0;
    MARS_CUDA_OK(device_sm_count(&num_sms));
    MARS_CUDA_OK(cudaFuncSetAttribute(emd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t lps = (int64_t)E * P;
    const int npw = ceil_div(N, 32);
    int64_t* key = reinterpret_cast<int64_t*>(workspace);
    int32_t* order = reinterpret_cast<int32_t*>(key + lps);
    int32_t* dup_of = order + lps;
    int32_t* ovf_list = dup_of + lps;
    int32_t* counter = ovf_list + lps;  // [0] fast-path queue head, [1] overflow count, [2] global-state queue head
    const int64_t head = (lps * 8 + lps * 4 + lps * 4 + lps * 4 + 256 + 255) / 256 * 256;
    unsigned char* gstate = reinterpret_cast<unsigned char*>(workspace) + head;
    MARS_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    MARS_CUDA_OK(cudaMemsetAsync(counter, 0, 4 * sizeof(int32_t), s));
    emd_sizes_kernel<<<(unsigned)ceil_div64(lps * 32, 256), 256, 0, s>>>(row_fg, pooled, P, m_rows, npw, lps, key);
    MARS_LAUNCH_OK();
    emd_dedupe_kernel<<<(unsigned)ceil_div64(lps * 32, 256), 256, 0, s>>>(pooled, P, npw, lps, key, dup_of);
    MARS_LAUNCH_OK();
    emd_mark_dups_kernel<<<(unsigned)ceil_div64(lps, 256), 256, 0, s>>>(lps, dup_of, key);
    MARS_LAUNCH_OK();
    emd_rank_kernel<<<(unsigned)ceil_div64(lps, 256), 256, 0, s>>>(key, lps, order);
    MARS_LAUNCH_OK();
    const unsigned grid = (unsigned)std::min<int64_t>(lps, (int64_t)num_sms * emd_ctas_per_sm(t_cap, m_cap));
    emd_kernel<<<grid, EMD_THREADS, smem, s>>>(cost, row_fg, pooled, P, m_rows, N, npw, t_cap, m_cap, (int)lps, order,
                                                      counter, dup_of, out, status, counter + 1, ovf_list, nullptr, 0);
    MARS_LAUNCH_OK();
    // problems beyond the fast path's caps (none in the usual case: the launch finds an empty list and returns)
    const int t_full = emd_full_t_cap(m_rows, N);
    emd_global_state_kernel<<<EMD_GLOBAL_CTAS, EMD_THREADS, 0, s>>>(cost, row_fg, pooled, P, m_rows, N, npw, t_full, N, (int)lps, order,
                                                            counter + 2, dup_of, out, status, counter + 1, ovf_list, gstate,
                                                            emd_gstate_stride(m_rows, N));
    MARS_LAUNCH_OK();
    emd_copy_dups_kernel<<<(unsigned)ceil_div64(lps, 256), 256, 0, s>>>(lps, dup_of, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"
