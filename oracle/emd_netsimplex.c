/* TEST INFRASTRUCTURE (oracle/): a CPU restatement of the exact transport LP the reference solves with POT 0.9.4
 * `ot.emd2` (mars/components/FilteringMergingModule.py:160-166, matcher/Matcher.py:1187-1193).  POT is not vendored in the
 * reference tree and not installable here; its `emd2` is a NETWORK SIMPLEX (Bonneel's adaptation of LEMON's
 * NetworkSimplex) on the bipartite transportation graph.  This file restates that algorithm class for the one problem
 * family the path needs - uniform marginals 1/T over the sources and 1/M over the sinks, dense costs - as a primal
 * (transportation) simplex on a spanning-tree basis:
 *   - mass in integer units of 1/(T M (T + 1)): sources supply M (T + 1) + 1, sinks demand T (T + 1), the last sink
 *     T (T + 1) + T.  The "+ 1" terms are Orden's perturbation: no proper subset of sources and sinks balances, so every
 *     basic solution is non-degenerate, every pivot moves mass and strictly lowers the objective, and the method cannot
 *     cycle.  The perturbation is smaller than one original unit (T + 1 perturbed units), so the optimal basis of the
 *     perturbed problem is feasible - hence optimal - for the original supplies, which the final tree solve uses;
 *   - costs scaled to integers (2^40; float32 inputs >= 2^-17 are exact, smaller ones round by <= 2^-41), so reduced costs
 *     and the optimality test are exact integer arithmetic;
 *   - start basis by the row-minimum rule, block pricing (most negative reduced cost inside a block of arcs, blocks taken
 *     cyclically - LEMON's BLOCK_SEARCH rule), the tree (parents, depths, duals) rebuilt by one breadth-first pass per pivot.
 * Nothing of the product links or calls this file: only tests/ and bench.py's CPU-baseline leg do (through oracle/emd_c.py).
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC -> oracle/_build/libmarsoracle.so)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef long long ll;

#ifdef MARS_ORACLE_STATS
/* profiles/emd_netsimplex_stats.py: pivots, cycle arcs (sum, max), nodes cut off by the leaving arc (smaller side, sum),
 * arcs priced, depth of the cut (sum) - the per-pivot work a CTA-parallel version would have to organise */
ll mars_oracle_stats[8];
#endif

typedef struct {
    int T, M, n;
    const ll* c;   /* integer costs [T * M] */
    int* arc_i;    /* basic arcs, n - 1 of them: source index */
    int* arc_j;    /*                            sink index  */
    ll* flow;      /* perturbed units */
    int* parent;   /* tree rooted at node 0 (source 0); nodes: sources 0 .. T-1, sinks T .. T+M-1 */
    int* parc;     /* basic arc joining a node to its parent */
    int* depth;
    int* order;    /* breadth-first order of the nodes */
    ll* pot;       /* duals: u_i for sources, v_j for sinks, c_ij = u_i + v_j on basic arcs */
    int* adj_off;  /* CSR adjacency of the tree */
    int* adj_node;
    int* adj_arc;
    int* fill;
} Tree;

static void rebuild(Tree* t) {
    const int n = t->n, T = t->T;
    memset(t->adj_off, 0, sizeof(int) * (size_t)(n + 1));
    for (int k = 0; k < n - 1; ++k) {
        t->adj_off[t->arc_i[k] + 1]++;
        t->adj_off[T + t->arc_j[k] + 1]++;
    }
    for (int v = 0; v < n; ++v) t->adj_off[v + 1] += t->adj_off[v];
    memcpy(t->fill, t->adj_off, sizeof(int) * (size_t)n);
    for (int k = 0; k < n - 1; ++k) {
        const int a = t->arc_i[k], b = T + t->arc_j[k];
        t->adj_node[t->fill[a]] = b; t->adj_arc[t->fill[a]++] = k;
        t->adj_node[t->fill[b]] = a; t->adj_arc[t->fill[b]++] = k;
    }
    int head = 0, tail = 0;
    t->order[tail++] = 0;
    t->parent[0] = -1; t->parc[0] = -1; t->depth[0] = 0; t->pot[0] = 0;
    while (head < tail) {
        const int v = t->order[head++];
        for (int e = t->adj_off[v]; e < t->adj_off[v + 1]; ++e) {
            const int w = t->adj_node[e], k = t->adj_arc[e];
            if (w == t->parent[v]) continue;
            t->parent[w] = v; t->parc[w] = k; t->depth[w] = t->depth[v] + 1;
            t->pot[w] = t->c[(size_t)t->arc_i[k] * t->M + t->arc_j[k]] - t->pot[v];
            t->order[tail++] = w;
        }
    }
}

/* Returns 0 on success.  cost [T * M] row-major doubles in [0, 2^20); obj = optimal transport cost for marginals 1/T, 1/M,
 * summed with the caller's double costs; obj_int = the same optimum on the integer-scaled costs; pivots = basis changes. */
int mars_oracle_emd_netsimplex(const double* cost, int T, int M, double* obj, double* obj_int, ll* pivots) {
    if (T <= 0 || M <= 0) { if (obj) *obj = 0.0; if (obj_int) *obj_int = 0.0; if (pivots) *pivots = 0; return 0; }
    const int n = T + M;
    const size_t arcs = (size_t)T * M;
    const double scale = 1099511627776.0; /* 2^40 */
    ll* c = (ll*)malloc(sizeof(ll) * arcs);
    if (!c) return -1;
    for (size_t e = 0; e < arcs; ++e) {
        if (!(cost[e] >= 0.0) || cost[e] >= 1048576.0) { free(c); return -2; }
        c[e] = llround(cost[e] * scale);
    }
    Tree t;
    t.T = T; t.M = M; t.n = n; t.c = c;
    t.arc_i = (int*)malloc(sizeof(int) * n); t.arc_j = (int*)malloc(sizeof(int) * n); t.flow = (ll*)malloc(sizeof(ll) * n);
    t.parent = (int*)malloc(sizeof(int) * n); t.parc = (int*)malloc(sizeof(int) * n); t.depth = (int*)malloc(sizeof(int) * n);
    t.order = (int*)malloc(sizeof(int) * n); t.pot = (ll*)malloc(sizeof(ll) * n);
    t.adj_off = (int*)malloc(sizeof(int) * (n + 1)); t.adj_node = (int*)malloc(sizeof(int) * 2 * n);
    t.adj_arc = (int*)malloc(sizeof(int) * 2 * n); t.fill = (int*)malloc(sizeof(int) * n);
    ll* supply = (ll*)malloc(sizeof(ll) * T);
    ll* demand = (ll*)malloc(sizeof(ll) * M);
    int* path = (int*)malloc(sizeof(int) * 2 * n);  /* basic arcs of the cycle, from the entering arc's sink back to its source */
    int* up_i = (int*)malloc(sizeof(int) * n);
    int rc = 0;
    ll npiv = 0;

    /* ---- perturbed marginals and the row-minimum start basis */
    const ll K = (ll)T + 1;
    for (int i = 0; i < T; ++i) supply[i] = (ll)M * K + 1;
    for (int j = 0; j < M; ++j) demand[j] = (ll)T * K;
    demand[M - 1] += T;
    int nb = 0;
    for (int i = 0; i < T; ++i) {
        while (supply[i] > 0) {
            int best = -1;
            for (int j = 0; j < M; ++j)
                if (demand[j] > 0 && (best < 0 || c[(size_t)i * M + j] < c[(size_t)i * M + best])) best = j;
            if (best < 0 || nb >= n - 1) { rc = -3; goto done; }
            const ll x = supply[i] < demand[best] ? supply[i] : demand[best];
            t.arc_i[nb] = i; t.arc_j[nb] = best; t.flow[nb] = x; ++nb;
            supply[i] -= x; demand[best] -= x;
        }
    }
    if (nb != n - 1) { rc = -3; goto done; }
    rebuild(&t);

    /* ---- pivots: block pricing over the arcs e = i * M + j, taken cyclically */
    {
        size_t block = (size_t)sqrt((double)arcs);
        if (block < 64) block = 64;
        if (block > arcs) block = arcs;
        size_t next = 0, scanned_clean = 0;
        int ni = 0, nj = 0;  /* next = ni * M + nj */
        while (scanned_clean < arcs) {
            ll best_d = 0; size_t best_e = 0;
            size_t cnt = block < arcs - scanned_clean ? block : arcs - scanned_clean;
            for (size_t s = 0; s < cnt; ++s) {
                const ll d = c[next] - t.pot[ni] - t.pot[T + nj];
                if (d < best_d) { best_d = d; best_e = next; }
                ++next;
                if (++nj == M) { nj = 0; if (++ni == T) { ni = 0; next = 0; } }
            }
#ifdef MARS_ORACLE_STATS
            mars_oracle_stats[4] += (ll)cnt;
#endif
            if (best_d >= 0) { scanned_clean += cnt; continue; }
            scanned_clean = 0;
            const int ei = (int)(best_e / M), ej = (int)(best_e % M);
            /* cycle: tree path from the sink T + ej to the source ei; arcs alternate -, +, -, ... along it */
            int a = T + ej, b = ei, na = 0, nbk = 0;
            while (a != b) {
                if (t.depth[a] >= t.depth[b]) { path[na++] = t.parc[a]; a = t.parent[a]; }
                else { up_i[nbk++] = t.parc[b]; b = t.parent[b]; }
            }
            for (int k = nbk - 1; k >= 0; --k) path[na++] = up_i[k];
            if ((na & 1) == 0) { rc = -4; goto done; }  /* bipartite: the path has odd length */
            ll theta = -1; int leave = -1;
            for (int k = 0; k < na; k += 2)
                if (theta < 0 || t.flow[path[k]] < theta) { theta = t.flow[path[k]]; leave = path[k]; }
            if (theta <= 0) { rc = -5; goto done; }      /* non-degenerate by construction */
            for (int k = 0; k < na; ++k) t.flow[path[k]] += (k & 1) ? theta : -theta;
            if (t.flow[leave] != 0) { rc = -5; goto done; }
#ifdef MARS_ORACLE_STATS
            {
                int child = -1, size = 0;
                for (int v = 1; v < n; ++v) if (t.parc[v] == leave) child = v;
                for (int v = 0; v < n; ++v) {  /* nodes below the leaving arc */
                    int w = v;
                    while (w != -1 && w != child) w = t.parent[w];
                    size += (w == child);
                }
                mars_oracle_stats[0] += 1; mars_oracle_stats[1] += na + 1;
                if (na + 1 > mars_oracle_stats[2]) mars_oracle_stats[2] = na + 1;
                mars_oracle_stats[3] += size < n - size ? size : n - size;
                mars_oracle_stats[5] += t.depth[child];
            }
#endif
            t.arc_i[leave] = ei; t.arc_j[leave] = ej; t.flow[leave] = theta;
            rebuild(&t);
            ++npiv;
        }
    }

    /* ---- the optimal basis with the ORIGINAL marginals (supply M, demand T, units of 1 / (T M)): tree solve, leaves first */
    {
        ll* net = t.pot;  /* the duals are no longer needed */
        for (int v = 0; v < n; ++v) net[v] = v < T ? (ll)M : -(ll)T;
        long double sum = 0.0L;
        __int128 sum_int = 0;
        for (int q = n - 1; q >= 1; --q) {
            const int v = t.order[q], k = t.parc[v];
            const ll f = v < T ? net[v] : -net[v];  /* a source sends its subtree's surplus up, a sink draws its deficit */
            if (f < 0) { rc = -6; goto done; }
            net[t.parent[v]] += net[v];
            const size_t e = (size_t)t.arc_i[k] * M + t.arc_j[k];
            sum += (long double)f * (long double)cost[e];
            sum_int += (__int128)f * c[e];
        }
        if (net[0] != 0) { rc = -6; goto done; }
        const long double tm = (long double)T * (long double)M;
        if (obj) *obj = (double)(sum / tm);
        if (obj_int) *obj_int = (double)((long double)sum_int / (long double)scale / tm);
    }
done:
    if (pivots) *pivots = npiv;
    free(c); free(t.arc_i); free(t.arc_j); free(t.flow); free(t.parent); free(t.parc); free(t.depth); free(t.order);
    free(t.pot); free(t.adj_off); free(t.adj_node); free(t.adj_arc); free(t.fill); free(supply); free(demand); free(path);
    free(up_i);
    return rc;
}
