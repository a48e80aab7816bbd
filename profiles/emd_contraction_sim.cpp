// Simulator: primal-dual phases, then contraction of the flow forest into supernodes for the tail.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <vector>
using namespace std;
typedef long long ll;

struct Stats { ll phases = 0, waves = 0, relax = 0, augments = 0, K = 0, small_phases = 0, small_waves = 0, viol = 0, contracted = 0, pivots = 0, cutwork = 0, neg0 = 0, negsum = 0; };
static int SWITCH_PHASE = 1000000, SWITCH_ROOTS = 0;

struct State {
    int T, M; const vector<double>* c;
    vector<double> u, v; vector<int> supply, demand; vector<vector<pair<int,int>>> flow; ll left;
};


struct Stats;
static void repair(State& s, Stats& st);
// solve the supernode transshipment by SSP; returns false if infeasible
// w: K x K arc costs (>= 0 under zero potentials), b: imbalance (sum 0).  Output f (K x K).
static void small_ssp(int K, const vector<double>& w, vector<ll> b, vector<ll>& f, vector<double>& pi, Stats& st) {
    f.assign((size_t)K * K, 0); pi.assign(K, 0.0);
    ll left = 0; for (int a = 0; a < K; ++a) if (b[a] > 0) left += b[a];
    while (left > 0) {
        st.small_phases++;
        // multi-source Dijkstra on reduced costs w(a,b) + pi[a] - pi[b]; residual arcs: forward any (a,b), backward (b,a) if f(a,b) > 0 with cost -w
        vector<double> d(K, 1e300); vector<int> pred(K, -1), ptype(K, 0), done(K, 0);
        for (int a = 0; a < K; ++a) if (b[a] > 0) d[a] = 0;
        int target = -1;
        while (true) {
            int x = -1; double dm = 1e300; for (int a = 0; a < K; ++a) if (!done[a] && d[a] < dm) { dm = d[a]; x = a; }
            if (x < 0) break;
            done[x] = 1; st.small_waves++;
            if (b[x] < 0) { target = x; break; }
            for (int y = 0; y < K; ++y) if (!done[y] && y != x) {
                double rc = w[(size_t)x * K + y] + pi[x] - pi[y];
                if (dm + rc < d[y]) { d[y] = dm + rc; pred[y] = x; ptype[y] = 0; }
                if (f[(size_t)y * K + x] > 0) { double rb = -w[(size_t)y * K + x] + pi[x] - pi[y]; if (dm + rb < d[y]) { d[y] = dm + rb; pred[y] = x; ptype[y] = 1; } }
            }
        }
        if (target < 0) { fprintf(stderr, "small infeasible\n"); exit(1); }
        double D = d[target];
        for (int a = 0; a < K; ++a) if (done[a]) pi[a] += d[a] - D;  // keeps reduced costs >= 0
        // augment
        ll delta = -b[target]; int x = target;
        while (pred[x] >= 0) { int p = pred[x]; if (ptype[x] == 1) delta = min(delta, f[(size_t)x * K + p]); x = p; }
        delta = min(delta, b[x]);
        b[x] -= delta; b[target] += delta; left -= delta;
        x = target;
        while (pred[x] >= 0) { int p = pred[x]; if (ptype[x] == 0) f[(size_t)p * K + x] += delta; else f[(size_t)x * K + p] -= delta; x = p; }
    }
}


// dual network simplex on the flow forest: drive negative flows out
static void repair(State& s, Stats& st) {
    int T = s.T, M = s.M, n = T + M; const vector<double>& c = *s.c;
    // check dual feasibility + tightness
    double worst = 0, tight = 0;
    for (int i = 0; i < T; ++i) for (int j = 0; j < M; ++j) worst = min(worst, c[(size_t)i * M + j] - s.u[i] - s.v[j]);
    for (int j = 0; j < M; ++j) for (auto& p : s.flow[j]) tight = max(tight, fabs(c[(size_t)p.first * M + j] - s.u[p.first] - s.v[j]));
    if (worst < -1e-12 || tight > 1e-12) { fprintf(stderr, "dual infeasible %g tight %g\n", worst, tight); }
    int nneg0 = 0; ll negsum = 0; for (int j = 0; j < M; ++j) for (auto& p : s.flow[j]) if (p.second < 0) { nneg0++; negsum -= p.second; }
    st.neg0 += nneg0; st.negsum += negsum;
    while (true) {
        // most negative arc
        int bj = -1, bk = -1, bf = 0;
        for (int j = 0; j < M; ++j) for (size_t k = 0; k < s.flow[j].size(); ++k) if (s.flow[j][k].second < bf) { bf = s.flow[j][k].second; bj = j; bk = (int)k; }
        if (bj < 0) break;
        st.pivots++;
        int li = s.flow[bj][bk].first, lj = bj;
        // adjacency without the leaving arc
        vector<vector<int>> adj(n);
        for (int j = 0; j < M; ++j) for (auto& p : s.flow[j]) { if (j == lj && p.first == li) continue; adj[p.first].push_back(T + j); adj[T + j].push_back(p.first); }
        vector<int> inX(n, 0), parent(n, -1), stack{li}; inX[li] = 1;
        while (!stack.empty()) { int x = stack.back(); stack.pop_back(); for (int y : adj[x]) if (!inX[y]) { inX[y] = 1; parent[y] = x; stack.push_back(y); } }
        if (inX[T + lj]) { fprintf(stderr, "cycle in support\n"); exit(1); }
        // entering arc: source in Y (component of lj... any node not in X that is connected? use all not-in-X), sink in X
        double best = 1e300; int ei = -1, ej = -1; ll cut = 0;
        for (int i = 0; i < T; ++i) if (!inX[i]) for (int j = 0; j < M; ++j) if (inX[T + j]) { cut++; double rc = c[(size_t)i * M + j] - s.u[i] - s.v[j]; if (rc < best) { best = rc; ei = i; ej = j; } }
        st.cutwork += cut;
        if (ei < 0) { fprintf(stderr, "no entering arc\n"); exit(1); }
        for (int i = 0; i < T; ++i) if (inX[i]) s.u[i] -= best;
        for (int j = 0; j < M; ++j) if (inX[T + j]) s.v[j] += best;
        // push d = -bf around: entering arc ei->ej gets +d; path in X from ej to li; leaving arc removed; path in Y from lj to ei.
        int d = -bf;
        // generic: set leaving flow to 0 (remove), add entering with d, then rebalance both trees by path updates.
        s.flow[lj].erase(s.flow[lj].begin() + bk);
        auto path_update = [&](int from, int to, const vector<int>& par_) {
            // walk from 'from' up to 'to' using parents (tree rooted at 'to'); flow moves from 'from' towards 'to'
            int x = from;
            while (x != to) { int p = par_[x];
                // moving d units from x to p
                if (x < T) { /* source x -> sink p: forward +d */ for (auto& q : s.flow[p - T]) if (q.first == x) { q.second += d; break; } }
                else { /* sink x -> source p: reduce flow p->x by d */ for (auto& q : s.flow[x - T]) if (q.first == p) { q.second -= d; break; } }
                x = p; }
        };
        // X side: rooted at li (parent[] from DFS).  Flow d arrives at sink ej (in X) and must reach li (which lost its negative outflow: li was sending bf<0, i.e. receiving d)
        // balance at li: before it had outflow bf on the leaving arc (i.e. net inflow d from lj). Now that is gone, so li needs d from elsewhere: from ej through the tree.
        path_update(T + ej, li, parent);
        // Y side: root at lj
        vector<int> parY(n, -1), seen(n, 0); stack = {T + lj}; seen[T + lj] = 1;
        while (!stack.empty()) { int x = stack.back(); stack.pop_back(); for (int y : adj[x]) if (!seen[y]) { seen[y] = 1; parY[y] = x; stack.push_back(y); } }
        if (!seen[ei]) { fprintf(stderr, "entering source not in Y tree\n"); exit(1); }
        // lj was sending d to li via negative arc (i.e. lj received bf = -d... ) now lj has d surplus inflow missing: lj must send d to ei?  lj previously "returned" d to li; now it must push d towards ei
        // moving d from lj to ei: walk from ei up to lj reversing direction: equivalent to moving -d from ei to lj
        d = -d; path_update(ei, T + lj, parY); d = -d;
        s.flow[ej].push_back({ei, d});
        // drop zero arcs? keep (degenerate) - they stay in the tree
        if (st.pivots > 100000) { fprintf(stderr, "pivot guard\n"); exit(1); }
    }
    // verify balances
    vector<ll> out(T, 0), in(M, 0);
    for (int j = 0; j < M; ++j) for (auto& p : s.flow[j]) { out[p.first] += p.second; in[j] += p.second; }
    for (int i = 0; i < T; ++i) if (out[i] != M) { fprintf(stderr, "bad supply %d %lld\n", i, out[i]); exit(1); }
    for (int j = 0; j < M; ++j) if (in[j] != T) { fprintf(stderr, "bad demand\n"); exit(1); }
}

// returns true if the contraction finished the LP without violating a capacity
static bool contract_finish(State& s, Stats& st) {
    int T = s.T, M = s.M; const vector<double>& c = *s.c;
    int n = T + M;
    vector<int> par(n); iota(par.begin(), par.end(), 0);
    function<int(int)> find = [&](int x) { while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; } return x; };
    for (int j = 0; j < M; ++j) for (auto& p : s.flow[j]) par[find(p.first)] = find(T + j);
    vector<int> comp(n, -1); int K = 0;
    for (int x = 0; x < n; ++x) { int r = find(x); if (comp[r] < 0) comp[r] = K++; comp[x] = comp[r]; }
    st.K += K; st.contracted++;
    vector<ll> b(K, 0);
    for (int i = 0; i < T; ++i) b[comp[i]] += s.supply[i];
    for (int j = 0; j < M; ++j) b[comp[T + j]] -= s.demand[j];
    vector<double> w((size_t)K * K, 1e300); vector<int> wi((size_t)K * K, -1), wj((size_t)K * K, -1);
    st.relax += (ll)T * M;
    for (int i = 0; i < T; ++i) for (int j = 0; j < M; ++j) {
        int a = comp[i], bb = comp[T + j]; if (a == bb) continue;
        double rc = c[(size_t)i * M + j] - s.u[i] - s.v[j];
        size_t k = (size_t)a * K + bb;
        if (rc < w[k]) { w[k] = rc; wi[k] = i; wj[k] = j; }
    }
    vector<ll> f; vector<double> pi;
    small_ssp(K, w, b, f, pi, st);
    // expand: new arcs, then tree rebalancing
    vector<ll> surplus(n, 0);
    for (int i = 0; i < T; ++i) surplus[i] = s.supply[i];
    for (int j = 0; j < M; ++j) surplus[T + j] = -s.demand[j];
    vector<pair<pair<int,int>, ll>> newarcs;
    for (int a = 0; a < K; ++a) for (int bb = 0; bb < K; ++bb) { ll x = f[(size_t)a * K + bb]; if (x > 0) { int i = wi[(size_t)a * K + bb], j = wj[(size_t)a * K + bb]; surplus[i] -= x; surplus[T + j] += x; newarcs.push_back({{i, j}, x}); } }
    // forest adjacency
    vector<vector<int>> adj(n);
    for (int j = 0; j < M; ++j) for (auto& p : s.flow[j]) { adj[p.first].push_back(T + j); adj[T + j].push_back(p.first); }
    vector<int> parent(n, -2), order; order.reserve(n);
    for (int r = 0; r < n; ++r) if (parent[r] == -2) {
        parent[r] = -1; size_t head = order.size(); order.push_back(r);
        while (head < order.size()) { int x = order[head++]; for (int y : adj[x]) if (parent[y] == -2) { parent[y] = x; order.push_back(y); } }
    }
    vector<ll> sub = surplus;
    bool ok = true;
    vector<vector<pair<int,int>>> nf = s.flow;
    for (int k = n - 1; k >= 0; --k) {
        int x = order[k], p = parent[x]; if (p < 0) { if (sub[x] != 0) { fprintf(stderr, "imbalance %lld\n", sub[x]); exit(1); } continue; }
        ll E = sub[x]; sub[p] += E;
        if (E == 0) continue;
        int src = x < T ? x : p, snk = (x < T ? p : x) - T;
        ll dlt = x < T ? E : -E;
        for (auto& q : nf[snk]) if (q.first == src) { q.second += (int)dlt; if (q.second < 0) ok = false; break; }
    }
    for (auto& a : newarcs) nf[a.first.second].push_back({a.first.first, (int)a.second});
    s.flow = nf; s.left = 0;
    for (int i = 0; i < T; ++i) s.u[i] -= pi[comp[i]];
    for (int j = 0; j < M; ++j) s.v[j] += pi[comp[T + j]];
    if (!ok) { st.viol++; repair(s, st); }
    return true;
}

double solve_pd(const vector<double>& c, int T, int M, Stats& st) {
    State s; s.T = T; s.M = M; s.c = &c;
    vector<double>& u = s.u; vector<double>& v = s.v; u.assign(T, 0.0); v.assign(M, 0.0);
    vector<double> dist(M), dsrc(T);
    vector<int>& supply = s.supply; vector<int>& demand = s.demand; supply.assign(T, M); demand.assign(M, T);
    vector<int> pred_src(M), pred_sink(T), reached(T), scanned(M);
    vector<vector<pair<int, int>>>& flow = s.flow; flow.assign(M, {});
    ll& left = s.left; left = (ll)T * M;
    for (int j = 0; j < M; ++j) { double b = 1e300; for (int i = 0; i < T; ++i) b = min(b, c[(size_t)i * M + j]); v[j] = b; }
    for (int i = 0; i < T; ++i) { double b = 1e300; for (int j = 0; j < M; ++j) b = min(b, c[(size_t)i * M + j] - v[j]); u[i] = b; }
    auto flow_of = [&](int j, int i) -> int* { for (auto& p : flow[j]) if (p.first == i) return &p.second; return nullptr; };
    bool tried = false;
    while (left > 0) {
        int nroots = 0, nopen = 0; for (int i = 0; i < T; ++i) nroots += supply[i] > 0; for (int j = 0; j < M; ++j) nopen += demand[j] > 0;
        if (!tried && (st.phases >= SWITCH_PHASE || (nroots <= SWITCH_ROOTS && nopen <= SWITCH_ROOTS))) {
            tried = true;
            if (getenv("VERBOSE")) printf("   switch at phase %lld roots=%d open=%d left=%lld\n", st.phases, nroots, nopen, left);
            if (contract_finish(s, st)) break;
        }
        st.phases++;
        vector<int> news;
        for (int i = 0; i < T; ++i) { reached[i] = supply[i] > 0; dsrc[i] = 0; pred_sink[i] = -1; if (reached[i]) news.push_back(i); }
        for (int j = 0; j < M; ++j) { scanned[j] = 0; dist[j] = 1e300; pred_src[j] = -1; }
        for (int i : news) { st.relax += M; for (int j = 0; j < M; ++j) { double d = c[(size_t)i * M + j] - u[i] - v[j]; if (d < dist[j]) { dist[j] = d; pred_src[j] = i; } } }
        vector<int> capflow(T, 0);
        double D = 0;
        ll open_unscanned = 0; for (int j = 0; j < M; ++j) open_unscanned += demand[j] > 0;
        ll w0 = st.waves, l0 = left;
        while (true) {
            double dmin = 1e300; for (int j = 0; j < M; ++j) if (!scanned[j]) dmin = min(dmin, dist[j]);
            if (dmin >= 1e300) break;
            st.waves++;
            D = dmin;
            vector<int> batch, newsrc;
            for (int j = 0; j < M; ++j) if (!scanned[j] && dist[j] == dmin) { scanned[j] = 1; batch.push_back(j); }
            auto expand = [&](int j) { for (auto& p : flow[j]) { int i = p.first; if (!reached[i]) { reached[i] = 1; dsrc[i] = dmin; pred_sink[i] = j; capflow[i] = p.second; newsrc.push_back(i); } } };
            for (int j : batch) if (demand[j] == 0) expand(j);
            for (int j : batch) if (demand[j] > 0) {
                open_unscanned--;
                int delta = demand[j]; int i = pred_src[j];
                while (pred_sink[i] >= 0) { delta = min(delta, capflow[i]); i = pred_src[pred_sink[i]]; }
                delta = min(delta, supply[i]);
                if (delta > 0) {
                    st.augments++;
                    supply[i] -= delta; demand[j] -= delta; left -= delta;
                    int jj = j;
                    while (true) {
                        int src = pred_src[jj];
                        int* f = flow_of(jj, src);
                        if (f) *f += delta; else flow[jj].push_back({src, delta});
                        int jp = pred_sink[src];
                        if (jp < 0) break;
                        capflow[src] -= delta;
                        for (size_t k = 0; k < flow[jp].size(); ++k) if (flow[jp][k].first == src) { flow[jp][k].second -= delta; if (flow[jp][k].second == 0) { flow[jp].erase(flow[jp].begin() + k); } break; }
                        jj = jp;
                    }
                }
                expand(j);
            }
            if (left <= 0) break;
            if (open_unscanned == 0) break;
            for (int i : newsrc) { st.relax += M; for (int j = 0; j < M; ++j) if (!scanned[j]) { double d = dmin + c[(size_t)i * M + j] - u[i] - v[j]; if (d < dist[j]) { dist[j] = d; pred_src[j] = i; } } }
        }
        for (int i = 0; i < T; ++i) if (reached[i]) u[i] += D - dsrc[i];
        for (int j = 0; j < M; ++j) if (scanned[j]) v[j] -= D - dist[j];
        if (getenv("VERBOSE")) printf("   phase %lld roots=%d open=%d waves=%lld moved=%lld left=%lld\n", st.phases, nroots, nopen, st.waves - w0, l0 - left, left);
    }
    double acc = 0;
    for (int j = 0; j < M; ++j) for (auto& p : flow[j]) acc += (double)p.second * c[(size_t)p.first * M + j];
    return acc / ((double)T * M);
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "ep40.bin";
    SWITCH_PHASE = argc > 2 ? atoi(argv[2]) : 1000000;
    SWITCH_ROOTS = argc > 3 ? atoi(argv[3]) : 0;
    int maxlp = argc > 4 ? atoi(argv[4]) : 32;
    FILE* f = fopen(path, "rb");
    int hdr[3]; if (fread(hdr, 4, 3, f) != 3) return 1;
    int R = hdr[0], N = hdr[1], P = hdr[2];
    vector<float> cost((size_t)R * N); if (fread(cost.data(), 4, cost.size(), f) != cost.size()) return 1;
    vector<uint8_t> sup(R); if (fread(sup.data(), 1, R, f) != (size_t)R) return 1;
    vector<uint8_t> pooled((size_t)P * N); if (fread(pooled.data(), 1, pooled.size(), f) != pooled.size()) return 1;
    fclose(f);
    vector<int> rows; for (int r = 0; r < R; ++r) if (sup[r]) rows.push_back(r);
    int T0 = rows.size();
    Stats tot, tot0;
    for (int p = 0; p < P && p < maxlp; ++p) {
        vector<int> cols; for (int j = 0; j < N; ++j) if (pooled[(size_t)p * N + j]) cols.push_back(j);
        int M0 = cols.size();
        bool swapped = 3 * M0 < T0;
        int T = swapped ? M0 : T0, M = swapped ? T0 : M0;
        vector<double> c((size_t)T * M);
        for (int i = 0; i < T; ++i) for (int j = 0; j < M; ++j) c[(size_t)i * M + j] = swapped ? cost[(size_t)rows[j] * N + cols[i]] : cost[(size_t)rows[i] * N + cols[j]];
        Stats st, st0;
        int sp = SWITCH_PHASE, sr = SWITCH_ROOTS;
        double emd = solve_pd(c, T, M, st);
        SWITCH_PHASE = 1000000; SWITCH_ROOTS = 0;
        char* vb = getenv("VERBOSE"); if (vb) unsetenv("VERBOSE");
        double emd0 = solve_pd(c, T, M, st0);
        if (vb) setenv("VERBOSE", "1", 1);
        SWITCH_PHASE = sp; SWITCH_ROOTS = sr;
        printf("lp %3d T=%4d M=%4d err=%.2e | exact: phases=%lld waves=%lld relax/TM=%.1f | contract: phases=%lld waves=%lld relax/TM=%.1f K=%lld small_phases=%lld small_waves=%lld viol=%lld neg0=%lld negsum=%lld pivots=%lld cut/TM=%.1f\n", p, T, M, emd - emd0,
               st0.phases, st0.waves, (double)st0.relax / ((double)T * M), st.phases, st.waves, (double)st.relax / ((double)T * M), st.K, st.small_phases, st.small_waves, st.viol, st.neg0, st.negsum, st.pivots, (double)st.cutwork/((double)T*M));
        tot.pivots += st.pivots; tot.cutwork += st.cutwork; tot.neg0 += st.neg0;
        tot.phases += st.phases; tot.waves += st.waves; tot.relax += st.relax; tot.viol += st.viol; tot.small_phases += st.small_phases; tot.small_waves += st.small_waves; tot.K += st.K;
        tot0.phases += st0.phases; tot0.waves += st0.waves; tot0.relax += st0.relax;
    }
    printf("TOTAL exact phases=%lld waves=%lld relax=%lld | contract phases=%lld waves=%lld relax=%lld viol=%lld small_phases=%lld small_waves=%lld K=%lld neg0=%lld pivots=%lld cutwork=%lld\n", tot0.phases, tot0.waves, tot0.relax, tot.phases, tot.waves, tot.relax, tot.viol, tot.small_phases, tot.small_waves, tot.K, tot.neg0, tot.pivots, tot.cutwork);
}
