// Simulator: primal-dual phases with a forest kept across phases (only invalidated subtrees are regrown).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace std;
typedef long long ll;
struct Stats { ll phases = 0, waves = 0, relax = 0, augments = 0, rescan = 0, expandwalk = 0; };
static int PERSIST = 1;

double solve_pd(const vector<double>& c, int T, int M, Stats& st) {
    vector<double> u(T, 0.0), v(M, 0.0), dist(M), dsrc(T);
    vector<int> supply(T, M), demand(M, T), pred_src(M), pred_sink(T), reached(T, 0), scanned(M, 0), capflow(T, 0), valid(T, 0);
    vector<vector<pair<int, int>>> flow(M);
    ll left = (ll)T * M;
    for (int j = 0; j < M; ++j) { double b = 1e300; for (int i = 0; i < T; ++i) b = min(b, c[(size_t)i * M + j]); v[j] = b; }
    for (int i = 0; i < T; ++i) { double b = 1e300; for (int j = 0; j < M; ++j) b = min(b, c[(size_t)i * M + j] - v[j]); u[i] = b; }
    auto flow_of = [&](int j, int i) -> int* { for (auto& p : flow[j]) if (p.first == i) return &p.second; return nullptr; };
    bool have_forest = false; double Dprev = 0;
    while (left > 0) {
        st.phases++;
        vector<int> news;
        if (!have_forest || !PERSIST) {
            for (int i = 0; i < T; ++i) { reached[i] = supply[i] > 0; dsrc[i] = 0; pred_sink[i] = -1; if (reached[i]) news.push_back(i); }
            for (int j = 0; j < M; ++j) { scanned[j] = 0; dist[j] = 1e300; pred_src[j] = -1; }
            for (int i : news) { st.relax += M; for (int j = 0; j < M; ++j) { double d = c[(size_t)i * M + j] - u[i] - v[j]; if (d < dist[j]) { dist[j] = d; pred_src[j] = i; } } }
        } else {
            // validity of every reached source: walk up
            for (int i = 0; i < T; ++i) {
                bool ok = reached[i]; int x = i;
                while (ok) {
                    int jp = pred_sink[x];
                    if (jp < 0) { ok = supply[x] > 0; break; }
                    int* f = flow_of(jp, x); if (!f || *f <= 0) { ok = false; break; }
                    x = pred_src[jp];
                    if (!reached[x]) { ok = false; break; }
                }
                valid[i] = ok;
            }
            // a root that was not reached before cannot appear (supplies only shrink)
            for (int j = 0; j < M; ++j) {
                bool vs = scanned[j] && pred_src[j] >= 0 && valid[pred_src[j]];
                if (vs) { dist[j] = 0; }
                else {
                    bool keep = !scanned[j] && pred_src[j] >= 0 && valid[pred_src[j]];
                    if (keep) dist[j] -= Dprev;
                    else { // rescan over valid sources
                        dist[j] = 1e300; pred_src[j] = -1; st.rescan += T; st.relax += T;
                        for (int i = 0; i < T; ++i) if (valid[i]) { double d = c[(size_t)i * M + j] - u[i] - v[j]; if (d < dist[j]) { dist[j] = d; pred_src[j] = i; } }
                    }
                    scanned[j] = 0;
                }
            }
            for (int i = 0; i < T; ++i) { reached[i] = valid[i]; dsrc[i] = 0; if (valid[i] && pred_sink[i] >= 0) capflow[i] = *flow_of(pred_sink[i], i); }
            // valid scanned sinks re-expand feeders
            for (int j = 0; j < M; ++j) if (scanned[j]) for (auto& p : flow[j]) { st.expandwalk++; int i = p.first; if (!reached[i]) { reached[i] = 1; dsrc[i] = 0; pred_sink[i] = j; capflow[i] = p.second; news.push_back(i); } }
            for (int i : news) { st.relax += M; for (int j = 0; j < M; ++j) if (!scanned[j]) { double d = c[(size_t)i * M + j] - u[i] - v[j]; if (d < dist[j]) { dist[j] = d; pred_src[j] = i; } } }
        }
        double D = 0;
        ll open_unscanned = 0; for (int j = 0; j < M; ++j) open_unscanned += (!scanned[j] && demand[j] > 0);
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
        Dprev = D; have_forest = true;
    }
    double acc = 0;
    for (int j = 0; j < M; ++j) for (auto& p : flow[j]) acc += (double)p.second * c[(size_t)p.first * M + j];
    return acc / ((double)T * M);
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "ep40.bin";
    int maxlp = argc > 2 ? atoi(argv[2]) : 32;
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
        PERSIST = 1; double emd = solve_pd(c, T, M, st);
        PERSIST = 0; double emd0 = solve_pd(c, T, M, st0);
        printf("lp %3d T=%4d M=%4d err=%.2e | restart: phases=%lld waves=%lld relax/TM=%.1f | persist: phases=%lld waves=%lld relax/TM=%.1f (rescan/TM=%.1f)\n", p, T, M, emd - emd0,
               st0.phases, st0.waves, (double)st0.relax / ((double)T * M), st.phases, st.waves, (double)st.relax / ((double)T * M), (double)st.rescan / ((double)T * M));
        tot.phases += st.phases; tot.waves += st.waves; tot.relax += st.relax;
        tot0.phases += st0.phases; tot0.waves += st0.waves; tot0.relax += st0.relax;
    }
    printf("TOTAL restart phases=%lld waves=%lld relax=%lld | persist phases=%lld waves=%lld relax=%lld\n", tot0.phases, tot0.waves, tot0.relax, tot.phases, tot.waves, tot.relax);
}
