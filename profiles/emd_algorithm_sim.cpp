// CPU simulator of the transport-LP solvers considered for emd.cu: counts phases / waves / relax work.
// g++ -O2 -o sim sim.cpp && ./sim ep40.bin
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace std;
typedef long long ll;

struct Stats { ll phases = 0, waves = 0, relax = 0, augments = 0, settled = 0; };

// variant flags
static int EARLY = 0;      // stop a phase when no unscanned sink has open demand / no root supply left
static int SCALE_BITS = 0; // 0 = double costs; else integer costs c * 2^bits

// Primal-dual with multi-source phases and equal-distance waves (the emd.cu algorithm).  Returns EMD.
double solve_pd(const vector<double>& c, int T, int M, Stats& st) {
    // sources i in [0,T) supply M; sinks j in [0,M) demand T
    vector<double> u(T, 0.0), v(M, 0.0), dist(M), dsrc(T);
    vector<int> supply(T, M), demand(M, T), pred_src(M), pred_sink(T), reached(T), scanned(M);
    vector<vector<pair<int, int>>> flow(M);  // per sink: (src, f)
    ll left = (ll)T * M;
    // init duals
    for (int j = 0; j < M; ++j) { double b = 1e300; for (int i = 0; i < T; ++i) b = min(b, c[(size_t)i * M + j]); v[j] = b; }
    for (int i = 0; i < T; ++i) { double b = 1e300; for (int j = 0; j < M; ++j) b = min(b, c[(size_t)i * M + j] - v[j]); u[i] = b; }
    auto flow_of = [&](int j, int i) -> int* { for (auto& p : flow[j]) if (p.first == i) return &p.second; return nullptr; };
    while (left > 0) {
        st.phases++;
        vector<int> news;
        for (int i = 0; i < T; ++i) { reached[i] = supply[i] > 0; dsrc[i] = 0; pred_sink[i] = -1; if (reached[i]) news.push_back(i); }
        for (int j = 0; j < M; ++j) { scanned[j] = 0; dist[j] = 1e300; pred_src[j] = -1; }
        for (int i : news) { st.relax += M; for (int j = 0; j < M; ++j) { double d = c[(size_t)i * M + j] - u[i] - v[j]; if (d < dist[j]) { dist[j] = d; pred_src[j] = i; } } }
        vector<int> capflow(T, 0);
        double D = 0;
        ll open_unscanned = 0; for (int j = 0; j < M; ++j) open_unscanned += demand[j] > 0;
        while (true) {
            double dmin = 1e300; for (int j = 0; j < M; ++j) if (!scanned[j]) dmin = min(dmin, dist[j]);
            if (dmin >= 1e300) break;
            st.waves++;
            D = dmin;
            vector<int> batch, newsrc;
            for (int j = 0; j < M; ++j) if (!scanned[j] && dist[j] == dmin) { scanned[j] = 1; st.settled++; batch.push_back(j); }
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
            if (EARLY) {
                ll rootsup = 0; // any root supply left that could reach?
                if (open_unscanned == 0) break;
                (void)rootsup;
            }
            for (int i : newsrc) { st.relax += M; for (int j = 0; j < M; ++j) if (!scanned[j]) { double d = dmin + c[(size_t)i * M + j] - u[i] - v[j]; if (d < dist[j]) { dist[j] = d; pred_src[j] = i; } } }
        }
        for (int i = 0; i < T; ++i) if (reached[i]) u[i] += D - dsrc[i];
        for (int j = 0; j < M; ++j) if (scanned[j]) v[j] -= D - dist[j];
    }
    double acc = 0;
    for (int j = 0; j < M; ++j) for (auto& p : flow[j]) acc += (double)p.second * c[(size_t)p.first * M + j];
    return acc / ((double)T * M);
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "ep40.bin";
    EARLY = argc > 2 ? atoi(argv[2]) : 0;
    int maxlp = argc > 3 ? atoi(argv[3]) : 32;
    FILE* f = fopen(path, "rb");
    int hdr[3]; fread(hdr, 4, 3, f);
    int R = hdr[0], N = hdr[1], P = hdr[2];
    vector<float> cost((size_t)R * N); fread(cost.data(), 4, cost.size(), f);
    vector<uint8_t> sup(R); fread(sup.data(), 1, R, f);
    vector<uint8_t> pooled((size_t)P * N); fread(pooled.data(), 1, pooled.size(), f);
    fclose(f);
    vector<int> rows; for (int r = 0; r < R; ++r) if (sup[r]) rows.push_back(r);
    int T0 = rows.size();
    Stats tot;
    for (int p = 0; p < P && p < maxlp; ++p) {
        vector<int> cols; for (int j = 0; j < N; ++j) if (pooled[(size_t)p * N + j]) cols.push_back(j);
        int M0 = cols.size();
        bool swapped = 3 * M0 < T0;
        int T = swapped ? M0 : T0, M = swapped ? T0 : M0;
        vector<double> c((size_t)T * M);
        for (int i = 0; i < T; ++i) for (int j = 0; j < M; ++j) c[(size_t)i * M + j] = swapped ? cost[(size_t)rows[j] * N + cols[i]] : cost[(size_t)rows[i] * N + cols[j]];
        Stats st;
        double emd = solve_pd(c, T, M, st);
        printf("lp %3d T=%4d M=%4d emd=%.12f phases=%lld waves=%lld relax/TM=%.1f aug=%lld\n", p, T, M, emd, st.phases, st.waves, (double)st.relax / ((double)T * M), st.augments);
        tot.phases += st.phases; tot.waves += st.waves; tot.relax += st.relax; tot.augments += st.augments;
    }
    printf("TOTAL phases=%lld waves=%lld relax=%lld aug=%lld\n", tot.phases, tot.waves, tot.relax, tot.augments);
}
