// Simulator: synchronous epsilon-scaling push-relabel (auction-like) for the uniform-marginal transport LP.
// Counts rounds and dense-scan work.  g++ -O2 -o sim_pr sim_pr.cpp && ./sim_pr ep40.bin
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace std;
typedef long long ll;

struct Res { double cost; ll rounds, scans, relabels; double gap; };

// sources i: supply M units; sinks j: demand T units.  rc(i,j) = c - u_i - v_j.
// eps-optimal: rc >= -eps for all arcs; rc <= eps on arcs with flow.
Res solve(const vector<double>& c, int T, int M, double eps_final, double alpha) {
    vector<double> u(T, 0.0), v(M, 0.0);
    vector<vector<pair<int,int>>> flow(M);  // per sink: (source, units)
    double cmax = 0; for (double x : c) cmax = max(cmax, x);
    for (int j = 0; j < M; ++j) { double b = 1e300; for (int i = 0; i < T; ++i) b = min(b, c[(size_t)i*M+j]); v[j] = b; }
    Res r{0, 0, 0, 0, 0};
    double eps = cmax / 4;
    vector<int> ex_src(T), ex_snk(M);
    while (true) {
        // refine(eps): drop flow on arcs violating eps-CS (rc > eps with flow): simplest: reset all flow
        for (int j = 0; j < M; ++j) flow[j].clear();
        for (int i = 0; i < T; ++i) ex_src[i] = M;
        for (int j = 0; j < M; ++j) ex_snk[j] = -T;
        // make all arcs satisfy rc >= -eps: lower u_i so that min_j rc(i,j) >= 0 (keeps feasibility direction)
        for (int i = 0; i < T; ++i) { double mn = 1e300; for (int j = 0; j < M; ++j) mn = min(mn, c[(size_t)i*M+j] - u[i] - v[j]); if (mn < 0) u[i] += mn; }
        ll guard = 0;
        while (true) {
            bool any = false;
            // ---- round A: every source with excess pushes to its best sink or relabels
            vector<int> bestj(T, -1); vector<double> best(T), second(T);
            for (int i = 0; i < T; ++i) if (ex_src[i] > 0) {
                any = true; r.scans += M;
                double b1 = 1e300, b2 = 1e300; int j1 = -1;
                for (int j = 0; j < M; ++j) { double rc = c[(size_t)i*M+j] - u[i] - v[j]; if (rc < b1) { b2 = b1; b1 = rc; j1 = j; } else if (rc < b2) b2 = rc; }
                bestj[i] = j1; best[i] = b1; second[i] = b2;
            }
            for (int i = 0; i < T; ++i) if (ex_src[i] > 0) {
                if (best[i] >= 0) { u[i] += best[i] + eps; r.relabels++; }   // relabel: best arc becomes admissible (rc = -eps)
                // push everything along the best arc (infinite capacity)
                int j = bestj[i]; int d = ex_src[i];
                bool found = false; for (auto& p : flow[j]) if (p.first == i) { p.second += d; found = true; break; }
                if (!found) flow[j].push_back({i, d});
                ex_src[i] = 0; ex_snk[j] += d;
            }
            // ---- round B: every sink with excess pushes back along arcs with the largest rc (least wanted), relabelling
            for (int j = 0; j < M; ++j) while (ex_snk[j] > 0) {
                any = true;
                // arcs with flow: reverse admissible if rc(i,j) > 0
                int bi = -1; double brc = -1e300;
                for (size_t k = 0; k < flow[j].size(); ++k) { double rc = c[(size_t)flow[j][k].first*M+j] - u[flow[j][k].first] - v[j]; if (rc > brc) { brc = rc; bi = (int)k; } }
                if (brc <= 0) { v[j] -= (-brc) + eps; r.relabels++; continue; }  // relabel: lower v_j so the worst arc becomes admissible
                int d = min(ex_snk[j], flow[j][bi].second);
                flow[j][bi].second -= d; ex_snk[j] -= d; ex_src[flow[j][bi].first] += d;
                if (flow[j][bi].second == 0) flow[j].erase(flow[j].begin() + bi);
            }
            if (!any) break;
            r.rounds++;
            if (++guard > 2000000) { printf("  [guard]\n"); break; }
        }
        if (eps <= eps_final) break;
        eps = max(eps / alpha, eps_final);
    }
    double acc = 0; for (int j = 0; j < M; ++j) for (auto& p : flow[j]) acc += (double)p.second * c[(size_t)p.first*M+j];
    r.cost = acc / ((double)T * M);
    // dual bound with feasible duals: v_j' = min_i (c - u_i)
    double dual = 0; for (int i = 0; i < T; ++i) dual += u[i] * M;
    for (int j = 0; j < M; ++j) { double b = 1e300; for (int i = 0; i < T; ++i) b = min(b, c[(size_t)i*M+j] - u[i]); dual += b * T; }
    r.gap = r.cost - dual / ((double)T * M);
    return r;
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "ep40.bin";
    int maxlp = argc > 2 ? atoi(argv[2]) : 6;
    FILE* f = fopen(path, "rb");
    int hdr[3]; if (fread(hdr, 4, 3, f) != 3) return 1;
    int R = hdr[0], N = hdr[1], P = hdr[2];
    vector<float> cost((size_t)R * N); if (fread(cost.data(), 4, cost.size(), f) != cost.size()) return 1;
    vector<uint8_t> sup(R); if (fread(sup.data(), 1, R, f) != (size_t)R) return 1;
    vector<uint8_t> pooled((size_t)P * N); if (fread(pooled.data(), 1, pooled.size(), f) != pooled.size()) return 1;
    fclose(f);
    vector<int> rows; for (int r = 0; r < R; ++r) if (sup[r]) rows.push_back(r);
    int T0 = rows.size();
    for (int p = 0; p < P && p < maxlp; ++p) {
        vector<int> cols; for (int j = 0; j < N; ++j) if (pooled[(size_t)p * N + j]) cols.push_back(j);
        int M0 = cols.size();
        bool swapped = 3 * M0 < T0;
        int T = swapped ? M0 : T0, M = swapped ? T0 : M0;
        vector<double> c((size_t)T * M);
        for (int i = 0; i < T; ++i) for (int j = 0; j < M; ++j) c[(size_t)i * M + j] = swapped ? cost[(size_t)rows[j] * N + cols[i]] : cost[(size_t)rows[i] * N + cols[j]];
        for (double alpha : {4.0, 8.0}) {
            Res r = solve(c, T, M, 1e-7, alpha);
            printf("lp %3d T=%4d M=%4d alpha=%g emd=%.10f gap=%.2e rounds=%lld scans/TM=%.1f relabels=%lld\n", p, T, M, alpha, r.cost, r.gap, r.rounds, (double)r.scans / ((double)T * M), r.relabels);
        }
    }
}
