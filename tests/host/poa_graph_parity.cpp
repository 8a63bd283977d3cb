// CPU parity test of the product's host POA graph (ccs_b200/csrc/host/poa_graph.h: CommitAdd threading, topological
// export, FindConsensus) against the oracle's independent restatement (oracle/poa_oracle.*).  The alignment itself is
// taken from the oracle's CPU DP and re-encoded in the GPU kernel's move format, so no device is needed.
// Built and run by tests/test_cpu_host.py::test_host_poa_graph_matches_oracle.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../ccs_b200/csrc/host/poa_graph.h"
#include "../../oracle/poa_oracle.h"

using namespace ccs;

static int fail(const char* what, int trial, int round) {
    std::fprintf(stderr, "MISMATCH: %s (trial %d, round %d)\n", what, trial, round);
    return 1;
}

int main(int argc, char** argv) {
    const int trials = argc > 1 ? std::atoi(argv[1]) : 40;
    std::mt19937 rng(20261017u);
    long cells = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int L = 150 + (int)(rng() % 900);
        std::vector<uint8_t> truth(L);
        for (auto& b : truth) b = (uint8_t)(rng() & 3);
        auto noisy = [&]() {     // ~12 % errors, indel dominated, like a subread
            std::vector<uint8_t> r;
            for (int i = 0; i < L; ++i) {
                const unsigned u = rng() % 1000;
                if (u < 40) continue;                                        // deletion
                if (u < 90) r.push_back((uint8_t)(rng() & 3));               // insertion before the base
                r.push_back(u < 120 ? (uint8_t)((truth[i] + 1 + rng() % 3) & 3) : truth[i]);
            }
            return r;
        };
        std::vector<uint8_t> seed = noisy();
        oracle::PoaGraph og;
        og.add_first(seed.data(), (int)seed.size());
        HostPoaGraph hg;
        hg.init(seed.data(), (int)seed.size());
        for (int round = 1; round <= 5; ++round) {
            // ---- exports must agree in rank space
            std::vector<int> ord, rank;
            og.order(ord, rank);
            const int V = (int)ord.size();
            if (V != hg.size()) return fail("vertex count", trial, round);
            std::vector<int32_t> order, poff(V + 1), preds(hg.n_edges() + 1);
            std::vector<uint8_t> base(V);
            hg.export_topo(order, base.data(), poff.data(), preds.data());
            size_t n_edges = 0;
            for (int t = 0; t < V; ++t) {
                const auto& vx = og.v[ord[t]];
                if (base[t] != vx.base) return fail("base", trial, round);
                std::vector<int> pr;
                for (int u : vx.in) pr.push_back(rank[u]);
                std::sort(pr.begin(), pr.end());
                if ((int)pr.size() != poff[t + 1] - poff[t]) return fail("in-degree", trial, round);
                for (size_t k = 0; k < pr.size(); ++k) if (preds[poff[t] + k] != pr[k]) return fail("predecessor", trial, round);
                n_edges += pr.size();
                ++cells;
            }
            if ((int)n_edges != hg.n_edges()) return fail("edge count", trial, round);
            // ---- consensus must agree for the minCov values the draft stage uses
            for (int mc = 1; mc <= 2; ++mc) {
                std::vector<uint8_t> hc;
                hg.consensus(mc, hc);
                const std::vector<int> oc = og.consensus(mc);
                if (hc.size() != oc.size()) return fail("consensus length", trial, round);
                for (size_t k = 0; k < oc.size(); ++k) if (hc[k] != og.v[oc[k]].base) return fail("consensus base", trial, round);
            }
            if (round == 5) break;
            // ---- align the next read with the oracle, re-encode the path as the kernel's traceback would
            const std::vector<uint8_t> read = noisy();
            const oracle::PoaAlignment a = og.align(read.data(), (int)read.size());
            if (a.path.empty()) return fail("empty alignment", trial, round);
            std::vector<uint8_t> moves;     // end -> start
            int end_t = -1, end_i = -1;
            for (size_t k = a.path.size(); k-- > 0;) {
                const oracle::PathStep& st = a.path[k];
                if (st.move == oracle::PM_INS) { moves.push_back(3u); continue; }
                const int t = rank[st.vertex];
                if (end_t < 0) {
                    if (st.move != oracle::PM_MATCH) return fail("path does not end with a match", trial, round);
                    end_t = t; end_i = st.readpos + 1;
                }
                // predecessor on the path = the previous step that has a vertex
                int pt = -1;
                for (size_t q = k; q-- > 0;) if (a.path[q].move != oracle::PM_INS) { pt = rank[a.path[q].vertex]; break; }
                unsigned ordk = 63u;
                if (pt >= 0) {
                    if (pt == t - 1) ordk = 62u;
                    else {
                        ordk = 64u;
                        for (int e = poff[t]; e < poff[t + 1]; ++e) if (preds[e] == pt) ordk = (unsigned)(e - poff[t]);
                        if (ordk == 64u) return fail("path step without an edge", trial, round);
                    }
                }
                moves.push_back((uint8_t)((st.move == oracle::PM_MATCH ? 1u : 2u) | (ordk << 2)));
            }
            og.commit(a, read.data());
            hg.commit(moves.data(), (int)moves.size(), end_t, end_i, order, poff.data(), preds.data(), read.data());
        }
    }
    std::printf("ok: host POA graph == oracle on %d graphs x 5 rounds (%ld vertices compared)\n", trials, cells);
    return 0;
}
