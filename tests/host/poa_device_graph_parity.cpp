// CPU parity test of the Draft Stage's device graph routines (ccs_b200/csrc/cuda/poa_graph_ops.cuh: seed chain,
// CommitAdd on the id/order/rank layout, FindConsensus, the hashed k-mer vote) against the oracle's independent
// restatement (oracle/poa_oracle.*).  The routines are the very code the CUDA kernels run, instantiated here over a
// host execution context: a single thread, and a team of std::threads meeting at a barrier wherever the CTA would
// __syncthreads() (built with -fsanitize=thread by the test driver, this checks the phase structure for races).
// The alignment itself is taken from the oracle's CPU DP, so no device is needed.
// Built and run by tests/test_cpu_host.py::test_device_poa_graph_ops_match_oracle.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <random>
#include <thread>
#include <vector>
#include <pthread.h>
#include "../../ccs_b200/csrc/cuda/poa_graph_ops.cuh"
#include "../../oracle/poa_oracle.h"

using namespace ccs;

struct TeamExec {
    int tid_, n_;
    pthread_barrier_t* bar;
    int tid() const { return tid_; }
    int nthreads() const { return n_; }
    void sync() const { if (n_ > 1) pthread_barrier_wait(bar); }
    uint32_t cas(uint32_t* p, uint32_t c, uint32_t v) const {
        __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
        return c;
    }
};

template <class F>
static void run_team(int n, F&& f) {
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, (unsigned)n);
    std::vector<std::thread> th;
    for (int t = 1; t < n; ++t) th.emplace_back([&, t]() { TeamExec x{t, n, &bar}; f(x); });
    TeamExec x{0, n, &bar};
    f(x);
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&bar);
}

static int fail(const char* what, int trial, int round) {
    std::fprintf(stderr, "MISMATCH: %s (trial %d, round %d)\n", what, trial, round);
    return 1;
}

int main(int argc, char** argv) {
    const int trials = argc > 1 ? std::atoi(argv[1]) : 30;
    const int team = argc > 2 ? std::atoi(argv[2]) : 1;
    std::mt19937 rng(20261017u);
    long cells = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int L = 150 + (int)(rng() % 900);
        std::vector<uint8_t> truth(L);
        for (auto& b : truth) b = (uint8_t)(rng() & 3);
        auto noisy = [&]() {     // ~12 % errors, indel dominated, like a subread; emission codes carry a pulse width
            std::vector<uint8_t> r;
            for (int i = 0; i < L; ++i) {
                const unsigned u = rng() % 1000;
                if (u < 40) continue;
                if (u < 90) r.push_back((uint8_t)(rng() & 3));
                r.push_back(u < 120 ? (uint8_t)((truth[i] + 1 + rng() % 3) & 3) : truth[i]);
            }
            return r;
        };
        auto to_codes = [&](const std::vector<uint8_t>& bases, bool rev) {    // native-orientation codes of an oriented read
            const int n = (int)bases.size();
            std::vector<uint8_t> c(n);
            for (int i = 0; i < n; ++i) {
                const int b = rev ? 3 - bases[n - 1 - i] : bases[i];
                c[i] = (uint8_t)(4 * (rng() % 3) + b);
            }
            return c;
        };
        std::vector<std::vector<uint8_t>> reads, codes;
        std::vector<int> revs;
        int cap = 8;
        for (int k = 0; k < 5; ++k) {
            reads.push_back(noisy());
            revs.push_back(k == 0 ? 0 : (int)(rng() & 1));
            codes.push_back(to_codes(reads.back(), revs.back() != 0));
            cap += k == 0 ? (int)reads.back().size() : poa_new_vertex_bound((int)reads.back().size());
        }
        // device-layout graph on the host
        PoaGraphHdr hdr{};
        hdr.voff = 16; hdr.cap = cap;            // a non-zero pool offset exercises the addressing
        std::vector<uint32_t> meta(hdr.voff + cap);
        std::vector<int32_t> pred0(hdr.voff + cap), predx(7 * (hdr.voff + cap)), rank(hdr.voff + cap), ordA(hdr.voff + cap), ordB(hdr.voff + cap), col(hdr.voff + cap);
        PoaGraphView G{&hdr, meta.data(), pred0.data(), predx.data(), rank.data(), {ordA.data(), ordB.data()}, col.data()};
        std::vector<int32_t> scratch(5 * 2000 + 12 * cap + 16), sm(64), sh(kPoaConsShared);
        std::vector<uint8_t> cons(cap);
        int32_t cons_len = 0;

        oracle::PoaGraph og;
        og.add_first(reads[0].data(), (int)reads[0].size());
        {
            PoaReadAcc R{codes[0].data(), (int)codes[0].size(), 0};
            run_team(team, [&](TeamExec& x) { poa_graph_init(x, G, 0, R); });
        }
        for (int round = 1; round <= 5; ++round) {
            std::vector<int> ord, orank;
            og.order(ord, orank);
            const int V = (int)ord.size();
            if (V != hdr.V) return fail("vertex count", trial, round);
            const int32_t* dord = poa_order(G, hdr.order_sel) + hdr.voff;
            for (int t = 0; t < V; ++t) {
                const int id = ord[t];
                if (dord[t] != id) return fail("order (vertex ids are handed out in path order)", trial, round);
                if (rank[hdr.voff + id] != t) return fail("rank", trial, round);
                if (col[hdr.voff + id] != og.v[id].col) return fail("seed coordinate", trial, round);
                const auto& vx = og.v[id];
                const uint32_t m = meta[hdr.voff + id];
                if (poa_meta_base(m) != vx.base) return fail("base", trial, round);
                if (poa_meta_nreads(m) != vx.nreads) return fail("nreads", trial, round);
                std::vector<int> pr;
                for (int u : vx.in) pr.push_back(orank[u]);
                std::sort(pr.begin(), pr.end());
                if ((int)pr.size() != poa_meta_nin(m)) return fail("in-degree", trial, round);
                for (size_t k = 0; k < pr.size(); ++k) {
                    const int p = k == 0 ? pred0[hdr.voff + id] : predx[7 * (hdr.voff + id) + k - 1];
                    if (rank[hdr.voff + p] != pr[k]) return fail("predecessor (sorted by rank)", trial, round);
                }
                ++cells;
            }
            // consensus (minCov follows from the number of threaded reads, as in the draft stage)
            {
                const int n = og.n_reads;
                if (n != hdr.n_reads) return fail("n_reads", trial, round);
                const std::vector<int> oc = og.consensus(n < 5 ? 1 : (n + 1) / 2 - 1);
                run_team(team, [&](TeamExec& x) { poa_graph_consensus(x, G, 0, scratch.data(), cons.data(), &cons_len, sh.data()); });
                if ((int)oc.size() != cons_len) return fail("consensus length", trial, round);
                for (size_t k = 0; k < oc.size(); ++k) if (cons[k] != og.v[oc[k]].base) return fail("consensus base", trial, round);
            }
            if (round == 5) break;
            // align the next read with the oracle; steps end -> start, deletions left out (the traceback kernel's output)
            const std::vector<uint8_t>& read = reads[round];
            const oracle::PoaAlignment a = og.align(read.data(), (int)read.size());
            if (a.path.empty()) return fail("empty alignment", trial, round);
            std::vector<PoaStep> steps;
            for (size_t k = a.path.size(); k-- > 0;) {
                const oracle::PathStep& st = a.path[k];
                if (st.move == oracle::PM_DEL) continue;
                steps.push_back(PoaStep{st.move == oracle::PM_MATCH ? st.vertex : -1, st.readpos});
            }
            og.commit(a, read.data());
            PoaReadAcc R{codes[round].data(), (int)codes[round].size(), revs[round]};
            for (int i = 0; i < R.n; ++i) if (R.base(i) != read[i]) return fail("read orientation accessor", trial, round);
            run_team(team, [&](TeamExec& x) {
                poa_graph_commit(x, G, 0, steps.data(), (int)steps.size(), R, scratch.data(), sm.data());
            });
            if (hdr.error) return fail("capacity bound exceeded", trial, round);
        }
        // k-mer vote: hashed set vs the oracle's sorted lists, both orientations and an unrelated read
        for (int k = 1; k < 4; ++k) {
            std::vector<uint8_t> fwd_codes = (k == 3) ? to_codes(noisy(), (rng() & 1) != 0) : codes[k];
            if (k == 3) for (auto& c : fwd_codes) c = (uint8_t)(4 * (c >> 2) + (rng() & 3));   // unrelated
            const int n = (int)fwd_codes.size();
            std::vector<uint8_t> fb(n);
            for (int i = 0; i < n; ++i) fb[i] = fwd_codes[i] & 3;
            const bool want = oracle::kmer_vote_reverse(reads[0].data(), (int)reads[0].size(), fb.data(), n);
            const int tcap = poa_kmer_table_cap((int)reads[0].size());
            std::vector<uint32_t> tab(tcap, 0u);
            PoaBaseAcc ref{codes[0].data(), 1};
            run_team(team, [&](TeamExec& x) { poa_kmer_build(x, tab.data(), tcap, ref, (int)codes[0].size()); });
            const int nv = std::min(n, kPoaVoteBases);
            int f = 0, c = 0;
            // split over "lanes" exactly as the kernel does
            const int npos = nv - (kPoaKmer - 1);
            if (npos > 0) {
                const int per = (npos + 31) / 32;
                for (int lane = 0; lane < 32; ++lane) {
                    const int b = kPoaKmer - 1 + lane * per, e = std::min(b + per, nv);
                    int lf, lc;
                    poa_kmer_count(tab.data(), tcap, fwd_codes.data(), b, e, lf, lc);
                    f += lf; c += lc;
                }
            }
            if ((c > f) != want) return fail("k-mer orientation vote", trial, k);
            if (k < 3 && L >= 600 && (c > f) != (revs[k] != 0)) return fail("k-mer vote does not recover the simulated strand", trial, k);
        }
    }
    std::printf("ok: device POA graph routines == oracle on %d graphs x 5 rounds, team of %d (%ld vertices compared)\n",
                trials, team, cells);
    return 0;
}
