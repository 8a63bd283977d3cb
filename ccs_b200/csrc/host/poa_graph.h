// Host-side partial-order graph of the Draft Stage (PoaGraph / SparsePoa bookkeeping,
// SURVEY.md 8a rows a2, a4 and Appendix B).  The alignment DP runs on the GPU
// (cuda/poa_align.cu); this class owns what is inherently sequential and tiny: threading an
// aligned read into the graph (CommitAdd), keeping the vertex list in topological order, and
// the consensus path (FindConsensus).
//
// Order maintenance: vertices live in a doubly linked list that is always a topological order;
// a new vertex is linked immediately after its predecessor on the read's path, so no re-sort is
// ever needed (DESIGN.md "Draft stage").
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#include "../cuda/poa_device.h"

namespace ccs {

class HostPoaGraph {
public:
    void init(const uint8_t* seq, int n) {
        base_.assign(seq, seq + n);
        nreads_.assign(n, 1);
        next_.resize(n); prev_.resize(n);
        in_.resize((size_t)n * kPoaMaxPred);      // only the first nin_ entries of a vertex are ever read
        nin_.assign(n, 0);
        for (int i = 0; i < n; ++i) {
            prev_[i] = i - 1; next_[i] = (i + 1 < n) ? i + 1 : -1;
            if (i > 0) { in_[(size_t)i * kPoaMaxPred] = i - 1; nin_[i] = 1; }
        }
        head_ = n ? 0 : -1;
        n_edges_ = n ? n - 1 : 0;
        spans_.clear();
        if (n) spans_.push_back({0, n - 1});
        n_reads_ = 1;
    }
    int size() const { return (int)base_.size(); }
    int n_reads() const { return n_reads_; }

    int n_edges() const { return n_edges_; }

    // Topological export into caller buffers: order[t] = vertex id; base[V]; pred_off[V+1] (relative
    // to this graph's list); preds[n_edges] = predecessor ranks, ascending per vertex.
    void export_topo(std::vector<int32_t>& order, uint8_t* base, int32_t* pred_off, int32_t* preds) const {
        const int V = size();
        order.resize(V);
        rank_.resize(V);
        // one pass in list (= topological) order: a vertex's predecessors precede it, so their ranks are already known
        int32_t np = 0;
        int t = 0;
        for (int x = head_; x >= 0; x = next_[x], ++t) {
            rank_[x] = t;
            order[t] = x;
            base[t] = base_[x];
            pred_off[t] = np;
            const int n = nin_[x];
            const int32_t* e = &in_[(size_t)x * kPoaMaxPred];
            if (n == 1) preds[np++] = rank_[e[0]];                 // the common case: a chain vertex
            else if (n > 1) {
                const int b = np;
                for (int k = 0; k < n; ++k) preds[np++] = rank_[e[k]];
                std::sort(preds + b, preds + np);
            }
        }
        pred_off[V] = np;
    }

    // CommitAdd: replay the traceback (moves end -> start from the GPU) against the exported
    // topology and thread the read.  order/pred_off/preds are this graph's export of the round.
    void commit(const uint8_t* moves_rev, int len, int end_t, int end_i, const std::vector<int32_t>& order,
                const int32_t* pred_off, const int32_t* preds, const uint8_t* seq) {
        // rebuild the path start -> end: (vertex id or -1, read position or -1)
        steps_.clear();
        int t = end_t, i = end_i;
        for (int k = 0; k < len; ++k) {
            const unsigned m = moves_rev[k], kind = m & 3u, ord = m >> 2;
            if (kind == 1u) {
                steps_.push_back({order[t], i - 1});
                if (ord == 63u) break;
                t = (ord == 62u) ? t - 1 : preds[pred_off[t] + ord]; i -= 1;
            } else if (kind == 2u) {
                t = (ord == 62u) ? t - 1 : preds[pred_off[t] + ord];   // vertex skipped: nothing to thread
            } else {
                steps_.push_back({-1, i - 1});
                i -= 1;
            }
        }
        int prevV = -1, first = -1, last = -1;
        for (size_t k = steps_.size(); k-- > 0;) {
            const int vx = steps_[k].first, rp = steps_[k].second;
            int cur;
            if (vx >= 0 && base_[vx] == seq[rp]) { nreads_[vx]++; cur = vx; }
            else cur = new_vertex_after(prevV, seq[rp]);
            add_edge(prevV, cur);
            prevV = cur;
            if (first < 0) first = cur;
            last = cur;
        }
        if (first >= 0) spans_.push_back({first, last});
        ++n_reads_;
    }

    // FindConsensus: score(v) = 2*nReads - max(spanning, minCov); best-scoring path
    void consensus(int min_cov, std::vector<uint8_t>& out) const {
        const int V = size();
        std::vector<int32_t> order(V);
        rank_.resize(V);
        { int t = 0; for (int x = head_; x >= 0; x = next_[x]) { rank_[x] = t; order[t++] = x; } }
        std::vector<int32_t> cov(V + 1, 0);
        for (auto& s : spans_) { cov[rank_[s.first]]++; cov[rank_[s.second] + 1]--; }
        for (int t = 1; t <= V; ++t) cov[t] += cov[t - 1];
        std::vector<int64_t> reach(V, 0);
        std::vector<int32_t> bp(V, -1);
        int64_t best = 0;
        int bt = -1;
        for (int t = 0; t < V; ++t) {
            const int x = order[t];
            const int64_t sc = 2ll * nreads_[x] - std::max(cov[t], min_cov);
            int64_t m = 0;
            int mp = -1;
            for (int k = 0; k < nin_[x]; ++k) {
                const int p = rank_[in_[(size_t)x * kPoaMaxPred + k]];
                if (reach[p] > m || (reach[p] == m && mp >= 0 && p < mp && reach[p] > 0)) { m = reach[p]; mp = p; }
            }
            reach[t] = sc + m;
            bp[t] = mp;
            if (bt < 0 || reach[t] > best) { best = reach[t]; bt = t; }
        }
        out.clear();
        for (int t = bt; t >= 0; t = bp[t]) out.push_back(base_[order[t]]);
        std::reverse(out.begin(), out.end());
    }

private:
    int new_vertex_after(int after, uint8_t b) {
        const int id = size();
        base_.push_back(b); nreads_.push_back(1); nin_.push_back(0);
        in_.resize(in_.size() + kPoaMaxPred);
        if (after < 0) { prev_.push_back(-1); next_.push_back(head_); if (head_ >= 0) prev_[head_] = id; head_ = id; }
        else {
            prev_.push_back(after); next_.push_back(next_[after]);
            if (next_[after] >= 0) prev_[next_[after]] = id;
            next_[after] = id;
        }
        return id;
    }
    void add_edge(int u, int w) {
        if (u < 0) return;
        int32_t* e = &in_[(size_t)w * kPoaMaxPred];
        for (int k = 0; k < nin_[w]; ++k) if (e[k] == u) return;
        if (nin_[w] < kPoaMaxPred) { e[nin_[w]++] = u; ++n_edges_; }
    }
    std::vector<uint8_t> base_;
    std::vector<int32_t> nreads_, next_, prev_, in_;
    std::vector<uint8_t> nin_;
    int head_ = -1, n_reads_ = 0, n_edges_ = 0;
    std::vector<std::pair<int32_t, int32_t>> spans_;
    mutable std::vector<int32_t> rank_;
    std::vector<std::pair<int32_t, int32_t>> steps_;
};

}  // namespace ccs
