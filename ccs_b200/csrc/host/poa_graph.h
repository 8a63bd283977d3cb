// Host-side partial-order graph of the Draft Stage (PoaGraph / SparsePoa bookkeeping,
// SURVEY.md 8a rows a2, a4 and Appendix B).  The alignment DP runs on the GPU
// (cuda/poa_align.cu); this class owns what is inherently sequential and tiny: threading an
// aligned read into the graph (CommitAdd), keeping the vertex list in topological order, and
// the consensus path (FindConsensus).
//
// Order maintenance: vertices live in a singly linked list that is always a topological order;
// a new vertex is linked immediately after its predecessor on the read's path, so no re-sort is
// ever needed (DESIGN.md "Draft stage").
//
// Layout: one 20-byte node per vertex {next, nreads, first predecessor, block, base, in-degree};
// almost every vertex is a chain vertex with exactly one predecessor, so export, CommitAdd and
// consensus touch one cache line per 3 vertices; further predecessors (up to kPoaMaxPred) live in
// 7-int blocks of a pool that only branch points own.  Per-call scratch (path steps, ranks, the
// consensus DP arrays) is thread-local and reused across graphs: a fresh graph per ZMW would
// otherwise page-fault a few hundred KB of temporaries on every call.  This host code is the
// multi-GPU scaling limit of the whole path (DESIGN.md "Multi-GPU"), hence the care.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#include "../cuda/poa_device.h"

namespace ccs {

class HostPoaGraph {
public:
    void init(const uint8_t* seq, int n) {
        node_.clear();
        node_.reserve((size_t)n + (size_t)n / 2 + 64);   // four more reads add ~10 % new vertices each
        node_.resize((size_t)n);
        for (int i = 0; i < n; ++i) {
            Node& v = node_[i];
            v.next = (i + 1 < n) ? i + 1 : -1;
            v.nreads = 1;
            v.in0 = i - 1;
            v.more = -1;
            v.base = seq[i];
            v.nin = i > 0 ? 1 : 0;
            v.pad_ = 0;
        }
        pool_.clear();
        head_ = n ? 0 : -1;
        n_edges_ = n ? n - 1 : 0;
        spans_.clear();
        if (n) spans_.push_back({0, n - 1});
        n_reads_ = 1;
    }
    int size() const { return (int)node_.size(); }
    int n_reads() const { return n_reads_; }
    int n_edges() const { return n_edges_; }

    // Topological export into caller buffers: order[t] = vertex id; base[V]; pred_off[V+1] (relative
    // to this graph's list); preds[n_edges] = predecessor ranks, ascending per vertex.
    void export_topo(std::vector<int32_t>& order, uint8_t* base, int32_t* pred_off, int32_t* preds) const {
        const int V = size();
        order.resize(V);
        std::vector<int32_t>& rank_ = scratch().rank;
        rank_.resize(V);
        // one pass in list (= topological) order: a vertex's predecessors precede it, so their ranks are already known
        int32_t np = 0;
        int t = 0;
        for (int x = head_; x >= 0; ++t) {
            const Node& v = node_[x];
            rank_[x] = t;
            order[t] = x;
            base[t] = v.base;
            pred_off[t] = np;
            if (v.nin == 1) preds[np++] = rank_[v.in0];            // the common case: a chain vertex
            else if (v.nin > 1) {
                const int b = np;
                preds[np++] = rank_[v.in0];
                const int32_t* e = &pool_[(size_t)v.more * kMore];
                for (int k = 0; k + 1 < v.nin; ++k) preds[np++] = rank_[e[k]];
                std::sort(preds + b, preds + np);
            }
            x = v.next;
        }
        pred_off[V] = np;
    }

    // CommitAdd: replay the traceback (moves end -> start from the GPU) against the exported
    // topology and thread the read.  order/pred_off/preds are this graph's export of the round.
    void commit(const uint8_t* moves_rev, int len, int end_t, int end_i, const std::vector<int32_t>& order,
                const int32_t* pred_off, const int32_t* preds, const uint8_t* seq) {
        // rebuild the path start -> end: (vertex id or -1, read position)
        std::vector<std::pair<int32_t, int32_t>>& steps_ = scratch().steps;
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
            if (vx >= 0 && node_[vx].base == seq[rp]) { node_[vx].nreads++; cur = vx; }
            else cur = new_vertex_after(prevV, seq[rp]);
            add_edge(prevV, cur);
            prevV = cur;
            if (first < 0) first = cur;
            last = cur;
        }
        if (first >= 0) spans_.push_back({first, last});
        ++n_reads_;
    }

    // FindConsensus: score(v) = 2*nReads - max(spanning, minCov); best-scoring path, ties to the lowest rank.
    // One pass in list (= topological) order: ranks, the number of reads spanning each vertex (running count of span
    // starts minus span ends, marked per vertex id) and the best-path DP together.
    void consensus(int min_cov, std::vector<uint8_t>& out) const {
        const int V = size();
        Scratch& S = scratch();
        std::vector<int32_t>& order_ = S.order; std::vector<int32_t>& rank_ = S.rank;
        std::vector<int32_t>& mark_ = S.cov; std::vector<int32_t>& bp_ = S.bp;
        std::vector<int64_t>& reach_ = S.reach;
        order_.resize(V);
        rank_.resize(V);
        reach_.resize((size_t)V);
        bp_.resize((size_t)V);
        mark_.assign((size_t)2 * V, 0);                    // [2x] span starts at vertex x, [2x+1] span ends at x
        for (auto& s : spans_) { mark_[2 * (size_t)s.first]++; mark_[2 * (size_t)s.second + 1]++; }
        int64_t best = 0;
        int bt = -1, t = 0, running = 0;
        for (int x = head_; x >= 0; ++t) {
            const Node& v = node_[x];
            rank_[x] = t;
            order_[t] = x;
            running += mark_[2 * (size_t)x];
            const int64_t sc = 2ll * v.nreads - std::max(running, min_cov);
            running -= mark_[2 * (size_t)x + 1];
            int64_t m = 0;
            int mp = -1;
            if (v.nin >= 1) {
                const int p = rank_[v.in0];
                if (reach_[p] > 0) { m = reach_[p]; mp = p; }
                const int32_t* e = pool_.data() + (size_t)(v.nin > 1 ? v.more : 0) * kMore;
                for (int k = 0; k + 1 < v.nin; ++k) {
                    const int q = rank_[e[k]];
                    if (reach_[q] > m || (reach_[q] == m && mp >= 0 && q < mp && reach_[q] > 0)) { m = reach_[q]; mp = q; }
                }
            }
            reach_[t] = sc + m;
            bp_[t] = mp;
            if (bt < 0 || reach_[t] > best) { best = reach_[t]; bt = t; }
            x = v.next;
        }
        out.clear();
        for (int k = bt; k >= 0; k = bp_[k]) out.push_back(node_[order_[k]].base);
        std::reverse(out.begin(), out.end());
    }

private:
    struct Node { int32_t next; int32_t nreads; int32_t in0; int32_t more; uint8_t base; uint8_t nin; uint16_t pad_; };
    static_assert(sizeof(Node) == 20, "Node layout");
    static constexpr int kMore = kPoaMaxPred - 1;     // predecessors beyond the first: one pool block per branch point
    struct Scratch {
        std::vector<std::pair<int32_t, int32_t>> steps;
        std::vector<int32_t> rank, order, cov, bp;
        std::vector<int64_t> reach;
    };
    static Scratch& scratch() { static thread_local Scratch s; return s; }

    int new_vertex_after(int after, uint8_t b) {
        const int id = size();
        Node v;
        v.nreads = 1; v.in0 = -1; v.more = -1; v.base = b; v.nin = 0; v.pad_ = 0;
        if (after < 0) { v.next = head_; head_ = id; }
        else { v.next = node_[after].next; node_[after].next = id; }
        node_.push_back(v);
        return id;
    }
    void add_edge(int u, int w) {
        if (u < 0) return;
        Node& v = node_[w];
        if (v.nin == 0) { v.in0 = u; v.nin = 1; ++n_edges_; return; }
        if (v.in0 == u) return;
        if (v.more < 0) { v.more = (int32_t)(pool_.size() / kMore); pool_.resize(pool_.size() + kMore); }
        int32_t* e = &pool_[(size_t)v.more * kMore];
        for (int k = 0; k + 1 < v.nin; ++k) if (e[k] == u) return;
        if (v.nin < kPoaMaxPred) { e[v.nin - 1] = u; ++v.nin; ++n_edges_; }
    }
    std::vector<Node> node_;
    std::vector<int32_t> pool_;        // blocks of kMore ints: predecessors 2..kPoaMaxPred of the branch points
    int head_ = -1, n_reads_ = 0, n_edges_ = 0;
    std::vector<std::pair<int32_t, int32_t>> spans_;
};

}  // namespace ccs
