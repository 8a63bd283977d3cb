// TEST INFRASTRUCTURE -- see poa_oracle.h.
#include "poa_oracle.h"
#include <algorithm>
#include <cassert>

namespace oracle {

// ---------------------------------------------------------------------------------------
// graph
// ---------------------------------------------------------------------------------------
void PoaGraph::add_first(const uint8_t* seq, int n) {
    v.clear(); spans.clear();
    v.resize(n);
    for (int i = 0; i < n; ++i) {
        v[i].base = seq[i]; v[i].nreads = 1;
        v[i].prev = i - 1; v[i].next = (i + 1 < n) ? i + 1 : -1;
        v[i].col = i;
        if (i > 0) { v[i].in.push_back(i - 1); v[i - 1].out.push_back(i); }
    }
    head = n ? 0 : -1; tail = n - 1;
    if (n) spans.push_back({0, n - 1});
    n_reads = 1;
}

void PoaGraph::order(std::vector<int>& ord, std::vector<int>& rank) const {
    ord.clear();
    rank.assign(v.size(), -1);
    for (int x = head; x >= 0; x = v[x].next) { rank[x] = (int)ord.size(); ord.push_back(x); }
}

// PoaGraph::TryAddRead: banded local alignment of a read against the DAG (SURVEY.md Appendix B;
// band rule / tie breaks per DESIGN.md "Draft stage")
PoaAlignment PoaGraph::align(const uint8_t* seq, int n) const {
    PoaAlignment res;
    std::vector<int> ord, rank;
    order(ord, rank);
    const int V = (int)ord.size(), W = POA_BAND;
    if (V == 0 || n == 0) return res;
    std::vector<int> H((size_t)V * W, 0), lo(V, 0), besti(V, 0), rowmax(V, 0);
    std::vector<uint8_t> mv((size_t)V * W, 0);
    auto hget = [&](int t, int i) -> int {
        const int c = i - lo[t];
        return (c < 0 || c >= W) ? 0 : H[(size_t)t * W + c];
    };
    int gbest = 0, gt = -1, gi = -1;
    std::vector<int> preds;
    for (int t = 0; t < V; ++t) {
        const Vertex& vx = v[ord[t]];
        preds.clear();
        for (int u : vx.in) preds.push_back(rank[u]);
        std::sort(preds.begin(), preds.end());
        assert(preds.size() <= 8);
        // Band rule: the rows of a block of POA_BLOCK share one anchor, the last row of the previous block.  If the anchor
        // carries an alignment (best score >= POA_ANCHOR_MIN) the band is centred on the anchor's best cell, moved by the
        // difference of the seed coordinates; otherwise it moves POA_BAND_DECAY cells back towards the read start.  The
        // first block is anchored on the cell in front of the first row.
        const int a = (t / POA_BLOCK) * POA_BLOCK - 1;
        int l;
        if (a < 0) l = (vx.col + 1) - W / 2;
        else if (rowmax[a] >= POA_ANCHOR_MIN) l = besti[a] + (vx.col - v[ord[a]].col) - W / 2;
        else l = lo[a] - POA_BAND_DECAY;
        l = std::max(0, std::min(l, std::max(0, n + 1 - W)));
        lo[t] = l;
        int* h = &H[(size_t)t * W];
        uint8_t* m = &mv[(size_t)t * W];
        for (int c = 0; c < W; ++c) {
            const int i = l + c;
            int best = 0;
            uint8_t move = PM_STOP;
            if (i <= n) {
                if (i >= 1) {
                    const int sc = (seq[i - 1] == vx.base) ? POA_MATCH : POA_MISMATCH;
                    if (preds.empty()) {
                        if (sc > best) { best = sc; move = PM_MATCH | (63 << 2); }
                    } else {
                        for (size_t k = 0; k < preds.size(); ++k) {
                            const int cand = hget(preds[k], i - 1) + sc;
                            if (cand > best) { best = cand; move = (uint8_t)(PM_MATCH | (k << 2)); }
                        }
                    }
                }
                for (size_t k = 0; k < preds.size(); ++k) {
                    const int cand = hget(preds[k], i) + POA_DEL;
                    if (cand > best) { best = cand; move = (uint8_t)(PM_DEL | (k << 2)); }
                }
            }
            h[c] = best;
            m[c] = move;
        }
        for (int c = 1; c < W; ++c) {
            if (l + c > n) break;
            const int cand = h[c - 1] + POA_INS;
            if (cand > h[c]) { h[c] = cand; m[c] = PM_INS; }
        }
        int bi = l, bv = -1;
        for (int c = 0; c < W; ++c) if (l + c <= n && h[c] > bv) { bv = h[c]; bi = l + c; }
        besti[t] = bi;
        rowmax[t] = std::max(bv, 0);
        if (bv > gbest) { gbest = bv; gt = t; gi = bi; }
    }
    res.score = gbest;
    if (gt < 0) return res;
    // traceback
    int t = gt, i = gi;
    while (t >= 0) {
        const int c = i - lo[t];
        if (c < 0 || c >= W) break;
        const int hv = H[(size_t)t * W + c];
        if (hv <= 0) break;
        const uint8_t m = mv[(size_t)t * W + c];
        const int kind = m & 3, k = m >> 2;
        const Vertex& vx = v[ord[t]];
        if (kind == PM_MATCH) {
            res.path.push_back({ord[t], i - 1, PM_MATCH});
            if (k == 63) break;
            std::vector<int> pr;
            for (int u : vx.in) pr.push_back(rank[u]);
            std::sort(pr.begin(), pr.end());
            t = pr[k]; i -= 1;
        } else if (kind == PM_DEL) {
            res.path.push_back({ord[t], -1, PM_DEL});
            std::vector<int> pr;
            for (int u : vx.in) pr.push_back(rank[u]);
            std::sort(pr.begin(), pr.end());
            t = pr[k];
        } else if (kind == PM_INS) {
            res.path.push_back({-1, i - 1, PM_INS});
            i -= 1;
        } else break;
    }
    std::reverse(res.path.begin(), res.path.end());
    return res;
}

// PoaGraph::CommitAdd: thread the aligned read into the graph.  New vertices are placed in the
// vertex order immediately after their predecessor on the read's path, which keeps the list a
// topological order without re-sorting.
void PoaGraph::commit(const PoaAlignment& a, const uint8_t* seq) {
    int prevV = -1, first = -1, last = -1;
    auto add_edge = [&](int u, int w) {
        if (u < 0) return;
        for (int x : v[u].out) if (x == w) return;
        if (v[w].in.size() >= 8) return;    // a vertex keeps at most 8 predecessors; later edges are ignored (spec)
        v[u].out.push_back(w);
        v[w].in.push_back(u);
    };
    auto new_vertex_after = [&](int after, uint8_t base) -> int {
        Vertex nv;
        nv.base = base; nv.nreads = 1;
        const int id = (int)v.size();
        if (after < 0) {   // cannot happen for local alignments (paths start with a match); keep well defined
            nv.prev = -1; nv.next = head;
            v.push_back(nv);
            if (head >= 0) v[head].prev = id;
            head = id;
            if (tail < 0) tail = id;
        } else {
            nv.prev = after; nv.next = v[after].next;
            nv.col = v[after].col;
            v.push_back(nv);
            if (v[after].next >= 0) v[v[after].next].prev = id; else tail = id;
            v[after].next = id;
        }
        return id;
    };
    for (const PathStep& st : a.path) {
        int cur = -1;
        if (st.move == PM_MATCH) {
            if (v[st.vertex].base == seq[st.readpos]) { v[st.vertex].nreads++; cur = st.vertex; }
            else cur = new_vertex_after(prevV, seq[st.readpos]);
        } else if (st.move == PM_INS) {
            cur = new_vertex_after(prevV, seq[st.readpos]);
        } else continue;   // deletion: vertex skipped
        add_edge(prevV, cur);
        prevV = cur;
        if (first < 0) first = cur;
        last = cur;
    }
    if (first >= 0) spans.push_back({first, last});
    ++n_reads;
}

// PoaGraph::FindConsensus
std::vector<int> PoaGraph::consensus(int min_cov) const {
    std::vector<int> ord, rank;
    order(ord, rank);
    const int V = (int)ord.size();
    std::vector<int> cov(V + 1, 0);
    for (auto& s : spans) { cov[rank[s.first]]++; cov[rank[s.second] + 1]--; }
    for (int t = 1; t <= V; ++t) cov[t] += cov[t - 1];
    std::vector<long long> reach(V, 0);
    std::vector<int> bp(V, -1);
    long long best = 0;
    int bt = -1;
    for (int t = 0; t < V; ++t) {
        const Vertex& vx = v[ord[t]];
        const long long sc = 2ll * vx.nreads - std::max(cov[t], min_cov);
        long long m = 0;
        int mp = -1;
        std::vector<int> pr;
        for (int u : vx.in) pr.push_back(rank[u]);
        std::sort(pr.begin(), pr.end());
        for (int p : pr) if (reach[p] > m) { m = reach[p]; mp = p; }
        reach[t] = sc + m;
        bp[t] = mp;
        if (bt < 0 || reach[t] > best) { best = reach[t]; bt = t; }
    }
    std::vector<int> path;
    for (int t = bt; t >= 0; t = bp[t]) path.push_back(ord[t]);
    std::reverse(path.begin(), path.end());
    return path;
}

// ---------------------------------------------------------------------------------------
// orientation vote + mapping
// ---------------------------------------------------------------------------------------
static void kmers(const uint8_t* s, int n, std::vector<uint32_t>& out) {
    out.clear();
    if (n < POA_KMER) return;
    uint32_t k = 0;
    const uint32_t mask = (1u << (2 * POA_KMER)) - 1;
    for (int i = 0; i < n; ++i) {
        k = ((k << 2) | s[i]) & mask;
        if (i >= POA_KMER - 1) out.push_back(k);
    }
}

// content sampling: a k-mer takes part iff the three top bits of k * 0x9E3779B1 (mod 2^32) are clear
static bool kmer_sampled(uint32_t k) { return ((uint32_t)(k * 0x9E3779B1u) >> 29) == 0u; }

// only the first POA_VOTE_BASES bases of the read vote (hundreds of sampled k-mers: the decision is not close)
bool kmer_vote_reverse(const uint8_t* ref, int nref, const uint8_t* read, int n_read) {
    const int n = std::min(n_read, POA_VOTE_BASES);
    std::vector<uint32_t> rk, fk, ck;
    kmers(ref, nref, rk);
    std::sort(rk.begin(), rk.end());
    rk.erase(std::unique(rk.begin(), rk.end()), rk.end());
    std::vector<uint8_t> rc(n);
    for (int i = 0; i < n; ++i) rc[i] = (uint8_t)(3 - read[n - 1 - i]);
    kmers(read, n, fk);
    kmers(rc.data(), n, ck);
    long long f = 0, c = 0;
    for (uint32_t k : fk) if (kmer_sampled(k)) f += std::binary_search(rk.begin(), rk.end(), k);
    for (uint32_t k : ck) if (kmer_sampled(k)) c += std::binary_search(rk.begin(), rk.end(), k);
    return c > f;
}

ReadMapping map_to_template(const uint8_t* tpl, int J, const uint8_t* read_bases, int n) {
    ReadMapping m;
    PoaGraph g;
    g.add_first(tpl, J);
    PoaAlignment a = g.align(read_bases, n);
    m.score = a.score;
    if (a.path.empty()) return m;
    int fv = -1, lv = -1, fr = -1, lr = -1;
    for (const PathStep& st : a.path) {
        if (st.move == PM_MATCH) {
            if (fv < 0) { fv = st.vertex; fr = st.readpos; }
            lv = st.vertex; lr = st.readpos;
        }
    }
    if (fv < 0) return m;
    m.grid.assign((size_t)J / WINDOW_GRID + 1, -1);
    int consumed = fr;
    for (const PathStep& st : a.path) {
        if (st.move == PM_MATCH) {
            if (st.vertex % WINDOW_GRID == 0) m.grid[st.vertex / WINDOW_GRID] = st.readpos;
            consumed = st.readpos + 1;
        } else if (st.move == PM_DEL) {
            if (st.vertex % WINDOW_GRID == 0) m.grid[st.vertex / WINDOW_GRID] = consumed;
        } else if (st.move == PM_INS) consumed = st.readpos + 1;
    }
    m.tstart = fv; m.tend = lv + 1; m.rstart = fr; m.rend = lr + 1;
    m.mapped = a.score >= n;
    return m;
}

}  // namespace oracle
