// Graph bookkeeping of the Draft Stage on the device-resident partial-order graph (poa_device.h):
// seed chain, CommitAdd (threading an aligned read), FindConsensus and the k-mer orientation vote
// (PoaGraph::CommitAdd / FindConsensus / SdpRangeFinder seeding -- SURVEY.md 8a rows a2-a4, Appendix B;
// /root/reference/docs/how-does-ccs-work.md:34-47).
//
// Every routine is written once, as phases of index-parallel loops separated by X.sync(), over an execution
// context X: on the device X is a CTA (poa_graph.cu: one CTA per graph), on the host X is a single thread
// (tests/host/poa_device_graph_parity.cpp checks the very same code against the oracle's graph on the CPU; the
// product never runs the host context).
//
// CommitAdd without a linked list: a read's path visits old vertices in increasing rank, and every new vertex
// goes "immediately after its path predecessor" (spec), so the new vertices form runs, each anchored behind one
// old path vertex (or at the head).  New ranks follow from one prefix sum over the number of new vertices
// anchored before each old rank; ids are handed out in path order (= the oracle's numbering).
#pragma once
#include <cstdint>
#include "poa_device.h"

#if defined(__CUDACC__)
#define CCS_HD __device__ __forceinline__
#else
#define CCS_HD inline
#endif

namespace ccs {

struct PoaReadAcc {      // oriented bases of a read stored as native-orientation emission codes
    const uint8_t* codes;
    int n;
    int rev;
    CCS_HD int base(int i) const { return rev ? 3 - (codes[n - 1 - i] & 3) : (codes[i] & 3); }
};

CCS_HD uint32_t poa_meta(int base, int nin, int nreads) { return (uint32_t)base | ((uint32_t)nin << 2) | ((uint32_t)nreads << 8); }
CCS_HD int poa_meta_base(uint32_t m) { return (int)(m & 3u); }
CCS_HD int poa_meta_nin(uint32_t m) { return (int)((m >> 2) & 15u); }
CCS_HD int poa_meta_nreads(uint32_t m) { return (int)(m >> 8); }

// Block-wide inclusive scan of a[0..n) in place (sum or max), a in global memory; sm = X.nthreads() ints of scratch.
template <class X, bool kMax>
CCS_HD void poa_block_scan(X& x, int32_t* a, int n, int32_t* sm) {
    const int T = x.nthreads(), tid = x.tid();
    const int per = (n + T - 1) / T;
    const int b = tid * per, e = (b + per < n) ? b + per : n;
    int acc = kMax ? -1 : 0;
    for (int i = b; i < e; ++i) acc = kMax ? (a[i] > acc ? a[i] : acc) : acc + a[i];
    sm[tid] = acc;
    x.sync();
    if (tid == 0) {
        int run = kMax ? -1 : 0;
        for (int t = 0; t < T; ++t) { const int v = sm[t]; sm[t] = run; run = kMax ? (v > run ? v : run) : run + v; }
    }
    x.sync();
    acc = sm[tid];
    for (int i = b; i < e; ++i) { acc = kMax ? (a[i] > acc ? a[i] : acc) : acc + a[i]; a[i] = acc; }
    x.sync();
}

// Seed chain: the first read becomes vertices 0..n-1.
template <class X>
CCS_HD void poa_graph_init(X& x, const PoaGraphView& G, int g, const PoaReadAcc& R) {
    PoaGraphHdr& H = G.hdr[g];
    const int n = R.n;
    for (int i = x.tid(); i < n; i += x.nthreads()) {
        G.meta[H.voff + i] = poa_meta(R.base(i), i > 0 ? 1 : 0, 1);
        G.pred0[H.voff + i] = i - 1;
        G.rank[H.voff + i] = i;
        G.order[0][H.voff + i] = i;
        G.col[H.voff + i] = i;
    }
    if (x.tid() == 0) {
        H.V = n; H.n_reads = 1; H.n_spans = n ? 1 : 0; H.order_sel = 0; H.error = 0;
        H.span_first[0] = 0; H.span_last[0] = n - 1;
    }
    x.sync();
}

// CommitAdd.  steps[0..L) = the read's path, END -> START, deletions left out ({vertex id or -1, read position}).
// scratch: 5 * L + (V + 1) ints.
template <class X>
CCS_HD void poa_graph_commit(X& x, const PoaGraphView& G, int g, const PoaStep* steps, int L, const PoaReadAcc& R,
                             int32_t* scratch, int32_t* sm) {
    PoaGraphHdr& H = G.hdr[g];
    if (L <= 0) return;
    const int64_t o = H.voff;
    const int Vold = H.V;
    const int sel = H.order_sel;
    int32_t* newidx = scratch;            // [L] flag -> inclusive count of new steps
    int32_t* anchor = scratch + L;        // [L] last step <= f that reuses an old vertex (-1: none)
    int32_t* cur = scratch + 2 * (int64_t)L;      // [L] vertex id of the step
    int32_t* arank = scratch + 3 * (int64_t)L;    // [L] old rank of the anchor vertex (-1: head)
    int32_t* isnew = scratch + 4 * (int64_t)L;    // [L]
    int32_t* add = scratch + 5 * (int64_t)L;      // [Vold + 1]: new vertices anchored behind old rank (index - 1)
    const int T = x.nthreads(), tid = x.tid();
    // A: classify the steps (forward order f = L-1-k)
    for (int f = tid; f < L; f += T) {
        const PoaStep s = steps[L - 1 - f];
        const bool nw = s.vertex < 0 || poa_meta_base(G.meta[o + s.vertex]) != R.base(s.readpos);
        isnew[f] = nw ? 1 : 0;
        newidx[f] = nw ? 1 : 0;
        anchor[f] = nw ? -1 : f;
    }
    for (int t = tid; t <= Vold; t += T) add[t] = 0;
    x.sync();
    poa_block_scan<X, false>(x, newidx, L, sm);
    poa_block_scan<X, true>(x, anchor, L, sm);
    const int n_new = newidx[L - 1];
    if (Vold + n_new > H.cap) {           // cannot happen (poa_new_vertex_bound); keep the graph intact
        if (tid == 0) H.error = 1;
        x.sync();
        return;
    }
    // C: vertex ids (new ids in path order)
    for (int f = tid; f < L; f += T) {
        const PoaStep s = steps[L - 1 - f];
        cur[f] = isnew[f] ? Vold + newidx[f] - 1 : s.vertex;
        arank[f] = (anchor[f] >= 0) ? G.rank[o + steps[L - 1 - anchor[f]].vertex] : -1;
    }
    x.sync();
    // D: thread the read -- every step touches only its own vertex
    for (int f = tid; f < L; f += T) {
        const PoaStep s = steps[L - 1 - f];
        const int c = cur[f];
        const int prev = f > 0 ? cur[f - 1] : -1;
        if (isnew[f]) {
            G.meta[o + c] = poa_meta(R.base(s.readpos), prev >= 0 ? 1 : 0, 1);
            G.pred0[o + c] = prev;
            G.col[o + c] = (anchor[f] >= 0) ? G.col[o + steps[L - 1 - anchor[f]].vertex] : 0;
            if (f == L - 1 || !isnew[f + 1]) add[arank[f] + 1] = f - anchor[f];   // run length (anchor -1: f + 1)
        } else {
            uint32_t m = G.meta[o + c] + (1u << 8);       // nReads + 1
            const int nin = poa_meta_nin(m);
            if (prev >= 0) {
                // sort key: old vertex 2 * rank; new vertex 2 * (anchor rank) + 1 (it sits right behind its anchor)
                const int kp = isnew[f - 1] ? 2 * arank[f - 1] + 1 : 2 * G.rank[o + prev];
                bool have = false;
                int pos = 0;                               // number of predecessors sorting before prev
                for (int k = 0; k < nin; ++k) {
                    const int p = (k == 0) ? G.pred0[o + c] : G.predx[7 * (o + c) + k - 1];
                    if (p == prev) have = true;
                    if (2 * G.rank[o + p] < kp) ++pos;
                }
                if (!have && nin < kPoaMaxPred) {
                    for (int k = nin; k > pos; --k) {      // shift the tail up by one
                        const int p = (k - 1 == 0) ? G.pred0[o + c] : G.predx[7 * (o + c) + k - 2];
                        G.predx[7 * (o + c) + k - 1] = p;
                    }
                    if (pos == 0) G.pred0[o + c] = prev; else G.predx[7 * (o + c) + pos - 1] = prev;
                    m += (1u << 2);
                }
            }
            G.meta[o + c] = m;
        }
    }
    x.sync();
    // E: new ranks.  S[t] = new vertices anchored before old rank t
    poa_block_scan<X, false>(x, add, Vold + 1, sm);
    const int32_t* ord_old = poa_order(G, sel) + o;
    int32_t* ord_new = poa_order(G, sel ^ 1) + o;
    for (int t = tid; t < Vold; t += T) {
        const int id = ord_old[t];
        const int nr = t + add[t];
        ord_new[nr] = id;
        G.rank[o + id] = nr;
    }
    for (int f = tid; f < L; f += T) {
        if (!isnew[f]) continue;
        const int a = arank[f];
        const int k = f - anchor[f];                      // 1-based position in its run (head run: anchor = -1)
        const int nr = (a >= 0) ? a + add[a] + k : k - 1;
        ord_new[nr] = cur[f];
        G.rank[o + cur[f]] = nr;
    }
    x.sync();
    if (tid == 0) {
        H.V = Vold + n_new;
        H.order_sel = sel ^ 1;
        if (H.n_spans < kPoaMaxReads) { H.span_first[H.n_spans] = cur[0]; H.span_last[H.n_spans] = cur[L - 1]; H.n_spans++; }
        H.n_reads++;
    }
    x.sync();
}

// FindConsensus: score(v) = 2 * nReads - max(spanning reads, minCov), best-scoring path, ties to the lowest rank.
// scratch: 12 * V ints.  sh: kPoaConsShared ints of fast scratch (shared memory on the device).  Writes the consensus
// bases to out[0..*out_len); *out_len <= V.
//
// The best-path recurrence is a serial walk over the ranks; everything it needs is prepared in parallel first -- per rank
// the score, the kind of vertex and, for a general vertex, the RANKS of its predecessors -- and handed to the walking
// thread in chunks staged through `sh`, together with a window of the reach values it produced itself: the walk touches
// global memory only for a predecessor more than a chunk behind.
constexpr int kPoaConsChunk = 512;
constexpr int kPoaConsShared = kPoaConsChunk * (2 + kPoaMaxPred) + 2 * kPoaConsChunk;

template <class X>
CCS_HD void poa_graph_consensus(X& x, const PoaGraphView& G, int g, int32_t* scratch, uint8_t* out, int32_t* out_len,
                                int32_t* sh) {
    const PoaGraphHdr& H = G.hdr[g];
    const int64_t o = H.voff;
    const int V = H.V;
    const int n = H.n_reads;
    const int min_cov = n < 5 ? 1 : (n + 1) / 2 - 1;
    const int32_t* ord = poa_order(G, H.order_sel) + o;
    int32_t* sc = scratch;                    // [V] vertex score by rank
    int32_t* kind = scratch + V;              // [V] -1: single predecessor at rank t-1; 0: none; k > 0: k predecessors (general)
    int32_t* reach = scratch + 2 * (int64_t)V;
    int32_t* bp = scratch + 3 * (int64_t)V;
    int32_t* prk = scratch + 4 * (int64_t)V;  // [V][8] predecessor ranks of general vertices, in predecessor order
    const int T = x.nthreads(), tid = x.tid();
    for (int t = tid; t < V; t += T) {
        const int id = ord[t];
        const uint32_t m = G.meta[o + id];
        int cov = 0;
        for (int s = 0; s < H.n_spans; ++s)
            cov += (G.rank[o + H.span_first[s]] <= t && t <= G.rank[o + H.span_last[s]]) ? 1 : 0;
        sc[t] = 2 * poa_meta_nreads(m) - (cov > min_cov ? cov : min_cov);
        const int nin = poa_meta_nin(m);
        int kd = nin;
        if (nin == 1 && G.rank[o + G.pred0[o + id]] == t - 1) kd = -1;
        else
            for (int e = 0; e < nin; ++e) {
                const int p = (e == 0) ? G.pred0[o + id] : G.predx[7 * (o + id) + e - 1];
                prk[(int64_t)t * kPoaMaxPred + e] = G.rank[o + p];
            }
        kind[t] = kd;
    }
    x.sync();
    int32_t* s_sc = sh;                                        // [chunk]
    int32_t* s_kind = sh + kPoaConsChunk;                      // [chunk]
    int32_t* s_prk = sh + 2 * kPoaConsChunk;                   // [chunk][8]
    int32_t* s_reach = sh + (2 + kPoaMaxPred) * kPoaConsChunk; // [2 * chunk] reach of ranks >= c0 - chunk, slot = rank mod (2 chunk)
    int best = 0, bt = -1, prev = 0;                           // state of the walking thread
    for (int c0 = 0; c0 < V; c0 += kPoaConsChunk) {
        const int nc = (V - c0 < kPoaConsChunk) ? V - c0 : kPoaConsChunk;
        for (int j = tid; j < nc; j += T) {
            s_sc[j] = sc[c0 + j];
            const int kd = kind[c0 + j];
            s_kind[j] = kd;
            for (int e = 0; e < kd; ++e) s_prk[j * kPoaMaxPred + e] = prk[(int64_t)(c0 + j) * kPoaMaxPred + e];
        }
        x.sync();
        if (tid == 0) {
            for (int j = 0; j < nc; ++j) {
                const int t = c0 + j;
                int m = 0, mp = -1;
                const int k = s_kind[j];
                if (k < 0) { if (prev > 0) { m = prev; mp = t - 1; } }
                else
                    for (int e = 0; e < k; ++e) {              // predecessors are sorted by rank: first maximum = lowest rank
                        const int q = s_prk[j * kPoaMaxPred + e];
                        const int rq = (q >= c0 - kPoaConsChunk) ? s_reach[q & (2 * kPoaConsChunk - 1)] : reach[q];
                        if (rq > m) { m = rq; mp = q; }
                    }
                prev = s_sc[j] + m;
                s_reach[t & (2 * kPoaConsChunk - 1)] = prev;
                reach[t] = prev;
                bp[t] = mp;
                if (bt < 0 || prev > best) { best = prev; bt = t; }
            }
        }
        x.sync();
    }
    // Backtrack into sc[] (no longer needed), reversed.  Back pointers lead to lower ranks, mostly the next lower one: the
    // chunks are staged again from the last one down (back pointers and vertex ids), and the walking thread follows the
    // path through shared memory instead of chasing one global load per vertex.
    {
        int k = bt, len = 0;                                   // state of the walking thread
        int32_t* s_bp = s_kind;
        int32_t* s_id = s_prk;
        int32_t* s_stop = s_reach;                             // [0]: the path has ended (or never began)
        for (int c0 = (V > 0 ? ((V - 1) / kPoaConsChunk) * kPoaConsChunk : 0); c0 >= 0 && V > 0; c0 -= kPoaConsChunk) {
            const int nc = (V - c0 < kPoaConsChunk) ? V - c0 : kPoaConsChunk;
            for (int j = tid; j < nc; j += T) { s_bp[j] = bp[c0 + j]; s_id[j] = ord[c0 + j]; }
            x.sync();
            if (tid == 0) {
                while (k >= c0) { sc[len++] = s_id[k - c0]; k = s_bp[k - c0]; }
                s_stop[0] = (k < 0) ? 1 : 0;
            }
            x.sync();
            const int stop = s_stop[0];
            x.sync();
            if (stop) break;
        }
        if (tid == 0) *out_len = len;
    }
    x.sync();
    const int len = *out_len;
    for (int j = tid; j < len; j += T) out[j] = (uint8_t)poa_meta_base(G.meta[o + sc[len - 1 - j]]);
    x.sync();
}

// ---- k-mer orientation vote (content-sampled 11-mers, DESIGN.md "Draft stage") ---------------------------------
CCS_HD uint32_t poa_kmer_hash(uint32_t k) { k *= 0x9E3779B1u; return k ^ (k >> 15); }
CCS_HD bool poa_kmer_sampled(uint32_t k) { return ((k * 0x9E3779B1u) >> 29) == 0u; }

// table capacity (power of two) for a reference of n bases: ~n/8 sampled k-mers -> load <= 1/2
CCS_HD int poa_kmer_table_cap(int n) {
    int cap = 256;
    while (cap * 4 < n) cap <<= 1;
    return cap;
}

struct PoaBaseAcc {      // plain bases or emission codes
    const uint8_t* p;
    int is_codes;
    CCS_HD int base(int i) const { return is_codes ? (p[i] & 3) : p[i]; }
};

// Inserts the sampled k-mers of ref into tab (cap entries, zeroed by the caller; entry = k-mer + 1).
template <class X>
CCS_HD void poa_kmer_build(X& x, uint32_t* tab, int cap, const PoaBaseAcc& ref, int n) {
    const uint32_t kmask = (1u << (2 * kPoaKmer)) - 1, mask = (uint32_t)cap - 1;
    const int T = x.nthreads();
    // thread-contiguous chunks with a rolling k-mer
    const int npos = n - kPoaKmer + 1;                   // k-mer end positions kPoaKmer-1 .. n-1
    if (npos > 0) {
        const int per = (npos + T - 1) / T;
        const int b = x.tid() * per, e = (b + per < npos) ? b + per : npos;
        uint32_t k = 0;
        if (b < e) for (int i = b; i < b + kPoaKmer - 1; ++i) k = ((k << 2) | (uint32_t)ref.base(i)) & kmask;
        for (int q = b; q < e; ++q) {
            k = ((k << 2) | (uint32_t)ref.base(q + kPoaKmer - 1)) & kmask;
            if (!poa_kmer_sampled(k)) continue;
            uint32_t h = poa_kmer_hash(k) & mask;
            for (;;) {
                const uint32_t old = x.cas(&tab[h], 0u, k + 1);
                if (old == 0u || old == k + 1) break;
                h = (h + 1) & mask;
            }
        }
    }
    x.sync();
}

CCS_HD bool poa_kmer_has(const uint32_t* tab, int cap, uint32_t k) {
    const uint32_t mask = (uint32_t)cap - 1;
    uint32_t h = poa_kmer_hash(k) & mask;
    for (;;) {
        const uint32_t e = tab[h];
        if (e == 0u) return false;
        if (e == k + 1) return true;
        h = (h + 1) & mask;
    }
}

// Sampled k-mers of the read's first kPoaVoteBases bases (forward) and of their reverse complement that occur in the
// table, over k-mer end positions [b, e).
CCS_HD void poa_kmer_count(const uint32_t* tab, int cap, const uint8_t* codes, int b, int e, int& fwd, int& rev) {
    const uint32_t kmask = (1u << (2 * kPoaKmer)) - 1;
    uint32_t kf = 0, kr = 0;
    fwd = 0; rev = 0;
    if (b >= e) return;
    for (int i = b - (kPoaKmer - 1); i < b; ++i) {
        const uint32_t c = codes[i] & 3u;
        kf = ((kf << 2) | c) & kmask;
        kr = (kr >> 2) | ((3u - c) << (2 * (kPoaKmer - 1)));
    }
    for (int i = b; i < e; ++i) {
        const uint32_t c = codes[i] & 3u;
        kf = ((kf << 2) | c) & kmask;
        kr = (kr >> 2) | ((3u - c) << (2 * (kPoaKmer - 1)));
        if (poa_kmer_sampled(kf)) fwd += poa_kmer_has(tab, cap, kf) ? 1 : 0;
        if (poa_kmer_sampled(kr)) rev += poa_kmer_has(tab, cap, kr) ? 1 : 0;
    }
}

}  // namespace ccs
