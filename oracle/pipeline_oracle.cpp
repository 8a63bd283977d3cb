// TEST INFRASTRUCTURE -- see pipeline_oracle.h.
#include "pipeline_oracle.h"
#include <algorithm>
#include <cmath>

namespace oracle {

int filter_reads(const std::vector<int>& lens, const uint8_t* cx, int top_passes, std::vector<char>& keep) {
    const int n = (int)lens.size();
    keep.assign(n, 0);
    if (n == 0) return 0;
    std::vector<int> s(lens);
    std::sort(s.begin(), s.end());
    const int median = (n & 1) ? s[n / 2] : (s[n / 2 - 1] + s[n / 2]) / 2;
    int nfull = 0;
    for (int r = 0; r < n; ++r) {
        if (2 * lens[r] < median || lens[r] > 2 * median) continue;   // <50 % or >200 % of the median
        const bool full = (cx[r] & 3) == 3;
        if (full) {
            if (nfull >= top_passes) continue;
            ++nfull;
        }
        keep[r] = 1;
    }
    return nfull;
}

static std::vector<uint8_t> bases_of(const uint8_t* codes, int n, bool rc) {
    std::vector<uint8_t> b(n);
    if (!rc) for (int i = 0; i < n; ++i) b[i] = codes[i] & 3;
    else for (int i = 0; i < n; ++i) b[i] = (uint8_t)(3 - (codes[n - 1 - i] & 3));
    return b;
}

void draft_zmw(const CcsConfig& cfg, int nreads, const uint8_t* codes, const int64_t* read_off, const uint8_t* cx,
               const float snr[4], CcsZmwResult& out, std::vector<char>& keep) {
    out = CcsZmwResult();
    out.maps.assign(nreads, ReadMapping());
    keep.assign(nreads, 0);
    if (nreads == 0) { out.status = Z_NO_SUBREADS; return; }
    if (std::min(std::min(snr[0], snr[1]), std::min(snr[2], snr[3])) < cfg.min_snr) { out.status = Z_POOR_SNR; return; }
    std::vector<int> lens(nreads);
    for (int r = 0; r < nreads; ++r) lens[r] = (int)(read_off[r + 1] - read_off[r]);
    const int nfull = filter_reads(lens, cx, cfg.top_passes, keep);
    if (nfull < cfg.min_passes) { out.status = Z_TOO_FEW_PASSES; return; }
    // Draft cascade (docs/faq/accuracy-vs-passes.md:41-46): generator 0 = SparsePoa over the first max_poa_reads
    // full-length reads, seeded by the first one; if its draft cannot be generated or too few subreads map back to it,
    // generator 1 = SparsePoa seeded by the full-length read closest to the median length, over up to
    // 2 * max_poa_reads - 1 reads in order of closeness (the seed that just failed is left out).  Length gates are final.
    std::vector<int> full;
    for (int r = 0; r < nreads; ++r) if (keep[r] && (cx[r] & 3) == 3) full.push_back(r);
    int kept = 0;
    for (int r = 0; r < nreads; ++r) kept += keep[r] ? 1 : 0;
    std::vector<int> first_sel;
    for (int gen = 0; gen < 2; ++gen) {
        std::vector<int> sel;
        if (gen == 0) {
            for (int r : full) if ((int)sel.size() < std::max(1, std::min(cfg.max_poa_reads, 16))) sel.push_back(r);
            first_sel = sel;
        } else {
            std::vector<int> sl;
            for (int r : full) sl.push_back(lens[r]);
            std::sort(sl.begin(), sl.end());
            const int med = sl.empty() ? 0 : sl[sl.size() / 2];
            for (int r : full) if (first_sel.empty() || r != first_sel[0]) sel.push_back(r);   // not the seed that just failed
            std::stable_sort(sel.begin(), sel.end(), [&](int a, int b) { return std::abs(lens[a] - med) < std::abs(lens[b] - med); });
            const size_t cap = (size_t)std::max(1, std::min(2 * cfg.max_poa_reads - 1, 16));
            if (sel.size() > cap) sel.resize(cap);
            if (sel.empty()) break;               // nothing new to try
        }
        out.draft.clear();
        out.maps.assign(nreads, ReadMapping());
        if (sel.empty()) { out.status = Z_DRAFT_FAILURE; return; }
        PoaGraph g;
        std::vector<uint8_t> seed;
        for (size_t k = 0; k < sel.size(); ++k) {
            const int r = sel[k];
            const uint8_t* c = codes + read_off[r];
            if (k == 0) {
                seed = bases_of(c, lens[r], false);
                g.add_first(seed.data(), lens[r]);
            } else {
                std::vector<uint8_t> fwd = bases_of(c, lens[r], false);
                const bool rev = kmer_vote_reverse(seed.data(), (int)seed.size(), fwd.data(), lens[r]);
                std::vector<uint8_t> b = rev ? bases_of(c, lens[r], true) : fwd;
                PoaAlignment a = g.align(b.data(), lens[r]);
                if (a.score >= lens[r]) g.commit(a, b.data());
            }
        }
        const int n = g.n_reads;
        const int min_cov = n < 5 ? 1 : (n + 1) / 2 - 1;
        std::vector<int> path = g.consensus(min_cov);
        out.draft.resize(path.size());
        for (size_t k = 0; k < path.size(); ++k) out.draft[k] = g.v[path[k]].base;
        const int J = (int)out.draft.size();
        if (J == 0) { out.status = Z_DRAFT_FAILURE; continue; }
        if (J < cfg.min_length) { out.status = Z_TOO_SHORT; return; }
        if (cfg.max_length > 0 && J > cfg.max_length) { out.status = Z_TOO_LONG; return; }
        // subread -> draft mapping of every kept read
        int mapped_full = 0, mapped = 0;
        for (int r = 0; r < nreads; ++r) {
            if (!keep[r]) continue;
            const uint8_t* c = codes + read_off[r];
            std::vector<uint8_t> fwd = bases_of(c, lens[r], false);
            const bool rev = kmer_vote_reverse(out.draft.data(), J, fwd.data(), lens[r]);
            std::vector<uint8_t> b = rev ? bases_of(c, lens[r], true) : fwd;
            ReadMapping m = map_to_template(out.draft.data(), J, b.data(), lens[r]);
            m.strand = rev ? 1 : 0;
            if (rev) { const int rs = lens[r] - m.rend, re = lens[r] - m.rstart; m.rstart = rs; m.rend = re; }
            if (m.mapped && (m.tend - m.tstart < 2 || m.rend - m.rstart < 2)) m.mapped = false;
            out.maps[r] = m;
            if (m.mapped) ++mapped;
            if (m.mapped && (cx[r] & 3) == 3) ++mapped_full;
        }
        // fewer than --min-passes full-length reads, or not more than half of the subreads, map back to the draft
        // (docs/faq/accuracy-vs-passes.md:31-39)
        if (mapped_full < cfg.min_passes || 2 * mapped <= kept) { out.status = Z_TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT; continue; }
        out.status = Z_SUCCESS;   // draft stage passed
        return;
    }
}


std::vector<Window> make_windows(int J, int W, int OV) {
    std::vector<Window> ws;
    if (W <= 0 || J < 2 * W) { ws.push_back(Window{0, J, 0, J}); return ws; }
    const int n = J / W;
    for (int k = 0; k < n; ++k) {
        Window w;
        w.c0 = k * W;
        w.c1 = (k == n - 1) ? J : (k + 1) * W;
        w.a = std::max(0, w.c0 - OV);
        w.b = std::min(J, w.c1 + OV);
        ws.push_back(w);
    }
    return ws;
}

std::vector<WindowRead> window_reads(const Window& w, const std::vector<ReadMapping>& maps, const std::vector<int>& lens) {
    std::vector<WindowRead> out;
    for (int r = 0; r < (int)maps.size(); ++r) {
        const ReadMapping& m = maps[r];
        if (!m.mapped) continue;
        if (std::min(m.tend, w.c1) - std::max(m.tstart, w.c0) < 1) continue;      // does not reach the core
        const int lo = std::max(w.a, m.tstart), hi = std::min(w.b, m.tend);
        if (hi - lo < 2) continue;
        const int n = lens[r];
        const int rs_o = m.strand ? n - m.rend : m.rstart, re_o = m.strand ? n - m.rstart : m.rend;   // oriented read
        const int s_o = (lo == m.tstart) ? rs_o : m.grid[lo / WINDOW_GRID];
        const int e_o = (hi == m.tend) ? re_o : m.grid[hi / WINDOW_GRID];
        if (e_o - s_o < 2 || s_o < 0 || e_o > n) continue;
        WindowRead x;
        x.parent = r; x.strand = m.strand; x.ts = lo - w.a; x.te = hi - w.a;
        x.ns = m.strand ? n - e_o : s_o; x.ne = m.strand ? n - s_o : e_o;                              // native slice
        out.push_back(x);
    }
    return out;
}

// Polish Stage of one ZMW: every window is an Arrow problem of its own (template = the padded draft slice, reads = the
// slices of the mapped subreads that the subread -> draft alignment places on it); the polished cores are concatenated
// (docs/how-does-ccs-work.md:57-61,108-110).  A read is judged (POOR_ZSCORE) on the sum over its windows.
void ccs_zmw(const ccs::ArrowModelParams& model, const CcsConfig& cfg, int nreads, const uint8_t* codes,
             const int64_t* read_off, const uint8_t* cx, const float snr[4], CcsZmwResult& out) {
    std::vector<char> keep;
    draft_zmw(cfg, nreads, codes, read_off, cx, snr, out, keep);
    out.read_ll.assign(nreads, NAN);
    out.read_status.assign(nreads, 4);
    if (out.status != Z_SUCCESS) return;
    const int J = (int)out.draft.size();
    const std::vector<Window> wins = make_windows(J, cfg.window_size, cfg.window_overlap);
    const int nw = (int)wins.size();
    PolishConfig pc = cfg.polish;
    pc.min_zscore = -1e300;                       // the z-score is taken over all windows of a read, below
    pc.growth_min = nw > 1 ? 128 : 512;
    std::vector<int> lens(nreads);
    for (int r = 0; r < nreads; ++r) lens[r] = (int)(read_off[r + 1] - read_off[r]);
    std::vector<Integrator<double>> ais((size_t)nw);
    std::vector<std::vector<int>> parent((size_t)nw);
    for (int k = 0; k < nw; ++k) {
        const Window& w = wins[k];
        Integrator<double>& ai = ais[k];
        ai.init(model, snr, out.draft.data() + w.a, w.b - w.a, pc);
        ai.mark_b = w.c0 - w.a;
        ai.mark_e = w.c1 - w.a;
        for (const WindowRead& x : window_reads(w, out.maps, lens)) {
            MappedRead mr;
            mr.codes.assign(codes + read_off[x.parent] + x.ns, codes + read_off[x.parent] + x.ne);
            mr.strand = x.strand; mr.tstart = x.ts; mr.tend = x.te;
            ai.add_read(mr);
            parent[k].push_back(x.parent);
        }
        if (parent[k].empty()) { out.status = Z_EMPTY_WINDOW_DURING_POLISHING; return; }
    }
    // POOR_ZSCORE (Integrator::AddRead): z of the read's summed log-likelihood against its summed expectation
    {
        std::vector<double> ll(nreads, 0.0), mean(nreads, 0.0), var(nreads, 0.0);
        std::vector<int> cnt(nreads, 0);
        for (int k = 0; k < nw; ++k)
            for (size_t x = 0; x < parent[k].size(); ++x) {
                if (!ais[k].active[x]) continue;
                double mu, va;
                ais[k].zmoments(x, mu, va);
                const int r = parent[k][x];
                ll[r] += ais[k].recs[x].ll(); mean[r] += mu; var[r] += va; ++cnt[r];
            }
        for (int k = 0; k < nw; ++k)
            for (size_t x = 0; x < parent[k].size(); ++x) {
                const int r = parent[k][x];
                if (ais[k].active[x] && cnt[r] > 0 && (ll[r] - mean[r]) / std::sqrt(var[r]) < cfg.polish.min_zscore) {
                    ais[k].active[x] = 0;
                    ais[k].recs[x].status = READ_POOR_ZSCORE;
                }
            }
    }
    auto usable = [&]() {
        for (int k = 0; k < nw; ++k) {
            const int a = ais[k].n_active();
            if (!(a > 0 && a >= cfg.min_active_fraction * (double)parent[k].size())) return false;
        }
        return true;
    };
    bool failed = !usable();
    out.pr = PolishResult();
    if (!failed) {
        out.pr.converged = true;
        for (int k = 0; k < nw; ++k) {
            const PolishResult p = polish(ais[k]);
            out.pr.converged = out.pr.converged && p.converged;
            out.pr.iterations = std::max(out.pr.iterations, p.iterations);
            out.pr.n_tested += p.n_tested;
            out.pr.n_applied += p.n_applied;
        }
        failed = !usable();
    }
    // per read: the first failure among its windows, else the sum of its window log-likelihoods
    {
        std::vector<int> seen(nreads, 0);
        for (int k = 0; k < nw; ++k)
            for (size_t x = 0; x < parent[k].size(); ++x) {
                const int r = parent[k][x];
                const int st = ais[k].recs[x].status;
                if (!seen[r]) { seen[r] = 1; out.read_status[r] = READ_VALID; out.read_ll[r] = 0.0; }
                if (out.read_status[r] != READ_VALID) continue;
                if (st != READ_VALID || !ais[k].active[x]) { out.read_status[r] = st; out.read_ll[r] = NAN; }
                else out.read_ll[r] += ais[k].recs[x].ll();
            }
    }
    if (failed) { out.status = Z_TOO_MANY_UNUSABLE; return; }
    out.seq.clear(); out.qv.clear();
    for (int k = 0; k < nw; ++k) {
        std::vector<uint8_t> q;
        consensus_qvs(ais[k], q);
        const int b = std::max(0, ais[k].mark_b), e = std::min((int)ais[k].fwd.size(), ais[k].mark_e);
        if (e > b) {
            out.seq.insert(out.seq.end(), ais[k].fwd.begin() + b, ais[k].fwd.begin() + e);
            out.qv.insert(out.qv.end(), q.begin() + b, q.begin() + e);
        }
    }
    out.rq = predicted_accuracy(out.qv);
    out.np = 0;
    for (int r = 0; r < nreads; ++r) if (out.read_status[r] == READ_VALID && (cx[r] & 3) == 3) ++out.np;
    const int Jc = (int)out.seq.size();
    if (!out.pr.converged) out.status = Z_NON_CONVERGENT;
    else if (Jc < cfg.min_length) out.status = Z_TOO_SHORT;
    else if (cfg.max_length > 0 && Jc > cfg.max_length) out.status = Z_TOO_LONG;
    else if (out.rq < cfg.min_rq) out.status = Z_POOR_QUALITY;
    else out.status = Z_SUCCESS;
}

}  // namespace oracle
