// TEST INFRASTRUCTURE -- C ABI over the CPU oracle for ctypes (tests/, smoke(), bench.py
// cpu_baseline / --impl reference only).  See arrow_oracle.h for provenance.
#include "arrow_oracle.h"
#include "pipeline_oracle.h"
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>

using namespace oracle;

namespace {
int g_margin = 2, g_edge_log2 = -60;  // test knobs: band rule overrides (defaults = spec)
template <class Real>
static int fill_impl(const void* model, const float* snr, const uint8_t* tpl, int J, const uint8_t* codes, int I, int W,
                     double* ll_alpha, double* ll_beta, int64_t* cells, float* alpha_out, float* beta_out,
                     int32_t* start_out, int64_t* aexp_out, int64_t* bexp_out) {
    Tables t;
    t.build(*(const ccs::ArrowModelParams*)model, snr);
    Recursor<Real> r;
    r.margin = g_margin; r.edge_log2 = g_edge_log2;
    r.fill(&t, tpl, J, codes, I, W);
    *ll_alpha = r.ll_alpha;
    *ll_beta = r.ll_beta;
    *cells = r.cells;
    if (r.status == READ_VALID || r.status == READ_DEAD) {
        if (alpha_out) for (size_t k = 0; k < r.alpha.v.size(); ++k) alpha_out[k] = (float)r.alpha.v[k];
        if (beta_out) for (size_t k = 0; k < r.beta.v.size(); ++k) beta_out[k] = (float)r.beta.v[k];
        if (start_out) std::memcpy(start_out, r.alpha.start.data(), sizeof(int32_t) * r.alpha.start.size());
        if (aexp_out) std::memcpy(aexp_out, r.alpha.cumexp.data(), sizeof(int64_t) * r.alpha.cumexp.size());
        if (bexp_out && !r.beta.cumexp.empty()) std::memcpy(bexp_out, r.beta.cumexp.data(), sizeof(int64_t) * r.beta.cumexp.size());
    }
    return r.status;
}

template <class Real>
static int score_impl(const void* model, const float* snr, const uint8_t* tpl, int J, const uint8_t* codes, int I, int W,
                      int nmut, const int32_t* types, const int32_t* pos, const int32_t* bases, double* ll_inc, double* ll_full) {
    Tables t;
    t.build(*(const ccs::ArrowModelParams*)model, snr);
    Recursor<Real> r;
    r.fill(&t, tpl, J, codes, I, W);
    if (r.status != READ_VALID) return r.status;
    std::vector<uint8_t> base(tpl, tpl + J);
    for (int k = 0; k < nmut; ++k) {
        Mutation m{types[k], pos[k], bases[k]};
        ll_inc[k] = r.ll_mutated(m);
        if (ll_full) {
            std::vector<Mutation> one(1, m);
            std::vector<uint8_t> mt = apply_mutations(base, one);
            Recursor<Real> f;
            f.fill(&t, mt.data(), (int)mt.size(), codes, I, W);
            ll_full[k] = f.ll_alpha;
        }
    }
    return 0;
}

// One ZMW: Integrator + Polish + ConsensusQualities on a given draft with mapped reads.
// stats[8]: converged, iterations, n_tested, n_applied, n_active, cells(total of last fill), 0, 0
// delta_out (optional): [ (J0+1) * 9 ] doubles of the FIRST scoring round on the draft, slot layout
// {SUB A,C,G,T, DEL, INS A,C,G,T}; NaN where not scored.
template <class Real>
static int polish_impl(const void* model, const float* snr, const uint8_t* draft, int J, int nreads,
                       const uint8_t* codes, const int64_t* read_off, const int32_t* strand, const int32_t* tstart,
                       const int32_t* tend, int W, int max_iter, uint8_t* cons, int cons_cap, int32_t* cons_len,
                       uint8_t* qv, double* rq, int64_t* stats, double* read_ll, int32_t* read_status) {
    PolishConfig cfg;
    cfg.band_width = W;
    if (max_iter >= 0) cfg.max_iterations = max_iter;
    Integrator<Real> ai;
    ai.init(*(const ccs::ArrowModelParams*)model, snr, draft, J, cfg);
    for (int r = 0; r < nreads; ++r) {
        MappedRead mr;
        mr.codes.assign(codes + read_off[r], codes + read_off[r + 1]);
        mr.strand = strand[r]; mr.tstart = tstart[r]; mr.tend = tend[r];
        ai.add_read(mr);
    }
    PolishResult pr{};
    if (max_iter != 0) pr = polish(ai);
    std::vector<uint8_t> q;
    consensus_qvs(ai, q);
    const int Jc = (int)ai.fwd.size();
    *cons_len = Jc;
    if (Jc > cons_cap) return -1;
    std::memcpy(cons, ai.fwd.data(), Jc);
    std::memcpy(qv, q.data(), Jc);
    *rq = predicted_accuracy(q);
    int64_t cells = 0;
    for (int r = 0; r < nreads; ++r) {
        read_ll[r] = ai.active[r] ? ai.recs[r].ll() : NAN;
        read_status[r] = ai.recs[r].status;
        cells += ai.recs[r].cells;
    }
    stats[0] = pr.converged; stats[1] = pr.iterations; stats[2] = pr.n_tested; stats[3] = pr.n_applied;
    stats[4] = ai.n_active(); stats[5] = cells; stats[6] = stats[7] = 0;
    return 0;
}
}  // namespace

extern "C" {

void oracle_set_band_rule(int margin, int edge_log2) { g_margin = margin; g_edge_log2 = edge_log2; }

int oracle_model_sizeof() { return (int)sizeof(ccs::ArrowModelParams); }
void oracle_synthetic_model(void* out) { ccs::synthetic_model(*(ccs::ArrowModelParams*)out); }

// em_match[36*16], em_ins[17*16], tr[36*4], log_cw
void oracle_get_tables(const void* model, const float* snr, double* em_match, double* em_ins, double* tr, double* log_cw) {
    Tables t;
    t.build(*(const ccs::ArrowModelParams*)model, snr);
    std::memcpy(em_match, t.em_match, sizeof(t.em_match));
    std::memcpy(em_ins, t.em_ins, sizeof(t.em_ins));
    std::memcpy(tr, t.tr, sizeof(t.tr));
    *log_cw = t.log_cw;
}

// folded factors the recursion uses: mm[36*16] = fl32(em_match * match), gg[17*16] = fl32(em_ins * (branch|stick))
void oracle_get_folded(const void* model, const float* snr, double* mm, double* gg) {
    Tables t;
    t.build(*(const ccs::ArrowModelParams*)model, snr);
    std::memcpy(mm, t.mm, sizeof(t.mm));
    std::memcpy(gg, t.gg, sizeof(t.gg));
}


// Recursor::FillAlphaBeta of one (read, template) pair.  precision: 0 = double, 1 = float cells.
// Optional dumps (may be NULL): alpha/beta cells [J*W] (slot = row mod W), start[J], cumexp[J].
int oracle_fill(const void* model, const float* snr, const uint8_t* tpl, int J, const uint8_t* codes, int I, int W,
                int precision, double* ll_alpha, double* ll_beta, int64_t* cells, float* alpha_out, float* beta_out,
                int32_t* start_out, int64_t* aexp_out, int64_t* bexp_out) {
    return precision ? fill_impl<float>(model, snr, tpl, J, codes, I, W, ll_alpha, ll_beta, cells, alpha_out, beta_out, start_out, aexp_out, bexp_out)
                     : fill_impl<double>(model, snr, tpl, J, codes, I, W, ll_alpha, ll_beta, cells, alpha_out, beta_out, start_out, aexp_out, bexp_out);
}


// Evaluator::LL(Mutation) incrementally (ll_inc) and by refilling on the mutated template (ll_full).
int oracle_score(const void* model, const float* snr, const uint8_t* tpl, int J, const uint8_t* codes, int I, int W,
                 int precision, int nmut, const int32_t* types, const int32_t* pos, const int32_t* bases, double* ll_inc,
                 double* ll_full) {
    return precision ? score_impl<float>(model, snr, tpl, J, codes, I, W, nmut, types, pos, bases, ll_inc, ll_full)
                     : score_impl<double>(model, snr, tpl, J, codes, I, W, nmut, types, pos, bases, ll_inc, ll_full);
}


int oracle_polish(const void* model, const float* snr, const uint8_t* draft, int J, int nreads, const uint8_t* codes,
                  const int64_t* read_off, const int32_t* strand, const int32_t* tstart, const int32_t* tend, int W,
                  int max_iter, int precision, uint8_t* cons, int cons_cap, int32_t* cons_len, uint8_t* qv, double* rq,
                  int64_t* stats, double* read_ll, int32_t* read_status) {
    return precision ? polish_impl<float>(model, snr, draft, J, nreads, codes, read_off, strand, tstart, tend, W, max_iter, cons, cons_cap, cons_len, qv, rq, stats, read_ll, read_status)
                     : polish_impl<double>(model, snr, draft, J, nreads, codes, read_off, strand, tstart, tend, W, max_iter, cons, cons_cap, cons_len, qv, rq, stats, read_ll, read_status);
}

// Integrator::LL(Mutation) for every slot of every position: out[(J+1)*9], slots
// {SUB A,C,G,T, DEL, INS A,C,G,T}; entries no active read covers (or SUB of the same base) are 0.
int oracle_score_all(const void* model, const float* snr, const uint8_t* draft, int J, int nreads, const uint8_t* codes,
                     const int64_t* read_off, const int32_t* strand, const int32_t* tstart, const int32_t* tend, int W,
                     double* out, double* read_ll) {
    PolishConfig cfg;
    cfg.band_width = W;
    Integrator<double> ai;
    ai.init(*(const ccs::ArrowModelParams*)model, snr, draft, J, cfg);
    for (int r = 0; r < nreads; ++r) {
        MappedRead mr;
        mr.codes.assign(codes + read_off[r], codes + read_off[r + 1]);
        mr.strand = strand[r]; mr.tstart = tstart[r]; mr.tend = tend[r];
        ai.add_read(mr);
    }
    for (int r = 0; r < nreads; ++r) read_ll[r] = ai.active[r] ? ai.recs[r].ll() : NAN;
    for (int p = 0; p <= J; ++p)
        for (int s = 0; s < 9; ++s) {
            double v = 0;
            Mutation m{s < 4 ? MUT_SUB : (s == 4 ? MUT_DEL : MUT_INS), p, s < 4 ? s : (s == 4 ? 0 : s - 5)};
            const bool ok = (m.type == MUT_INS) ? (p >= 1 && p <= J - 1) : (p < J && !(m.type == MUT_SUB && m.base == draft[p]));
            if (ok) v = ai.delta_ll(m);
            out[(size_t)p * 9 + s] = v;
        }
    return 0;
}


// ---- draft stage / whole pipeline ------------------------------------------------------
// cfg_i[8] = {min_passes, top_passes, max_poa_reads, min_length, max_length, max_iterations(-1 default),
//            window_size (<0: default; rounded up to a multiple of 64), window_overlap (likewise)}
// cfg_d[4] = {min_snr, min_rq, min_active_fraction, min_zscore}
static CcsConfig make_cfg(const int32_t* ci, const double* cd) {
    CcsConfig c;
    if (ci) {
        c.min_passes = ci[0]; c.top_passes = ci[1]; c.max_poa_reads = ci[2]; c.min_length = ci[3]; c.max_length = ci[4];
        if (ci[5] >= 0) c.polish.max_iterations = ci[5];
        if (ci[6] >= 0) c.window_size = (ci[6] + WINDOW_GRID - 1) / WINDOW_GRID * WINDOW_GRID;
        if (ci[7] >= 0) c.window_overlap = (ci[7] + WINDOW_GRID - 1) / WINDOW_GRID * WINDOW_GRID;
    }
    if (cd) { c.min_snr = cd[0]; c.min_rq = cd[1]; c.min_active_fraction = cd[2]; c.polish.min_zscore = cd[3]; }
    return c;
}

// Draft stage only: draft + per-read mapping.  maps[r*6..] = {mapped, strand, tstart, tend, rstart, rend}
int oracle_draft_zmw(const int32_t* cfg_i, const double* cfg_d, const float* snr, int nreads, const uint8_t* codes,
                     const int64_t* read_off, const uint8_t* cx, uint8_t* draft, int draft_cap, int32_t* draft_len,
                     int32_t* maps, int32_t* status) {
    CcsConfig cfg = make_cfg(cfg_i, cfg_d);
    CcsZmwResult res;
    std::vector<char> keep;
    draft_zmw(cfg, nreads, codes, read_off, cx, snr, res, keep);
    *status = res.status;
    *draft_len = (int32_t)res.draft.size();
    if ((int)res.draft.size() > draft_cap) return -1;
    if (!res.draft.empty()) std::memcpy(draft, res.draft.data(), res.draft.size());
    for (int r = 0; r < nreads; ++r) {
        const ReadMapping& m = res.maps[r];
        int32_t* o = maps + 6 * r;
        o[0] = m.mapped; o[1] = m.strand; o[2] = m.tstart; o[3] = m.tend; o[4] = m.rstart; o[5] = m.rend;
    }
    return 0;
}

// Whole per-ZMW path.  stats[8] = {status, np, converged, iterations, n_tested, n_applied, draft_len, 0}
int oracle_ccs_zmw(const void* model, const int32_t* cfg_i, const double* cfg_d, const float* snr, int nreads,
                   const uint8_t* codes, const int64_t* read_off, const uint8_t* cx, uint8_t* seq, int seq_cap,
                   int32_t* seq_len, uint8_t* qv, double* rq, int64_t* stats, double* read_ll, int32_t* read_status) {
    CcsConfig cfg = make_cfg(cfg_i, cfg_d);
    CcsZmwResult res;
    ccs_zmw(*(const ccs::ArrowModelParams*)model, cfg, nreads, codes, read_off, cx, snr, res);
    *seq_len = (int32_t)res.seq.size();
    if ((int)res.seq.size() > seq_cap) return -1;
    if (!res.seq.empty()) { std::memcpy(seq, res.seq.data(), res.seq.size()); std::memcpy(qv, res.qv.data(), res.qv.size()); }
    *rq = res.rq;
    stats[0] = res.status; stats[1] = res.np; stats[2] = res.pr.converged; stats[3] = res.pr.iterations;
    stats[4] = res.pr.n_tested; stats[5] = res.pr.n_applied; stats[6] = (int64_t)res.draft.size(); stats[7] = 0;
    for (int r = 0; r < nreads; ++r) { read_ll[r] = res.read_ll[r]; read_status[r] = res.read_status[r]; }
    return 0;
}

}  // extern "C"
