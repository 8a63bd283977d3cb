// TEST INFRASTRUCTURE -- see arrow_oracle.h for scope, provenance and the "parity unpinned" note.
#include "arrow_oracle.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <set>

namespace oracle {

static const double LN2 = 0.6931471805599453094;
static const double NEG_INF = -std::numeric_limits<double>::infinity();

// ---------------------------------------------------------------------------------------
// Tables: restates ModelConfig::Populate (docs/how-does-ccs-work.md:90-94)
// ---------------------------------------------------------------------------------------
void Tables::build(const ccs::ArrowModelParams& m, const float snr[4]) {
    std::memset(this, 0, sizeof(*this));
    log_cw = std::log(m.counter_weight);
    for (int ctx = 0; ctx < 16; ++ctx) {
        const int ch = ctx & 3;
        double s = (double)snr[ch];
        if (s < m.snr_lo[ch]) s = m.snr_lo[ch];
        if (s > m.snr_hi[ch]) s = m.snr_hi[ch];
        double x[3], denom = 1.0;
        for (int t = 0; t < 3; ++t) {
            const double* c = m.trans[ctx][t];
            x[t] = std::exp(c[0] + s * (c[1] + s * (c[2] + s * c[3])));
            denom += x[t];
        }
        tr[ctx][0] = (double)(float)(1.0 / denom);
        tr[ctx][1] = (double)(float)(x[ccs::TR_DELETION] / denom);
        tr[ctx][2] = (double)(float)(x[ccs::TR_BRANCH] / denom);
        tr[ctx][3] = (double)(float)(x[ccs::TR_STICK] / denom);
        for (int code = 0; code < 12; ++code) {
            const double em = (double)(float)(m.counter_weight * m.emission[ccs::MOVE_MATCH][ctx][code]);
            em_match[ctx][code] = em;
            em_match[CTX_END + ctx][code] = em;
            const bool cognate = (code & 3) == (ctx & 3);
            em_ins[ctx][code] =
                (double)(float)(m.counter_weight * m.emission[cognate ? ccs::MOVE_BRANCH : ccs::MOVE_STICK][ctx][code]);
        }
    }
    for (int b = 0; b < 4; ++b)
        for (int code = 0; code < 12; ++code)
            em_match[CTX_START + b][code] = (double)(float)(m.counter_weight * m.emission[ccs::MOVE_MATCH][5 * b][code]);
    for (int r = CTX_START; r < N_MROWS; ++r) tr[r][0] = 1.0;
    // folded factors: one fp32 multiply of the fp32 table values (no contraction: -ffp-contract=off)
    for (int r = 0; r < N_MROWS; ++r)
        for (int code = 0; code < CODE_STRIDE; ++code) {
            const float p = (float)em_match[r][code] * (float)tr[r][0];
            mm[r][code] = (double)p;
        }
    for (int ctx = 0; ctx < 16; ++ctx)
        for (int code = 0; code < CODE_STRIDE; ++code) {
            const bool cognate = (code & 3) == (ctx & 3);
            const float p = (float)em_ins[ctx][code] * (float)(cognate ? tr[ctx][2] : tr[ctx][3]);
            gg[ctx][code] = (double)p;
        }
    // z-score moments from the model's own probabilities in double (no fp32 rounding, no counter weight)
    for (int ctx = 0; ctx < 16; ++ctx) {
        const int ch = ctx & 3;
        double s = (double)snr[ch];
        if (s < m.snr_lo[ch]) s = m.snr_lo[ch];
        if (s > m.snr_hi[ch]) s = m.snr_hi[ch];
        double x[3], denom = 1.0;
        for (int t = 0; t < 3; ++t) {
            const double* c = m.trans[ctx][t];
            x[t] = std::exp(c[0] + s * (c[1] + s * (c[2] + s * c[3])));
            denom += x[t];
        }
        const double pM = 1.0 / denom, pD = x[ccs::TR_DELETION] / denom, pB = x[ccs::TR_BRANCH] / denom, pS = x[ccs::TR_STICK] / denom;
        const double q = pB + pS;                      // another insertion
        double ex = 0, ex2 = 0, ey = 0, ey2 = 0;
        for (int code = 0; code < 12; ++code) {
            const double pb = pB * m.emission[ccs::MOVE_BRANCH][ctx][code], ps = pS * m.emission[ccs::MOVE_STICK][ctx][code];
            const double pm = pM * m.emission[ccs::MOVE_MATCH][ctx][code];
            if (pb > 0) { const double l = std::log(pb); ex += pb * l; ex2 += pb * l * l; }
            if (ps > 0) { const double l = std::log(ps); ex += ps * l; ex2 += ps * l * l; }
            if (pm > 0) { const double l = std::log(pm); ey += pm * l; ey2 += pm * l * l; }
        }
        { const double l = std::log(pD); ey += pD * l; ey2 += pD * l * l; }
        ex /= q; ex2 /= q; ey /= (1.0 - q); ey2 /= (1.0 - q);
        const double en = q / (1.0 - q), vn = q / ((1.0 - q) * (1.0 - q));
        zs_mean[ctx] = en * ex + ey;
        zs_var[ctx] = en * (ex2 - ex * ex) + vn * ex * ex + (ey2 - ey * ey);
    }
    for (int b = 0; b < 4; ++b) {
        double e1 = 0, e2 = 0;
        for (int code = 0; code < 12; ++code) {
            const double p = m.emission[ccs::MOVE_MATCH][5 * b][code];
            if (p > 0) { const double l = std::log(p); e1 += p * l; e2 += p * l * l; }
        }
        zs_first_mean[b] = e1;
        zs_first_var[b] = e2 - e1 * e1;
    }
}

// Column normalisation + band tracking (DESIGN.md "Band rule").
//  * scale: divide by 2^k, k = fp32 exponent of the column maximum (exact, power of two);
//  * edge:  largest row whose UNSCALED value is >= 2^edge_log2 (the leading edge of the
//           probability mass; the forward mass trails behind the true path inside insertion
//           bursts, so the band is anchored at its leading edge, not at its maximum).
//  * schedule: only columns j with j % kScaleEvery == 0 are rescaled (k = 0 elsewhere): the values stay far inside
//           the fp32 range over three unscaled columns, and the kernels save the octet-wide maximum on those columns.
constexpr int kScaleEvery = 4;
template <class Real>
static void scale_column(Real* col, int W, int s, int edge_log2, int& edge_row, int& k_out, bool& dead, bool rescale) {
    double best_val = 0.0;
    const double thr = std::ldexp(1.0, edge_log2);
    edge_row = s - 1;
    for (int rel = 0; rel < W; ++rel) {
        const int row = s + rel;
        const double a = (double)col[row % W];
        if (a > best_val) best_val = a;
        if (a >= thr) edge_row = row;
    }
    dead = !(best_val > 0.0);
    int k = 0;
    if (!dead && rescale) {
        float f = (float)best_val;
        uint32_t u;
        std::memcpy(&u, &f, 4);
        k = (int)((u >> 23) & 255u) - 127;
        const Real sc = (Real)std::ldexp(1.0, -k);
        for (int i = 0; i < W; ++i) col[i] *= sc;
    }
    k_out = k;
}

template <class Real>
void Recursor<Real>::fill(const Tables* t, const uint8_t* tpl_, int J_, const uint8_t* codes_, int I_, int W_) {
    tab = t;
    tpl.assign(tpl_, tpl_ + J_);
    codes.assign(codes_, codes_ + I_);
    W = W_;
    status = READ_VALID;
    cells = 0;
    ll_alpha = ll_beta = NEG_INF;
    if (J_ < 2 || I_ < 2) { status = READ_TEMPLATE_TOO_SMALL; return; }
    fill_alpha();
    if (status != READ_VALID) return;
    fill_beta();
}

// Recursor::FillAlpha -- SURVEY.md A.4, DESIGN.md "Recursion"
template <class Real>
void Recursor<Real>::fill_alpha() {
    const int Jn = J(), In = I();
    const Tables& T = *tab;
    alpha.init(Jn, W);
    alpha.at(0, 0) = Real(1);
    int edge = 0;   // leading edge of column 0 is row 0
    for (int j = 1; j < Jn; ++j) {
        // the start advances by one `slide` step whenever the leading edge comes within `margin` + 1 rows of the
        // band's last row (quantised slide: band starts stay multiples of `slide`)
        const int s = alpha.start[j - 1] + ((edge + 2 + margin - W > alpha.start[j - 1]) ? slide : 0);
        alpha.start[j] = s;
        const int cm = (j == 1) ? CTX_START + tpl[0] : 4 * tpl[j - 2] + tpl[j - 1];
        const int ci = 4 * tpl[j - 1] + tpl[j];
        Real run = Real(0);
        for (int rel = 0; rel < W; ++rel) {
            const int i = s + rel;
            const int code = code_at_row(i);
            const Real up = alpha.get(j - 1, i - 1), pv = alpha.get(j - 1, i);
            const Real C = (Real)T.mm[cm][code] * up + (Real)T.tr[cm][1] * pv;
            const Real g = (Real)T.gg[ci][code];
            run = C + g * (rel == 0 ? Real(0) : run);
            alpha.at(j, i) = run;
        }
        int k;
        bool dead;
        scale_column(&alpha.v[(size_t)j * W], W, s, edge_log2, edge, k, dead, j % kScaleEvery == 0);
        alpha.cumexp[j] = alpha.cumexp[j - 1] + k;
        cells += W;
        if (dead) { status = READ_DEAD; return; }
    }
    const int ctxl = 4 * tpl[Jn - 2] + tpl[Jn - 1];
    const double a = (double)alpha.get(Jn - 1, In - 1) * T.em_match[CTX_END + ctxl][codes[In - 1]];
    if (!(a > 0.0)) { status = READ_DEAD; return; }
    ll_alpha = std::log(a) + LN2 * (double)alpha.cumexp[Jn - 1] - In * T.log_cw;
}

// Recursor::FillBeta -- mirror image on the same band
template <class Real>
void Recursor<Real>::fill_beta() {
    const int Jn = J(), In = I();
    const Tables& T = *tab;
    beta.init(Jn, W);
    beta.start = alpha.start;
    for (int j = Jn - 1; j >= 1; --j) {
        const int s = beta.start[j];
        const int ci = 4 * tpl[j - 1] + tpl[j];   // == cm of column j+1
        Real run = Real(0);
        for (int rel = W - 1; rel >= 0; --rel) {
            const int i = s + rel;
            const int code1 = code_at_row(i + 1);
            Real C;
            if (j == Jn - 1) {
                C = (i == In - 1) ? (Real)T.em_match[CTX_END + ci][codes[In - 1]] : Real(0);
            } else {
                const Real nd = beta.get(j + 1, i + 1), nx = beta.get(j + 1, i);
                C = (Real)T.mm[ci][code1] * nd + (Real)T.tr[ci][1] * nx;
            }
            const Real g = (Real)T.gg[ci][code1];
            run = C + g * (rel == W - 1 ? Real(0) : run);
            beta.at(j, i) = run;
        }
        int k, m;
        bool dead;
        scale_column(&beta.v[(size_t)j * W], W, s, edge_log2, m, k, dead, j % kScaleEvery == 0);
        beta.cumexp[j] = (j == Jn - 1 ? 0 : beta.cumexp[j + 1]) + k;
        if (dead) { status = READ_DEAD; return; }
    }
    const double b = T.em_match[CTX_START + tpl[0]][codes[0]] * (double)beta.get(1, 1);
    if (!(b > 0.0)) { status = READ_DEAD; return; }
    ll_beta = std::log(b) + LN2 * (double)beta.cumexp[1] - In * T.log_cw;
}

// Evaluator::LL(Mutation) = ExtendAlpha over the columns whose context changed, then
// LinkAlphaBeta (or run to the pinned end)  -- SURVEY.md A.6, DESIGN.md "Mutation scoring"
template <class Real>
double Recursor<Real>::ll_mutated(const Mutation& mu) const {
    const int Jn = J(), In = I();
    const Tables& T = *tab;
    const int q = mu.pos;
    const int delta = mu.type == MUT_INS ? 1 : (mu.type == MUT_DEL ? -1 : 0);
    const int Jp = Jn + delta;
    if (Jp < 2) return NEG_INF;
    auto tv = [&](int j) -> int {
        if (mu.type == MUT_SUB) return j == q ? mu.base : tpl[j];
        if (mu.type == MUT_INS) return j < q ? tpl[j] : (j == q ? mu.base : tpl[j - 1]);
        return j < q ? tpl[j] : tpl[j + 1];
    };
    const int a = std::max(1, q);
    const int bp = (mu.type == MUT_DEL) ? q + 1 : q + 2;
    const int borig = bp - delta;
    const bool terminal = bp > Jp - 1;
    const int last = terminal ? Jp - 1 : bp - 1;
    std::vector<Real> prev(W), cur(W);
    int sprev = alpha.start[a - 1];
    for (int rel = 0; rel < W; ++rel) prev[rel] = alpha.get(a - 1, sprev + rel);
    auto getp = [&](int row) -> Real { return (row < sprev || row >= sprev + W) ? Real(0) : prev[row - sprev]; };
    for (int jp = a; jp <= last; ++jp) {
        const int S = std::max(sprev, alpha.start[std::min(jp, Jn - 1)]);
        const int cm = (jp == 1) ? CTX_START + tv(0) : 4 * tv(jp - 2) + tv(jp - 1);
        const int ci = 4 * tv(jp - 1) + tv(jp);
        Real run = Real(0);
        for (int rel = 0; rel < W; ++rel) {
            const int i = S + rel;
            const int code = code_at_row(i);
            const Real C = (Real)T.mm[cm][code] * getp(i - 1) + (Real)T.tr[cm][1] * getp(i);
            const Real g = (Real)T.gg[ci][code];
            run = C + g * (rel == 0 ? Real(0) : run);
            cur[rel] = run;
        }
        prev.swap(cur);
        sprev = S;
    }
    if (terminal) {
        const int ctxl = 4 * tv(Jp - 2) + tv(Jp - 1);
        const double v = (double)getp(In - 1) * T.em_match[CTX_END + ctxl][codes[In - 1]];
        if (!(v > 0.0)) return NEG_INF;
        return std::log(v) + LN2 * (double)alpha.cumexp[a - 1] - In * T.log_cw;
    }
    const int cmL = (bp == 1) ? CTX_START + tv(0) : 4 * tv(bp - 2) + tv(bp - 1);
    Real sum = Real(0);
    for (int rel = 0; rel < W; ++rel) {
        const int i = sprev + rel;
        const int code1 = code_at_row(i + 1);
        const Real w = (Real)T.mm[cmL][code1] * beta.get(borig, i + 1) + (Real)T.tr[cmL][1] * beta.get(borig, i);
        sum += prev[rel] * w;
    }
    if (!((double)sum > 0.0)) return NEG_INF;
    return std::log((double)sum) + LN2 * (double)(alpha.cumexp[a - 1] + beta.cumexp[borig]) - In * T.log_cw;
}

// ---------------------------------------------------------------------------------------
// Integrator
// ---------------------------------------------------------------------------------------
static void revcomp(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    const size_t n = in.size();
    out.resize(n);
    for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)(3 - in[n - 1 - i]);
}

template <class Real>
void Integrator<Real>::init(const ccs::ArrowModelParams& m, const float snr[4], const uint8_t* tpl, int J,
                            const PolishConfig& c) {
    tab.build(m, snr);
    cfg = c;
    fwd.assign(tpl, tpl + J);
    revcomp(fwd, rev);
    reads.clear(); recs.clear(); active.clear();
    mark_b = 0; mark_e = J;
}

template <class Real>
void Integrator<Real>::add_read(const MappedRead& r) {
    reads.push_back(r);
    recs.emplace_back();
    active.push_back(1);
    const size_t k = reads.size() - 1;
    refill(k);
    // POOR_ZSCORE: only when the read is added (its LL against the draft), never after later refills
    if (active[k] && zscore(k) < cfg.min_zscore) { active[k] = 0; recs[k].status = READ_POOR_ZSCORE; }
}

template <class Real>
void Integrator<Real>::zmoments(size_t r, double& mean, double& var) const {
    const MappedRead& rd = reads[r];
    const int J = (int)fwd.size();
    const int len = rd.tend - rd.tstart;
    const uint8_t* t = rd.strand ? rev.data() + (J - rd.tend) : fwd.data() + rd.tstart;
    mean = tab.zs_first_mean[t[0]]; var = tab.zs_first_var[t[0]];
    for (int j = 1; j < len; ++j) { const int ctx = 4 * t[j - 1] + t[j]; mean += tab.zs_mean[ctx]; var += tab.zs_var[ctx]; }
}

template <class Real>
double Integrator<Real>::zscore(size_t r) const {
    double mean, var;
    zmoments(r, mean, var);
    return (recs[r].ll() - mean) / std::sqrt(var);
}

template <class Real>
void Integrator<Real>::refill(size_t r) {
    const MappedRead& rd = reads[r];
    const int J = (int)fwd.size();
    const int len = rd.tend - rd.tstart;
    if (!active[r]) return;
    if (len < 2 || rd.tstart < 0 || rd.tend > J) { active[r] = 0; recs[r].status = READ_TEMPLATE_TOO_SMALL; return; }
    const uint8_t* t = rd.strand ? rev.data() + (J - rd.tend) : fwd.data() + rd.tstart;
    recs[r].fill(&tab, t, len, rd.codes.data(), (int)rd.codes.size(), cfg.band_width);
    if (recs[r].status == READ_VALID) {
        const double la = recs[r].ll_alpha, lb = recs[r].ll_beta;
        if (!(std::fabs(1.0 - la / lb) <= cfg.ab_mismatch_tol)) recs[r].status = READ_ALPHA_BETA_MISMATCH;
    }
    if (recs[r].status != READ_VALID) active[r] = 0;
}

template <class Real>
void Integrator<Real>::refill_all() { for (size_t r = 0; r < reads.size(); ++r) refill(r); }

template <class Real>
double Integrator<Real>::ll() const {
    double s = 0;
    for (size_t r = 0; r < reads.size(); ++r) if (active[r]) s += recs[r].ll();
    return s;
}

template <class Real>
bool Integrator<Real>::read_delta(size_t r, const Mutation& m, double& d) const {
    d = 0;
    if (!active[r]) return false;
    const MappedRead& rd = reads[r];
    Mutation loc = m;
    if (m.type == MUT_INS) {
        if (!(m.pos > rd.tstart && m.pos < rd.tend)) return false;
        loc.pos = rd.strand ? rd.tend - m.pos : m.pos - rd.tstart;
    } else {
        if (!(m.pos >= rd.tstart && m.pos < rd.tend)) return false;
        loc.pos = rd.strand ? rd.tend - 1 - m.pos : m.pos - rd.tstart;
    }
    if (rd.strand) loc.base = 3 - m.base;
    d = recs[r].ll_mutated(loc) - recs[r].ll();
    return true;
}

template <class Real>
double Integrator<Real>::delta_ll(const Mutation& m) const {
    double s = 0, d;
    for (size_t r = 0; r < reads.size(); ++r) if (read_delta(r, m, d)) s += d;
    return s;
}

std::vector<uint8_t> apply_mutations(const std::vector<uint8_t>& tpl, const std::vector<Mutation>& muts) {
    std::vector<uint8_t> out;
    out.reserve(tpl.size() + muts.size());
    size_t k = 0;
    const int J = (int)tpl.size();
    for (int j = 0; j <= J; ++j) {
        bool skip = false;
        while (k < muts.size() && muts[k].pos == j) {
            const Mutation& m = muts[k++];
            if (m.type == MUT_INS) out.push_back((uint8_t)m.base);
            else if (m.type == MUT_SUB) { out.push_back((uint8_t)m.base); skip = true; }
            else skip = true;
        }
        if (j < J && !skip) out.push_back(tpl[j]);
    }
    return out;
}

template <class Real>
void Integrator<Real>::apply(const std::vector<Mutation>& muts) {
    fwd = apply_mutations(fwd, muts);
    revcomp(fwd, rev);
    for (auto& rd : reads) {
        int ds = 0, de = 0;
        for (const auto& m : muts) {
            if (m.type == MUT_INS) { if (m.pos <= rd.tstart) ++ds; if (m.pos < rd.tend) ++de; }
            else if (m.type == MUT_DEL) { if (m.pos < rd.tstart) --ds; if (m.pos < rd.tend) --de; }
        }
        rd.tstart += ds;
        rd.tend += de;
    }
    {
        int db = 0, de = 0;
        for (const auto& m : muts) {
            if (m.type == MUT_INS) { if (m.pos <= mark_b) ++db; if (m.pos <= mark_e) ++de; }
            else if (m.type == MUT_DEL) { if (m.pos < mark_b) --db; if (m.pos < mark_e) --de; }
        }
        mark_b += db; mark_e += de;
    }
    refill_all();
}

// ---------------------------------------------------------------------------------------
// Polish / QVs  (docs/how-does-ccs-work.md:96-106)
// ---------------------------------------------------------------------------------------
bool mutation_is_canonical(const std::vector<uint8_t>& tpl, const Mutation& m) {
    const int J = (int)tpl.size();
    if (m.type == MUT_SUB) return m.pos >= 0 && m.pos < J && m.base != tpl[m.pos];
    if (m.type == MUT_DEL) return m.pos >= 0 && m.pos < J && !(m.pos > 0 && tpl[m.pos] == tpl[m.pos - 1]);
    return m.pos >= 1 && m.pos <= J - 1 && m.base != tpl[m.pos - 1];
}

static int type_rank(int t) { return t == MUT_DEL ? 0 : (t == MUT_INS ? 1 : 2); }

std::vector<Mutation> best_mutations(std::vector<std::pair<double, Mutation>>& scored, int separation) {
    std::sort(scored.begin(), scored.end(), [](const auto& a, const auto& b) {
        if (a.first != b.first) return a.first > b.first;
        if (a.second.pos != b.second.pos) return a.second.pos < b.second.pos;
        if (a.second.type != b.second.type) return type_rank(a.second.type) < type_rank(b.second.type);
        return a.second.base < b.second.base;
    });
    std::vector<Mutation> chosen;
    for (const auto& sm : scored) {
        bool ok = true;
        for (const auto& c : chosen) if (std::abs(c.pos - sm.second.pos) < separation) { ok = false; break; }
        if (ok) chosen.push_back(sm.second);
    }
    std::sort(chosen.begin(), chosen.end(), [](const Mutation& a, const Mutation& b) { return a.pos < b.pos; });
    return chosen;
}

static uint64_t tpl_hash(const std::vector<uint8_t>& t) {
    uint64_t h = 1469598103934665603ull;
    for (uint8_t b : t) { h ^= b; h *= 1099511628211ull; }
    return h ^ (t.size() * 0x9E3779B97F4A7C15ull);
}

template <class Real>
PolishResult polish(Integrator<Real>& ai) {
    PolishResult res;
    std::set<uint64_t> seen;
    seen.insert(tpl_hash(ai.fwd));
    std::vector<int> sites;   // positions (current coordinates) of last-applied mutations
    // growth cap (DESIGN.md "Known deviations"): a template that would outgrow its initial length by more than
    // max(growth_min = 512, J/8) bases (rounded up to a multiple of 16) stops being refined and counts as not converged
    const int J0 = (int)ai.fwd.size();
    const int cap = ((J0 + std::max(ai.cfg.growth_min, J0 / 8)) + 15) & ~15;
    for (int it = 0; it < ai.cfg.max_iterations; ++it) {
        res.iterations = it + 1;
        const int J = (int)ai.fwd.size();
        // positions tested this round: round 0 everything, later +-neighborhood around the sites just edited -- always
        // within test_margin bases of the window core
        const int tb = std::max(0, ai.mark_b - ai.cfg.test_margin), te = std::min(J, ai.mark_e + ai.cfg.test_margin);
        std::vector<char> want(J + 1, 0);
        if (it == 0) for (int p = tb; p < te; ++p) want[p] = 1;
        else
            for (int s : sites)
                for (int p = std::max(tb, s - ai.cfg.neighborhood); p <= std::min(te - 1, s + ai.cfg.neighborhood); ++p) want[p] = 1;
        std::vector<std::pair<double, Mutation>> scored;
        for (int p = 0; p <= J; ++p) {
            if (!want[p]) continue;
            for (int t = 0; t < 3; ++t)
                for (int b = 0; b < (t == MUT_DEL ? 1 : 4); ++b) {
                    Mutation m{t, p, b};
                    if (!mutation_is_canonical(ai.fwd, m)) continue;
                    ++res.n_tested;
                    const double d = ai.delta_ll(m);
                    if (d > 0.0) scored.push_back({d, m});
                }
        }
        if (scored.empty()) { res.converged = true; break; }
        std::vector<Mutation> best = best_mutations(scored, ai.cfg.separation);
        std::vector<uint8_t> next = apply_mutations(ai.fwd, best);
        if (seen.count(tpl_hash(next))) {   // cycle guard: fall back to the single best
            best.assign(1, scored.front().second);
            next = apply_mutations(ai.fwd, best);
        }
        seen.insert(tpl_hash(next));
        if ((int)next.size() > cap) break;
        // sites in new coordinates
        sites.clear();
        int off = 0;
        for (const auto& m : best) {
            sites.push_back(m.pos + off);
            off += m.type == MUT_INS ? 1 : (m.type == MUT_DEL ? -1 : 0);
        }
        res.n_applied += (int)best.size();
        ai.apply(best);
        if (ai.n_active() == 0) break;
    }
    return res;
}

template <class Real>
void consensus_qvs(const Integrator<Real>& ai, std::vector<uint8_t>& qv) {
    const int J = (int)ai.fwd.size();
    qv.assign(J, 0);
    for (int p = 0; p < J; ++p) {
        double s = 0;
        for (int t = 0; t < 3; ++t)
            for (int b = 0; b < (t == MUT_DEL ? 1 : 4); ++b) {
                if (t == MUT_SUB && b == ai.fwd[p]) continue;
                if (t == MUT_INS && !(p >= 1 && p <= J - 1)) continue;
                s += std::exp(ai.delta_ll(Mutation{t, p, b}));
            }
        double q = (s > 0) ? -10.0 * std::log10(s / (1.0 + s)) : 93.0;
        if (!(q < 93.0)) q = 93.0;
        if (q < 0.0) q = 0.0;
        qv[p] = (uint8_t)std::lround(q);
    }
}

double predicted_accuracy(const std::vector<uint8_t>& qv) {
    if (qv.empty()) return 0.0;
    double e = 0;
    for (uint8_t q : qv) e += std::pow(10.0, -0.1 * q);
    return 1.0 - e / (double)qv.size();
}

template struct Recursor<double>;
template struct Recursor<float>;
template struct Integrator<double>;
template struct Integrator<float>;
template PolishResult polish<double>(Integrator<double>&);
template PolishResult polish<float>(Integrator<float>&);
template void consensus_qvs<double>(const Integrator<double>&, std::vector<uint8_t>&);
template void consensus_qvs<float>(const Integrator<float>&, std::vector<uint8_t>&);

}  // namespace oracle
