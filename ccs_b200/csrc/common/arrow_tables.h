// Per-ZMW Arrow tables (host side, product + simulator).
//
// Turns ArrowModelParams + the ZMW's four channel SNRs into the transition table the
// kernels index by dinucleotide context (behaviour documented at
// /root/reference/docs/how-does-ccs-work.md:90-94), and the chemistry-wide emission tables
// with the pinned-start / pinned-end pseudo contexts appended (DESIGN.md "Recursion").
#pragma once
#include "arrow_model.h"
#include <algorithm>

namespace ccs {

struct TransProb { double match, branch, stick, deletion; };

// tp[ctx]: probabilities of the four moves when the next template base has context ctx
// (ctx = 4*prev + next).  SNR channel = the next base's channel.
inline void transition_probs(const ArrowModelParams& m, const float snr[4], TransProb tp[kNumCtx]) {
    for (int ctx = 0; ctx < kNumCtx; ++ctx) {
        const int ch = ctx & 3;
        const double s = std::min(std::max((double)snr[ch], m.snr_lo[ch]), m.snr_hi[ch]);
        double e[3], sum = 1.0;
        for (int t = 0; t < 3; ++t) {
            const double* c = m.trans[ctx][t];
            e[t] = std::exp(c[0] + s * (c[1] + s * (c[2] + s * c[3])));
            sum += e[t];
        }
        tp[ctx] = {1.0 / sum, e[TR_BRANCH] / sum, e[TR_STICK] / sum, e[TR_DELETION] / sum};
    }
}

// Context rows used by the kernels.  Rows 0..15 are the dinucleotide contexts; 16..19 the
// pinned first move (template base b, no transition factor); 20..35 the pinned last move
// (context c, no transition factor); kCtxZero is an all-zero row (no insertion allowed).
constexpr int kCtxStart = 16;
constexpr int kCtxEnd = 20;
constexpr int kNumMatchRows = 36;
constexpr int kCtxZero = 16;      // in the insertion table
constexpr int kNumInsRows = 17;
constexpr int kCodeStride = 16;   // 12 real codes + 4 zero "sentinel" codes
constexpr int kCodeSentinel = 12;

// Chemistry-wide emission tables in fp32, pre-multiplied by the counter weight.
//   em_match[row][code]  row in [0,36)
//   em_ins  [row][code]  row in [0,17): Branch pmf where code's base == context's cur base,
//                        Stick pmf otherwise (the transition factor is applied per ZMW)
struct EmissionTables {
    float em_match[kNumMatchRows][kCodeStride];
    float em_ins[kNumInsRows][kCodeStride];
};

inline void build_emission_tables(const ArrowModelParams& m, EmissionTables& t) {
    std::memset(&t, 0, sizeof(t));
    const double cw = m.counter_weight;
    for (int ctx = 0; ctx < kNumCtx; ++ctx)
        for (int code = 0; code < kNumCodes; ++code) {
            t.em_match[ctx][code] = (float)(cw * m.emission[MOVE_MATCH][ctx][code]);
            t.em_match[kCtxEnd + ctx][code] = (float)(cw * m.emission[MOVE_MATCH][ctx][code]);
            const bool cognate = (code & 3) == (ctx & 3);
            t.em_ins[ctx][code] = (float)(cw * m.emission[cognate ? MOVE_BRANCH : MOVE_STICK][ctx][code]);
        }
    for (int b = 0; b < 4; ++b)
        for (int code = 0; code < kNumCodes; ++code)
            t.em_match[kCtxStart + b][code] = (float)(cw * m.emission[MOVE_MATCH][4 * b + b][code]);
}

// Per-ZMW transition scalars in fp32, one float4-like row per match-table row:
//   tr[row] = {match, deletion, branch, stick};  start/end rows have match = 1, rest 0.
struct ZmwTransitions { float tr[kNumMatchRows][4]; };

inline void build_zmw_transitions(const ArrowModelParams& m, const float snr[4], ZmwTransitions& z) {
    TransProb tp[kNumCtx];
    transition_probs(m, snr, tp);
    std::memset(&z, 0, sizeof(z));
    for (int ctx = 0; ctx < kNumCtx; ++ctx) {
        z.tr[ctx][0] = (float)tp[ctx].match;
        z.tr[ctx][1] = (float)tp[ctx].deletion;
        z.tr[ctx][2] = (float)tp[ctx].branch;
        z.tr[ctx][3] = (float)tp[ctx].stick;
    }
    for (int r = kCtxStart; r < kNumMatchRows; ++r) z.tr[r][0] = 1.0f;
}

// Expected log-likelihood moments of one template position under the generative HMM, for the POOR_ZSCORE read filter
// (Integrator::AddRead; DESIGN.md "z-score filter"): at a position with context ctx the read emits a
// geometric number N of insertions (probability q = branch + stick each) and then a match or a deletion, so
//   E[L] = E[N] E[X] + E[Y],  Var[L] = E[N] Var[X] + Var[N] E[X]^2 + Var[Y],
// X = log-probability of one insertion event, Y = log-probability of the closing match / deletion event.
struct ZscoreMoments { double mean[kNumCtx], var[kNumCtx], first_mean[4], first_var[4]; };

inline void zscore_moments(const ArrowModelParams& m, const float snr[4], ZscoreMoments& z) {
    TransProb tp[kNumCtx];
    transition_probs(m, snr, tp);
    for (int ctx = 0; ctx < kNumCtx; ++ctx) {
        const double q = tp[ctx].branch + tp[ctx].stick;
        double sx = 0, sxx = 0, sy = 0, syy = 0;
        auto acc = [](double p, double& s1, double& s2) { if (p > 0) { const double l = std::log(p); s1 += p * l; s2 += p * l * l; } };
        for (int code = 0; code < kNumCodes; ++code) {
            acc(tp[ctx].branch * m.emission[MOVE_BRANCH][ctx][code], sx, sxx);
            acc(tp[ctx].stick * m.emission[MOVE_STICK][ctx][code], sx, sxx);
            acc(tp[ctx].match * m.emission[MOVE_MATCH][ctx][code], sy, syy);
        }
        acc(tp[ctx].deletion, sy, syy);
        const double ex = sx / q, ex2 = sxx / q, ey = sy / (1.0 - q), ey2 = syy / (1.0 - q);
        const double en = q / (1.0 - q), vn = q / ((1.0 - q) * (1.0 - q));
        z.mean[ctx] = en * ex + ey;
        z.var[ctx] = en * (ex2 - ex * ex) + vn * ex * ex + (ey2 - ey * ey);
    }
    for (int b = 0; b < 4; ++b) {
        double s1 = 0, s2 = 0;
        for (int code = 0; code < kNumCodes; ++code) {
            const double p = m.emission[MOVE_MATCH][5 * b][code];
            if (p > 0) { const double l = std::log(p); s1 += p * l; s2 += p * l * l; }
        }
        z.first_mean[b] = s1;
        z.first_var[b] = s2 - s1 * s1;
    }
}

}  // namespace ccs
