// TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the per-ZMW Arrow polish path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this; the product (ccs_b200/) never links or calls it.
//
// PARITY UNPINNED: /root/reference holds documentation only; the real
// implementation (closed `pbccs`, last public source PacificBiosciences/unanimity @
// 6f11a13e1472b8c00337ba8c5e94bf83bdab31d6, /root/reference/docs/faq/source-code.md:7-15) is
// not available, and the reference ships no golden vectors for this path.  This file
// therefore restates the algorithm from the documented behaviour
// (/root/reference/docs/how-does-ccs-work.md:87-106) and the normative spec in DESIGN.md
// (SURVEY.md Appendix A), and is pinned by first-principles tests (tests/test_oracle_*.py):
// brute-force path enumeration, alpha/beta agreement, incremental-vs-refill mutation
// scoring, recovery of a known template.
//
// Function <-> reference concept map (unanimity class names as used by BASELINE.json):
//   Tables               ModelConfig::Populate / TemplatePosition   docs/how-does-ccs-work.md:90-94
//   Recursor::fill_alpha Recursor::FillAlpha                        docs/how-does-ccs-work.md:94-96
//   Recursor::fill_beta  Recursor::FillBeta                         (same)
//   Recursor::ll_mutated Evaluator::LL(Mutation): ExtendAlpha + LinkAlphaBeta   :96-99
//   Integrator           Integrator::{AddRead,LL,ApplyMutations}    :15 ("using all subreads")
//   polish()             Polish()                                   :96-101
//   consensus_qvs()      ConsensusQualities()                       :103-106
#pragma once
#include <cstdint>
#include <vector>
#include <string>
#include "../ccs_b200/csrc/common/arrow_model.h"

namespace oracle {

constexpr int CTX_START = 16, CTX_END = 20, N_MROWS = 36, N_IROWS = 17, CODE_STRIDE = 16, CODE_SENTINEL = 12;

enum MutType : int { MUT_SUB = 0, MUT_INS = 1, MUT_DEL = 2 };
struct Mutation { int type; int pos; int base; };   // INS: inserted before pos

enum ReadStatus : int { READ_VALID = 0, READ_ALPHA_BETA_MISMATCH = 1, READ_TEMPLATE_TOO_SMALL = 2, READ_DEAD = 3,
                        READ_POOR_ZSCORE = 5 };

// per-ZMW tables; values are fp32-rounded (the spec says the tables are fp32) held in double
struct Tables {
    double em_match[N_MROWS][CODE_STRIDE];
    double em_ins[N_IROWS][CODE_STRIDE];
    double tr[N_MROWS][4];   // match, deletion, branch, stick
    double log_cw;
    // The model the recursion uses (DESIGN.md "Arrow model"): per-ZMW FOLDED factors, each one fp32 product of
    // the fp32 emission and the fp32 transition, rounded once to fp32 and held in double:
    //   mm[row][code] = fl32(em_match[row][code] * match[row])          (start/end rows: match = 1)
    //   gg[ctx][code] = fl32(em_ins[ctx][code] * (cognate ? branch[ctx] : stick[ctx]))
    double mm[N_MROWS][CODE_STRIDE];
    double gg[N_IROWS][CODE_STRIDE];
    // Expected log-likelihood moments of one template position under the generative HMM (DESIGN.md "z-score filter";
    // Integrator's POOR_ZSCORE read filter): a geometric number of insertions, then a match or a
    // deletion.  zs_*[ctx] for the positions 1..J-1, zs_first_*[base] for the pinned first match.
    double zs_mean[16], zs_var[16], zs_first_mean[4], zs_first_var[4];
    void build(const ccs::ArrowModelParams& m, const float snr[4]);
};

template <class Real>
struct Banded {
    int J = 0, W = 32;
    std::vector<Real> v;            // J columns x W slots, slot = row mod W
    std::vector<int32_t> start;     // band start row per column
    std::vector<int64_t> cumexp;    // cumulative power-of-two scale exponent through column j
    void init(int J_, int W_) { J = J_; W = W_; v.assign((size_t)J * W, Real(0)); start.assign(J, 0); cumexp.assign(J, 0); }
    inline Real get(int j, int row) const {
        const int s = start[j];
        if (row < s || row >= s + W) return Real(0);
        return v[(size_t)j * W + (row % W)];
    }
    inline Real& at(int j, int row) { return v[(size_t)j * W + (row % W)]; }
};

template <class Real>
struct Recursor {
    const Tables* tab = nullptr;
    std::vector<uint8_t> tpl;     // read-oriented template slice, bases 0..3
    std::vector<uint8_t> codes;   // read emission codes
    int W = 32;
    int margin = 2;               // band rule: rows kept beyond the leading edge
    int edge_log2 = -60;          // band rule: leading-edge threshold 2^edge_log2 on unscaled cells
    int slide = 4;                // band rule: the band start advances in steps of `slide` rows (one lane of an octet)
    Banded<Real> alpha, beta;
    double ll_alpha = 0, ll_beta = 0;
    int status = READ_VALID;
    int64_t cells = 0;            // band cells defined (for the roofline's algorithmic bytes)

    inline int I() const { return (int)codes.size(); }
    inline int J() const { return (int)tpl.size(); }
    inline int code_at_row(int i) const { return (i >= 1 && i <= I() - 1) ? codes[i - 1] : CODE_SENTINEL; }

    void fill(const Tables* t, const uint8_t* tpl_, int J_, const uint8_t* codes_, int I_, int W_);
    void fill_alpha();
    void fill_beta();
    // LL of the read under the template mutated by m (m in this read's local orientation),
    // computed incrementally from the stored alpha/beta (generic extend + link).
    double ll_mutated(const Mutation& m) const;
    inline double ll() const { return ll_alpha; }
};

struct MappedRead {
    std::vector<uint8_t> codes;
    int strand = 0;          // 0 forward, 1 reverse complement of the template
    int tstart = 0, tend = 0;  // span on the forward template, [tstart,tend)
    int full_length = 1;
};

struct PolishConfig {
    int max_iterations = 40;
    int separation = 10;
    int neighborhood = 20;
    int band_width = 32;
    double ab_mismatch_tol = 1e-3;
    double min_zscore = -3.4;     // AddRead drops a read whose LL lies further below its expectation (POOR_ZSCORE)
    int growth_min = 512;         // growth cap: the template may outgrow its first length by max(growth_min, J/8) bases
    int test_margin = 48;         // mutations are tested within this many bases of the window core [mark_b, mark_e) only; the
                                  // rest of a window's padding is context (DESIGN.md "Windowing"); whole templates: everything
};

struct PolishResult {
    bool converged = false;
    int iterations = 0;
    int64_t n_tested = 0;
    int n_applied = 0;
};

template <class Real>
struct Integrator {
    Tables tab;
    PolishConfig cfg;
    std::vector<uint8_t> fwd, rev;
    std::vector<MappedRead> reads;
    std::vector<Recursor<Real>> recs;
    std::vector<int> active;

    void init(const ccs::ArrowModelParams& m, const float snr[4], const uint8_t* tpl, int J, const PolishConfig& c);
    void add_read(const MappedRead& r);
    double zscore(size_t r) const;           // (LL - E[LL]) / sd[LL] of read r on its template slice
    void zmoments(size_t r, double& mean, double& var) const;   // E[LL], Var[LL] of read r on its template slice
    // Two template borders tracked through ApplyMutations like a read's tstart (an insertion at the border goes to
    // its left, a deletion of the base at the border to its right): the window core of DESIGN.md "Windowing".
    int mark_b = 0, mark_e = 0;
    void refill_all();
    void refill(size_t r);
    double ll() const;                       // sum over active reads
    double delta_ll(const Mutation& m) const;  // sum over active reads covering m
    bool read_delta(size_t r, const Mutation& m, double& d) const;
    void apply(const std::vector<Mutation>& muts);  // muts sorted by position, non-overlapping
    int n_active() const { int n = 0; for (int a : active) n += a; return n; }
};

// dedup rule for the polish candidate set (QVs use the full set)
bool mutation_is_canonical(const std::vector<uint8_t>& tpl, const Mutation& m);
std::vector<Mutation> best_mutations(std::vector<std::pair<double, Mutation>>& scored, int separation);
std::vector<uint8_t> apply_mutations(const std::vector<uint8_t>& tpl, const std::vector<Mutation>& muts);

template <class Real> PolishResult polish(Integrator<Real>& ai);
template <class Real> void consensus_qvs(const Integrator<Real>& ai, std::vector<uint8_t>& qv);
double predicted_accuracy(const std::vector<uint8_t>& qv);

}  // namespace oracle
