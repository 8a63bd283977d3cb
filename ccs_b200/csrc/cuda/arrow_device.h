// Device-side data layout of the Arrow polish stage, shared by the CUDA kernels and the
// C++ host engine (plain PODs; no CUDA types).
//
// Band storage (DESIGN.md "Layout in HBM"): every (read, template) pair owns J columns;
// a column is one 128-byte line of 32 fp32 cells holding rows [start_j, start_j + 32) of the
// DP matrix at slot (row mod 32) -- the rotated layout lets a warp octet (8 lanes x 4 cells)
// keep each DP row in the same register from column to column while the band slides down.
// Reference concept: Recursor's ScaledMatrix/SparseMatrix columns.
#pragma once
#include <cstdint>

namespace ccs {

constexpr int kBandW = 32;            // band rows per column (spec)
constexpr int kBandMargin = 2;        // rows kept beyond the leading edge (spec)
constexpr int kEdgeLog2 = -60;        // leading-edge threshold 2^-60 on unscaled cells (spec)
constexpr int kScaleEvery = 4;        // columns j with j % 4 == 0 are rescaled by the power of two of their maximum (spec)
constexpr int kRowCodePad = 96;
constexpr int kDeltaStride = 16;       // doubles per template position in the delta store:
                                      // {SUB A,C,G,T, DEL, INS A,C,G,T, INS' A,C,G,T (reverse-strand share), pad x3}       // sentinel codes after the last real row code

struct alignas(8) ColInfo { int32_t start; int32_t cumexp; };   // per alpha column: band start, cumulative scale exponent

struct DevRead {
    int64_t code_off;    // rowcode[code_off + i] = 4 * emission code of DP row i (sentinel 48 at i = 0 and i >= I);
                         // a second copy shifted by one row (code of row i+1) follows at code_off + code_stride
    int64_t col_off;     // first column of this pair in the alpha / beta / colinfo stores
    int32_t I;           // read length
    int32_t J;           // template slice length (columns 0..J-1)
    int32_t tpl_off;     // offset of the read-oriented template slice in the template buffer
    int32_t zmw;         // owning ZMW (index into the transition tables)
    int32_t ts, te;      // slice on the forward template [ts,te)
    uint8_t strand;      // 0 forward, 1 reverse complement
    uint8_t active;      // 0: skip (filtered / dropped)
    uint8_t first_code;  // e_0
    uint8_t last_code;   // e_{I-1}
    int32_t code_stride; // bytes of one row-code copy (multiple of 16)
};
static_assert(sizeof(DevRead) == 48, "DevRead layout");

struct DevZmw {
    int64_t delta_off;   // first row of this ZMW in the delta store (rows = template positions)
    int32_t read_begin, read_end;   // reads [begin,end) in DevRead order
    int32_t fwd_off, rev_off;       // template buffer offsets (forward / reverse complement)
    int32_t J;                      // current template length
    int32_t pad_;
};
static_assert(sizeof(DevZmw) == 32, "DevZmw layout");

// A contiguous run of forward-template positions of one ZMW to score; `first` = index of the
// range's first position in the flattened work list.
struct ScoreRange { int32_t zmw; int32_t p_begin; int32_t p_end; int32_t pad_; int64_t first; };

// One read to encode into its two row-code copies (arrow_pack_rowcodes_kernel).
struct PackJob { int64_t src_off; int64_t dst_off; int32_t I; int32_t stride; };

// Delta rows of one edited ZMW to re-index (arrow_remap_delta_kernel).
struct RemapJob { int64_t delta_off; int64_t scratch_off; int32_t J_old, J_new; int32_t site_off, n_sites; };

// One positive-scoring canonical mutation found by the pick kernel.
struct Candidate { float score_hi; int32_t zmw; int32_t pos; int16_t type; int16_t base; double score; };

// Kernel-side view of one batch.
struct ArrowBatchView {
    const uint8_t* rowcode;
    const uint8_t* tpl;          // template buffer (bases 0..3)
    const float* em_match;       // [36][16] chemistry-wide, counter-weight folded in
    const float* em_ins;         // [17][16]
    const float* trans;          // [n_zmw][36][4] = {match, deletion, branch, stick}
    const DevRead* reads;
    const DevZmw* zmws;
    float* alpha;                // [total_cols][32]
    float* beta;                 // [total_cols][32]
    ColInfo* colinfo;            // [total_cols]
    int32_t* beta_exp;           // [total_cols] cumulative scale exponent of beta from the right
    double* ll_alpha;            // [n_reads] full log-likelihood (incl. counter-weight correction)
    double* ll_beta;             // [n_reads]
    double* base_ll;             // [n_reads] ll_alpha + I*log(cw): the scale bookkeeping delta-LLs subtract
    int32_t* status;             // [n_reads] ccs_read_status
    int32_t n_reads;
    int32_t n_zmws;
    double log_cw;
    double ab_tol;
};

}  // namespace ccs
