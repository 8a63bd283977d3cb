// Host side of the Polish Stage (the GPU box of /root/reference/docs/img/ccs-impl.png):
// packs a batch of ZMWs (drafts + mapped subreads) into the device layout of
// cuda/arrow_device.h, drives the Arrow kernels round by round, and owns the small sequential
// pieces of Polish() -- greedy mutation selection with separation, template editing, span
// bookkeeping (docs/how-does-ccs-work.md:96-112).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <functional>
#include "../common/arrow_tables.h"
#include "../cuda/arrow_device.h"
#include "cuda_util.h"

namespace ccs {

struct PolishInput {
    int32_t n_zmws = 0, n_reads = 0;
    const int32_t* zmw_read_off = nullptr;   // [n_zmws+1] -> reads
    const int64_t* read_off = nullptr;       // [n_reads+1] -> codes
    const uint8_t* codes = nullptr;          // emission codes 4*(pw-1)+base, native read orientation
    const float* snr = nullptr;              // [n_zmws*4] A,C,G,T
    const int64_t* tpl_off = nullptr;        // [n_zmws+1] -> tpl
    const uint8_t* tpl = nullptr;            // draft templates, bases 0..3
    const uint8_t* strand = nullptr;         // [n_reads] 0 fwd / 1 rev
    const int32_t* tstart = nullptr;         // [n_reads] span on the draft [tstart,tend); tend<=tstart: unmapped
    const int32_t* tend = nullptr;
    const int32_t* rstart = nullptr;         // optional clip of each read [rstart,rend), native orientation
    const int32_t* rend = nullptr;
    // optional: the same codes already resident on this device (the lane's Draft Stage uploaded them);
    // d_codes[k] is codes[read_off[0] + k].  NULL: the engine uploads them itself.
    const uint8_t* d_codes = nullptr;
    // ---- windowed input (window_host.h): every "ZMW" is a window of a draft, every "read" a slice of a subread --------
    // code_start / code_len (both or neither): the slice of `codes` each read occupies (absolute offsets), replacing
    // read_off / rstart / rend; needs d_codes with d_codes[k] = codes[code_base + k].
    const int64_t* code_start = nullptr;
    const int32_t* code_len = nullptr;
    int64_t code_base = 0;
    const int32_t* read_group = nullptr;     // [n_reads] reads of one group (a subread's windows) share one z-score verdict
    int32_t n_groups = 0;
    const int32_t* zmw_group = nullptr;      // [n_zmws] windows of one group (a draft) are consecutive; a group whose window is
                                             // unusable when polishing starts is not polished at all
    const int32_t* core_b = nullptr;         // [n_zmws] borders of the window core, tracked through the edits;
    const int32_t* core_e = nullptr;         //          NULL: the whole template
    const int32_t* growth_min = nullptr;     // [n_zmws] template growth room max(growth_min, J/8); NULL: 512
};

struct PolishParams {
    int32_t max_iterations = 40;
    int32_t separation = 10;
    int32_t neighborhood = 20;
    double ab_mismatch_tol = 1e-3;
    double min_rq = 0.99;          // --min-rq
    int32_t min_length = 10;       // --min-length
    int32_t max_length = 50000;    // --max-length
    double min_active_fraction = 0.5;   // TOO_MANY_UNUSABLE below this share of mapped reads
    double min_zscore = -3.4;           // POOR_ZSCORE: reads whose LL z-score against the draft is lower are dropped
    int32_t window_size = 1024;         // windowing (window_host.h), applied by ccsgpu_ccs
    int32_t window_overlap = 64;
    int32_t test_margin = 48;           // mutations are tested within this many bases of a window's core only
};

struct HostMutation { int32_t type, pos, base; double score; };

struct ZmwState {
    std::vector<uint8_t> tpl;          // current forward template
    int32_t read_begin = 0, read_end = 0;
    bool converged = false, failed = false, done = false;
    bool stale_scores = false;         // a read was dropped after scoring began: stored delta-LLs no longer add up
    std::vector<uint8_t> stale;        // per template position: 1 = the stored delta-LLs predate an edit within qv_halo
                                       // positions (or were never computed); only kept when reuse_scores is on
    std::vector<int32_t> remap_sites, remap_shifts;   // this round's edits (new coordinates, cumulative shift)
    int32_t J_before = 0;
    bool dirty = true;                 // template changed since the last alpha/beta fill of its reads
    int32_t iterations = 0, n_applied = 0;
    int64_t n_tested = 0;
    std::vector<uint64_t> seen;        // template hashes (cycle guard)
    std::vector<int32_t> sites;        // positions of the last-applied mutations (new coordinates)
    int32_t n_mapped = 0;
    int32_t core_b = 0, core_e = 0;    // borders of the window core in current template coordinates (ApplyMutations keeps them)
};

struct EngineStats {
    // per-kernel accumulated device time (ms, CUDA events on the engine stream), launches,
    // algorithmic bytes (DESIGN.md "Roofline") -- reset by reset_stats()
    double ms_fill_alpha = 0, ms_fill_beta = 0, ms_score = 0, ms_pick = 0, ms_qv = 0, ms_h2d = 0, ms_d2h = 0;
    int64_t n_fill_alpha = 0, n_fill_beta = 0, n_score = 0, n_pick = 0, n_qv = 0;
    int64_t bytes_fill_alpha = 0, bytes_fill_beta = 0, bytes_score = 0;
    int64_t cells_fill = 0, score_items = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int64_t rounds = 0;
    int64_t top_fill_alpha_bytes = 0;   // largest single arrow_fill_alpha launch: algorithmic bytes ...
    double top_fill_alpha_ms = 0;       // ... and its CUDA-event duration
    int64_t top_fill_beta_bytes = 0, top_score_bytes = 0;
    double top_fill_beta_ms = 0, top_score_ms = 0;
    int64_t n_pack = 0;
    double ms_resident = 0;   // CUDA-event time of polish() with inputs already in HBM (after load)
    double ms_e2e = 0;        // host wall time of whole stage calls (pack + H2D + kernels + D2H)
    int64_t n_zmws = 0;
};

class ArrowEngine {
public:
    ArrowEngine(int device, const ArrowModelParams& model, size_t budget_bytes);
    ~ArrowEngine();

    // ---- batch life cycle -------------------------------------------------------------
    void load(const PolishInput& in);        // pack + H2D; every mapped read starts active
    void fill();                             // alpha + beta for every active read, statuses updated
    void score_ranges(const std::vector<ScoreRange>& ranges, int64_t n_items);   // delta-LL of every slot
    void score_all_positions();
    int64_t pick(std::vector<Candidate>& out);   // canonical mutations with delta-LL > 0 in the scored ranges
    void polish(const PolishParams& pp);     // full Polish() loop + QVs
    void consensus_qvs();                    // ConsensusQualities on the current templates

    // ---- results ----------------------------------------------------------------------
    int32_t n_reads() const { return (int32_t)reads_.size(); }
    int32_t n_zmws() const { return (int32_t)zmws_.size(); }
    const std::vector<ZmwState>& zmw_states() const { return zstate_; }
    const std::vector<DevRead>& reads() const { return reads_; }
    void read_lls(double* ll_alpha, double* ll_beta, int32_t* status);
    void dump_pair(int r, float* alpha, float* beta, int32_t* start, int32_t* aexp, int32_t* bexp);
    void download_delta(int z, double* out /* [J*9] */);
    const std::vector<std::vector<uint8_t>>& qvs() const { return qv_; }

    EngineStats stats;
    void reset_stats() { stats = EngineStats(); }
    void release_buffers();       // free the (grow-only) per-batch device buffers, e.g. when the lane count changes
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }
    bool timing_enabled = true;
    int host_threads = 8;         // threads for the per-ZMW host pieces of a round
    // ConsensusQualities re-scores only the positions within qv_halo of an edit made after they were last scored (or
    // never scored) and reuses the stored delta-LLs elsewhere: an edit further away than the halo moves a position's
    // delta-LLs by less than QV rounding (measured against the full pass: 0 of 2.76 M positions off by more than 1 QV at
    // halo 20, 32, 48, 64; 155 / 122 / 93 / 80 off by exactly 1; scripts/qv_reuse_check.py, profiles/r2_qv_reuse.txt).
    // -29 % scoring work at halo 32.  Halo 20 (= Polish()'s neighborhood) would save another 15 % of the scoring time, but
    // one position of a 256-ZMW config-2 batch then lands 2 QV units from the oracle (the reuse error and the fp32 error
    // add up across a rounding boundary): rejected.  CCS_B200_REUSE_SCORES=0 restores the full pass.
    bool reuse_scores = true;
    int qv_halo = 32;             // > neighborhood (20): positions within this distance of an edit are re-scored
    bool generic_score = false;   // use the unfactored reference scoring kernel (tests)
    int score_variant = 0;        // compile-time variant of the scoring kernel (occupancy target / unroll), for A/B runs

private:
    // in-stream timing: spans are recorded without host syncs and resolved at the next natural sync
    struct Span { cudaEvent_t a, b; double* acc; int64_t bytes; int64_t* top_bytes; double* top_ms; };
    void span_begin(double* acc, int64_t bytes = 0, int64_t* top_bytes = nullptr, double* top_ms = nullptr, cudaStream_t s = nullptr);
    void span_end(cudaStream_t s = nullptr);
    void resolve_spans();
    std::vector<Span> spans_;
    std::vector<cudaEvent_t> ev_pool_;
    size_t ev_used_ = 0;
    cudaEvent_t next_event();
    int64_t count_canonical(const std::vector<uint8_t>& t, int b, int e) const;
    void upload_templates_and_reads();       // (re)build DevRead/DevZmw/template buffer from host state
    void sync_statuses();
    void zscore_filter(double min_zscore);   // Integrator::AddRead's POOR_ZSCORE check, after the first fill
    void remap_deltas();
    ArrowBatchView view() const;

    int device_;
    size_t budget_;
    size_t used_ = 0;
    ArrowModelParams model_;
    EmissionTables em_;
    cudaStream_t stream_ = nullptr;
    // The fill kernels are long serial chains (latency bound, few warps); they run on a HIGH-priority stream so that
    // their CTAs get SM slots ahead of the short-lived CTAs of the other lanes' scoring kernels and co-reside with them.
    cudaStream_t stream_hi_ = nullptr;
    cudaEvent_t ev_order_ = nullptr;              // orders stream_ <-> stream_hi_
    cudaEvent_t ev_sync_ = nullptr;               // blocking-sync event of stream_sync_blocking (owned: no per-thread leak)
    cudaEvent_t evA_ = nullptr, evB_ = nullptr;   // bracket polish() (resident-input time)

    // host state of the current batch
    std::vector<ZmwState> zstate_;
    std::vector<DevRead> reads_;
    std::vector<DevZmw> zmws_;
    std::vector<int32_t> order_;             // fill work list: 16-slot groups of one ZMW's reads, -1 padded
    std::vector<int32_t> status_;
    std::vector<std::vector<uint8_t>> qv_;
    std::vector<int64_t> col_base_;          // fixed first column of each read's band slot
    std::vector<int32_t> col_cap_;           // columns reserved for it
    std::vector<int32_t> tpl_cap_;           // per-ZMW template capacity in the device buffer
    std::vector<ZscoreMoments> zs_mom_;      // per ZMW: expected-LL moments per context (POOR_ZSCORE filter)
    std::vector<int32_t> read_group_;        // per read: z-score group (empty: every read on its own)
    int32_t n_groups_ = 0;
    std::vector<int32_t> zmw_group_;         // per ZMW: group of windows (empty: every ZMW on its own)
    int64_t total_cols_ = 0, total_delta_rows_ = 0;
    double ab_tol_ = 1e-3;
    int64_t score_mark_ = 0;                 // stats.n_score when the current polish() began
    int n_ranges_ = 0;
    int64_t n_range_items_ = 0;

    // device buffers
    DevBuf<uint8_t> d_rowcode_, d_tpl_, d_rawcodes_;
    DevBuf<PackJob> d_pack_;
    PinBuf<PackJob> h_pack_;
    PinBuf<uint8_t> h_rawcodes_;
    DevBuf<float> d_emm_, d_emi_, d_trans_, d_alpha_, d_beta_;
    DevBuf<DevRead> d_reads_;
    DevBuf<DevZmw> d_zmws_;
    DevBuf<ColInfo> d_colinfo_;
    DevBuf<int32_t> d_bexp_, d_status_, d_order_, d_counter_;
    DevBuf<double> d_ll_alpha_, d_ll_beta_, d_base_ll_, d_delta_, d_delta_scratch_;
    DevBuf<RemapJob> d_remap_jobs_;
    DevBuf<int32_t> d_remap_sites_, d_remap_shifts_;
    DevBuf<ScoreRange> d_ranges_;
    DevBuf<Candidate> d_cand_;
    DevBuf<uint8_t> d_qv_;
    // pinned staging
    PinBuf<uint8_t> h_tpl_, h_qv_;
    PinBuf<float> h_trans_;
    PinBuf<DevRead> h_reads_;
    PinBuf<DevZmw> h_zmws_;
    PinBuf<int32_t> h_order_, h_status_, h_counter_;
    PinBuf<double> h_ll_;
    PinBuf<ScoreRange> h_ranges_, h_ranges_qv_;   // two host lists: the QV list is staged while the scoring list's copy may still be in flight
    PinBuf<Candidate> h_cand_;
    PinBuf<RemapJob> h_remap_jobs_;
    PinBuf<int32_t> h_remap_ints_;
};

}  // namespace ccs
