// Host side of the Draft Stage (the first GPU box of /root/reference/docs/img/ccs-impl.png).
// The host only filters reads (a1) and lays out task descriptors; the k-mer orientation votes, the SparsePoa
// rounds (alignment, traceback, CommitAdd on the device-resident graph), FindConsensus and the subread -> draft
// mapping all run as kernels on one stream, with two synchronisation points per chunk: after the consensus
// (draft lengths size the mapping pass) and after the mapping (SURVEY.md 8a rows a1-a5;
// /root/reference/docs/how-does-ccs-work.md:19-55).
#pragma once
#include <cstdint>
#include <vector>
#include "../cuda/poa_device.h"
#include "draft_host.h"
#include "cuda_util.h"

namespace ccs {

struct DraftParams {
    double min_snr = 2.5;      // --min-snr
    int32_t min_passes = 3;    // --min-passes
    int32_t top_passes = 60;   // --top-passes
    int32_t max_poa_reads = 5; // "draft consensus from a few subreads"
    int32_t min_length = 10, max_length = 50000;   // max_length 0: unlimited
};

struct DraftInput {
    int32_t n_zmws = 0, n_reads = 0;
    const int32_t* zmw_read_off = nullptr;
    const int64_t* read_off = nullptr;
    const uint8_t* codes = nullptr;
    const float* snr = nullptr;
    const uint8_t* cx = nullptr;
};

struct DraftOutput {
    std::vector<int32_t> status;               // per ZMW: ccs_zmw_status (SUCCESS = draft stage passed)
    std::vector<std::vector<uint8_t>> draft;   // per ZMW
    std::vector<ReadMap> maps;                 // per read
    std::vector<uint8_t> keep;                 // per read: survived FilterReads
    // windowing: grid[grid_off[r] + k] = bases of the ORIENTED read r placed before draft position k * kWindowGrid on its
    // mapping path (-1 where the path does not pass); grid_off[r] < 0: read not mapped
    std::vector<int32_t> grid;
    std::vector<int64_t> grid_off;
};

struct DraftStats {
    double ms_align = 0;        // CUDA-event time of the POA-round poa_align + traceback launches
    double ms_map = 0;          // ... of the mapping launches (linear templates)
    double ms_graph = 0;        // ... of the graph kernels (init, CommitAdd, consensus) and the k-mer votes
    int64_t n_align_launches = 0, n_graph_launches = 0, n_tasks = 0, rows = 0;
    int64_t bytes_align = 0;    // algorithmic bytes of the align launches
    int64_t bytes_map = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int64_t top_align_bytes = 0, top_map_bytes = 0;   // largest single launch: algorithmic bytes, CUDA-event ms
    double top_align_ms = 0, top_map_ms = 0;
};

class DraftEngine {
public:
    DraftEngine(int device, size_t scratch_budget_bytes);
    ~DraftEngine();
    void run(const DraftInput& in, const DraftParams& dp, DraftOutput& out);
    DraftStats stats;
    int host_threads = 8;
    void release_buffers();      // free the (grow-only) device scratch
    void set_budget(size_t bytes) { if (bytes) budget_ = bytes; }
    // The uploaded read codes of the last run() stay resident for the Polish Stage of the same lane.
    const uint8_t* device_codes() const { return d_codes_.p; }
    cudaStream_t stream() const { return stream_; }

private:
    struct Zw {                           // per-ZMW host state of one run
        std::vector<int32_t> poa_reads;   // batch read indices, seed first
        bool alive = false;
    };
    void poa_chunk(const DraftInput& in, const DraftParams& dp, DraftOutput& out, const std::vector<int32_t>& lens,
                   std::vector<Zw>& work, const std::vector<int>& zlist);
    void span(double* acc, int64_t bytes = 0, int64_t* top_bytes = nullptr, double* top_ms = nullptr);
    void span_end();
    void resolve_spans();
    int device_;
    size_t budget_;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev_sync_ = nullptr;      // blocking-sync event (owned by the engine)
    struct Span { cudaEvent_t a, b; double* acc; int64_t bytes; int64_t* top_bytes; double* top_ms; };
    std::vector<Span> spans_;
    std::vector<cudaEvent_t> ev_pool_;
    size_t ev_used_ = 0;
    DevBuf<uint8_t> d_codes_, d_desc_, d_rev_, d_moves_, d_draft_;
    DevBuf<uint32_t> d_meta_;
    DevBuf<int32_t> d_pred0_, d_predx_, d_rank_, d_order_, d_lo_, d_hrows_, d_scratch_, d_draft_len_, d_grid_, d_col_;
    DevBuf<PoaStep> d_steps_;
    DevBuf<PoaResult> d_results_;
    PinBuf<uint8_t> h_codes_, h_desc_, h_draft_, h_rev_;
    PinBuf<int32_t> h_draft_len_, h_grid_;
    PinBuf<PoaResult> h_results_;
};

}  // namespace ccs
