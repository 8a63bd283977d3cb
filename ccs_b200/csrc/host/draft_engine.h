// Host side of the Draft Stage (the first GPU box of /root/reference/docs/img/ccs-impl.png):
// initial filtering, k-mer orientation vote, SparsePoa rounds (GPU alignment + host graph
// threading), consensus, and the subread -> draft mapping that feeds the Polish Stage
// (/root/reference/docs/how-does-ccs-work.md:19-55).
#pragma once
#include <cstdint>
#include <vector>
#include "../cuda/poa_device.h"
#include "cuda_util.h"

namespace ccs {

struct DraftParams {
    double min_snr = 2.5;      // --min-snr
    int32_t min_passes = 3;    // --min-passes
    int32_t top_passes = 60;   // --top-passes
    int32_t max_poa_reads = 5; // "draft consensus from a few subreads"
    int32_t min_length = 10, max_length = 50000;
};

struct DraftInput {
    int32_t n_zmws = 0, n_reads = 0;
    const int32_t* zmw_read_off = nullptr;
    const int64_t* read_off = nullptr;
    const uint8_t* codes = nullptr;
    const float* snr = nullptr;
    const uint8_t* cx = nullptr;
};

struct ReadMap { int32_t mapped = 0, strand = 0, tstart = 0, tend = 0, rstart = 0, rend = 0, score = 0; };

struct DraftOutput {
    std::vector<int32_t> status;               // per ZMW: ccs_zmw_status (SUCCESS = draft stage passed)
    std::vector<std::vector<uint8_t>> draft;   // per ZMW
    std::vector<ReadMap> maps;                 // per read
    std::vector<uint8_t> keep;                 // per read: survived FilterReads
};

struct DraftStats {
    double ms_align = 0;        // CUDA-event time of poa_align + traceback launches
    int64_t n_align_launches = 0, n_tasks = 0, rows = 0;
    int64_t bytes_align = 0;    // algorithmic bytes
    int64_t h2d_bytes = 0, d2h_bytes = 0;
};

class DraftEngine {
public:
    DraftEngine(int device, size_t scratch_budget_bytes);
    ~DraftEngine();
    void run(const DraftInput& in, const DraftParams& dp, DraftOutput& out);
    DraftStats stats;
    int host_threads = 8;

private:
    struct TaskHost { int zmw; int read; int rev; int V; int n; const uint8_t* bases; };   // bases: oriented read
    // staged in pinned memory by the caller: h_tasks_[0..nt), h_vbase_, h_poff_, h_preds_, h_reads_
    void align_tasks(int nt, bool any_dag, bool want_paths, int64_t rows, int64_t path_bytes, size_t n_vbase,
                     size_t n_poff, size_t n_preds, size_t n_reads);
    int device_;
    size_t budget_;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    DevBuf<PoaTask> d_tasks_;
    DevBuf<uint8_t> d_vbase_, d_reads_, d_moves_, d_paths_;
    DevBuf<int32_t> d_poff_, d_preds_, d_lo_, d_besti_, d_hrows_;
    DevBuf<PoaResult> d_results_;
    PinBuf<PoaTask> h_tasks_;
    PinBuf<uint8_t> h_vbase_, h_reads_, h_paths_;
    PinBuf<int32_t> h_poff_, h_preds_;
    PinBuf<PoaResult> h_results_;
};

}  // namespace ccs
