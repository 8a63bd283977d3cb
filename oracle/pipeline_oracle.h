// TEST INFRASTRUCTURE -- the whole per-ZMW path of `ccs` restated on the CPU: initial filtering,
// draft (SparsePoa), subread -> draft mapping, Arrow polish, QVs, final gates
// (/root/reference/docs/how-does-ccs-work.md:19-112, status order
// /root/reference/docs/faq/reports-aux-files.md:143-159).  PARITY UNPINNED, see arrow_oracle.h.
#pragma once
#include "arrow_oracle.h"
#include "poa_oracle.h"

namespace oracle {

enum ZmwStatus : int {
    Z_POOR_SNR = 0, Z_NO_SUBREADS, Z_TOO_FEW_PASSES, Z_LOW_PASS_SHORTCUT, Z_HETERODUPLEXES, Z_COVERAGE_DROPS,
    Z_INSUFFICIENT_SPANS, Z_TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT, Z_DRAFT_FAILURE, Z_TOO_LONG, Z_TOO_SHORT,
    Z_TOO_MANY_UNUSABLE, Z_EMPTY_WINDOW_DURING_POLISHING, Z_NON_CONVERGENT, Z_POOR_QUALITY, Z_EXCEPTION_THROWN, Z_SUCCESS
};

struct CcsConfig {
    double min_snr = 2.5;
    int min_passes = 3;
    int top_passes = 60;
    int max_poa_reads = 5;
    int min_length = 10, max_length = 50000;
    double min_rq = 0.99;
    double min_active_fraction = 0.5;
    // Windowing (docs/how-does-ccs-work.md:57-61; DESIGN.md "Windowing"): a draft of at least 2 * window_size bases is
    // polished as floor(J / window_size) windows -- cores [k W, (k+1) W), the last one running to the end -- each
    // padded by window_overlap bases on both sides; 0 = one window.  Both are multiples of WINDOW_GRID.
    int window_size = 1024, window_overlap = 64;
    PolishConfig polish;
};

struct Window { int a, b, c0, c1; };   // padded template range [a, b), core [c0, c1)
std::vector<Window> make_windows(int J, int window_size, int window_overlap);
// the slice of read `parent` (native coordinates [ns, ne)) that the mapping places on window template range [ts, te)
struct WindowRead { int parent, ns, ne, strand, ts, te; };
std::vector<WindowRead> window_reads(const Window& w, const std::vector<ReadMapping>& maps, const std::vector<int>& lens);

struct CcsZmwResult {
    int status = Z_EXCEPTION_THROWN;
    std::vector<uint8_t> draft, seq, qv;
    double rq = 0;
    int np = 0;
    std::vector<ReadMapping> maps;     // one per input read (mapped=false: filtered or not placed)
    std::vector<double> read_ll;
    std::vector<int> read_status;      // ReadStatus, 4 = not used
    PolishResult pr;
};

// FilterReads (docs/how-does-ccs-work.md:19-32): keep[r] and the number of full-length reads kept
int filter_reads(const std::vector<int>& lens, const uint8_t* cx, int top_passes, std::vector<char>& keep);

void draft_zmw(const CcsConfig& cfg, int nreads, const uint8_t* codes, const int64_t* read_off, const uint8_t* cx,
               const float snr[4], CcsZmwResult& out, std::vector<char>& keep);

void ccs_zmw(const ccs::ArrowModelParams& model, const CcsConfig& cfg, int nreads, const uint8_t* codes,
             const int64_t* read_off, const uint8_t* cx, const float snr[4], CcsZmwResult& out);

}  // namespace oracle
