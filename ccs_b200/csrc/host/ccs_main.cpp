// `ccs` command line: subreads.bam -> hifi_reads.bam through the GPU stages
//   ccs <in.subreads.bam> <out.bam> [flags]        (/root/reference/docs/index.md:52-64)
// Hot-path subset of the reference's flags: --min-snr --min-passes --min-length
// --max-length --min-rq --top-passes --chunk i/N -j --by-strand --report-file --report-json --metrics-json
// --hifi-summary-json --batch-size --log-level --refresh-rate, plus --gpus / --device.
//
// The process is the pipeline the reference draws (/root/reference/docs/img/ccs-impl.png):
//   reader thread (BGZF inflate on a worker pool) -> bounded queue of ZMW batches -> one stage worker per
//   (GPU, pipeline slot), each with its own ccsgpu_ctx -> in-order writer (BAM, metrics, report) on the main thread.
// Batches are finer than GPUs and are taken by whichever worker is free (work stealing by batch), yet written in
// input order, so the output does not depend on --gpus / --pipeline (/root/reference/docs/faq/parallelize.md:7-28).
// Per-ZMW failures are counted in <prefix>.ccs_report.txt, never fatal
// (/root/reference/docs/faq/reports-aux-files.md:16-72,143-159); a BAM without the chemistry triple is fatal
// (/root/reference/docs/changelog.md:66); so is a damaged input file or an output that cannot be written.
// Compute goes through the C ABI only (include/ccsgpu.h).
#include "../../../include/ccsgpu.h"
#include "bam_io.h"
#include "draft_host.h"
#include "parallel.h"
#include <cuda_runtime.h>
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

using namespace ccs;

namespace {

struct Options {
    std::string in, out, report, report_json, metrics, hifi_summary, model_path;
    ccs_draft_cfg d;
    ccs_polish_cfg p;
    int chunk_i = 1, chunk_n = 1, batch = 256, device = 0, gpus = 1, log_level = 1, pipeline = 2, threads = 0;
    double refresh_rate = 5.0;
    bool by_strand = false;
};

void usage() {
    std::fprintf(stderr,
                 "ccs (ccs-b200) - Generate circular consensus sequences (ccs) from subreads on B200 GPUs.\n"
                 "Usage: ccs [options] <IN.subreads.bam> <OUT.ccs.bam>\n"
                 "  --min-snr FLOAT      Minimum SNR of subreads to use for generating CCS. [2.5]\n"
                 "  --min-passes INT     Minimum number of full-length subreads required to generate CCS for a ZMW. [3]\n"
                 "  --top-passes INT     Pick at maximum the top N passes for each ZMW. [60]\n"
                 "  --min-length INT     Minimum draft length before polishing. [10]\n"
                 "  --max-length INT     Maximum draft length before polishing (0: no limit). [50000]\n"
                 "  --min-rq FLOAT       Minimum predicted accuracy in [0, 1]. [0.99]\n"
                 "  --by-strand          Generate a consensus for each strand.\n"
                 "  --window-size INT    Polish drafts of 2x this length and longer as independent windows of this many bases\n"
                 "                       (rounded up to a multiple of 64; 0: never split). [1024]\n"
                 "  --window-overlap INT Bases of padding on both sides of a window. [64]\n"
                 "  --chunk i/N          Operate on a single chunk. Format i/N, where i in [1,N].\n"
                 "  -j,--num-threads INT Number of host threads to use, 0 means autodetection. [0]\n"
                 "  --report-file FILE   Where to write the results report. [<out prefix>.ccs_report.txt]\n"
                 "  --report-json FILE   Where to write the results report as JSON. [none]\n"
                 "  --metrics-json FILE  Where to write the zmw_metrics JSON (gzip). [<out prefix>.zmw_metrics.json.gz]\n"
                 "  --hifi-summary-json FILE  Where to write the HiFi summary JSON. [none]\n"
                 "  --model-path FILE    Arrow model JSON (else $SMRT_CHEMISTRY_BUNDLE_DIR/arrow/model.json, else built-in synthetic).\n"
                 "  --batch-size INT     ZMWs per GPU batch. [256]\n"
                 "  --gpus INT           GPUs to use (devices --device .. --device+N-1), 0: all visible. [1]\n"
                 "  --device INT         First CUDA device. [0]\n"
                 "  --pipeline INT       Stage instances (batches in flight) per GPU. [2]\n"
                 "  --refresh-rate FLOAT Seconds between progress lines at --log-level INFO. [5]\n"
                 "  --log-level STR      Set log level: DEBUG INFO WARN. [WARN]\n");
}

bool parse(int argc, char** argv, Options& o) {
    ccs_draft_cfg_default(&o.d);
    ccs_polish_cfg_default(&o.p);
    std::vector<std::string> pos;
    for (int k = 1; k < argc; ++k) {
        std::string a = argv[k];
        auto val = [&](const char* name) -> const char* {
            if (k + 1 >= argc) { std::fprintf(stderr, "ccs: %s needs a value\n", name); std::exit(2); }
            return argv[++k];
        };
        if (a == "--min-snr") o.d.min_snr = std::atof(val("--min-snr"));
        else if (a == "--min-passes") o.d.min_passes = std::atoi(val("--min-passes"));
        else if (a == "--top-passes") o.d.top_passes = std::atoi(val("--top-passes"));
        else if (a == "--min-length") o.d.min_length = o.p.min_length = std::atoi(val("--min-length"));
        else if (a == "--max-length") o.d.max_length = o.p.max_length = std::atoi(val("--max-length"));
        else if (a == "--min-rq") o.p.min_rq = std::atof(val("--min-rq"));
        else if (a == "--by-strand") o.by_strand = true;
        else if (a == "--window-size") o.p.window_size = std::max(0, std::atoi(val("--window-size")));
        else if (a == "--window-overlap") o.p.window_overlap = std::max(0, std::atoi(val("--window-overlap")));
        else if (a == "--report-file") o.report = val("--report-file");
        else if (a == "--report-json") o.report_json = val("--report-json");
        else if (a == "--metrics-json") o.metrics = val("--metrics-json");
        else if (a == "--hifi-summary-json") o.hifi_summary = val("--hifi-summary-json");
        else if (a == "--model-path") o.model_path = val("--model-path");
        else if (a == "--batch-size") o.batch = std::max(1, std::atoi(val("--batch-size")));
        else if (a == "--device") o.device = std::atoi(val("--device"));
        else if (a == "--gpus") o.gpus = std::max(0, std::atoi(val("--gpus")));
        else if (a == "--pipeline") o.pipeline = std::max(1, std::min(4, std::atoi(val("--pipeline"))));
        else if (a == "-j" || a == "--num-threads") o.threads = std::max(0, std::atoi(val("-j")));
        else if (a == "--refresh-rate") o.refresh_rate = std::max(0.1, std::atof(val("--refresh-rate")));
        else if (a == "--log-level") { std::string l = val("--log-level"); o.log_level = l == "DEBUG" ? 3 : (l == "INFO" ? 2 : 1); }
        else if (a == "--chunk") {
            if (std::sscanf(val("--chunk"), "%d/%d", &o.chunk_i, &o.chunk_n) != 2 || o.chunk_i < 1 || o.chunk_i > o.chunk_n) {
                std::fprintf(stderr, "ccs: --chunk expects i/N with i in [1,N]\n");
                return false;
            }
        } else if (a == "-h" || a == "--help") { usage(); std::exit(0); }
        else if (!a.empty() && a[0] == '-') { std::fprintf(stderr, "ccs: unknown option %s\n", a.c_str()); return false; }
        else pos.push_back(a);
    }
    if (pos.size() != 2) { usage(); return false; }
    o.in = pos[0]; o.out = pos[1];
    std::string pre = o.out;
    const size_t dot = pre.rfind(".bam");
    if (dot != std::string::npos) pre = pre.substr(0, dot);
    if (o.report.empty()) o.report = pre + ".ccs_report.txt";
    if (o.metrics.empty()) o.metrics = pre + ".zmw_metrics.json.gz";
    return true;
}

struct ReadStat { int32_t len, np; float rq; int32_t q30_bases; };
struct Report {
    int64_t input = 0, pass = 0, counts[17] = {0};
    std::vector<ReadStat> reads;       // every written consensus read
};

std::string commas(int64_t v) {
    std::string s = std::to_string(v), o;
    for (size_t k = 0; k < s.size(); ++k) { if (k && (s.size() - k) % 3 == 0) o += ','; o += s[k]; }
    return o;
}

struct Block { int64_t n = 0, yield = 0, mean = 0, median = 0, n50 = 0; int qmed = 0, np_mean = 0; };

Block block_of(const std::vector<ReadStat>& rs, double rq_lo, double rq_hi) {
    Block b;
    std::vector<int32_t> sl;
    std::vector<float> sq;
    double np = 0;
    for (const ReadStat& r : rs) if (r.rq >= rq_lo && r.rq < rq_hi) { sl.push_back(r.len); sq.push_back(r.rq); b.yield += r.len; np += r.np; }
    b.n = (int64_t)sl.size();
    if (!b.n) return b;
    std::sort(sl.begin(), sl.end());
    std::sort(sq.begin(), sq.end());
    b.mean = b.yield / b.n;
    b.median = sl[sl.size() / 2];
    int64_t acc = 0;
    for (size_t k = sl.size(); k-- > 0;) { acc += sl[k]; if (2 * acc >= b.yield) { b.n50 = sl[k]; break; } }
    b.qmed = (int)std::lround(-10.0 * std::log10(std::max(1e-10, 1.0 - (double)sq[sq.size() / 2])));
    b.np_mean = (int)std::lround(np / b.n);
    return b;
}

const struct { const char* label; const char* key; int status; } kRows[] = {
    {"Below SNR threshold           ", "below_snr_threshold", CCS_ZMW_POOR_SNR},
    {"Median length filter          ", "median_length_filter", CCS_ZMW_NO_SUBREADS},
    {"Lacking full passes           ", "lacking_full_passes", CCS_ZMW_TOO_FEW_PASSES},
    {"Heteroduplex insertions       ", "heteroduplex_insertions", CCS_ZMW_HETERODUPLEXES},
    {"Coverage drops                ", "coverage_drops", CCS_ZMW_COVERAGE_DROPS},
    {"Insufficient draft cov        ", "insufficient_draft_cov", CCS_ZMW_INSUFFICIENT_SPANS},
    {"Draft too different           ", "draft_too_different", CCS_ZMW_TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT},
    {"Draft generation error        ", "draft_generation_error", CCS_ZMW_DRAFT_FAILURE},
    {"Draft above --max-length      ", "draft_above_max_length", CCS_ZMW_TOO_LONG},
    {"Draft below --min-length      ", "draft_below_min_length", CCS_ZMW_TOO_SHORT},
    {"Reads failed polishing        ", "reads_failed_polishing", CCS_ZMW_TOO_MANY_UNUSABLE},
    {"Empty coverage windows        ", "empty_coverage_windows", CCS_ZMW_EMPTY_WINDOW_DURING_POLISHING},
    {"CCS did not converge          ", "ccs_did_not_converge", CCS_ZMW_NON_CONVERGENT},
    {"CCS below minimum RQ          ", "ccs_below_minimum_rq", CCS_ZMW_POOR_QUALITY},
    {"Unknown error                 ", "unknown_error", CCS_ZMW_EXCEPTION_THROWN}};

// <prefix>.ccs_report.txt (docs/faq/reports-aux-files.md:16-72); unit = ZMWs, or single-strand reads with --by-strand
bool write_report(const std::string& path, const Report& r, bool by_strand) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    const int64_t fail = r.input - r.pass;
    auto pct = [](int64_t a, int64_t b) { return b ? 100.0 * a / b : 0.0; };
    const std::string unit = by_strand ? "Single-Strand Reads" : "ZMWs";
    auto label = [&](const char* what) { std::string l = unit + " " + what; if (l.size() < 30) l.resize(30, ' '); return l; };
    std::fprintf(f, "%s: %lld\n\n", label("input").c_str(), (long long)r.input);
    std::fprintf(f, "%s: %lld (%.2f%%)\n", label("pass filters").c_str(), (long long)r.pass, pct(r.pass, r.input));
    std::fprintf(f, "%s: %lld (%.2f%%)\n", label("fail filters").c_str(), (long long)fail, pct(fail, r.input));
    std::fprintf(f, "%s: 0 (0.00%%)\n\n", label("shortcut filters").c_str());
    std::fprintf(f, "ZMWs with tandem repeats      : 0 (0.00%%)\n\n");
    std::fprintf(f, "Exclusive failed counts\n");
    for (const auto& row : kRows)
        std::fprintf(f, "%s: %lld (%.2f%%)\n", row.label, (long long)r.counts[row.status], pct(r.counts[row.status], fail));
    std::fprintf(f, "\nAdditional passing metrics\nZMWs missing adapters         : 0 (0.000%%)\n");
    std::fprintf(f, "\n- - - - - - - - - - - - - - - : - - - - -\n\n");
    const Block hifi = block_of(r.reads, 0.99, 2.0), low = block_of(r.reads, -1.0, 0.99), q30 = block_of(r.reads, 0.999, 2.0);
    std::fprintf(f, "HiFi Reads                    : %s\n", commas(hifi.n).c_str());
    std::fprintf(f, "HiFi Yield (bp)               : %s\n", commas(hifi.yield).c_str());
    std::fprintf(f, "HiFi Read Length (mean, bp)   : %s\n", commas(hifi.mean).c_str());
    std::fprintf(f, "HiFi Read Length (median, bp) : %s\n", commas(hifi.median).c_str());
    std::fprintf(f, "HiFi Read Length N50 (bp)     : %s\n", commas(hifi.n50).c_str());
    std::fprintf(f, "HiFi Read Quality (median)    : %d\n", hifi.qmed);
    std::fprintf(f, "HiFi Number of Passes (mean)  : %d\n\n", hifi.np_mean);
    std::fprintf(f, "<Q20 Reads                    : %s\n", commas(low.n).c_str());
    std::fprintf(f, "<Q20 Yield (bp)               : %s\n", commas(low.yield).c_str());
    std::fprintf(f, "<Q20 Read Length (mean, bp)   : %s\n", commas(low.mean).c_str());
    std::fprintf(f, "<Q20 Read Length (median, bp) : %s\n", commas(low.median).c_str());
    std::fprintf(f, "<Q20 Read Quality (median)    : %d\n\n", low.qmed);
    std::fprintf(f, ">=Q30 Reads                   : %s\n", commas(q30.n).c_str());
    std::fprintf(f, ">=Q30 Yield (bp)              : %s\n", commas(q30.yield).c_str());
    std::fprintf(f, ">=Q30 Read Length (mean, bp)  : %s\n", commas(q30.mean).c_str());
    std::fprintf(f, ">=Q30 Read Length (median, bp): %s\n", commas(q30.median).c_str());
    std::fprintf(f, ">=Q30 Read Quality (median)   : %d\n\n", q30.qmed);
    int64_t bases = 0, b30 = 0;
    for (const ReadStat& x : r.reads) { bases += x.len; b30 += x.q30_bases; }
    std::fprintf(f, "Base quality >=Q30 (bp)       : %s (%.1f%%)\n", commas(b30).c_str(), pct(b30, bases));
    return std::fclose(f) == 0;
}

// --report-json: the same numbers, machine readable (docs/changelog.md:72, docs/faq/sqiie.md:42)
bool write_report_json(const std::string& path, const Report& r, bool by_strand) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    const Block hifi = block_of(r.reads, 0.99, 2.0), low = block_of(r.reads, -1.0, 0.99), q30 = block_of(r.reads, 0.999, 2.0);
    std::fprintf(f, "{\n  \"unit\": \"%s\",\n  \"input\": %lld,\n  \"pass_filters\": %lld,\n  \"fail_filters\": %lld,\n"
                    "  \"shortcut_filters\": 0,\n  \"exclusive_failed_counts\": {\n", by_strand ? "single_strand_reads" : "zmws",
                 (long long)r.input, (long long)r.pass, (long long)(r.input - r.pass));
    for (size_t k = 0; k < sizeof(kRows) / sizeof(kRows[0]); ++k)
        std::fprintf(f, "    \"%s\": %lld%s\n", kRows[k].key, (long long)r.counts[kRows[k].status], k + 1 < sizeof(kRows) / sizeof(kRows[0]) ? "," : "");
    auto blk = [&](const char* name, const Block& b, bool last) {
        std::fprintf(f, "  \"%s\": {\"reads\": %lld, \"yield_bp\": %lld, \"read_length_mean\": %lld, \"read_length_median\": %lld, "
                        "\"read_length_n50\": %lld, \"read_quality_median\": %d, \"passes_mean\": %d}%s\n",
                     name, (long long)b.n, (long long)b.yield, (long long)b.mean, (long long)b.median, (long long)b.n50, b.qmed, b.np_mean,
                     last ? "" : ",");
    };
    std::fprintf(f, "  },\n");
    blk("hifi", hifi, false); blk("below_q20", low, false); blk("q30_and_above", q30, true);
    std::fprintf(f, "}\n");
    return std::fclose(f) == 0;
}

// --hifi-summary-json (docs/faq/sqiie.md:45): yield statistics of the >= Q20 reads
bool write_hifi_summary(const std::string& path, const Report& r, double seconds) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    const Block hifi = block_of(r.reads, 0.99, 2.0);
    int64_t umy = 0;
    for (const ReadStat& x : r.reads) umy += x.len;
    std::fprintf(f, "{\n  \"zmws_input\": %lld,\n  \"zmws_written\": %lld,\n  \"unique_molecular_yield_bp\": %lld,\n"
                    "  \"hifi_reads\": %lld,\n  \"hifi_yield_bp\": %lld,\n  \"hifi_read_length_mean_bp\": %lld,\n"
                    "  \"hifi_read_quality_median\": %d,\n  \"elapsed_s\": %.3f\n}\n",
                 (long long)r.input, (long long)r.pass, (long long)umy, (long long)hifi.n, (long long)hifi.yield, (long long)hifi.mean,
                 hifi.qmed, seconds);
    return std::fclose(f) == 0;
}

// One batch travelling through the pipeline.
struct BatchIO {
    int64_t seq_no = 0;
    std::vector<RawZmw> raw;                   // as grouped by the reader thread (validated, not decoded)
    std::vector<ZmwSubreads> zmws;             // decoded by the stage worker; with --by-strand: one entry per strand bucket
    std::vector<uint8_t> strand_tag;           // 0: whole ZMW, 1: fwd bucket, 2: rev bucket
    std::vector<int32_t> zmw_read_off, hole, status, npass, iters, napp, rstatus;
    std::vector<int64_t> read_off, seq_off, ntest;
    std::vector<uint8_t> codes, cx, seq, qv;
    std::vector<float> snr, rq;
    std::vector<double> rll;
    std::string err;
    int rc = CCS_OK;
};

template <class T>
class BoundedQueue {
public:
    explicit BoundedQueue(size_t cap) : cap_(cap) {}
    void push(T&& v) {
        std::unique_lock<std::mutex> lk(m_);
        not_full_.wait(lk, [&] { return q_.size() < cap_ || closed_; });
        if (closed_) return;
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    bool pop(T& out) {
        std::unique_lock<std::mutex> lk(m_);
        not_empty_.wait(lk, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void close() { std::lock_guard<std::mutex> lk(m_); closed_ = true; not_empty_.notify_all(); not_full_.notify_all(); }
private:
    std::mutex m_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};

// --by-strand (docs/faq/mode-by-strand.md:16-24): orientation of every read with respect to the read closest to the
// median length (k-mer vote), then one bucket per strand, each treated like a ZMW of its own
void split_by_strand(const ZmwSubreads& z, ZmwSubreads& fwd, ZmwSubreads& rev) {
    fwd = ZmwSubreads(); rev = ZmwSubreads();
    fwd.hole = rev.hole = z.hole;
    std::memcpy(fwd.snr, z.snr, sizeof(z.snr)); std::memcpy(rev.snr, z.snr, sizeof(z.snr));
    if (z.reads.empty()) return;
    std::vector<int32_t> ls;
    for (const auto& r : z.reads) ls.push_back((int32_t)r.codes.size());
    std::vector<int32_t> sl(ls);
    std::sort(sl.begin(), sl.end());
    const int32_t med = sl[sl.size() / 2];
    size_t ref = 0;
    for (size_t k = 1; k < ls.size(); ++k) if (std::abs(ls[k] - med) < std::abs(ls[ref] - med)) ref = k;
    std::vector<uint8_t> rb(ls[ref]);
    orient(z.reads[ref].codes.data(), ls[ref], false, rb.data());
    KmerSet ks;
    ks.build(rb.data(), ls[ref]);
    for (size_t k = 0; k < z.reads.size(); ++k) {
        int64_t f = 0, c = 0;
        if (k != ref) ks.count(z.reads[k].codes.data(), std::min(ls[k], kPoaVoteBases), f, c);
        (c > f ? rev : fwd).reads.push_back(z.reads[k]);
    }
}

}  // namespace

int main(int argc, char** argv) {
    Options o;
    if (!parse(argc, argv, o)) return 2;
    const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
    const int host_threads = o.threads > 0 ? o.threads : (int)hc;
    SubreadBamReader reader;
    std::string err;
    if (!reader.open(o.in, err, std::max(1, std::min(16, host_threads / 2)))) { std::fprintf(stderr, "ccs: %s\n", err.c_str()); return 1; }
    if (!reader.chemistry_ok()) {
        std::fprintf(stderr, "ccs: missing chemistry information (BINDINGKIT / SEQUENCINGKIT / BASECALLERVERSION) in the "
                             "read group of %s; cannot select an Arrow model\n", o.in.c_str());
        return 1;
    }
    // --chunk i/N (docs/faq/parallelize.md:8-28): the ZMWs with index in [z_begin, z_end) counted from where the reader
    // stands after select_chunk() -- the chunk's first record when <in>.pbi exists, the file start otherwise
    int64_t z_begin = 0, z_end = INT64_MAX;
    if (o.chunk_n > 1) {
        bool used_index = false;
        if (!select_chunk(reader, o.in, o.chunk_i, o.chunk_n, z_begin, z_end, used_index, err)) {
            std::fprintf(stderr, "ccs: %s\n", err.c_str());
            return 1;
        }
    }
    int64_t zmws_expected = -1;                 // for the ETA of the progress line: needs the .pbi
    {
        PbiIndex pbi;
        if (pbi.read(o.in + ".pbi") && pbi.size() > 0) {
            const int64_t total = (int64_t)pbi.zmw_starts().size() - 1;
            zmws_expected = total * o.chunk_i / o.chunk_n - total * (o.chunk_i - 1) / o.chunk_n;
        }
    }
    std::vector<uint8_t> model(ccs_model_sizeof());
    ccs_model_synthetic(model.data());          // the only chemistry this build ships (DESIGN.md "Model")
    {   // model injection: --model-path, or the chemistry bundle directory (docs/faq/chemistry.md:28-56)
        std::string mp = o.model_path;
        if (mp.empty()) if (const char* b = std::getenv("SMRT_CHEMISTRY_BUNDLE_DIR")) {
            const std::string cand = std::string(b) + "/arrow/model.json";
            if (FILE* t = std::fopen(cand.c_str(), "rb")) { std::fclose(t); mp = cand; }
        }
        if (!mp.empty() && ccs_model_load_json(mp.c_str(), model.data()) != CCS_OK) {
            std::fprintf(stderr, "ccs: cannot load the Arrow model from %s\n", mp.c_str());
            return 1;
        }
    }
    // stage workers: `pipeline` instances (one ccsgpu_ctx each) on each of `gpus` devices
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess) n_dev = 0;
    int gpus = o.gpus == 0 ? std::max(1, n_dev - o.device) : o.gpus;
    if (n_dev > 0 && o.device + gpus > n_dev) { std::fprintf(stderr, "ccs: --device %d --gpus %d exceeds the %d visible devices\n", o.device, gpus, n_dev); return 1; }
    const int n_workers = gpus * o.pipeline;
    if (!std::getenv("CCS_B200_THREADS"))      // -j: the host threads are split between the stage instances
        setenv("CCS_B200_THREADS", std::to_string(std::max(1, host_threads / n_workers)).c_str(), 1);
    std::vector<ccsgpu_ctx*> ctxs;
    for (int g = 0; g < gpus; ++g) {
        size_t free_b = 0, total_b = 0;
        cudaSetDevice(o.device + g);
        cudaMemGetInfo(&free_b, &total_b);                      // fails harmlessly without a device: budget 0
        const size_t budget = (size_t)((double)free_b * 0.85 / o.pipeline);
        for (int k = 0; k < o.pipeline; ++k) {
            int cerr = 0;
            ccsgpu_ctx* c = ccsgpu_create(o.device + g, model.data(), budget, &cerr);
            if (!c) { std::fprintf(stderr, "ccs: %s\n", ccsgpu_last_error(nullptr)); return 1; }
            if (o.pipeline > 1) ccsgpu_set_lanes(c, 3);
            ctxs.push_back(c);
        }
    }
    std::string cl;
    for (int k = 0; k < argc; ++k) { if (k) cl += ' '; cl += argv[k]; }
    CcsBamWriter writer;
    if (!writer.open(o.out, reader.header_text(), reader.movie(), reader.read_group_id(), cl)) {
        std::fprintf(stderr, "ccs: cannot write %s\n", o.out.c_str());
        return 1;
    }
    const int decode_threads = std::max(1, host_threads / std::max(1, n_workers));
    auto process = [&](ccsgpu_ctx* ctx, BatchIO& B) -> int {
        // decode the records of the batch (SEQ + pw -> emission codes) here, in the stage worker: the reader thread only
        // inflates and groups, so decoding scales with the number of workers
        {
            std::vector<ZmwSubreads> dec(B.raw.size());
            std::string derr;
            std::mutex emu;
            parallel_for((int)B.raw.size(), decode_threads, [&](int k) {
                std::string e;
                if (!decode_zmw(B.raw[k], dec[k], e)) { std::lock_guard<std::mutex> g(emu); if (derr.empty()) derr = e; }
            }, /*min_items_per_thread=*/4);
            if (!derr.empty()) { B.err = derr; return CCS_ERR_ARG; }
            B.raw.clear(); B.raw.shrink_to_fit();
            for (auto& z : dec) {
                if (o.by_strand) {
                    ZmwSubreads f, r;
                    split_by_strand(z, f, r);
                    B.zmws.push_back(std::move(f)); B.strand_tag.push_back(1);
                    B.zmws.push_back(std::move(r)); B.strand_tag.push_back(2);
                } else { B.zmws.push_back(std::move(z)); B.strand_tag.push_back(0); }
            }
        }
        const int nz = (int)B.zmws.size();
        B.zmw_read_off.assign(1, 0); B.read_off.assign(1, 0);
        size_t maxlen = 1, total = 0;
        for (const auto& zz : B.zmws) for (const auto& rd : zz.reads) total += rd.codes.size();
        B.codes.reserve(total);
        for (const auto& zz : B.zmws) {
            for (const auto& rd : zz.reads) {
                B.codes.insert(B.codes.end(), rd.codes.begin(), rd.codes.end());
                B.read_off.push_back((int64_t)B.codes.size());
                B.cx.push_back(rd.cx);
                maxlen = std::max(maxlen, rd.codes.size());
            }
            B.zmw_read_off.push_back((int32_t)B.cx.size());
            B.snr.insert(B.snr.end(), zz.snr, zz.snr + 4);
            B.hole.push_back(zz.hole);
        }
        const int nr = (int)B.cx.size();
        ccs_batch b{nz, nr, B.zmw_read_off.data(), B.read_off.data(), B.codes.data(), B.snr.data(), B.cx.data(), B.hole.data()};
        int64_t cap = (int64_t)maxlen * 2 * nz + 1024;
        for (int attempt = 0; attempt < 2; ++attempt) {
            B.seq_off.assign(nz + 1, 0); B.seq.assign(cap, 0); B.qv.assign(cap, 0); B.rq.assign(nz, 0); B.status.assign(nz, 0);
            B.npass.assign(nz, 0); B.iters.assign(nz, 0); B.napp.assign(nz, 0); B.ntest.assign(nz, 0);
            B.rll.assign(nr, 0); B.rstatus.assign(nr, 0);
            ccs_results r{cap, B.seq_off.data(), B.seq.data(), B.qv.data(), B.rq.data(), B.status.data(), B.npass.data(),
                          B.iters.data(), B.napp.data(), B.ntest.data(), B.rll.data(), B.rstatus.data()};
            const int rc = ccsgpu_ccs(ctx, &b, &o.d, &o.p, &r);
            if (rc == CCS_ERR_CAPACITY) { cap = r.seq_cap + 1024; continue; }
            if (rc != CCS_OK) { B.err = ccsgpu_last_error(ctx); return rc; }
            return CCS_OK;
        }
        B.err = "result capacity";
        return CCS_ERR_CAPACITY;
    };

    // ---- reader -> queue -> workers -> ordered results ----------------------------------------------------------
    BoundedQueue<BatchIO> todo((size_t)2 * n_workers);
    std::mutex done_m;
    std::condition_variable done_cv;
    std::map<int64_t, BatchIO> done;
    std::atomic<int64_t> n_batches(-1);          // set by the reader when the input is exhausted
    std::atomic<bool> abort_flag(false);
    std::string reader_error;
    std::thread reader_thread([&]() {
        int64_t z_index = 0, seq = 0;
        BatchIO B;
        auto flush = [&]() {
            if (B.raw.empty()) return;
            B.seq_no = seq++;
            todo.push(std::move(B));
            B = BatchIO();
        };
        RawZmw z;
        while (!abort_flag.load() && z_index < z_end && reader.next_zmw_raw(z)) {
            if (z_index >= z_begin) {
                B.raw.push_back(std::move(z));
                z = RawZmw();
                if ((int)B.raw.size() >= o.batch) flush();
            }
            ++z_index;
        }
        if (z_index < z_end) reader_error = reader.error();   // stopped before the chunk's end: end of data, or damage
        flush();
        n_batches.store(seq);
        todo.close();
        { std::lock_guard<std::mutex> lk(done_m); }
        done_cv.notify_all();
    });
    std::vector<std::thread> workers;
    for (int w = 0; w < n_workers; ++w)
        workers.emplace_back([&, w]() {
            BatchIO B;
            while (todo.pop(B)) {
                B.rc = process(ctxs[w], B);
                { std::lock_guard<std::mutex> lk(done_m); { const int64_t key = B.seq_no; done.emplace(key, std::move(B)); } }
                done_cv.notify_all();
                B = BatchIO();
            }
        });

    // <prefix>.zmw_metrics.json.gz: one entry per input ZMW (docs/faq/reports-aux-files.md:100-171)
    static const char* kStatusName[17] = {"POOR_SNR", "NO_SUBREADS", "TOO_FEW_PASSES", "LOW_PASS_SHORTCUT", "HETERODUPLEXES",
        "COVERAGE_DROPS", "INSUFFICIENT_SPANS", "TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT", "DRAFT_FAILURE", "TOO_LONG", "TOO_SHORT",
        "TOO_MANY_UNUSABLE", "EMPTY_WINDOW_DURING_POLISHING", "NON_CONVERGENT", "POOR_QUALITY", "EXCEPTION_THROWN", "SUCCESS"};
    gzFile mz = gzopen(o.metrics.c_str(), "wb");
    if (mz) gzputs(mz, "{\n  \"zmws\": [");
    bool first_metric = true;
    Report rep;
    const auto t0 = std::chrono::steady_clock::now();
    auto last_log = t0;
    std::deque<std::pair<double, std::pair<int64_t, int64_t>>> hist;   // (time, (ZMWs, CCSs)) for the last-minute rates
    int exit_code = 0;
    std::string fatal;
    for (int64_t next = 0;; ++next) {
        BatchIO B;
        {
            std::unique_lock<std::mutex> lk(done_m);
            done_cv.wait(lk, [&] { return done.count(next) || (n_batches.load() >= 0 && next >= n_batches.load()); });
            auto it = done.find(next);
            if (it == done.end()) break;         // every batch written
            B = std::move(it->second);
            done.erase(it);
        }
        if (B.rc != CCS_OK) { fatal = B.err; exit_code = 1; abort_flag.store(true); break; }
        for (int zi = 0; zi < (int)B.zmws.size(); ++zi) {
            if (B.strand_tag[zi] != 0 && B.zmws[zi].reads.empty()) continue;    // by-strand: nothing on this strand
            ++rep.input;
            ++rep.counts[B.status[zi]];
            const char* suffix = B.strand_tag[zi] == 1 ? "ccs/fwd" : (B.strand_tag[zi] == 2 ? "ccs/rev" : "ccs");
            if (mz) {
                const ZmwSubreads& zz = B.zmws[zi];
                int64_t poly = 0;
                std::vector<int32_t> ls;
                for (const auto& rd : zz.reads) { poly += (int64_t)rd.codes.size(); ls.push_back((int32_t)rd.codes.size()); }
                std::sort(ls.begin(), ls.end());
                const int64_t clen = B.seq_off[zi + 1] - B.seq_off[zi];
                const int64_t insert = clen > 0 ? clen : (ls.empty() ? 0 : ls[ls.size() / 2]);
                const bool has_rq = clen > 0;
                gzprintf(mz, "%s\n    {\"effective_coverage\": %d, \"has_tandem_repeat\": false, \"insert_size\": %lld, "
                             "\"num_full_passes\": %d, \"polymerase_length\": %lld, \"predicted_accuracy\": %.6f, "
                             "\"status\": \"%s\", \"zmw\": \"%s/%d%s\"}",
                         first_metric ? "" : ",", B.npass[zi] + (B.npass[zi] > 0 ? 1 : 0), (long long)insert, B.npass[zi],
                         (long long)poly, has_rq ? (double)B.rq[zi] : -1.0, kStatusName[B.status[zi]], reader.movie().c_str(),
                         zz.hole, B.strand_tag[zi] == 1 ? "/fwd" : (B.strand_tag[zi] == 2 ? "/rev" : ""));
                first_metric = false;
            }
            if (B.status[zi] != CCS_ZMW_SUCCESS) continue;
            ++rep.pass;
            CcsRecord rec;
            rec.hole = B.hole[zi]; rec.np = B.npass[zi]; rec.rq = B.rq[zi]; rec.ec = (float)(B.npass[zi] + 1);
            std::memcpy(rec.snr, &B.snr[4 * zi], sizeof(rec.snr));
            rec.seq = B.seq.data() + B.seq_off[zi]; rec.qv = B.qv.data() + B.seq_off[zi];
            rec.len = (int32_t)(B.seq_off[zi + 1] - B.seq_off[zi]);
            writer.write(rec, suffix);
            int32_t q30 = 0;
            for (int32_t j = 0; j < rec.len; ++j) q30 += rec.qv[j] >= 30;
            rep.reads.push_back(ReadStat{rec.len, rec.np, rec.rq, q30});
        }
        if (o.log_level >= 2) {
            // progress line Z1/Z2/Z3 C1/C2/C3 ETA (docs/faq/reports-aux-files.md:176-192)
            const auto now = std::chrono::steady_clock::now();
            const double el = std::chrono::duration<double>(now - t0).count();
            hist.push_back({el, {rep.input, rep.pass}});
            while (hist.size() > 1 && hist.front().first < el - 60.0) hist.pop_front();
            if (std::chrono::duration<double>(now - last_log).count() >= o.refresh_rate) {
                last_log = now;
                const int64_t z2 = rep.input - hist.front().second.first, c2 = rep.pass - hist.front().second.second;
                std::string eta;
                if (zmws_expected > 0 && rep.input > 0) {
                    const double left = el * (double)(zmws_expected * (o.by_strand ? 2 : 1) - rep.input) / (double)rep.input;
                    char buf[64];
                    std::snprintf(buf, sizeof(buf), " %dh %dm", (int)(left / 3600), (int)(std::fmod(left, 3600) / 60));
                    eta = buf;
                }
                std::fprintf(stderr, "%lld/%lld/%.1f %lld/%lld/%.1f%s\n", (long long)rep.input, (long long)z2, (double)z2 / host_threads,
                             (long long)rep.pass, (long long)c2, (double)c2 / host_threads, eta.c_str());
            }
        }
    }
    if (exit_code != 0) todo.close();
    reader_thread.join();
    for (auto& t : workers) t.join();
    if (!fatal.empty()) std::fprintf(stderr, "ccs: %s\n", fatal.c_str());
    if (!reader_error.empty()) { std::fprintf(stderr, "ccs: %s: %s\n", o.in.c_str(), reader_error.c_str()); exit_code = 1; }
    if (!writer.close()) { std::fprintf(stderr, "ccs: error writing %s (disk full?)\n", o.out.c_str()); exit_code = 1; }
    if (mz) { gzputs(mz, "\n  ]\n}\n"); if (gzclose(mz) != Z_OK) { std::fprintf(stderr, "ccs: error writing %s\n", o.metrics.c_str()); exit_code = 1; } }
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!write_report(o.report, rep, o.by_strand)) { std::fprintf(stderr, "ccs: error writing %s\n", o.report.c_str()); exit_code = 1; }
    if (!o.report_json.empty() && !write_report_json(o.report_json, rep, o.by_strand)) { std::fprintf(stderr, "ccs: error writing %s\n", o.report_json.c_str()); exit_code = 1; }
    if (!o.hifi_summary.empty() && !write_hifi_summary(o.hifi_summary, rep, el)) { std::fprintf(stderr, "ccs: error writing %s\n", o.hifi_summary.c_str()); exit_code = 1; }
    if (o.log_level >= 1) {
        int64_t umy = 0, hifi_n = 0, hifi_y = 0;
        for (const ReadStat& x : rep.reads) { umy += x.len; if (x.rq >= 0.99f) { ++hifi_n; hifi_y += x.len; } }
        std::fprintf(stderr, "%s Input    : %lld\n%s Written  : %lld\nUMY           : %s Bases\nHiFi Yield    : %s Bases\nHiFi Reads    : %lld\n"
                             "Elapsed       : %.2f s  (%.1f ZMW/s on %d GPU%s)\n", o.by_strand ? "SS-Reads" : "ZMWs", (long long)rep.input,
                     o.by_strand ? "SS-Reads" : "ZMWs", (long long)rep.pass, commas(umy).c_str(), commas(hifi_y).c_str(), (long long)hifi_n,
                     el, rep.input / std::max(el, 1e-9), gpus, gpus > 1 ? "s" : "");
    }
    for (ccsgpu_ctx* c : ctxs) ccsgpu_destroy(c);
    return exit_code;
}
