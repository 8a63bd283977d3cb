// `ccs` command line: subreads.bam -> hifi_reads.bam through the GPU stages
//   ccs <in.subreads.bam> <out.bam> [flags]        (/root/reference/docs/index.md:52-64)
// Hot-path subset of the reference's flags: --min-snr --min-passes --min-length
// --max-length --min-rq --top-passes --chunk i/N --report-file --batch-size --log-level, plus --device.
// Per-ZMW failures are counted in <prefix>.ccs_report.txt, never fatal
// (/root/reference/docs/faq/reports-aux-files.md:16-72,143-159); a BAM without the chemistry triple is fatal
// (/root/reference/docs/changelog.md:66).  Compute goes through the C ABI only (include/ccsgpu.h).
#include "../../../include/ccsgpu.h"
#include "bam_io.h"
#include <cuda_runtime.h>
#include <zlib.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <thread>
#include <string>
#include <vector>

using namespace ccs;

namespace {

struct Options {
    std::string in, out, report, metrics, model_path;
    ccs_draft_cfg d;
    ccs_polish_cfg p;
    int chunk_i = 1, chunk_n = 1, batch = 512, device = 0, log_level = 1, pipeline = 2;
};

void usage() {
    std::fprintf(stderr,
                 "ccs (ccs-b200) - Generate circular consensus sequences (ccs) from subreads on a B200 GPU.\n"
                 "Usage: ccs [options] <IN.subreads.bam> <OUT.ccs.bam>\n"
                 "  --min-snr FLOAT      Minimum SNR of subreads to use for generating CCS. [2.5]\n"
                 "  --min-passes INT     Minimum number of full-length subreads required to generate CCS for a ZMW. [3]\n"
                 "  --top-passes INT     Pick at maximum the top N passes for each ZMW. [60]\n"
                 "  --min-length INT     Minimum draft length before polishing. [10]\n"
                 "  --max-length INT     Maximum draft length before polishing. [50000]\n"
                 "  --min-rq FLOAT       Minimum predicted accuracy in [0, 1]. [0.99]\n"
                 "  --chunk i/N          Operate on a single chunk. Format i/N, where i in [1,N].\n"
                 "  --report-file FILE   Where to write the results report. [<out prefix>.ccs_report.txt]\n"
                 "  --metrics-json FILE  Where to write the zmw_metrics JSON (gzip). [<out prefix>.zmw_metrics.json.gz]\n"
                 "  --model-path FILE    Arrow model JSON (else $SMRT_CHEMISTRY_BUNDLE_DIR/arrow/model.json, else built-in synthetic).\n"
                 "  --batch-size INT     ZMWs per GPU batch. [512]\n"
                 "  --device INT         CUDA device. [0]\n"
                 "  --pipeline INT       Batches in flight (GPU stage instances fed by the reader). [2]\n"
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
        else if (a == "--report-file") o.report = val("--report-file");
        else if (a == "--metrics-json") o.metrics = val("--metrics-json");
        else if (a == "--model-path") o.model_path = val("--model-path");
        else if (a == "--batch-size") o.batch = std::max(1, std::atoi(val("--batch-size")));
        else if (a == "--device") o.device = std::atoi(val("--device"));
        else if (a == "--pipeline") o.pipeline = std::max(1, std::min(4, std::atoi(val("--pipeline"))));
        else if (a == "-j" || a == "--num-threads") val("-j");     // accepted for drop-in compatibility; host threads follow the core count
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
    if (o.report.empty()) {
        std::string pre = o.out;
        const size_t dot = pre.rfind(".bam");
        if (dot != std::string::npos) pre = pre.substr(0, dot);
        o.report = pre + ".ccs_report.txt";
    }
    if (o.metrics.empty()) {
        std::string pre = o.out;
        const size_t dot = pre.rfind(".bam");
        if (dot != std::string::npos) pre = pre.substr(0, dot);
        o.metrics = pre + ".zmw_metrics.json.gz";
    }
    return true;
}

struct Report {
    int64_t input = 0, pass = 0, counts[17] = {0};
    std::vector<int32_t> lens, nps;
    std::vector<float> rqs;
};

std::string commas(int64_t v) {
    std::string s = std::to_string(v), o;
    for (size_t k = 0; k < s.size(); ++k) { if (k && (s.size() - k) % 3 == 0) o += ','; o += s[k]; }
    return o;
}

void write_report(const std::string& path, const Report& r) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return;
    const int64_t fail = r.input - r.pass;
    auto pct = [](int64_t a, int64_t b) { return b ? 100.0 * a / b : 0.0; };
    std::fprintf(f, "ZMWs input                    : %lld\n\n", (long long)r.input);
    std::fprintf(f, "ZMWs pass filters             : %lld (%.2f%%)\n", (long long)r.pass, pct(r.pass, r.input));
    std::fprintf(f, "ZMWs fail filters             : %lld (%.2f%%)\n", (long long)fail, pct(fail, r.input));
    std::fprintf(f, "ZMWs shortcut filters         : 0 (0.00%%)\n\n");
    std::fprintf(f, "Exclusive failed counts\n");
    const struct { const char* label; int status; } rows[] = {
        {"Below SNR threshold           ", CCS_ZMW_POOR_SNR}, {"Median length filter          ", CCS_ZMW_NO_SUBREADS},
        {"Lacking full passes           ", CCS_ZMW_TOO_FEW_PASSES}, {"Heteroduplex insertions       ", CCS_ZMW_HETERODUPLEXES},
        {"Coverage drops                ", CCS_ZMW_COVERAGE_DROPS}, {"Insufficient draft cov        ", CCS_ZMW_INSUFFICIENT_SPANS},
        {"Draft too different           ", CCS_ZMW_TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT},
        {"Draft generation error        ", CCS_ZMW_DRAFT_FAILURE}, {"Draft above --max-length      ", CCS_ZMW_TOO_LONG},
        {"Draft below --min-length      ", CCS_ZMW_TOO_SHORT}, {"Reads failed polishing        ", CCS_ZMW_TOO_MANY_UNUSABLE},
        {"Empty coverage windows        ", CCS_ZMW_EMPTY_WINDOW_DURING_POLISHING},
        {"CCS did not converge          ", CCS_ZMW_NON_CONVERGENT}, {"CCS below minimum RQ          ", CCS_ZMW_POOR_QUALITY},
        {"Unknown error                 ", CCS_ZMW_EXCEPTION_THROWN}};
    for (const auto& row : rows)
        std::fprintf(f, "%s: %lld (%.2f%%)\n", row.label, (long long)r.counts[row.status], pct(r.counts[row.status], fail));
    std::fprintf(f, "\n- - - - - - - - - - - - - - - : - - - - -\n\n");
    int64_t yield = 0;
    for (int32_t l : r.lens) yield += l;
    std::vector<int32_t> sl(r.lens), sn(r.nps);
    std::sort(sl.begin(), sl.end());
    std::vector<float> sq(r.rqs);
    std::sort(sq.begin(), sq.end());
    int64_t n50 = 0, acc = 0;
    for (size_t k = sl.size(); k-- > 0;) { acc += sl[k]; if (2 * acc >= yield) { n50 = sl[k]; break; } }
    double np_mean = 0;
    for (int32_t x : r.nps) np_mean += x;
    const double medq = sq.empty() ? 0 : -10.0 * std::log10(std::max(1e-10, 1.0 - (double)sq[sq.size() / 2]));
    std::fprintf(f, "HiFi Reads                    : %s\n", commas((int64_t)r.lens.size()).c_str());
    std::fprintf(f, "HiFi Yield (bp)               : %s\n", commas(yield).c_str());
    std::fprintf(f, "HiFi Read Length (mean, bp)   : %s\n", commas(r.lens.empty() ? 0 : yield / (int64_t)r.lens.size()).c_str());
    std::fprintf(f, "HiFi Read Length (median, bp) : %s\n", commas(sl.empty() ? 0 : sl[sl.size() / 2]).c_str());
    std::fprintf(f, "HiFi Read Length N50 (bp)     : %s\n", commas(n50).c_str());
    std::fprintf(f, "HiFi Read Quality (median)    : %d\n", (int)std::lround(medq));
    std::fprintf(f, "HiFi Number of Passes (mean)  : %d\n", r.nps.empty() ? 0 : (int)std::lround(np_mean / r.nps.size()));
    std::fclose(f);
}

}  // namespace

int main(int argc, char** argv) {
    Options o;
    if (!parse(argc, argv, o)) return 2;
    SubreadBamReader reader;
    std::string err;
    if (!reader.open(o.in, err)) { std::fprintf(stderr, "ccs: %s\n", err.c_str()); return 1; }
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
    // `pipeline` stage instances (one ccsgpu_ctx each); the reader deals batches to them round-robin, they run
    // concurrently, and results are written in input order (reader -> stages -> ordered writer, docs/img/ccs-impl.png)
    std::vector<ccsgpu_ctx*> ctxs;
    if (!std::getenv("CCS_B200_THREADS")) {   // split the host cores between the stage instances
        const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
        setenv("CCS_B200_THREADS", std::to_string(std::max(1u, hc / (unsigned)o.pipeline)).c_str(), 1);
    }
    size_t free_b = 0, total_b = 0;
    cudaSetDevice(o.device);
    cudaMemGetInfo(&free_b, &total_b);                          // fails harmlessly without a device: budget 0
    const size_t budget = (size_t)((double)free_b * 0.85 / o.pipeline);
    for (int k = 0; k < o.pipeline; ++k) {
        int cerr = 0;
        ccsgpu_ctx* c = ccsgpu_create(o.device, model.data(), budget, &cerr);
        if (!c) { std::fprintf(stderr, "ccs: %s\n", ccsgpu_last_error(nullptr)); return 1; }
        if (o.pipeline > 1) ccsgpu_set_lanes(c, 3);
        ctxs.push_back(c);
    }
    std::string cl;
    for (int k = 0; k < argc; ++k) { if (k) cl += ' '; cl += argv[k]; }
    CcsBamWriter writer;
    if (!writer.open(o.out, reader.header_text(), reader.movie(), reader.read_group_id(), cl)) {
        std::fprintf(stderr, "ccs: cannot write %s\n", o.out.c_str());
        return 1;
    }
    struct BatchIO {
        std::vector<ZmwSubreads> zmws;
        std::vector<int32_t> zmw_read_off, hole, status, npass, iters, napp, rstatus;
        std::vector<int64_t> read_off, seq_off, ntest;
        std::vector<uint8_t> codes, cx, seq, qv;
        std::vector<float> snr, rq;
        std::vector<double> rll;
        std::string err;
    };
    auto process = [&](ccsgpu_ctx* ctx, BatchIO& B) -> int {
        const int nz = (int)B.zmws.size();
        B.zmw_read_off.assign(1, 0); B.read_off.assign(1, 0);
        size_t maxlen = 1;
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
    // <prefix>.zmw_metrics.json.gz: one entry per input ZMW (docs/faq/reports-aux-files.md:100-171)
    static const char* kStatusName[17] = {"POOR_SNR", "NO_SUBREADS", "TOO_FEW_PASSES", "LOW_PASS_SHORTCUT", "HETERODUPLEXES",
        "COVERAGE_DROPS", "INSUFFICIENT_SPANS", "TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT", "DRAFT_FAILURE", "TOO_LONG", "TOO_SHORT",
        "TOO_MANY_UNUSABLE", "EMPTY_WINDOW_DURING_POLISHING", "NON_CONVERGENT", "POOR_QUALITY", "EXCEPTION_THROWN", "SUCCESS"};
    gzFile mz = gzopen(o.metrics.c_str(), "wb");
    if (mz) gzputs(mz, "{\n  \"zmws\": [");
    bool first_metric = true;
    Report rep;
    const auto t0 = std::chrono::steady_clock::now();
    int64_t z_index = 0;
    bool more = true;
    while (more) {
        std::vector<BatchIO> wave(ctxs.size());
        size_t n_wave = 0;
        for (; n_wave < ctxs.size() && more; ++n_wave) {
            BatchIO& B = wave[n_wave];
            ZmwSubreads z;
            while ((int)B.zmws.size() < o.batch && (more = reader.next_zmw(z))) {
                if (z_index >= z_begin && z_index < z_end) B.zmws.push_back(std::move(z));
                ++z_index;
                if (z_index >= z_end) { more = false; break; }
            }
            if (B.zmws.empty()) break;
        }
        if (n_wave == 0) break;
        std::vector<std::future<int>> fut;
        for (size_t k = 0; k < n_wave; ++k)
            fut.push_back(std::async(std::launch::async, [&, k]() { return process(ctxs[k], wave[k]); }));
        for (size_t k = 0; k < n_wave; ++k) {
            if (fut[k].get() != CCS_OK) { std::fprintf(stderr, "ccs: %s\n", wave[k].err.c_str()); return 1; }
            BatchIO& B = wave[k];
            for (int zi = 0; zi < (int)B.zmws.size(); ++zi) {
                ++rep.input;
                ++rep.counts[B.status[zi]];
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
                                 "\"status\": \"%s\", \"zmw\": \"%s/%d\"}",
                             first_metric ? "" : ",", B.npass[zi] + (B.npass[zi] > 0 ? 1 : 0), (long long)insert, B.npass[zi],
                             (long long)poly, has_rq ? (double)B.rq[zi] : -1.0, kStatusName[B.status[zi]], reader.movie().c_str(),
                             zz.hole);
                    first_metric = false;
                }
                if (B.status[zi] != CCS_ZMW_SUCCESS) continue;
                ++rep.pass;
                CcsRecord rec;
                rec.hole = B.hole[zi]; rec.np = B.npass[zi]; rec.rq = B.rq[zi]; rec.ec = (float)(B.npass[zi] + 1);
                std::memcpy(rec.snr, &B.snr[4 * zi], sizeof(rec.snr));
                rec.seq = B.seq.data() + B.seq_off[zi]; rec.qv = B.qv.data() + B.seq_off[zi];
                rec.len = (int32_t)(B.seq_off[zi + 1] - B.seq_off[zi]);
                writer.write(rec);
                rep.lens.push_back(rec.len); rep.nps.push_back(rec.np); rep.rqs.push_back(rec.rq);
            }
        }
        if (o.log_level >= 2) {
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            std::fprintf(stderr, "| ZMWs %lld  HiFi %lld  %.1f ZMW/s\n", (long long)rep.input, (long long)rep.pass, rep.input / el);
        }
    }
    writer.close();
    if (mz) { gzputs(mz, "\n  ]\n}\n"); gzclose(mz); }
    write_report(o.report, rep);
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (o.log_level >= 1)
        std::fprintf(stderr, "ZMWs input: %lld  ZMWs pass filters: %lld  elapsed: %.2f s  (%.1f ZMW/s)\n", (long long)rep.input,
                     (long long)rep.pass, el, rep.input / std::max(el, 1e-9));
    for (ccsgpu_ctx* c : ctxs) ccsgpu_destroy(c);
    return 0;
}
