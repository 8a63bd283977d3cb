// Synthetic Sequel-II-shape ZMW generator: samples subreads from the Arrow HMM itself so
// that the model is well specified (read structure -- alternating
// strands, partial first/last pass -- per /root/reference/docs/how-does-ccs-work.md:9-32).
// Counter-based: ZMW k of a config is a pure function of (seed, k), so any shard can be
// generated independently (the `--chunk i/N` analogue, docs/faq/parallelize.md:15-20).
#pragma once
#include "arrow_tables.h"
#include <vector>
#include <cstdint>

namespace ccs {

struct SimConfig {
    int insert_mean = 10000, insert_sd = 0;
    int passes_min = 10, passes_max = 10;   // full-length passes, uniform in [min,max]
    int partials = 1;                       // add leading + trailing partial pass (40-60 %)
    double snr_mean[4] = {9.0, 16.0, 8.5, 13.0};
    double snr_sd = 0.0;
    double frac_low_snr = 0.0;              // ZMWs forced below --min-snr
    double frac_few_passes = 0.0;           // ZMWs with < 3 full passes
    uint64_t seed = 0xCC5;
};

// BASELINE.json configs 1,2,3,5 (4 = config 3 range-sharded).
inline SimConfig sim_config(int id) {
    SimConfig c;
    c.seed = 0xCC5ull ^ (uint64_t)id;
    switch (id) {
        case 1: break;
        case 2: c.insert_sd = 500; c.snr_sd = 1.0; c.frac_low_snr = 0.05; c.frac_few_passes = 0.05; break;
        case 3: case 4: c.insert_mean = 15000; c.insert_sd = 1500; c.passes_min = 5; c.passes_max = 20; c.snr_sd = 1.0; break;
        case 5: c.insert_mean = 25000; c.insert_sd = 1000; c.passes_min = 4; c.passes_max = 4; c.snr_sd = 1.0; break;
        default: break;
    }
    return c;
}

struct SimRead {
    std::vector<uint8_t> codes;   // 4*(pw-1)+base per read base, native orientation
    uint8_t cx = 0;               // local context flags: 1 adapter before, 2 adapter after
    uint8_t strand = 0;           // truth: 0 = same strand as `tpl`, 1 = reverse complement
    int32_t tstart = 0, tend = 0; // truth: span on `tpl` (forward coordinates)
};

struct SimZmw {
    int32_t hole = 0;
    float snr[4] = {0, 0, 0, 0};
    std::vector<uint8_t> tpl;     // truth insert, bases 0..3
    std::vector<SimRead> reads;
};

struct SimRng {
    uint64_t s;
    explicit SimRng(uint64_t seed) : s(seed) {}
    uint64_t next() { return splitmix64(s); }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    int below(int n) { return (int)(uniform() * n); }
};

inline int sample_pmf(SimRng& r, const double* p, int n) {
    double u = r.uniform(), acc = 0.0, tot = 0.0;
    for (int i = 0; i < n; ++i) tot += p[i];
    u *= tot;
    for (int i = 0; i < n; ++i) { acc += p[i]; if (u < acc) return i; }
    return n - 1;
}

// One pass of the HMM over tpl[0..J): pinned first and last base (both matches).
inline void sample_read(SimRng& r, const ArrowModelParams& m, const TransProb tp[kNumCtx],
                        const uint8_t* tpl, int J, std::vector<uint8_t>& codes) {
    codes.clear();
    if (J < 2) return;
    codes.push_back((uint8_t)sample_pmf(r, m.emission[MOVE_MATCH][4 * tpl[0] + tpl[0]], kNumCodes));
    int j = 1;  // template bases consumed
    while (true) {
        const int ctx = 4 * tpl[j - 1] + tpl[j];
        const double p[4] = {tp[ctx].match, tp[ctx].branch, tp[ctx].stick, tp[ctx].deletion};
        const int mv = sample_pmf(r, p, 4);
        if (mv == 1) { codes.push_back((uint8_t)sample_pmf(r, m.emission[MOVE_BRANCH][ctx], kNumCodes)); continue; }
        if (mv == 2) { codes.push_back((uint8_t)sample_pmf(r, m.emission[MOVE_STICK][ctx], kNumCodes)); continue; }
        if (j == J - 1) {  // last template base: forced match
            codes.push_back((uint8_t)sample_pmf(r, m.emission[MOVE_MATCH][ctx], kNumCodes));
            return;
        }
        if (mv == 0) codes.push_back((uint8_t)sample_pmf(r, m.emission[MOVE_MATCH][ctx], kNumCodes));
        ++j;
    }
}

inline void revcomp_bases(const uint8_t* in, int n, std::vector<uint8_t>& out) {
    out.resize(n);
    for (int i = 0; i < n; ++i) out[i] = (uint8_t)(3 - in[n - 1 - i]);
}

inline void simulate_zmw(const ArrowModelParams& m, const SimConfig& c, int64_t index, SimZmw& z) {
    uint64_t mix = c.seed * 0x9E3779B97F4A7C15ull + (uint64_t)index * 0xD1B54A32D192ED03ull + 0x1234567ull;
    SimRng r(splitmix64(mix));
    z.hole = (int32_t)index;
    int J = c.insert_mean + (c.insert_sd > 0 ? (int)std::lround(r.normal() * c.insert_sd) : 0);
    if (J < 64) J = 64;
    const bool low_snr = r.uniform() < c.frac_low_snr;
    const bool few = r.uniform() < c.frac_few_passes;
    for (int ch = 0; ch < 4; ++ch) {
        double s = c.snr_mean[ch] + (c.snr_sd > 0 ? r.normal() * c.snr_sd : 0.0);
        s = std::min(std::max(s, m.snr_lo[ch] + 0.5), m.snr_hi[ch]);
        if (low_snr) s = 1.5 + 0.5 * r.uniform();
        z.snr[ch] = (float)s;
    }
    z.tpl.resize(J);
    for (int j = 0; j < J; ++j) z.tpl[j] = (uint8_t)(r.next() >> 62);
    std::vector<uint8_t> rc;
    revcomp_bases(z.tpl.data(), J, rc);
    TransProb tp[kNumCtx];
    transition_probs(m, z.snr, tp);
    int passes = c.passes_min + (c.passes_max > c.passes_min ? r.below(c.passes_max - c.passes_min + 1) : 0);
    if (few) passes = 1 + r.below(2);
    int strand = (int)(r.next() & 1);
    z.reads.clear();
    auto emit = [&](int tstart, int tend, uint8_t cx) {
        // read of strand `strand` covering forward span [tstart,tend)
        SimRead rd;
        rd.cx = cx; rd.strand = (uint8_t)strand; rd.tstart = tstart; rd.tend = tend;
        const uint8_t* src = strand ? rc.data() + (J - tend) : z.tpl.data() + tstart;
        sample_read(r, m, tp, src, tend - tstart, rd.codes);
        z.reads.push_back(std::move(rd));
        strand ^= 1;
    };
    if (c.partials) {  // polymerase starts mid-insert: suffix of the strand being read
        const int len = (int)(J * (0.40 + 0.20 * r.uniform()));
        if (strand == 0) emit(J - len, J, 2); else emit(0, len, 2);
    }
    for (int p = 0; p < passes; ++p) emit(0, J, 3);
    if (c.partials) {  // polymerase stops mid-insert: prefix of the strand being read
        const int len = (int)(J * (0.40 + 0.20 * r.uniform()));
        if (strand == 0) emit(0, len, 1); else emit(J - len, J, 1);
    }
}

// Corrupt the truth into a draft-like template (for polish-only tests/benches that bypass
// the draft stage) and remap each read's span onto it.  rate = per-base error probability,
// split evenly between substitution, insertion and deletion.
inline void corrupt_template(const std::vector<uint8_t>& truth, double rate, uint64_t seed,
                             std::vector<uint8_t>& out, std::vector<int32_t>& map /* truth pos -> out pos, size J+1 */) {
    SimRng r(seed * 0x2545F4914F6CDD1Dull + 99);
    out.clear();
    map.assign(truth.size() + 1, 0);
    const int J = (int)truth.size();
    for (int j = 0; j < J; ++j) {
        map[j] = (int32_t)out.size();
        const bool edge = j < 2 || j >= J - 2;
        const double u = edge ? 1.0 : r.uniform();
        if (u < rate / 3) { out.push_back((uint8_t)((truth[j] + 1 + r.below(3)) & 3)); }
        else if (u < 2 * rate / 3) { out.push_back((uint8_t)r.below(4)); out.push_back(truth[j]); }
        else if (u < rate) { /* deletion */ }
        else out.push_back(truth[j]);
    }
    map[J] = (int32_t)out.size();
}

}  // namespace ccs
