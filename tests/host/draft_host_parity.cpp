// CPU parity test of the Draft Stage's device-free host helpers (ccs_b200/csrc/host/draft_host.h: FilterReads, the
// hashed k-mer orientation vote, read orientation) against the oracle's independent restatements
// (oracle/pipeline_oracle.cpp filter_reads, oracle/poa_oracle.cpp kmer_vote_reverse -- sorted k-mer lists and binary
// search instead of a hash set).  Built and run by tests/test_cpu_host.py::test_draft_host_helpers_match_oracle.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../ccs_b200/csrc/host/draft_host.h"
#include "../../oracle/pipeline_oracle.h"

int main(int argc, char** argv) {
    const int trials = argc > 1 ? std::atoi(argv[1]) : 200;
    std::mt19937 rng(4242u);
    long votes = 0, filters = 0;
    for (int trial = 0; trial < trials; ++trial) {
        // ---- FilterReads on ragged read sets (partials, missed adapters, more than top_passes full passes)
        {
            const int n = 1 + (int)(rng() % 70);
            std::vector<int32_t> lens(n);
            std::vector<int> lens_o(n);
            std::vector<uint8_t> cx(n);
            const int base = 200 + (int)(rng() % 3000);
            for (int r = 0; r < n; ++r) {
                const unsigned u = rng() % 100;
                int len = base + (int)(rng() % 100) - 50;
                if (u < 10) len = len / 3;            // short partial
                else if (u < 15) len = len * 3;       // missed adapter
                lens[r] = lens_o[r] = std::max(len, 1);
                cx[r] = (uint8_t)((rng() % 100 < 85) ? 3 : (rng() & 3));
            }
            const int top = (trial & 1) ? 60 : 5;
            std::vector<uint8_t> keep(n);
            std::vector<char> keep_o;
            const int nf = ccs::filter_reads(lens.data(), cx.data(), n, top, keep.data());
            const int nf_o = oracle::filter_reads(lens_o, cx.data(), top, keep_o);
            if (nf != nf_o) { std::fprintf(stderr, "MISMATCH: full-length count (trial %d)\n", trial); return 1; }
            for (int r = 0; r < n; ++r) if ((keep[r] != 0) != (keep_o[r] != 0)) { std::fprintf(stderr, "MISMATCH: keep[%d] (trial %d)\n", r, trial); return 1; }
            ++filters;
        }
        // ---- orientation vote: reference + noisy forward / reverse-complement / unrelated reads, short and long
        {
            const int L = 30 + (int)(rng() % 6000);
            std::vector<uint8_t> ref(L);
            for (auto& b : ref) b = (uint8_t)(rng() & 3);
            ccs::KmerSet ks;
            ks.build(ref.data(), L);
            for (int rep = 0; rep < 6; ++rep) {
                std::vector<uint8_t> read;
                if (rep == 5) { read.resize(5 + rng() % 3000); for (auto& b : read) b = (uint8_t)(rng() & 3); }
                else {
                    for (int i = 0; i < L; ++i) {
                        const unsigned u = rng() % 1000;
                        if (u < 40) continue;
                        if (u < 90) read.push_back((uint8_t)(rng() & 3));
                        read.push_back(u < 120 ? (uint8_t)((ref[i] + 1 + rng() % 3) & 3) : ref[i]);
                    }
                    if (rep & 1) { std::vector<uint8_t> rc(read.rbegin(), read.rend()); for (auto& b : rc) b = (uint8_t)(3 - b); read.swap(rc); }
                }
                // the product sees emission codes (base in the low two bits) and votes on the first kPoaVoteBases bases
                std::vector<uint8_t> codes(read.size());
                for (size_t i = 0; i < read.size(); ++i) codes[i] = (uint8_t)(read[i] | ((rng() % 3) << 2));
                int64_t f = 0, c = 0;
                ks.count(codes.data(), std::min((int)codes.size(), ccs::kPoaVoteBases), f, c);
                const bool rev = c > f;
                const bool rev_o = oracle::kmer_vote_reverse(ref.data(), L, read.data(), (int)read.size());
                if (rev != rev_o) { std::fprintf(stderr, "MISMATCH: orientation vote (trial %d rep %d)\n", trial, rep); return 1; }
                // orient() = bases of the codes, reverse-complemented on demand
                std::vector<uint8_t> o(read.size());
                ccs::orient(codes.data(), (int)codes.size(), rev, o.data());
                for (size_t i = 0; i < read.size(); ++i) {
                    const uint8_t want = rev ? (uint8_t)(3 - read[read.size() - 1 - i]) : read[i];
                    if (o[i] != want) { std::fprintf(stderr, "MISMATCH: orient (trial %d)\n", trial); return 1; }
                }
                ++votes;
            }
        }
    }
    std::printf("ok: %ld FilterReads cases and %ld orientation votes agree with the oracle\n", filters, votes);
    return 0;
}
