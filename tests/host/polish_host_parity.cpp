// CPU parity test of the device-free pieces of Polish() (ccs_b200/csrc/host/polish_host.h: candidate order +
// BestMutations with separation, Template::ApplyMutations, the de-duplicated candidate count) against the oracle
// (oracle/arrow_oracle.cpp best_mutations / apply_mutations / mutation_is_canonical).
// Built and run by tests/test_cpu_host.py::test_polish_host_helpers_match_oracle.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../oracle/arrow_oracle.h"

// polish_host.h only needs HostMutation from the engine header; give it that without pulling in CUDA
#define CCS_POLISH_HOST_STANDALONE
namespace ccs { struct HostMutation { int32_t type, pos, base; double score; }; }
#include "../../ccs_b200/csrc/host/polish_host.h"

int main(int argc, char** argv) {
    const int trials = argc > 1 ? std::atoi(argv[1]) : 300;
    std::mt19937 rng(777u);
    long n_sel = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int J = 2 + (int)(rng() % 400);
        std::vector<uint8_t> tpl(J);
        for (auto& b : tpl) b = (uint8_t)(rng() % (trial % 3 == 0 ? 2 : 4));      // some homopolymer-rich templates
        // canonical count == number of canonical mutations the oracle enumerates
        {
            int64_t want = 0;
            for (int p = 0; p <= J; ++p)
                for (int type = 0; type < 3; ++type)
                    for (int b = 0; b < (type == oracle::MUT_DEL ? 1 : 4); ++b) {
                        if (p == J && type != oracle::MUT_INS) continue;
                        if (oracle::mutation_is_canonical(tpl, oracle::Mutation{type, p, b})) ++want;
                    }
            // the engine counts per position range; insertions before p belong to position p (p <= J-1)
            if (ccs::count_canonical_mutations(tpl, 0, J) != want) { std::fprintf(stderr, "MISMATCH: canonical count (trial %d): %lld vs %lld\n", trial, (long long)ccs::count_canonical_mutations(tpl, 0, J), (long long)want); return 1; }
        }
        // random scored candidates (ties on purpose), at most one per (pos, type, base)
        const int nc = 1 + (int)(rng() % 60);
        std::vector<ccs::HostMutation> sc;
        std::vector<std::pair<double, oracle::Mutation>> so;
        for (int k = 0; k < nc; ++k) {
            const int type = (int)(rng() % 3);
            const int pos = (type == oracle::MUT_INS) ? 1 + (int)(rng() % std::max(1, J - 1)) : (int)(rng() % J);
            const int base = (type == oracle::MUT_DEL) ? 0 : (int)(rng() & 3);
            bool dup = false;
            for (auto& m : sc) if (m.type == type && m.pos == pos && m.base == base) dup = true;
            if (dup) continue;
            const double score = 0.25 * (double)(1 + rng() % 12);
            sc.push_back({type, pos, base, score});
            so.push_back({score, oracle::Mutation{type, pos, base}});
        }
        const int sep = 1 + (int)(rng() % 12);
        std::vector<ccs::HostMutation> best = ccs::select_best_mutations(sc, J, sep);
        std::vector<oracle::Mutation> want = oracle::best_mutations(so, sep);
        if (best.size() != want.size()) { std::fprintf(stderr, "MISMATCH: selection size (trial %d)\n", trial); return 1; }
        for (size_t k = 0; k < best.size(); ++k)
            if (best[k].type != want[k].type || best[k].pos != want[k].pos || best[k].base != want[k].base) { std::fprintf(stderr, "MISMATCH: selected mutation %zu (trial %d)\n", k, trial); return 1; }
        if (sc.front().type != so.front().second.type || sc.front().pos != so.front().second.pos || sc.front().base != so.front().second.base) { std::fprintf(stderr, "MISMATCH: single best (trial %d)\n", trial); return 1; }
        // ApplyMutations (the selection has at most one edit per site when sep >= 1)
        const std::vector<uint8_t> a = ccs::apply_to_template(tpl, best);
        const std::vector<uint8_t> b = oracle::apply_mutations(tpl, want);
        if (a != b) { std::fprintf(stderr, "MISMATCH: applied template (trial %d)\n", trial); return 1; }
        ++n_sel;
    }
    std::printf("ok: %ld selections / applications agree with the oracle\n", n_sel);
    return 0;
}
