// Device-free pieces of Polish() (/root/reference/docs/how-does-ccs-work.md:96-101): the
// candidate order, BestMutations (greedy by score with a minimum separation), Template::ApplyMutations, the
// homopolymer-deduplicated candidate count and the template hash of the cycle guard.  Header-only so that
// tests/host/polish_host_parity.cpp can check them against the oracle on the CPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#ifndef CCS_POLISH_HOST_STANDALONE   // the CPU parity test supplies HostMutation itself (no CUDA headers there)
#include "polish_engine.h"
#endif

namespace ccs {

inline uint64_t tpl_hash(const std::vector<uint8_t>& t) {
    uint64_t h = 1469598103934665603ull;
    for (uint8_t b : t) { h ^= b; h *= 1099511628211ull; }
    return h ^ ((uint64_t)t.size() * 0x9E3779B97F4A7C15ull);
}

inline int type_rank(int t) { return t == 2 ? 0 : (t == 1 ? 1 : 2); }   // DEL < INS < SUB

// Sorts `sc` by (score desc, pos, DEL < INS < SUB, base) -- sc.front() is the single best afterwards -- and returns
// the greedy selection with chosen sites at least `separation` apart, ordered by position.
inline std::vector<HostMutation> select_best_mutations(std::vector<HostMutation>& sc, int J, int separation) {
    std::sort(sc.begin(), sc.end(), [](const HostMutation& a, const HostMutation& b) {
        if (a.score != b.score) return a.score > b.score;
        if (a.pos != b.pos) return a.pos < b.pos;
        if (a.type != b.type) return type_rank(a.type) < type_rank(b.type);
        return a.base < b.base;
    });
    std::vector<uint8_t> blocked((size_t)J + 2, 0);
    std::vector<HostMutation> best;
    for (const auto& m : sc) {
        if (blocked[m.pos]) continue;
        best.push_back(m);
        const int lo = std::max(0, m.pos - separation + 1), hi = std::min(J + 1, m.pos + separation - 1);
        for (int x = lo; x <= hi; ++x) blocked[x] = 1;
    }
    std::sort(best.begin(), best.end(), [](const HostMutation& a, const HostMutation& b) { return a.pos < b.pos; });
    return best;
}

// Template::ApplyMutations: muts sorted by position; at most one of SUB/DEL per position
inline std::vector<uint8_t> apply_to_template(const std::vector<uint8_t>& tpl, const std::vector<HostMutation>& muts) {
    std::vector<uint8_t> out;
    out.reserve(tpl.size() + muts.size());
    size_t k = 0;
    const int J = (int)tpl.size();
    for (int j = 0; j <= J; ++j) {
        bool skip = false;
        while (k < muts.size() && muts[k].pos == j) {
            const HostMutation& m = muts[k++];
            if (m.type == 1) out.push_back((uint8_t)m.base);
            else if (m.type == 0) { out.push_back((uint8_t)m.base); skip = true; }
            else skip = true;
        }
        if (j < J && !skip) out.push_back(tpl[j]);
    }
    return out;
}

// number of canonical (de-duplicated) single-base mutations of template positions [b, e)
inline int64_t count_canonical_mutations(const std::vector<uint8_t>& t, int b, int e) {
    const int J = (int)t.size();
    int64_t n = 0;
    for (int p = b; p < e; ++p) {
        n += 3;                                            // substitutions
        if (!(p > 0 && t[p] == t[p - 1])) ++n;             // deletion
        if (p >= 1 && p <= J - 1) n += 3;                  // insertions except the one equal to t[p-1]
    }
    return n;
}

}  // namespace ccs
