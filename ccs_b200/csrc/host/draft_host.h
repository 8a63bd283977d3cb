// Host-side pieces of the Draft Stage that need no device: FilterReads, the k-mer orientation vote and read
// orientation (/root/reference/docs/how-does-ccs-work.md:19-32).  Header-only so that
// tests/host/draft_host_parity.cpp can check them against the oracle on the CPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#include "../cuda/poa_device.h"

namespace ccs {

// Placement of one subread on the draft (ExtractMappedRead): strand, draft span [tstart, tend), aligned part of the
// read [rstart, rend) in its native orientation.
struct ReadMap { int32_t mapped = 0, strand = 0, tstart = 0, tend = 0, rstart = 0, rend = 0, score = 0; };

// FilterReads (docs/how-does-ccs-work.md:19-32): drop reads <50 % or >200 % of the median length,
// cap the full-length passes at top_passes; returns the number of full-length reads kept.
inline int filter_reads(const int32_t* lens, const uint8_t* cx, int n, int top_passes, uint8_t* keep) {
    std::vector<int32_t> s(lens, lens + n);
    std::sort(s.begin(), s.end());
    const int median = (n & 1) ? s[n / 2] : (s[n / 2 - 1] + s[n / 2]) / 2;
    int nfull = 0;
    for (int r = 0; r < n; ++r) {
        keep[r] = 0;
        if (2 * lens[r] < median || lens[r] > 2 * median) continue;
        if ((cx[r] & 3) == 3) {
            if (nfull >= top_passes) continue;
            ++nfull;
        }
        keep[r] = 1;
    }
    return nfull;
}

// K-mer presence set of a reference + hit counting: the seeding half of SdpRangeFinder, used to
// orient a read before it is aligned.
struct KmerSet {
    std::vector<uint32_t> tab;
    uint32_t mask = 0;
    // content sampling: only k-mers whose hash has its three top bits clear take part (1 in 8), on both the
    // reference and the read side, so the table and the number of probes shrink 8x at the same vote statistic
    static uint32_t hash(uint32_t k) { k *= 0x9E3779B1u; return k ^ (k >> 15); }
    static bool sampled(uint32_t k) { return ((k * 0x9E3779B1u) >> 29) == 0u; }
    void build(const uint8_t* s, int n) {
        size_t cap = 256;
        while (cap < (size_t)n) cap <<= 1;          // ~n/8 entries expected: load factor <= 1/8
        tab.assign(cap, 0u);
        mask = (uint32_t)cap - 1;
        if (n < kPoaKmer) return;
        const uint32_t kmask = (1u << (2 * kPoaKmer)) - 1;
        uint32_t k = 0;
        for (int i = 0; i < n; ++i) {
            k = ((k << 2) | s[i]) & kmask;
            if (i >= kPoaKmer - 1 && sampled(k)) insert(k);
        }
    }
    void insert(uint32_t k) {
        uint32_t h = hash(k) & mask;
        while (tab[h] != 0u && tab[h] != k + 1) h = (h + 1) & mask;
        tab[h] = k + 1;
    }
    bool has(uint32_t k) const {
        uint32_t h = hash(k) & mask;
        while (tab[h] != 0u) { if (tab[h] == k + 1) return true; h = (h + 1) & mask; }
        return false;
    }
    // sampled k-mers of seq (forward) and of its reverse complement that occur in the reference
    void count(const uint8_t* codes, int n, int64_t& fwd, int64_t& rev) const {
        fwd = rev = 0;
        if (n < kPoaKmer) return;
        const uint32_t kmask = (1u << (2 * kPoaKmer)) - 1;
        uint32_t kf = 0, kr = 0;
        for (int i = 0; i < n; ++i) {
            const uint32_t b = codes[i] & 3u;
            kf = ((kf << 2) | b) & kmask;
            kr = (kr >> 2) | ((3u - b) << (2 * (kPoaKmer - 1)));
            if (i >= kPoaKmer - 1) {
                if (sampled(kf)) fwd += has(kf);
                if (sampled(kr)) rev += has(kr);
            }
        }
    }
};

inline void orient(const uint8_t* codes, int n, bool rev, uint8_t* out) {
    if (!rev) for (int i = 0; i < n; ++i) out[i] = codes[i] & 3;
    else for (int i = 0; i < n; ++i) out[i] = (uint8_t)(3 - (codes[n - 1 - i] & 3));
}


}  // namespace ccs
