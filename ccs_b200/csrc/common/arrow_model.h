// Arrow chemistry model: parameter container (the "chemistry bundle" data) and the
// deterministic synthetic parameter set this repo ships.
//
// The reference's trained parameter tables are not public (they live in the closed
// binary / in a PacBio chemistry bundle, /root/reference/docs/faq/chemistry.md:28-56), so
// the FORM follows the documented model -- "emission and transition parameters are
// estimated by a dinucleotide template context ... transition parameters ... only
// [depend on] the pulse width of a base call, the dinucleotide context of the template,
// and the SNR of the ZMW" (/root/reference/docs/how-does-ccs-work.md:88-94) -- and the NUMBERS
// are synthetic (DESIGN.md "Model").
//
// This header is input data shared by the product, the simulator and the oracle; the
// arithmetic that turns it into per-ZMW tables is restated independently on each side.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

namespace ccs {

enum Move { MOVE_MATCH = 0, MOVE_BRANCH = 1, MOVE_STICK = 2 };
enum Trans { TR_BRANCH = 0, TR_STICK = 1, TR_DELETION = 2 };

constexpr int kNumCtx = 16;    // dinucleotide contexts: 4*prev + cur
constexpr int kNumCodes = 12;  // emission codes: 4*(min(pw,3)-1) + base

struct ArrowModelParams {
    char chemistry[64];
    double snr_lo[4], snr_hi[4];      // per channel (A,C,G,T) clip range
    double trans[kNumCtx][3][4];      // [ctx][Branch,Stick,Deletion][cubic coefficient d]
    double emission[3][kNumCtx][kNumCodes];  // [Match,Branch,Stick][ctx][code] pmfs
    double counter_weight;            // constant pre-multiplier of every emission
};

inline uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

inline double model_jitter(uint32_t a, uint32_t b, uint32_t c) {
    uint64_t s = 0xCC5B200ull + a * 1000003ull + b * 10007ull + c * 101ull;
    uint64_t r = splitmix64(s);
    return (double)(r >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;  // [-1,1)
}

// Deterministic synthetic chemistry "S/P0-C0/synthetic": ~11-13 % subread error at
// SNR ~ (9,16,8.5,13), indel dominated, homopolymer contexts worse (DESIGN.md "Model").
inline void synthetic_model(ArrowModelParams& m) {
    std::memset(&m, 0, sizeof(m));
    std::strncpy(m.chemistry, "S/P0-C0/synthetic", sizeof(m.chemistry) - 1);
    for (int c = 0; c < 4; ++c) { m.snr_lo[c] = 4.0; m.snr_hi[c] = 20.0; }
    m.counter_weight = 2.0;
    for (int ctx = 0; ctx < kNumCtx; ++ctx) {
        const bool hp = (ctx >> 2) == (ctx & 3);
        double* br = m.trans[ctx][TR_BRANCH];
        double* st = m.trans[ctx][TR_STICK];
        double* de = m.trans[ctx][TR_DELETION];
        br[0] = -2.55 + (hp ? 0.55 : 0.0) + 0.10 * model_jitter(ctx, 1, 0);
        br[1] = -0.055; br[2] = 0.0006; br[3] = 0.0;
        st[0] = -3.30 + 0.10 * model_jitter(ctx, 2, 0);
        st[1] = -0.050; st[2] = 0.0005; st[3] = 0.0;
        de[0] = -2.45 + (hp ? 0.45 : 0.0) + 0.10 * model_jitter(ctx, 3, 0);
        de[1] = -0.050; de[2] = 0.0005; de[3] = 0.0;
        const int cur = ctx & 3;
        // pulse-width pmfs per move, lightly context dependent
        double pwm[3] = {0.25 + 0.05 * model_jitter(ctx, 4, 0), 0.35 + 0.05 * model_jitter(ctx, 4, 1),
                         0.40 + 0.05 * model_jitter(ctx, 4, 2)};
        double pwb[3] = {0.55 + 0.05 * model_jitter(ctx, 5, 0), 0.30 + 0.05 * model_jitter(ctx, 5, 1),
                         0.15 + 0.03 * model_jitter(ctx, 5, 2)};
        double pws[3] = {0.50 + 0.05 * model_jitter(ctx, 6, 0), 0.30 + 0.05 * model_jitter(ctx, 6, 1),
                         0.20 + 0.03 * model_jitter(ctx, 6, 2)};
        auto norm3 = [](double* p) { double s = p[0] + p[1] + p[2]; p[0] /= s; p[1] /= s; p[2] /= s; };
        norm3(pwm); norm3(pwb); norm3(pws);
        const double miscall = 0.015 + 0.003 * model_jitter(ctx, 7, 0);
        for (int pw = 0; pw < 3; ++pw)
            for (int b = 0; b < 4; ++b) {
                const int code = 4 * pw + b;
                m.emission[MOVE_MATCH][ctx][code] = (b == cur ? (1.0 - miscall) : miscall / 3.0) * pwm[pw];
                m.emission[MOVE_BRANCH][ctx][code] = (b == cur ? 1.0 : 0.0) * pwb[pw];
                m.emission[MOVE_STICK][ctx][code] = (b == cur ? 0.0 : 1.0 / 3.0) * pws[pw];
            }
    }
}

}  // namespace ccs
