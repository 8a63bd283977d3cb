// Arrow model <-> JSON: the chemistry bundle mechanism of the reference
// ("$SMRT_CHEMISTRY_BUNDLE_DIR/arrow/*.json", /root/reference/docs/faq/chemistry.md:28-56).  Schema (ours):
//   { "ChemistryName": str, "ModelForm": "PwSnr", "CounterWeight": x, "SnrRanges": [[lo,hi] x4],
//     "TransitionParameters": [16][3][4], "EmissionParameters": [3][16][12] }
// A small hand-written reader: numbers are collected in document order per key, nesting is only checked by count.
#include "../common/arrow_model.h"
#include "../../../include/ccsgpu.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace ccs;

namespace {

bool read_file(const char* path, std::string& out) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    char buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    std::fclose(f);
    return true;
}

// position just after `"key"` and the following ':' (or npos)
size_t find_key(const std::string& s, const char* key) {
    const std::string k = std::string("\"") + key + "\"";
    size_t p = s.find(k);
    if (p == std::string::npos) return p;
    p = s.find(':', p + k.size());
    return p == std::string::npos ? p : p + 1;
}

bool numbers_after(const std::string& s, size_t p, size_t want, std::vector<double>& out) {
    out.clear();
    int depth = 0;
    bool opened = false;
    while (p < s.size()) {
        const char c = s[p];
        if (c == '[') { ++depth; opened = true; ++p; }
        else if (c == ']') { --depth; ++p; if (opened && depth == 0) break; }
        else if ((c >= '0' && c <= '9') || c == '-' || c == '+' || c == '.') {
            char* e = nullptr;
            out.push_back(std::strtod(s.c_str() + p, &e));
            p = (size_t)(e - s.c_str());
            if (!opened) break;   // scalar value
        } else ++p;
    }
    return out.size() == want;
}

}  // namespace

extern "C" {

int ccs_model_sizeof(void) { return (int)sizeof(ccs::ArrowModelParams); }

void ccs_model_synthetic(void* model_out) { ccs::synthetic_model(*(ccs::ArrowModelParams*)model_out); }

int ccs_model_save_json(const void* model, const char* path) {
    const ArrowModelParams& m = *(const ArrowModelParams*)model;
    FILE* f = std::fopen(path, "w");
    if (!f) return CCS_ERR_IO;
    std::fprintf(f, "{\n  \"ChemistryName\": \"%s\",\n  \"ModelForm\": \"PwSnr\",\n  \"CounterWeight\": %.17g,\n", m.chemistry,
                 m.counter_weight);
    std::fprintf(f, "  \"SnrRanges\": [");
    for (int c = 0; c < 4; ++c) std::fprintf(f, "%s[%.17g, %.17g]", c ? ", " : "", m.snr_lo[c], m.snr_hi[c]);
    std::fprintf(f, "],\n  \"TransitionParameters\": [\n");
    for (int ctx = 0; ctx < kNumCtx; ++ctx) {
        std::fprintf(f, "    [");
        for (int t = 0; t < 3; ++t) {
            std::fprintf(f, "%s[", t ? ", " : "");
            for (int d = 0; d < 4; ++d) std::fprintf(f, "%s%.17g", d ? ", " : "", m.trans[ctx][t][d]);
            std::fprintf(f, "]");
        }
        std::fprintf(f, "]%s\n", ctx + 1 < kNumCtx ? "," : "");
    }
    std::fprintf(f, "  ],\n  \"EmissionParameters\": [\n");
    for (int mv = 0; mv < 3; ++mv) {
        std::fprintf(f, "    [\n");
        for (int ctx = 0; ctx < kNumCtx; ++ctx) {
            std::fprintf(f, "      [");
            for (int c = 0; c < kNumCodes; ++c) std::fprintf(f, "%s%.17g", c ? ", " : "", m.emission[mv][ctx][c]);
            std::fprintf(f, "]%s\n", ctx + 1 < kNumCtx ? "," : "");
        }
        std::fprintf(f, "    ]%s\n", mv < 2 ? "," : "");
    }
    std::fprintf(f, "  ]\n}\n");
    std::fclose(f);
    return CCS_OK;
}

int ccs_model_load_json(const char* path, void* model_out) {
    std::string s;
    if (!read_file(path, s)) return CCS_ERR_IO;
    ArrowModelParams m;
    std::memset(&m, 0, sizeof(m));
    size_t p = find_key(s, "ChemistryName");
    if (p != std::string::npos) {
        const size_t a = s.find('"', p), b = a == std::string::npos ? a : s.find('"', a + 1);
        if (b != std::string::npos) std::strncpy(m.chemistry, s.substr(a + 1, b - a - 1).c_str(), sizeof(m.chemistry) - 1);
    }
    p = find_key(s, "ModelForm");
    if (p == std::string::npos || s.find("PwSnr", p) == std::string::npos) return CCS_ERR_CHEMISTRY;   // only form implemented
    std::vector<double> v;
    if ((p = find_key(s, "CounterWeight")) == std::string::npos || !numbers_after(s, p, 1, v)) return CCS_ERR_CHEMISTRY;
    m.counter_weight = v[0];
    if ((p = find_key(s, "SnrRanges")) == std::string::npos || !numbers_after(s, p, 8, v)) return CCS_ERR_CHEMISTRY;
    for (int c = 0; c < 4; ++c) { m.snr_lo[c] = v[2 * c]; m.snr_hi[c] = v[2 * c + 1]; }
    if ((p = find_key(s, "TransitionParameters")) == std::string::npos || !numbers_after(s, p, 16 * 3 * 4, v)) return CCS_ERR_CHEMISTRY;
    std::memcpy(m.trans, v.data(), sizeof(m.trans));
    if ((p = find_key(s, "EmissionParameters")) == std::string::npos || !numbers_after(s, p, 3 * 16 * 12, v)) return CCS_ERR_CHEMISTRY;
    std::memcpy(m.emission, v.data(), sizeof(m.emission));
    if (!(m.counter_weight > 0)) return CCS_ERR_CHEMISTRY;
    std::memcpy(model_out, &m, sizeof(m));
    return CCS_OK;
}

}  // extern "C"
