// C ABI over the synthetic ZMW generator and the chemistry model container
// (declared in include/ccsgpu.h, "Synthetic data" section).
#include "../common/sim.h"
#include "../../../include/ccsgpu.h"
#include <cstring>

using namespace ccs;

static_assert(sizeof(ccs_sim_config) == sizeof(SimConfig), "ccs_sim_config must mirror ccs::SimConfig");

extern "C" {

int ccs_model_sizeof(void) { return (int)sizeof(ArrowModelParams); }

void ccs_model_synthetic(void* model_out) { synthetic_model(*(ArrowModelParams*)model_out); }

void ccs_sim_get_config(int config_id, ccs_sim_config* out) {
    SimConfig c = sim_config(config_id);
    std::memcpy(out, &c, sizeof(c));
}

int ccs_sim_zmw(const void* model, const ccs_sim_config* cfg, int64_t index, float* snr, uint8_t* tpl, int32_t tpl_cap,
                int32_t* tpl_len, uint8_t* codes, int64_t codes_cap, int32_t max_reads, int32_t* n_reads,
                int64_t* read_off, uint8_t* cx, uint8_t* strand, int32_t* tstart, int32_t* tend) {
    SimConfig c;
    std::memcpy(&c, cfg, sizeof(c));
    SimZmw z;
    simulate_zmw(*(const ArrowModelParams*)model, c, index, z);
    *tpl_len = (int32_t)z.tpl.size();
    *n_reads = (int32_t)z.reads.size();
    if ((int)z.tpl.size() > tpl_cap || (int)z.reads.size() > max_reads) return CCS_ERR_CAPACITY;
    int64_t tot = 0;
    for (auto& r : z.reads) tot += (int64_t)r.codes.size();
    if (tot > codes_cap) return CCS_ERR_CAPACITY;
    std::memcpy(snr, z.snr, sizeof(z.snr));
    std::memcpy(tpl, z.tpl.data(), z.tpl.size());
    int64_t off = 0;
    for (size_t k = 0; k < z.reads.size(); ++k) {
        const SimRead& r = z.reads[k];
        read_off[k] = off;
        std::memcpy(codes + off, r.codes.data(), r.codes.size());
        off += (int64_t)r.codes.size();
        cx[k] = r.cx; strand[k] = r.strand; tstart[k] = r.tstart; tend[k] = r.tend;
    }
    read_off[z.reads.size()] = off;
    return CCS_OK;
}

int ccs_sim_corrupt(const uint8_t* tpl, int32_t len, double rate, uint64_t seed, uint8_t* out, int32_t out_cap,
                    int32_t* out_len, int32_t* map) {
    std::vector<uint8_t> t(tpl, tpl + len), o;
    std::vector<int32_t> m;
    corrupt_template(t, rate, seed, o, m);
    *out_len = (int32_t)o.size();
    if ((int)o.size() > out_cap) return CCS_ERR_CAPACITY;
    std::memcpy(out, o.data(), o.size());
    std::memcpy(map, m.data(), sizeof(int32_t) * m.size());
    return CCS_OK;
}

}  // extern "C"
