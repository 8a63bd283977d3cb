// C ABI of libccssim.so: the synthetic ZMW generator (include/ccssim.h).  Test / bench infrastructure, built
// separately from the product library.
#include "../common/sim.h"
#include "../../../include/ccssim.h"
#include "../../../include/ccsgpu.h"   // error codes only
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

// ---- batch generation (multi-threaded): opaque handle + copy-out ----------------------------
#include <thread>
#include <atomic>

namespace {
struct SimBatch {
    std::vector<SimZmw> z;
    std::vector<std::vector<uint8_t>> draft;
    std::vector<std::vector<int32_t>> map;
};
}  // namespace

extern "C" {

void* ccs_sim_batch_create(const void* model, const ccs_sim_config* cfg, int64_t first_index, int32_t n_zmws,
                           double draft_error_rate, int32_t n_threads) {
    SimConfig c;
    std::memcpy((void*)&c, cfg, sizeof(c));
    auto* b = new SimBatch();
    b->z.resize(n_zmws); b->draft.resize(n_zmws); b->map.resize(n_zmws);
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= n_zmws) break;
            simulate_zmw(*(const ArrowModelParams*)model, c, first_index + k, b->z[k]);
            if (draft_error_rate >= 0)
                corrupt_template(b->z[k].tpl, draft_error_rate, (uint64_t)(first_index + k) + 17, b->draft[k], b->map[k]);
        }
    };
    const int nt = std::max(1, std::min<int>(n_threads, n_zmws));
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return b;
}

void ccs_sim_batch_free(void* h) { delete (SimBatch*)h; }

// sizes[4] = {n_reads, total_codes, total_truth_bases, total_draft_bases}
void ccs_sim_batch_sizes(const void* h, int64_t* sizes) {
    const SimBatch* b = (const SimBatch*)h;
    int64_t nr = 0, nc = 0, nt = 0, nd = 0;
    for (size_t k = 0; k < b->z.size(); ++k) {
        nr += (int64_t)b->z[k].reads.size();
        for (auto& r : b->z[k].reads) nc += (int64_t)r.codes.size();
        nt += (int64_t)b->z[k].tpl.size();
        nd += (int64_t)b->draft[k].size();
    }
    sizes[0] = nr; sizes[1] = nc; sizes[2] = nt; sizes[3] = nd;
}

// Copies the batch into caller arrays (any pointer may be NULL).  Draft spans are the truth
// spans mapped onto the corrupted draft.
void ccs_sim_batch_copy(const void* h, int32_t* zmw_read_off, int64_t* read_off, uint8_t* codes, float* snr, uint8_t* cx,
                        int32_t* hole, int64_t* truth_off, uint8_t* truth, uint8_t* strand, int32_t* tstart,
                        int32_t* tend, int64_t* draft_off, uint8_t* draft, int32_t* dstart, int32_t* dend) {
    const SimBatch* b = (const SimBatch*)h;
    int64_t r = 0, c = 0, t = 0, d = 0;
    for (size_t k = 0; k < b->z.size(); ++k) {
        const SimZmw& z = b->z[k];
        if (zmw_read_off) zmw_read_off[k] = (int32_t)r;
        if (snr) std::memcpy(snr + 4 * k, z.snr, sizeof(z.snr));
        if (hole) hole[k] = z.hole;
        if (truth_off) truth_off[k] = t;
        if (truth) std::memcpy(truth + t, z.tpl.data(), z.tpl.size());
        if (draft_off) draft_off[k] = d;
        if (draft && !b->draft[k].empty()) std::memcpy(draft + d, b->draft[k].data(), b->draft[k].size());
        for (auto& rd : z.reads) {
            if (read_off) read_off[r] = c;
            if (codes) std::memcpy(codes + c, rd.codes.data(), rd.codes.size());
            if (cx) cx[r] = rd.cx;
            if (strand) strand[r] = rd.strand;
            if (tstart) tstart[r] = rd.tstart;
            if (tend) tend[r] = rd.tend;
            if (dstart && !b->map[k].empty()) dstart[r] = b->map[k][rd.tstart];
            if (dend && !b->map[k].empty()) dend[r] = b->map[k][rd.tend];
            c += (int64_t)rd.codes.size();
            ++r;
        }
        t += (int64_t)z.tpl.size();
        d += (int64_t)b->draft[k].size();
    }
    if (zmw_read_off) zmw_read_off[b->z.size()] = (int32_t)r;
    if (read_off) read_off[r] = c;
    if (truth_off) truth_off[b->z.size()] = t;
    if (draft_off) draft_off[b->z.size()] = d;
}

}  // extern "C"

// ---- synthetic subreads.bam (tests / demos of the `ccs` command line) --------------------------
#include "../host/bam_io.h"

extern "C" {

int ccs_sim_write_subreads_bam(const char* path, const char* movie, const void* model, const ccs_sim_config* cfg,
                               int64_t first_index, int32_t n_zmws, int32_t with_chemistry) {
    SimConfig c;
    std::memcpy((void*)&c, cfg, sizeof(c));
    SubreadBamWriter w;
    if (!w.open(path, movie, with_chemistry != 0)) return CCS_ERR_IO;
    for (int32_t k = 0; k < n_zmws; ++k) {
        SimZmw z;
        simulate_zmw(*(const ArrowModelParams*)model, c, first_index + k, z);
        int32_t q = 0;
        for (const SimRead& r : z.reads) {
            SubreadOut s{z.hole + 1, q, q + (int32_t)r.codes.size(), z.snr, r.cx, r.codes.data(), (int32_t)r.codes.size()};
            w.write(s);
            q += (int32_t)r.codes.size() + 45;   // adapter gap
        }
    }
    w.close();
    return CCS_OK;
}

}  // extern "C"
