// extern "C" boundary of libccsgpu.so (include/ccsgpu.h).  No exception crosses it.
#include "../../../include/ccsgpu.h"
#include "polish_engine.h"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <string>
#include <chrono>

using namespace ccs;

struct ccsgpu_ctx {
    std::unique_ptr<ArrowEngine> engine;
    ArrowModelParams model;
    std::string last_error;
};

static thread_local std::string g_create_error;

namespace {

template <class F>
int guarded(ccsgpu_ctx* ctx, F&& f) {
    if (!ctx || !ctx->engine) return CCS_ERR_ARG;
    try {
        return f();
    } catch (const OomError& e) {
        ctx->last_error = e.what();
        return CCS_ERR_OOM;
    } catch (const CudaError& e) {
        ctx->last_error = e.what();
        return CCS_ERR_CUDA;
    } catch (const std::exception& e) {
        ctx->last_error = e.what();
        return CCS_ERR_ARG;
    }
}

PolishInput make_input(const ccs_batch* in, const ccs_drafts* d) {
    PolishInput p;
    p.n_zmws = in->n_zmws; p.n_reads = in->n_reads;
    p.zmw_read_off = in->zmw_read_off; p.read_off = in->read_off; p.codes = in->codes; p.snr = in->snr;
    p.tpl_off = d->tpl_off; p.tpl = d->tpl; p.strand = d->strand; p.tstart = d->tstart; p.tend = d->tend;
    return p;
}

}  // namespace

extern "C" {

ccsgpu_ctx* ccsgpu_create(int device, const void* model, size_t device_bytes_budget, int* err) {
    if (err) *err = CCS_OK;
    if (!model) { if (err) *err = CCS_ERR_ARG; g_create_error = "model is NULL"; return nullptr; }
    auto* ctx = new ccsgpu_ctx();
    std::memcpy(&ctx->model, model, sizeof(ArrowModelParams));
    try {
        ctx->engine.reset(new ArrowEngine(device, ctx->model, device_bytes_budget));
        if (const char* e = std::getenv("CCS_B200_GENERIC_SCORE")) ctx->engine->generic_score = (e[0] == '1');
    } catch (const std::exception& e) {
        g_create_error = e.what();
        if (err) *err = CCS_ERR_NO_DEVICE;
        delete ctx;
        return nullptr;
    }
    return ctx;
}

void ccsgpu_destroy(ccsgpu_ctx* ctx) { delete ctx; }

const char* ccsgpu_last_error(const ccsgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

void ccs_polish_cfg_default(ccs_polish_cfg* c) {
    PolishParams p;
    c->max_iterations = p.max_iterations; c->separation = p.separation; c->neighborhood = p.neighborhood;
    c->min_length = p.min_length; c->max_length = p.max_length; c->min_rq = p.min_rq;
    c->ab_mismatch_tol = p.ab_mismatch_tol; c->min_active_fraction = p.min_active_fraction;
}

int ccsgpu_fill_alpha_beta(ccsgpu_ctx* ctx, int32_t n_pairs, const int64_t* tpl_off, const uint8_t* tpl,
                           const int64_t* read_off, const uint8_t* codes, const float* snr, double* ll_alpha,
                           double* ll_beta, int32_t* status, int32_t dump_pair, float* alpha_out, float* beta_out,
                           int32_t* start_out, int32_t* aexp_out, int32_t* bexp_out) {
    return guarded(ctx, [&]() {
        std::vector<int32_t> zoff(n_pairs + 1), ts(n_pairs, 0), te(n_pairs);
        std::vector<uint8_t> strand(n_pairs, 0);
        for (int k = 0; k <= n_pairs; ++k) zoff[k] = k;
        for (int k = 0; k < n_pairs; ++k) te[k] = (int32_t)(tpl_off[k + 1] - tpl_off[k]);
        PolishInput p;
        p.n_zmws = n_pairs; p.n_reads = n_pairs; p.zmw_read_off = zoff.data(); p.read_off = read_off; p.codes = codes;
        p.snr = snr; p.tpl_off = tpl_off; p.tpl = tpl; p.strand = strand.data(); p.tstart = ts.data(); p.tend = te.data();
        ctx->engine->load(p);
        ctx->engine->fill();
        ctx->engine->read_lls(ll_alpha, ll_beta, status);
        if (dump_pair >= 0 && dump_pair < n_pairs)
            ctx->engine->dump_pair(dump_pair, alpha_out, beta_out, start_out, aexp_out, bexp_out);
        return (int)CCS_OK;
    });
}

int ccsgpu_score_all(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_drafts* drafts, double* delta, double* read_ll,
                     int32_t* read_status) {
    return guarded(ctx, [&]() {
        ArrowEngine& E = *ctx->engine;
        E.load(make_input(in, drafts));
        E.fill();
        E.read_lls(read_ll, nullptr, read_status);
        E.score_all_positions();
        for (int z = 0; z < in->n_zmws; ++z) E.download_delta(z, delta + drafts->tpl_off[z] * 9);
        return (int)CCS_OK;
    });
}

int ccsgpu_polish(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_drafts* drafts, const ccs_polish_cfg* cfg,
                  ccs_results* out) {
    return guarded(ctx, [&]() {
        ArrowEngine& E = *ctx->engine;
        const auto t_begin = std::chrono::steady_clock::now();
        PolishParams pp;
        if (cfg) {
            pp.max_iterations = cfg->max_iterations; pp.separation = cfg->separation; pp.neighborhood = cfg->neighborhood;
            pp.min_length = cfg->min_length; pp.max_length = cfg->max_length; pp.min_rq = cfg->min_rq;
            pp.ab_mismatch_tol = cfg->ab_mismatch_tol; pp.min_active_fraction = cfg->min_active_fraction;
        }
        E.load(make_input(in, drafts));
        E.polish(pp);
        const auto& zs = E.zmw_states();
        const auto& qv = E.qvs();
        int64_t need = 0;
        for (const auto& z : zs) need += (int64_t)z.tpl.size();
        if (need > out->seq_cap) { out->seq_cap = need; return (int)CCS_ERR_CAPACITY; }
        if (out->read_ll || out->read_status) E.read_lls(out->read_ll, nullptr, out->read_status);
        int64_t off = 0;
        for (int z = 0; z < in->n_zmws; ++z) {
            const ZmwState& s = zs[z];
            out->seq_off[z] = off;
            const int J = (int)s.tpl.size();
            int status = CCS_ZMW_SUCCESS;
            double rq = 0.0;
            if (s.failed) status = CCS_ZMW_TOO_MANY_UNUSABLE;
            else {
                std::memcpy(out->seq + off, s.tpl.data(), (size_t)J);
                double e = 0;
                for (int j = 0; j < J; ++j) { out->qv[off + j] = qv[z][j]; e += std::pow(10.0, -0.1 * qv[z][j]); }
                rq = J ? 1.0 - e / J : 0.0;
                if (!s.converged) status = CCS_ZMW_NON_CONVERGENT;
                else if (J < pp.min_length) status = CCS_ZMW_TOO_SHORT;
                else if (J > pp.max_length) status = CCS_ZMW_TOO_LONG;
                else if (rq < pp.min_rq) status = CCS_ZMW_POOR_QUALITY;
                off += J;
            }
            if (out->rq) out->rq[z] = (float)rq;
            if (out->status) out->status[z] = status;
            if (out->iterations) out->iterations[z] = s.iterations;
            if (out->n_applied) out->n_applied[z] = s.n_applied;
            if (out->n_tested) out->n_tested[z] = s.n_tested;
            if (out->n_passes) {
                int np = 0;
                for (int r = s.read_begin; r < s.read_end; ++r)
                    if (E.reads()[r].active && in->cx && (in->cx[r] & 3) == 3) ++np;
                out->n_passes[z] = np;
            }
        }
        out->seq_off[in->n_zmws] = off;
        E.stats.ms_e2e += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        return (int)CCS_OK;
    });
}

int ccsgpu_get_stats(ccsgpu_ctx* ctx, ccs_stats* out, int reset) {
    return guarded(ctx, [&]() {
        const EngineStats& s = ctx->engine->stats;
        std::memset(out, 0, sizeof(*out));
        out->ms_fill_alpha = s.ms_fill_alpha; out->ms_fill_beta = s.ms_fill_beta; out->ms_score = s.ms_score;
        out->ms_pick = s.ms_pick; out->ms_qv = s.ms_qv; out->ms_h2d = s.ms_h2d;
        out->launches_fill_alpha = s.n_fill_alpha; out->launches_fill_beta = s.n_fill_beta;
        out->launches_score = s.n_score; out->launches_pick = s.n_pick; out->launches_qv = s.n_qv;
        out->bytes_fill_alpha = s.bytes_fill_alpha; out->bytes_fill_beta = s.bytes_fill_beta;
        out->cells_fill = s.cells_fill; out->score_items = s.score_items; out->rounds = s.rounds;
        out->h2d_bytes = s.h2d_bytes; out->d2h_bytes = s.d2h_bytes;
        out->ms_resident = s.ms_resident; out->ms_e2e = s.ms_e2e; out->n_zmws = s.n_zmws;
        if (reset) ctx->engine->reset_stats();
        return (int)CCS_OK;
    });
}

}  // extern "C"
