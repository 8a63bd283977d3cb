// extern "C" boundary of libccsgpu.so (include/ccsgpu.h).  No exception crosses it.
#include "../../../include/ccsgpu.h"
#include "polish_engine.h"
#include "draft_engine.h"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <string>
#include <chrono>
#include <thread>
#include <algorithm>

using namespace ccs;

struct ccsgpu_ctx {
    std::unique_ptr<ArrowEngine> engine;
    std::unique_ptr<DraftEngine> draft;
    int device = 0;
    double ms_draft = 0;   // wall time of the Draft Stage calls (host graph work + GPU alignment)
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
    p.rstart = d->rstart; p.rend = d->rend;
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
        ctx->draft.reset(new DraftEngine(device, 0));
        ctx->device = device;
        { int hc = (int)std::thread::hardware_concurrency(); if (hc < 1) hc = 8;
          if (const char* e = std::getenv("CCS_B200_THREADS")) hc = std::max(1, std::atoi(e));
          ctx->engine->host_threads = hc; ctx->draft->host_threads = hc; }
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

static PolishParams to_params(const ccs_polish_cfg* cfg) {
    PolishParams pp;
    if (cfg) {
        pp.max_iterations = cfg->max_iterations; pp.separation = cfg->separation; pp.neighborhood = cfg->neighborhood;
        pp.min_length = cfg->min_length; pp.max_length = cfg->max_length; pp.min_rq = cfg->min_rq;
        pp.ab_mismatch_tol = cfg->ab_mismatch_tol; pp.min_active_fraction = cfg->min_active_fraction;
    }
    return pp;
}

// Writes the engine's polish results; draft_status (optional) carries the Draft Stage verdicts.
static int write_results(ArrowEngine& E, const ccs_batch* in, const PolishParams& pp, const int32_t* draft_status,
                         ccs_results* out) {
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
        if (draft_status && draft_status[z] != CCS_ZMW_SUCCESS) status = draft_status[z];
        else if (s.failed || (int)qv[z].size() != J) status = CCS_ZMW_TOO_MANY_UNUSABLE;
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
    return (int)CCS_OK;
}

int ccsgpu_polish(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_drafts* drafts, const ccs_polish_cfg* cfg,
                  ccs_results* out) {
    return guarded(ctx, [&]() {
        ArrowEngine& E = *ctx->engine;
        const auto t_begin = std::chrono::steady_clock::now();
        const PolishParams pp = to_params(cfg);
        E.load(make_input(in, drafts));
        E.polish(pp);
        const int rc = write_results(E, in, pp, nullptr, out);
        E.stats.ms_e2e += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        return rc;
    });
}

void ccs_draft_cfg_default(ccs_draft_cfg* c) {
    DraftParams d;
    c->min_snr = d.min_snr; c->min_passes = d.min_passes; c->top_passes = d.top_passes; c->max_poa_reads = d.max_poa_reads;
    c->min_length = d.min_length; c->max_length = d.max_length;
}

static DraftParams to_draft_params(const ccs_draft_cfg* c) {
    DraftParams d;
    if (c) {
        d.min_snr = c->min_snr; d.min_passes = c->min_passes; d.top_passes = c->top_passes;
        d.max_poa_reads = c->max_poa_reads; d.min_length = c->min_length; d.max_length = c->max_length;
    }
    return d;
}

static DraftInput to_draft_input(const ccs_batch* in) {
    DraftInput di;
    di.n_zmws = in->n_zmws; di.n_reads = in->n_reads; di.zmw_read_off = in->zmw_read_off; di.read_off = in->read_off;
    di.codes = in->codes; di.snr = in->snr; di.cx = in->cx;
    return di;
}

int ccsgpu_draft(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_draft_cfg* cfg, ccs_drafts_out* out) {
    return guarded(ctx, [&]() {
        DraftOutput d;
        ctx->draft->run(to_draft_input(in), to_draft_params(cfg), d);
        int64_t need = 0;
        for (auto& t : d.draft) need += (int64_t)t.size();
        if (need > out->tpl_cap) { out->tpl_cap = need; return (int)CCS_ERR_CAPACITY; }
        int64_t off = 0;
        for (int z = 0; z < in->n_zmws; ++z) {
            out->tpl_off[z] = off;
            if (!d.draft[z].empty()) std::memcpy(out->tpl + off, d.draft[z].data(), d.draft[z].size());
            off += (int64_t)d.draft[z].size();
            out->status[z] = d.status[z];
        }
        out->tpl_off[in->n_zmws] = off;
        for (int r = 0; r < in->n_reads; ++r) {
            const ReadMap& m = d.maps[r];
            out->strand[r] = (uint8_t)m.strand;
            out->tstart[r] = m.mapped ? m.tstart : 0; out->tend[r] = m.mapped ? m.tend : 0;
            out->rstart[r] = m.mapped ? m.rstart : 0; out->rend[r] = m.mapped ? m.rend : 0;
        }
        return (int)CCS_OK;
    });
}

int ccsgpu_ccs(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_draft_cfg* dcfg, const ccs_polish_cfg* pcfg,
               ccs_results* out) {
    return guarded(ctx, [&]() {
        const auto t_begin = std::chrono::steady_clock::now();
        DraftOutput d;
        ctx->draft->run(to_draft_input(in), to_draft_params(dcfg), d);
        ctx->ms_draft += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        // Polish Stage input from the Draft Stage output; ZMWs that failed the draft get an empty template
        const int nz = in->n_zmws, nr = in->n_reads;
        std::vector<int64_t> tpl_off(nz + 1, 0);
        for (int z = 0; z < nz; ++z)
            tpl_off[z + 1] = tpl_off[z] + (d.status[z] == CCS_ZMW_SUCCESS ? (int64_t)d.draft[z].size() : 0);
        std::vector<uint8_t> tpl((size_t)tpl_off[nz] + 1), strand(nr);
        std::vector<int32_t> ts(nr), te(nr), rs(nr), re(nr);
        for (int z = 0; z < nz; ++z) {
            const bool ok = d.status[z] == CCS_ZMW_SUCCESS;
            if (ok) std::memcpy(tpl.data() + tpl_off[z], d.draft[z].data(), d.draft[z].size());
            for (int r = in->zmw_read_off[z]; r < in->zmw_read_off[z + 1]; ++r) {
                const ReadMap& m = d.maps[r];
                const bool use = ok && m.mapped;
                strand[r] = (uint8_t)m.strand;
                ts[r] = use ? m.tstart : 0; te[r] = use ? m.tend : 0;
                rs[r] = use ? m.rstart : 0; re[r] = use ? m.rend : 0;
            }
        }
        PolishInput p;
        p.n_zmws = nz; p.n_reads = nr; p.zmw_read_off = in->zmw_read_off; p.read_off = in->read_off; p.codes = in->codes;
        p.snr = in->snr; p.tpl_off = tpl_off.data(); p.tpl = tpl.data(); p.strand = strand.data();
        p.tstart = ts.data(); p.tend = te.data(); p.rstart = rs.data(); p.rend = re.data();
        ArrowEngine& E = *ctx->engine;
        const PolishParams pp = to_params(pcfg);
        E.load(p);
        E.polish(pp);
        const int rc = write_results(E, in, pp, d.status.data(), out);
        E.stats.ms_e2e += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        return rc;
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
        { const DraftStats& ds = ctx->draft->stats;
          out->ms_poa_align = ds.ms_align; out->launches_poa = ds.n_align_launches; out->poa_tasks = ds.n_tasks;
          out->ms_draft = ctx->ms_draft; out->poa_rows = ds.rows; out->bytes_poa_align = ds.bytes_align; out->launches_draft = ds.n_align_launches; }
        out->ms_resident = s.ms_resident; out->ms_e2e = s.ms_e2e; out->n_zmws = s.n_zmws;
        if (reset) { ctx->engine->reset_stats(); ctx->draft->stats = DraftStats(); ctx->ms_draft = 0; }
        return (int)CCS_OK;
    });
}

}  // extern "C"
