// extern "C" boundary of libccsgpu.so (include/ccsgpu.h).  No exception crosses it.
#include "../../../include/ccsgpu.h"
#include "polish_engine.h"
#include "parallel.h"
#include "draft_engine.h"
#include "window_host.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <string>
#include <chrono>
#include <thread>
#include <algorithm>
#include <vector>
#include <exception>

using namespace ccs;

// A lane = one Draft + one Polish engine with their own streams and buffers.  A stage call splits its
// batch into one contiguous ZMW chunk per lane and runs the lanes on concurrent host threads, so one
// lane's host round logic and latency-bound kernels overlap another lane's issue-bound kernels.
struct Lane {
    std::unique_ptr<ArrowEngine> engine;
    std::unique_ptr<DraftEngine> draft;
    double ms_draft = 0;
};

struct ccsgpu_ctx {
    std::unique_ptr<ArrowEngine> engine;     // lane 0 (also serves the single-lane hooks)
    std::unique_ptr<DraftEngine> draft;
    std::vector<Lane> extra;                 // lanes 1..n-1
    int n_lanes = 1;
    int host_threads = 8;
    size_t budget = 0;           // device bytes this ctx may use (0 at create -> 85 % of the free memory then)
    size_t lane_budget() const { return n_lanes > 0 ? budget / (size_t)n_lanes : budget; }
    bool generic_score = false;
    bool reuse_scores = true;
    int qv_halo = 32;
    int score_variant = 0;
    double ms_e2e = 0;
    int64_t n_zmws = 0;
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
          ctx->host_threads = hc; ctx->engine->host_threads = hc; ctx->draft->host_threads = hc; }
        if (const char* e = std::getenv("CCS_B200_GENERIC_SCORE")) ctx->generic_score = (e[0] == '1');
        ctx->engine->generic_score = ctx->generic_score;
        if (const char* e = std::getenv("CCS_B200_REUSE_SCORES")) ctx->reuse_scores = (e[0] != '0');
        ctx->engine->reuse_scores = ctx->reuse_scores;
        if (const char* e = std::getenv("CCS_B200_QV_HALO")) ctx->qv_halo = std::max(0, std::atoi(e));
        ctx->engine->qv_halo = ctx->qv_halo;
        if (const char* e = std::getenv("CCS_B200_SCORE_VARIANT")) ctx->score_variant = std::atoi(e);
        ctx->engine->score_variant = ctx->score_variant;
        ctx->budget = device_bytes_budget;
        if (ctx->budget == 0) {
            size_t fr = 0, tot = 0;
            if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) ctx->budget = fr - fr / 7;
        }
        int lanes = 4;
        if (const char* e = std::getenv("CCS_B200_LANES")) lanes = std::max(1, std::min(8, std::atoi(e)));
        ccsgpu_set_lanes(ctx, lanes);
    } catch (const std::exception& e) {
        g_create_error = e.what();
        if (err) *err = CCS_ERR_NO_DEVICE;
        delete ctx;
        return nullptr;
    }
    return ctx;
}

void ccsgpu_destroy(ccsgpu_ctx* ctx) { delete ctx; }

int ccsgpu_set_lanes(ccsgpu_ctx* ctx, int n_lanes) {
    if (!ctx || !ctx->engine || n_lanes < 1 || n_lanes > 8) return CCS_ERR_ARG;
    try {
        while ((int)ctx->extra.size() < n_lanes - 1) {
            Lane l;
            l.engine.reset(new ArrowEngine(ctx->device, ctx->model, ctx->budget));
            l.draft.reset(new DraftEngine(ctx->device, 0));
            l.engine->generic_score = ctx->generic_score;
            l.engine->reuse_scores = ctx->reuse_scores;
            l.engine->qv_halo = ctx->qv_halo;
            l.engine->score_variant = ctx->score_variant;
            ctx->extra.push_back(std::move(l));
        }
    } catch (const std::exception& e) {
        ctx->last_error = e.what();
        return CCS_ERR_CUDA;
    }
    if (n_lanes != ctx->n_lanes) {     // the per-lane share of the budget changes: drop the grow-only buffers sized for the old share
        ctx->engine->release_buffers(); ctx->draft->release_buffers();
        for (auto& l : ctx->extra) { l.engine->release_buffers(); l.draft->release_buffers(); }
    }
    ctx->n_lanes = n_lanes;
    // a lane's device share: ~80 % Polish Stage (bands, row codes, delta rows), ~20 % Draft Stage scratch
    const size_t draft_budget = std::max<size_t>(ctx->lane_budget() / 5, 256ull << 20);
    ctx->draft->set_budget(draft_budget);
    for (auto& l : ctx->extra) l.draft->set_budget(draft_budget);
    const int th = std::max(1, ctx->host_threads / n_lanes);
    ctx->engine->host_threads = th; ctx->draft->host_threads = th;
    for (auto& l : ctx->extra) { l.engine->host_threads = th; l.draft->host_threads = th; }
    return CCS_OK;
}

const char* ccsgpu_last_error(const ccsgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

void ccs_polish_cfg_default(ccs_polish_cfg* c) {
    PolishParams p;
    c->max_iterations = p.max_iterations; c->separation = p.separation; c->neighborhood = p.neighborhood;
    c->min_length = p.min_length; c->max_length = p.max_length; c->min_rq = p.min_rq;
    c->ab_mismatch_tol = p.ab_mismatch_tol; c->min_active_fraction = p.min_active_fraction; c->min_zscore = p.min_zscore;
    c->window_size = p.window_size; c->window_overlap = p.window_overlap;
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

// CCS_B200_WINDOW=size[,overlap] overrides the caller's windowing (A/B runs; 0 = one window per draft)
static void window_env_override(PolishParams& pp) {
    const char* e = std::getenv("CCS_B200_WINDOW");
    if (!e || !*e) return;
    int a = 0, b = -1;
    const int n = std::sscanf(e, "%d,%d", &a, &b);
    if (n >= 1) pp.window_size = round_to_window_grid(a);
    if (n >= 2 && b >= 0) pp.window_overlap = round_to_window_grid(b);
}

static PolishParams to_params(const ccs_polish_cfg* cfg) {
    PolishParams pp;
    if (cfg) {
        pp.max_iterations = cfg->max_iterations; pp.separation = cfg->separation; pp.neighborhood = cfg->neighborhood;
        pp.min_length = cfg->min_length; pp.max_length = cfg->max_length; pp.min_rq = cfg->min_rq;
        pp.ab_mismatch_tol = cfg->ab_mismatch_tol; pp.min_active_fraction = cfg->min_active_fraction; pp.min_zscore = cfg->min_zscore;
        pp.window_size = round_to_window_grid(cfg->window_size); pp.window_overlap = round_to_window_grid(cfg->window_overlap);
    }
    window_env_override(pp);
    return pp;
}

// Results of one lane's chunk, merged into the caller's ccs_results afterwards.
struct ChunkOut {
    std::vector<uint8_t> seq, qv;
    std::vector<int64_t> seq_len, n_tested;
    std::vector<float> rq;
    std::vector<int32_t> status, n_passes, iterations, n_applied, read_status;
    std::vector<double> read_ll;
};

// Collects the engine's polish results; draft_status (optional) carries the Draft Stage verdicts.
static void collect_results(ArrowEngine& E, int nz, int nr, const uint8_t* cx, const PolishParams& pp,
                            const int32_t* draft_status, ChunkOut& co) {
    const auto& zs = E.zmw_states();
    const auto& qv = E.qvs();
    co.seq.clear(); co.qv.clear();
    co.seq_len.assign(nz, 0); co.n_tested.assign(nz, 0); co.rq.assign(nz, 0.f);
    co.status.assign(nz, 0); co.n_passes.assign(nz, 0); co.iterations.assign(nz, 0); co.n_applied.assign(nz, 0);
    co.read_ll.assign(nr, NAN); co.read_status.assign(nr, 4);
    E.read_lls(co.read_ll.data(), nullptr, co.read_status.data());
    for (int z = 0; z < nz; ++z) {
        const ZmwState& s = zs[z];
        const int J = (int)s.tpl.size();
        int status = CCS_ZMW_SUCCESS;
        double rq = 0.0;
        if (draft_status && draft_status[z] != CCS_ZMW_SUCCESS) status = draft_status[z];
        else if (s.failed || (int)qv[z].size() != J) status = CCS_ZMW_TOO_MANY_UNUSABLE;
        else {
            co.seq.insert(co.seq.end(), s.tpl.begin(), s.tpl.end());
            co.qv.insert(co.qv.end(), qv[z].begin(), qv[z].end());
            static const std::vector<double> err_of_qv = [] { std::vector<double> t(256); for (int q = 0; q < 256; ++q) t[q] = std::pow(10.0, -0.1 * q); return t; }();
            double e = 0;
            for (int j = 0; j < J; ++j) e += err_of_qv[qv[z][j]];
            rq = J ? 1.0 - e / J : 0.0;
            if (!s.converged) status = CCS_ZMW_NON_CONVERGENT;
            else if (J < pp.min_length) status = CCS_ZMW_TOO_SHORT;
            else if (pp.max_length > 0 && J > pp.max_length) status = CCS_ZMW_TOO_LONG;
            else if (rq < pp.min_rq) status = CCS_ZMW_POOR_QUALITY;
            co.seq_len[z] = J;
        }
        co.rq[z] = (float)rq; co.status[z] = status; co.iterations[z] = s.iterations; co.n_applied[z] = s.n_applied;
        co.n_tested[z] = s.n_tested;
        int np = 0;
        for (int r = s.read_begin; r < s.read_end; ++r)
            if (E.reads()[r].active && cx && (cx[r] & 3) == 3) ++np;
        co.n_passes[z] = np;
    }
}

// Results of a windowed chunk: the polished window cores of every ZMW are concatenated (docs/how-does-ccs-work.md:108-110);
// a subread reports the first failure among its windows, else the sum of its window log-likelihoods.
static void collect_windowed(ArrowEngine& E, const WindowPlan& P, int nz, int nr, const int32_t* zmw_read_off,
                             const uint8_t* cx, const PolishParams& pp, const int32_t* draft_status, ChunkOut& co) {
    const auto& zs = E.zmw_states();
    const auto& qv = E.qvs();
    const int nwr = P.n_reads();
    co.seq.clear(); co.qv.clear();
    co.seq_len.assign(nz, 0); co.n_tested.assign(nz, 0); co.rq.assign(nz, 0.f);
    co.status.assign(nz, 0); co.n_passes.assign(nz, 0); co.iterations.assign(nz, 0); co.n_applied.assign(nz, 0);
    co.read_ll.assign(nr, NAN); co.read_status.assign(nr, 4);
    std::vector<double> wll((size_t)nwr + 1, NAN);
    std::vector<int32_t> wst((size_t)nwr + 1, 4);
    E.read_lls(wll.data(), nullptr, wst.data());
    {
        std::vector<uint8_t> seen((size_t)nr, 0);
        for (int k = 0; k < nwr; ++k) {
            const int r = P.parent[k];
            if (!seen[r]) { seen[r] = 1; co.read_status[r] = 0; co.read_ll[r] = 0.0; }
            if (co.read_status[r] != 0) continue;
            if (wst[k] != 0) { co.read_status[r] = wst[k]; co.read_ll[r] = NAN; }
            else co.read_ll[r] += wll[k];
        }
    }
    static const std::vector<double> err_of_qv = [] { std::vector<double> t(256); for (int q = 0; q < 256; ++q) t[q] = std::pow(10.0, -0.1 * q); return t; }();
    for (int z = 0; z < nz; ++z) {
        const int w0 = P.zmw_win_off[z], w1 = P.zmw_win_off[z + 1];
        int status = CCS_ZMW_SUCCESS;
        double rq = 0.0;
        bool failed = false, converged = true;
        for (int w = w0; w < w1; ++w) {
            const ZmwState& s = zs[w];
            failed |= s.failed || qv[w].size() != s.tpl.size();
            converged &= s.converged;
            co.iterations[z] = std::max(co.iterations[z], s.iterations);
            co.n_applied[z] += s.n_applied;
            co.n_tested[z] += s.n_tested;
        }
        if (draft_status && draft_status[z] != CCS_ZMW_SUCCESS) status = draft_status[z];
        else if (P.zmw_empty[z] || w1 == w0) status = CCS_ZMW_EMPTY_WINDOW_DURING_POLISHING;
        else if (failed) status = CCS_ZMW_TOO_MANY_UNUSABLE;
        else {
            const size_t s_mark = co.seq.size();
            for (int w = w0; w < w1; ++w) {
                const ZmwState& s = zs[w];
                const int b = std::max(0, s.core_b), e = std::min((int)s.tpl.size(), s.core_e);
                if (e > b) {
                    co.seq.insert(co.seq.end(), s.tpl.begin() + b, s.tpl.begin() + e);
                    co.qv.insert(co.qv.end(), qv[w].begin() + b, qv[w].begin() + e);
                }
            }
            const int J = (int)(co.seq.size() - s_mark);
            double e = 0;
            for (int j = 0; j < J; ++j) e += err_of_qv[co.qv[s_mark + j]];
            rq = J ? 1.0 - e / J : 0.0;
            if (!converged) status = CCS_ZMW_NON_CONVERGENT;
            else if (J < pp.min_length) status = CCS_ZMW_TOO_SHORT;
            else if (pp.max_length > 0 && J > pp.max_length) status = CCS_ZMW_TOO_LONG;
            else if (rq < pp.min_rq) status = CCS_ZMW_POOR_QUALITY;
            co.seq_len[z] = J;
        }
        co.rq[z] = (float)rq; co.status[z] = status;
    }
    // np: full-length subreads that took part in every window they cover
    for (int z = 0; z < nz; ++z) {
        int np = 0;
        for (int r = zmw_read_off[z]; r < zmw_read_off[z + 1]; ++r)
            if (co.read_status[r] == 0 && cx && (cx[r] & 3) == 3) ++np;
        co.n_passes[z] = np;
    }
}

// A contiguous ZMW range of a batch with offsets rebased to the chunk.
struct SubBatch {
    int z0 = 0, z1 = 0, r0 = 0, r1 = 0;
    std::vector<int32_t> zmw_read_off;
    std::vector<int64_t> read_off, tpl_off;
    ccs_batch b;
    ccs_drafts d;
};

static void make_sub(const ccs_batch* in, const ccs_drafts* dr, int z0, int z1, SubBatch& sb) {
    sb.z0 = z0; sb.z1 = z1; sb.r0 = in->zmw_read_off[z0]; sb.r1 = in->zmw_read_off[z1];
    const int nz = z1 - z0, nr = sb.r1 - sb.r0;
    sb.zmw_read_off.resize(nz + 1);
    for (int z = 0; z <= nz; ++z) sb.zmw_read_off[z] = in->zmw_read_off[z0 + z] - sb.r0;
    sb.read_off.resize(nr + 1);
    const int64_t c0 = in->read_off[sb.r0];
    for (int r = 0; r <= nr; ++r) sb.read_off[r] = in->read_off[sb.r0 + r] - c0;
    sb.b.n_zmws = nz; sb.b.n_reads = nr; sb.b.zmw_read_off = sb.zmw_read_off.data(); sb.b.read_off = sb.read_off.data();
    sb.b.codes = in->codes + c0; sb.b.snr = in->snr + 4 * (size_t)z0;
    sb.b.cx = in->cx ? in->cx + sb.r0 : nullptr; sb.b.hole = in->hole ? in->hole + z0 : nullptr;
    if (dr) {
        sb.tpl_off.resize(nz + 1);
        const int64_t t0 = dr->tpl_off[z0];
        for (int z = 0; z <= nz; ++z) sb.tpl_off[z] = dr->tpl_off[z0 + z] - t0;
        sb.d.tpl_off = sb.tpl_off.data(); sb.d.tpl = dr->tpl + t0; sb.d.strand = dr->strand + sb.r0;
        sb.d.tstart = dr->tstart + sb.r0; sb.d.tend = dr->tend + sb.r0;
        sb.d.rstart = dr->rstart ? dr->rstart + sb.r0 : nullptr; sb.d.rend = dr->rend ? dr->rend + sb.r0 : nullptr;
    }
}

// Contiguous ZMW chunks with ~equal numbers of read bases: at least one per lane, and more (processed in
// waves, lane k takes chunks k, k+n_lanes, ...) when a chunk's device footprint would exceed the lane's share of
// the budget.  Footprint estimate per read base: two 128-B band columns + column info (alpha, beta), two row-code
// copies, ~15 % growth room; plus the per-position delta rows -- ~330 B per read base, rounded up to 400; plus the
// Draft Stage scratch of the same lane (graph pools, DP rows, traceback moves: ~100 B per read base).
static std::vector<int> split_zmws(const ccs_batch* in, int n_lanes, size_t lane_budget_bytes) {
    std::vector<int> cut(1, 0);
    const int nz = in->n_zmws;
    const int64_t total = in->read_off[in->n_reads];
    int n_chunks = std::max(1, std::min(n_lanes, std::max(1, nz / 8)));
    if (lane_budget_bytes > 0) {
        const int64_t per_chunk = (int64_t)(lane_budget_bytes / 500);   // + the Draft Stage's share
        const int need = (int)std::min<int64_t>(nz, (total + per_chunk - 1) / std::max<int64_t>(per_chunk, 1));
        if (need > n_chunks) n_chunks = ((need + n_lanes - 1) / n_lanes) * n_lanes;
    }
    for (int c = 1; c < n_chunks; ++c) {
        const int64_t want = total * c / n_chunks;
        int z = cut.back();
        while (z < nz && in->read_off[in->zmw_read_off[z]] < want) ++z;
        cut.push_back(std::max(z, cut.back()));
    }
    cut.push_back(nz);
    return cut;
}

static int merge_chunks(const ccs_batch* in, const std::vector<int>& cut, const std::vector<ChunkOut>& cos, ccs_results* out) {
    int64_t need = 0;
    for (const auto& c : cos) need += (int64_t)c.seq.size();
    if (need > out->seq_cap) { out->seq_cap = need; return (int)CCS_ERR_CAPACITY; }
    int64_t off = 0;
    for (size_t k = 0; k < cos.size(); ++k) {
        const ChunkOut& c = cos[k];
        const int z0 = cut[k], nz = cut[k + 1] - cut[k], r0 = in->zmw_read_off[z0], nr = in->zmw_read_off[cut[k + 1]] - r0;
        if (!c.seq.empty()) { std::memcpy(out->seq + off, c.seq.data(), c.seq.size()); std::memcpy(out->qv + off, c.qv.data(), c.qv.size()); }
        for (int z = 0; z < nz; ++z) {
            out->seq_off[z0 + z] = off;
            off += c.seq_len[z];
            if (out->rq) out->rq[z0 + z] = c.rq[z];
            if (out->status) out->status[z0 + z] = c.status[z];
            if (out->n_passes) out->n_passes[z0 + z] = c.n_passes[z];
            if (out->iterations) out->iterations[z0 + z] = c.iterations[z];
            if (out->n_applied) out->n_applied[z0 + z] = c.n_applied[z];
            if (out->n_tested) out->n_tested[z0 + z] = c.n_tested[z];
        }
        for (int r = 0; r < nr; ++r) {
            if (out->read_ll) out->read_ll[r0 + r] = c.read_ll[r];
            if (out->read_status) out->read_status[r0 + r] = c.read_status[r];
        }
    }
    out->seq_off[in->n_zmws] = off;
    return (int)CCS_OK;
}

static ArrowEngine& lane_engine(ccsgpu_ctx* ctx, int k) { return k == 0 ? *ctx->engine : *ctx->extra[k - 1].engine; }
static DraftEngine& lane_draft(ccsgpu_ctx* ctx, int k) { return k == 0 ? *ctx->draft : *ctx->extra[k - 1].draft; }

}  // extern "C"

// Runs f(lane, chunk) for every chunk: lane k (its own host thread) takes chunks k, k+n_lanes, ... in turn;
// rethrows the first failure.
template <class F>
static void run_lanes(int n_lanes, int n_chunks, F&& f) {
    n_lanes = std::max(1, std::min(n_lanes, n_chunks));
    std::vector<std::exception_ptr> errs(n_lanes);
    std::vector<std::thread> th;
    auto body = [&](int lane) {
        try { for (int c = lane; c < n_chunks; c += n_lanes) f(lane, c); } catch (...) { errs[lane] = std::current_exception(); }
    };
    for (int k = 1; k < n_lanes; ++k) th.emplace_back(body, k);
    body(0);
    for (auto& t : th) t.join();
    for (auto& e : errs) if (e) std::rethrow_exception(e);
}

extern "C" {

int ccsgpu_polish(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_drafts* drafts, const ccs_polish_cfg* cfg,
                  ccs_results* out) {
    return guarded(ctx, [&]() {
        const auto t_begin = std::chrono::steady_clock::now();
        const PolishParams pp = to_params(cfg);
        const std::vector<int> cut = split_zmws(in, ctx->n_lanes, ctx->lane_budget());
        const int nc = (int)cut.size() - 1;
        std::vector<ChunkOut> cos(nc);
        std::vector<SubBatch> subs(nc);
        run_lanes(ctx->n_lanes, nc, [&](int lane, int k) {
            make_sub(in, drafts, cut[k], cut[k + 1], subs[k]);
            ArrowEngine& E = lane_engine(ctx, lane);
            E.load(make_input(&subs[k].b, &subs[k].d));
            E.polish(pp);
            collect_results(E, subs[k].b.n_zmws, subs[k].b.n_reads, subs[k].b.cx, pp, nullptr, cos[k]);
        });
        const int rc = merge_chunks(in, cut, cos, out);
        ctx->ms_e2e += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        ctx->n_zmws += in->n_zmws;
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

// Draft Stage output -> Polish Stage input for one chunk, then polish + collect.
static void ccs_chunk(ccsgpu_ctx* ctx, int lane, const ccs_batch* in, const DraftParams& dpar, const PolishParams& pp,
                      ChunkOut& co) {
    const auto t0 = std::chrono::steady_clock::now();
    DraftOutput d;
    { HostPhase hp("ccs.draft total"); lane_draft(ctx, lane).run(to_draft_input(in), dpar, d); }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (lane == 0) ctx->ms_draft += ms; else ctx->extra[lane - 1].ms_draft += ms;
    // Windowing (window_host.h): every window of every draft that passed is a ZMW of the Arrow engine, its reads are slices
    // of the resident subread codes; a short draft is one window, i.e. the plain per-ZMW problem
    const int nz = in->n_zmws, nr = in->n_reads;
    WindowPlan P;
    {
        HostPhase hp("ccs.window plan");
        std::vector<uint8_t> ok((size_t)nz);
        for (int z = 0; z < nz; ++z) ok[z] = d.status[z] == CCS_ZMW_SUCCESS;
        WindowParams wp;
        wp.size = pp.window_size; wp.overlap = pp.window_overlap;
        build_window_plan(nz, in->zmw_read_off, in->read_off, in->snr, d.draft, ok.data(), d.maps.data(), d.grid.data(),
                          d.grid_off.data(), wp, P);
    }
    PolishInput p;
    p.n_zmws = P.n_windows(); p.n_reads = P.n_reads(); p.zmw_read_off = P.win_read_off.data(); p.codes = in->codes;
    p.snr = P.snr.data(); p.tpl_off = P.tpl_off.data(); p.tpl = P.tpl.data(); p.strand = P.strand.data();
    p.tstart = P.ts.data(); p.tend = P.te.data();
    p.code_start = P.code_start.data(); p.code_len = P.code_len.data(); p.code_base = nr ? in->read_off[0] : 0;
    p.read_group = P.parent.data(); p.n_groups = nr; p.zmw_group = P.win_zmw.data();
    p.core_b = P.core_b.data(); p.core_e = P.core_e.data(); p.growth_min = P.growth_min.data();
    p.d_codes = lane_draft(ctx, lane).device_codes();   // uploaded once by the Draft Stage of this lane, still resident
    ArrowEngine& E = lane_engine(ctx, lane);
    E.load(p);
    E.polish(pp);
    HostPhase hp("ccs.collect_results");
    collect_windowed(E, P, nz, nr, in->zmw_read_off, in->cx, pp, d.status.data(), co);
}

int ccsgpu_ccs(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_draft_cfg* dcfg, const ccs_polish_cfg* pcfg,
               ccs_results* out) {
    return guarded(ctx, [&]() {
        const auto t_begin = std::chrono::steady_clock::now();
        const DraftParams dpar = to_draft_params(dcfg);
        const PolishParams pp = to_params(pcfg);
        const std::vector<int> cut = split_zmws(in, ctx->n_lanes, ctx->lane_budget());
        const int nc = (int)cut.size() - 1;
        std::vector<ChunkOut> cos(nc);
        std::vector<SubBatch> subs(nc);
        { HostPhase hp("ccs.lanes total");
        run_lanes(ctx->n_lanes, nc, [&](int lane, int k) {
            { HostPhase hp2("ccs.make_sub"); make_sub(in, nullptr, cut[k], cut[k + 1], subs[k]); }
            ccs_chunk(ctx, lane, &subs[k].b, dpar, pp, cos[k]);
        });
        }
        int rc;
        { HostPhase hp("ccs.merge_chunks"); rc = merge_chunks(in, cut, cos, out); }
        HostProf::dump();
        ctx->ms_e2e += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        ctx->n_zmws += in->n_zmws;
        return rc;
    });
}

int ccsgpu_get_stats(ccsgpu_ctx* ctx, ccs_stats* out, int reset) {
    return guarded(ctx, [&]() {
        std::memset(out, 0, sizeof(*out));
        const int n = 1 + (int)ctx->extra.size();
        for (int k = 0; k < n; ++k) {
            const EngineStats& s = lane_engine(ctx, k).stats;
            const DraftStats& ds = lane_draft(ctx, k).stats;
            out->ms_fill_alpha += s.ms_fill_alpha; out->ms_fill_beta += s.ms_fill_beta; out->ms_score += s.ms_score;
            out->ms_pick += s.ms_pick; out->ms_qv += s.ms_qv; out->ms_h2d += s.ms_h2d;
            out->launches_fill_alpha += s.n_fill_alpha; out->launches_fill_beta += s.n_fill_beta;
            out->launches_score += s.n_score; out->launches_pick += s.n_pick; out->launches_qv += s.n_qv;
            out->bytes_fill_alpha += s.bytes_fill_alpha; out->bytes_fill_beta += s.bytes_fill_beta;
            out->cells_fill += s.cells_fill; out->score_items += s.score_items; out->rounds += s.rounds;
            out->h2d_bytes += s.h2d_bytes + ds.h2d_bytes; out->d2h_bytes += s.d2h_bytes + ds.d2h_bytes;
            out->ms_resident += s.ms_resident;
            if (s.top_fill_alpha_bytes > out->top_fill_alpha_bytes) { out->top_fill_alpha_bytes = s.top_fill_alpha_bytes; out->top_fill_alpha_ms = s.top_fill_alpha_ms; }
            out->ms_poa_align += ds.ms_align; out->launches_poa += ds.n_align_launches; out->poa_tasks += ds.n_tasks;
            out->poa_rows += ds.rows; out->bytes_poa_align += ds.bytes_align;
            out->launches_draft += ds.n_align_launches + ds.n_graph_launches;
            out->ms_poa_map += ds.ms_map; out->ms_poa_graph += ds.ms_graph; out->launches_poa_graph += ds.n_graph_launches;
            out->bytes_poa_map += ds.bytes_map;
            out->bytes_score += s.bytes_score; out->launches_pack += s.n_pack;
            if (s.top_fill_beta_bytes > out->top_fill_beta_bytes) { out->top_fill_beta_bytes = s.top_fill_beta_bytes; out->top_fill_beta_ms = s.top_fill_beta_ms; }
            if (s.top_score_bytes > out->top_score_bytes) { out->top_score_bytes = s.top_score_bytes; out->top_score_ms = s.top_score_ms; }
            if (ds.top_align_bytes > out->top_poa_align_bytes) { out->top_poa_align_bytes = ds.top_align_bytes; out->top_poa_align_ms = ds.top_align_ms; }
            if (ds.top_map_bytes > out->top_poa_map_bytes) { out->top_poa_map_bytes = ds.top_map_bytes; out->top_poa_map_ms = ds.top_map_ms; }
            out->ms_draft += (k == 0 ? ctx->ms_draft : ctx->extra[k - 1].ms_draft);
        }
        out->ms_e2e = ctx->ms_e2e; out->n_zmws = ctx->n_zmws;
        if (reset) {
            for (int k = 0; k < n; ++k) { lane_engine(ctx, k).reset_stats(); lane_draft(ctx, k).stats = DraftStats(); }
            ctx->ms_draft = 0; for (auto& l : ctx->extra) l.ms_draft = 0;
            ctx->ms_e2e = 0; ctx->n_zmws = 0;
        }
        return (int)CCS_OK;
    });
}

}  // extern "C"
