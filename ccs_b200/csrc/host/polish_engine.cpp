// Polish Stage host engine -- see polish_engine.h.
#include "polish_engine.h"
#include "polish_host.h"
#include "../cuda/arrow_launch.h"
#include "parallel.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <atomic>

namespace ccs {

ArrowEngine::ArrowEngine(int device, const ArrowModelParams& model, size_t budget_bytes)
    : device_(device), budget_(budget_bytes), model_(model) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device >= n)
        throw CudaError(std::string("no usable CUDA device (ccs_b200 has no CPU fallback): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range"));
    CCS_CUDA(cudaSetDevice(device_));
    if (budget_ == 0) {
        size_t fr = 0, tot = 0;
        CCS_CUDA(cudaMemGetInfo(&fr, &tot));
        budget_ = fr - fr / 10;
    }
    CCS_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        const char* e = std::getenv("CCS_B200_PRIO");
        if (!(e && e[0] == '0') && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess && hi < lo) {
            CCS_CUDA(cudaStreamCreateWithPriority(&stream_hi_, cudaStreamNonBlocking, hi));
            CCS_CUDA(cudaEventCreateWithFlags(&ev_order_, cudaEventDisableTiming));
        }
    }
    CCS_CUDA(cudaEventCreate(&evA_));
    CCS_CUDA(cudaEventCreate(&evB_));
    CCS_CUDA(make_blocking_event(&ev_sync_));
    build_emission_tables(model_, em_);
    for (auto* b : {&d_emm_, &d_emi_, &d_trans_, &d_alpha_, &d_beta_}) b->budget_used = &used_;
    d_rowcode_.budget_used = &used_; d_tpl_.budget_used = &used_; d_colinfo_.budget_used = &used_;
    d_bexp_.budget_used = &used_; d_delta_.budget_used = &used_; d_qv_.budget_used = &used_;
    d_emm_.ensure(36 * 16);
    d_emi_.ensure(17 * 16);
    CCS_CUDA(cudaMemcpyAsync(d_emm_.p, em_.em_match, sizeof(em_.em_match), cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_emi_.p, em_.em_ins, sizeof(em_.em_ins), cudaMemcpyHostToDevice, stream_));
    d_counter_.ensure(4);
    h_counter_.ensure(4);
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
}

ArrowEngine::~ArrowEngine() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    if (evA_) cudaEventDestroy(evA_);
    if (evB_) cudaEventDestroy(evB_);
    for (cudaEvent_t e : ev_pool_) cudaEventDestroy(e);
    if (stream_) cudaStreamDestroy(stream_);
}

void ArrowEngine::release_buffers() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    d_rowcode_.release(); d_tpl_.release(); d_rawcodes_.release(); d_alpha_.release(); d_beta_.release();
    d_colinfo_.release(); d_bexp_.release(); d_delta_.release(); d_delta_scratch_.release(); d_qv_.release();
    d_cand_.release(); d_trans_.release();
}

cudaEvent_t ArrowEngine::next_event() {
    if (ev_used_ == ev_pool_.size()) {
        cudaEvent_t e;
        CCS_CUDA(cudaEventCreate(&e));
        ev_pool_.push_back(e);
    }
    return ev_pool_[ev_used_++];
}

void ArrowEngine::span_begin(double* acc, int64_t bytes, int64_t* top_bytes, double* top_ms, cudaStream_t s) {
    if (!timing_enabled) return;
    Span sp{next_event(), next_event(), acc, bytes, top_bytes, top_ms};
    CCS_CUDA(cudaEventRecord(sp.a, s ? s : stream_));
    spans_.push_back(sp);
}

void ArrowEngine::span_end(cudaStream_t s) {
    if (!timing_enabled) return;
    CCS_CUDA(cudaEventRecord(spans_.back().b, s ? s : stream_));
}

// call only when the stream is known to be idle (after a synchronize)
void ArrowEngine::resolve_spans() {
    for (const Span& sp : spans_) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            *sp.acc += ms;
            if (sp.top_bytes && sp.bytes > *sp.top_bytes) { *sp.top_bytes = sp.bytes; *sp.top_ms = ms; }
        }
    }
    spans_.clear();
    ev_used_ = 0;
}

ArrowBatchView ArrowEngine::view() const {
    ArrowBatchView V;
    V.rowcode = d_rowcode_.p; V.tpl = d_tpl_.p; V.em_match = d_emm_.p; V.em_ins = d_emi_.p; V.trans = d_trans_.p;
    V.reads = d_reads_.p; V.zmws = d_zmws_.p; V.alpha = d_alpha_.p; V.beta = d_beta_.p; V.colinfo = d_colinfo_.p;
    V.beta_exp = d_bexp_.p; V.ll_alpha = d_ll_alpha_.p; V.ll_beta = d_ll_beta_.p; V.base_ll = d_base_ll_.p;
    V.status = d_status_.p; V.n_reads = (int32_t)reads_.size(); V.n_zmws = (int32_t)zmws_.size();
    V.log_cw = std::log(model_.counter_weight); V.ab_tol = ab_tol_;
    return V;
}

// ---------------------------------------------------------------------------------------------
// load: pack one batch.  Row codes are stored row-indexed with sentinels so the kernels never
// bounds-check (Recursor::EncodeRead); transition tables per ZMW
// (ModelConfig::Populate, row a6).
// ---------------------------------------------------------------------------------------------
void ArrowEngine::load(const PolishInput& in) {
    HostPhase hp("polish.load pack");
    CCS_CUDA(cudaSetDevice(device_));
    const int nz = in.n_zmws, nr = in.n_reads;
    zstate_.assign(nz, ZmwState());
    reads_.assign(nr, DevRead());
    zmws_.assign(nz, DevZmw());
    status_.assign(nr, 4);
    qv_.assign(nz, {});
    tpl_cap_.assign(nz, 0);

    // row codes: two 16-byte aligned copies per read (row i, and row i+1 for the backward pass), pre-multiplied by 4 =
    // byte offset into an emission-table row -- encoded ON THE DEVICE from the resident codes (arrow_pack.cu)
    int64_t code_total = 0;
    const bool sliced = in.code_start != nullptr;
    if (sliced && (!in.code_len || !in.d_codes)) throw std::invalid_argument("sliced reads need code_len and resident codes");
    auto rlen = [&](int r) -> int64_t {
        if (sliced) return in.code_len[r];
        return in.rstart ? std::max(0, in.rend[r] - in.rstart[r]) : (in.read_off[r + 1] - in.read_off[r]);
    };
    read_group_.clear(); zmw_group_.clear(); n_groups_ = 0;
    if (in.read_group) { read_group_.assign(in.read_group, in.read_group + nr); n_groups_ = in.n_groups; }
    if (in.zmw_group) zmw_group_.assign(in.zmw_group, in.zmw_group + nz);
    std::vector<int64_t> coffs(nr + 1, 0);
    for (int r = 0; r < nr; ++r) {
        const bool mapped = in.tend[r] > in.tstart[r];        // unmapped reads are never touched: no row codes
        coffs[r + 1] = coffs[r] + (mapped ? 2 * ((rlen(r) + kRowCodePad + 1 + 15) & ~15ll) : 0);
    }
    code_total = coffs[nr];
    for (int z = 0; z < nz; ++z) {
        ZmwState& zs = zstate_[z];
        zs.read_begin = in.zmw_read_off[z];
        zs.read_end = in.zmw_read_off[z + 1];
        zs.tpl.assign(in.tpl + in.tpl_off[z], in.tpl + in.tpl_off[z + 1]);
        zs.seen.assign(1, tpl_hash(zs.tpl));
        zs.core_b = in.core_b ? in.core_b[z] : 0;
        zs.core_e = in.core_e ? in.core_e[z] : (int32_t)zs.tpl.size();
    }
    const int64_t raw_base = sliced ? in.code_base : (nr ? in.read_off[0] : 0);
    h_pack_.ensure((size_t)nr + 1);
    parallel_for(nr, host_threads, [&](int r) {
        DevRead& rd = reads_[r];
        std::memset(&rd, 0, sizeof(rd));
        const int64_t I = rlen(r);
        const int64_t soff = sliced ? in.code_start[r] : in.read_off[r] + (in.rstart ? in.rstart[r] : 0);
        const uint8_t* src = in.codes + soff;
        const int64_t stride = (coffs[r + 1] - coffs[r]) / 2;
        rd.code_off = coffs[r];
        rd.code_stride = (int32_t)std::max<int64_t>(stride, 16);
        rd.I = (int32_t)I;
        rd.strand = in.strand[r];
        rd.ts = in.tstart[r];
        rd.te = in.tend[r];
        rd.active = (rd.te > rd.ts) ? 1 : 0;
        rd.first_code = I >= 1 ? src[0] : 0;
        rd.last_code = I >= 1 ? src[I - 1] : 0;
        h_pack_.p[r] = PackJob{soff - raw_base, coffs[r], (int32_t)I, (int32_t)stride};   // stride 0: nothing to write
    });
    for (int z = 0; z < nz; ++z) {
        ZmwState& zs = zstate_[z];
        for (int r = zs.read_begin; r < zs.read_end; ++r) {
            reads_[r].zmw = z;
            if (reads_[r].active) ++zs.n_mapped;
        }
    }
    d_rowcode_.ensure((size_t)code_total + 64, budget_);
    d_pack_.ensure((size_t)nr + 1);
    // transitions and z-score moments per ZMW; the windows of one draft share their SNR: computed once per run of equal SNRs
    h_trans_.ensure((size_t)nz * 36 * 4);
    d_trans_.ensure((size_t)nz * 36 * 4, budget_);
    zs_mom_.resize((size_t)nz);
    {
        std::vector<int> leader((size_t)nz), leaders;
        for (int z = 0; z < nz; ++z) {
            leader[z] = (z > 0 && std::memcmp(in.snr + 4 * z, in.snr + 4 * (z - 1), 4 * sizeof(float)) == 0) ? leader[z - 1] : z;
            if (leader[z] == z) leaders.push_back(z);
        }
        parallel_for((int)leaders.size(), host_threads, [&](int k) {
            const int z = leaders[k];
            ZmwTransitions zt;
            build_zmw_transitions(model_, in.snr + 4 * z, zt);
            std::memcpy(h_trans_.p + (size_t)z * 36 * 4, zt.tr, sizeof(zt.tr));
            zscore_moments(model_, in.snr + 4 * z, zs_mom_[z]);
        });
        for (int z = 0; z < nz; ++z)
            if (leader[z] != z) {
                std::memcpy(h_trans_.p + (size_t)z * 36 * 4, h_trans_.p + (size_t)leader[z] * 36 * 4, 36 * 4 * sizeof(float));
                zs_mom_[z] = zs_mom_[leader[z]];
            }
    }
    span_begin(&stats.ms_h2d);
    const uint8_t* d_raw = in.d_codes;
    if (!d_raw) {     // stand-alone Polish Stage: the engine uploads the read codes itself
        const int64_t raw_total = nr ? in.read_off[nr] - raw_base : 0;
        h_rawcodes_.ensure((size_t)raw_total + 64);
        d_rawcodes_.ensure((size_t)raw_total + 64);
        const int64_t piece = 1 << 20;
        parallel_for((int)((raw_total + piece - 1) / piece), host_threads, [&](int k) {
            const int64_t b = k * piece, e = std::min<int64_t>(raw_total, b + piece);
            std::memcpy(h_rawcodes_.p + b, in.codes + raw_base + b, (size_t)(e - b));
        });
        if (raw_total) CCS_CUDA(cudaMemcpyAsync(d_rawcodes_.p, h_rawcodes_.p, (size_t)raw_total, cudaMemcpyHostToDevice, stream_));
        stats.h2d_bytes += raw_total;
        d_raw = d_rawcodes_.p;
    }
    CCS_CUDA(cudaMemcpyAsync(d_pack_.p, h_pack_.p, sizeof(PackJob) * nr, cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_trans_.p, h_trans_.p, (size_t)nz * 36 * 4 * sizeof(float), cudaMemcpyHostToDevice, stream_));
    launch_pack_rowcodes(d_pack_.p, nr, d_raw, d_rowcode_.p, stream_);
    ++stats.n_pack;
    stats.h2d_bytes += (int64_t)sizeof(PackJob) * nr + (int64_t)nz * 36 * 4 * 4;
    // template capacities: room for the template to grow during polishing
    for (int z = 0; z < nz; ++z) {
        const int J = (int)zstate_[z].tpl.size();
        tpl_cap_[z] = ((J + std::max(in.growth_min ? in.growth_min[z] : 512, J / 8)) + 15) & ~15;
    }
    // fixed column slots per read (span + the template's growth room), so unchanged ZMWs keep their bands
    col_base_.assign(nr, 0); col_cap_.assign(nr, 0);
    {
        int64_t cb = 0;
        for (int z = 0; z < nz; ++z) {
            const int J = (int)zstate_[z].tpl.size();
            const int grow = tpl_cap_[z] - J;
            zstate_[z].dirty = true;
            for (int r = zstate_[z].read_begin; r < zstate_[z].read_end; ++r) {
                col_base_[r] = cb;
                if (reads_[r].active) { col_cap_[r] = (reads_[r].te - reads_[r].ts) + grow; cb += col_cap_[r]; }
            }
        }
    }
    upload_templates_and_reads();
    span_end();
}

// (Re)derive DevRead / DevZmw / template buffer / column offsets from the host state and push
// them.  Called after load() and after every round that edited templates.
void ArrowEngine::upload_templates_and_reads() {
    HostPhase hp("polish.upload templates+reads");
    const int nz = (int)zstate_.size(), nr = (int)reads_.size();
    int64_t toff = 0, cols = 0, drows = 0;
    for (int z = 0; z < nz; ++z) {
        ZmwState& zs = zstate_[z];
        const int J = (int)zs.tpl.size();
        if (J > tpl_cap_[z]) throw OomError("template outgrew its capacity");
        DevZmw& dz = zmws_[z];
        dz.read_begin = zs.read_begin; dz.read_end = zs.read_end;
        dz.fwd_off = (int32_t)toff; dz.rev_off = (int32_t)(toff + tpl_cap_[z]);
        dz.J = J; dz.delta_off = drows; dz.pad_ = 0;
        toff += 2 * (int64_t)tpl_cap_[z];
        drows += tpl_cap_[z] + 2;   // fixed slot: delta rows survive template edits (re-indexed, not moved)
    }
    if (toff > 0x7fffffffll) throw OomError("template buffer exceeds 2 GiB; use smaller batches");
    h_tpl_.ensure((size_t)toff + 16);
    if (toff < 64) std::memset(h_tpl_.p + toff, 0, (size_t)(64 - toff));   // idle lanes of the kernels read the first bytes of the buffer
    // column offsets (serial prefix), then the per-ZMW copies in parallel
    for (int z = 0; z < nz; ++z) {
        const ZmwState& zs = zstate_[z];
        const int J = zmws_[z].J;
        for (int r = zs.read_begin; r < zs.read_end; ++r) {
            DevRead& rd = reads_[r];
            const int len = rd.te - rd.ts;
            if (rd.active && (len < 2 || rd.ts < 0 || rd.te > J || rd.I < 2)) { rd.active = 0; status_[r] = 2; }
            rd.J = rd.active ? len : 0;
            if (col_cap_[r] < rd.J) throw OomError("column capacity of a read exceeded");   // cannot happen: growth is bounded by the template capacity
            rd.col_off = col_base_[r];
            cols = std::max<int64_t>(cols, col_base_[r] + col_cap_[r]);
        }
    }
    parallel_for(nz, host_threads, [&](int z) {
        const ZmwState& zs = zstate_[z];
        const DevZmw& dz = zmws_[z];
        const int J = dz.J;
        if (!zs.dirty) return;       // template and spans unchanged since the last upload: its bytes in h_tpl_ still hold
        uint8_t* f = h_tpl_.p + dz.fwd_off;
        uint8_t* rv = h_tpl_.p + dz.rev_off;
        // template bytes carry the trinucleotide index 16*t[j-2] + 4*t[j-1] + t[j] (the base is byte & 3): the fill
        // kernels use the byte as the row index of their folded factor tables
        unsigned hf = 0, hr = 0;
        for (int j = 0; j < J; ++j) {
            hf = ((hf << 2) | zs.tpl[j]) & 63u;
            hr = ((hr << 2) | (3u - zs.tpl[J - 1 - j])) & 63u;
            f[j] = (uint8_t)hf;
            rv[j] = (uint8_t)hr;
        }
        for (int r = zs.read_begin; r < zs.read_end; ++r) {
            DevRead& rd = reads_[r];
            rd.tpl_off = rd.strand ? dz.rev_off + (J - rd.te) : dz.fwd_off + rd.ts;
        }
    });
    total_cols_ = cols;
    total_delta_rows_ = drows;
    // reads to (re)fill: every active read of a ZMW whose template changed since its last fill
    // fill work list: groups of 16 slots (one CTA each) holding reads of ONE ZMW, longest template first, padded
    // with -1; groups ordered longest first
    order_.clear();
    {
        std::vector<std::pair<int, int>> groups;   // (longest J, first slot in `slots`)
        std::vector<int32_t> slots, zr;
        // The windows of one draft share their transition table (same SNR), which is all a fill CTA has per "ZMW": their
        // reads are packed into groups together, so a 12-read window does not leave a quarter of its CTA idle.
        for (int z = 0; z < nz;) {
            int ze = z + 1;
            if (!zmw_group_.empty()) while (ze < nz && zmw_group_[ze] == zmw_group_[z]) ++ze;
            zr.clear();
            for (int w = z; w < ze; ++w) {
                if (!zstate_[w].dirty) continue;
                for (int r = zstate_[w].read_begin; r < zstate_[w].read_end; ++r) if (reads_[r].active) zr.push_back(r);
            }
            std::stable_sort(zr.begin(), zr.end(), [&](int a, int b) { return reads_[a].J > reads_[b].J; });
            for (size_t k = 0; k < zr.size(); k += 16) {
                groups.emplace_back(reads_[zr[k]].J, (int)slots.size());
                for (size_t x = 0; x < 16; ++x) slots.push_back(k + x < zr.size() ? zr[k + x] : -1);
            }
            z = ze;
        }
        std::stable_sort(groups.begin(), groups.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first > b.first; });
        order_.reserve(slots.size());
        for (const auto& gp : groups) order_.insert(order_.end(), slots.begin() + gp.second, slots.begin() + gp.second + 16);
    }

    d_tpl_.ensure((size_t)toff + 16, budget_);
    d_reads_.ensure((size_t)nr + 1);
    d_zmws_.ensure((size_t)nz + 1);
    d_order_.ensure(order_.size() + 16);
    d_status_.ensure((size_t)nr + 1);
    d_ll_alpha_.ensure((size_t)nr + 1); d_ll_beta_.ensure((size_t)nr + 1); d_base_ll_.ensure((size_t)nr + 1);
    d_alpha_.ensure((size_t)(cols + 2) * 32, budget_);
    d_beta_.ensure((size_t)(cols + 2) * 32, budget_);
    d_colinfo_.ensure((size_t)cols + 8, budget_);      // the staged scoring kernel copies 16-byte aligned pieces: slack at the end
    d_bexp_.ensure((size_t)cols + 8, budget_);
    h_reads_.ensure((size_t)nr + 1); h_zmws_.ensure((size_t)nz + 1); h_order_.ensure(order_.size() + 16);
    h_status_.ensure((size_t)nr + 1);
    std::memcpy(h_reads_.p, reads_.data(), sizeof(DevRead) * nr);
    std::memcpy(h_zmws_.p, zmws_.data(), sizeof(DevZmw) * nz);
    std::memcpy(h_order_.p, order_.data(), sizeof(int32_t) * order_.size());
    std::memcpy(h_status_.p, status_.data(), sizeof(int32_t) * nr);
    CCS_CUDA(cudaMemcpyAsync(d_tpl_.p, h_tpl_.p, (size_t)toff, cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_reads_.p, h_reads_.p, sizeof(DevRead) * nr, cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_zmws_.p, h_zmws_.p, sizeof(DevZmw) * nz, cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_order_.p, h_order_.p, sizeof(int32_t) * order_.size(), cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_status_.p, h_status_.p, sizeof(int32_t) * nr, cudaMemcpyHostToDevice, stream_));
    stats.h2d_bytes += toff + (int64_t)sizeof(DevRead) * nr + (int64_t)sizeof(DevZmw) * nz + 8ll * nr;
}

void ArrowEngine::fill() {
    CCS_CUDA(cudaSetDevice(device_));
    const ArrowBatchView V = view();
    const int n = (int)order_.size();
    int64_t cells = 0, in_bytes = 0;
    for (int r : order_) if (r >= 0) { cells += 32ll * (reads_[r].J - 1); in_bytes += reads_[r].I + reads_[r].J; }
    cudaStream_t fs = stream_;
    if (stream_hi_) {      // hand over to the high-priority stream for the two fill kernels, then back
        CCS_CUDA(cudaEventRecord(ev_order_, stream_));
        CCS_CUDA(cudaStreamWaitEvent(stream_hi_, ev_order_, 0));
        fs = stream_hi_;
    }
    span_begin(&stats.ms_fill_alpha, 4 * cells + 8 * (cells / 32) + in_bytes, &stats.top_fill_alpha_bytes, &stats.top_fill_alpha_ms, fs);
    launch_fill_alpha(V, d_order_.p, n, fs);
    span_end(fs);
    span_begin(&stats.ms_fill_beta, 4 * cells + 8 * (cells / 32) + in_bytes, &stats.top_fill_beta_bytes, &stats.top_fill_beta_ms, fs);
    launch_fill_beta(V, d_order_.p, n, fs);
    span_end(fs);
    if (stream_hi_) {
        CCS_CUDA(cudaEventRecord(ev_order_, stream_hi_));
        CCS_CUDA(cudaStreamWaitEvent(stream_, ev_order_, 0));
    }
    CCS_CUDA(cudaGetLastError());
    for (auto& zs : zstate_) zs.dirty = false;
    ++stats.n_fill_alpha; ++stats.n_fill_beta;
    stats.cells_fill += cells;
    stats.bytes_fill_alpha += 4 * cells + 8 * (cells / 32) + in_bytes;
    stats.bytes_fill_beta += 4 * cells + 8 * (cells / 32) + in_bytes;   // 4 B exponent out + 4 B band start in
    sync_statuses();
}

void ArrowEngine::sync_statuses() {
    HostPhase hp("polish.fill (wait)");
    const int nr = (int)reads_.size();
    CCS_CUDA(cudaMemcpyAsync(h_status_.p, d_status_.p, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, stream_));
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
    resolve_spans();
    stats.d2h_bytes += 4ll * nr;
    for (int r = 0; r < nr; ++r) {
        if (!reads_[r].active) continue;
        status_[r] = h_status_.p[r];
        if (status_[r] != 0) {
            reads_[r].active = 0;   // dropped for the rest of the polish (Integrator semantics)
            if (stats.n_score > score_mark_) zstate_[reads_[r].zmw].stale_scores = true;
        }
    }
}

// Integrator::AddRead: a read whose log-likelihood against the draft lies more than |min_zscore| standard deviations
// below its expectation is dropped (CCS_READ_POOR_ZSCORE).  Checked once, after the first fill.  The expectation sums
// per-position moments over the read's template slice: prefix sums over the ZMW's forward / reverse template.
void ArrowEngine::zscore_filter(double min_zscore) {
    if (!(min_zscore > -1e29)) return;
    HostPhase hp("polish.zscore");
    const int nr = (int)reads_.size(), nz = (int)zstate_.size();
    h_ll_.ensure((size_t)2 * nr + 2);
    CCS_CUDA(cudaMemcpyAsync(h_ll_.p, d_ll_alpha_.p, sizeof(double) * nr, cudaMemcpyDeviceToHost, stream_));
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
    stats.d2h_bytes += 8ll * nr;
    // expectation and variance of every live read's log-likelihood on its template slice
    std::vector<double> r_mean((size_t)nr, 0.0), r_var((size_t)nr, 0.0);
    std::vector<uint8_t> live((size_t)nr, 0);
    parallel_for(nz, host_threads, [&](int z) {
        const ZmwState& zs = zstate_[z];
        const ZscoreMoments& M = zs_mom_[z];
        const int J = (int)zs.tpl.size();
        bool any = false;
        for (int r = zs.read_begin; r < zs.read_end; ++r) any |= reads_[r].active && status_[r] == 0;
        if (!any || J < 2) return;
        // pm[s][j] / pv[s][j] = sums over template positions 1..j of strand s (0 forward, 1 reverse complement)
        std::vector<double> pm[2], pv[2];
        for (int s = 0; s < 2; ++s) {
            pm[s].assign((size_t)J, 0.0); pv[s].assign((size_t)J, 0.0);
            for (int j = 1; j < J; ++j) {
                const int a = s ? 3 - zs.tpl[J - j] : zs.tpl[j - 1], b = s ? 3 - zs.tpl[J - 1 - j] : zs.tpl[j];
                pm[s][j] = pm[s][j - 1] + M.mean[4 * a + b];
                pv[s][j] = pv[s][j - 1] + M.var[4 * a + b];
            }
        }
        for (int r = zs.read_begin; r < zs.read_end; ++r) {
            const DevRead& rd = reads_[r];
            if (!rd.active || status_[r] != 0) continue;
            const int s = rd.strand ? 1 : 0;
            const int b0 = s ? J - rd.te : rd.ts, b1 = s ? J - rd.ts : rd.te;      // slice [b0, b1) on strand s
            const int t0 = s ? 3 - zs.tpl[J - 1 - b0] : zs.tpl[b0];
            r_mean[r] = M.first_mean[t0] + (pm[s][b1 - 1] - pm[s][b0]);
            r_var[r] = M.first_var[t0] + (pv[s][b1 - 1] - pv[s][b0]);
            live[r] = 1;
        }
    });
    // a read is judged on the sum over its group (the windows of one subread), in read order
    int dropped = 0;
    if (read_group_.empty()) {
        for (int r = 0; r < nr; ++r)
            if (live[r] && (h_ll_.p[r] - r_mean[r]) / std::sqrt(r_var[r]) < min_zscore) { reads_[r].active = 0; status_[r] = 5; ++dropped; }
    } else {
        std::vector<double> g_ll((size_t)n_groups_, 0.0), g_mean((size_t)n_groups_, 0.0), g_var((size_t)n_groups_, 0.0);
        for (int r = 0; r < nr; ++r)
            if (live[r]) { const int g = read_group_[r]; g_ll[g] += h_ll_.p[r]; g_mean[g] += r_mean[r]; g_var[g] += r_var[r]; }
        for (int r = 0; r < nr; ++r) {
            if (!live[r]) continue;
            const int g = read_group_[r];
            if ((g_ll[g] - g_mean[g]) / std::sqrt(g_var[g]) < min_zscore) { reads_[r].active = 0; status_[r] = 5; ++dropped; }
        }
    }
    if (dropped > 0) {      // the scoring kernel reads the statuses on the device
        std::memcpy(h_status_.p, status_.data(), sizeof(int32_t) * nr);
        CCS_CUDA(cudaMemcpyAsync(d_status_.p, h_status_.p, sizeof(int32_t) * nr, cudaMemcpyHostToDevice, stream_));
        stats.h2d_bytes += 4ll * nr;
    }
}

void ArrowEngine::read_lls(double* ll_alpha, double* ll_beta, int32_t* status) {
    const int nr = (int)reads_.size();
    h_ll_.ensure((size_t)2 * nr + 2);
    CCS_CUDA(cudaMemcpyAsync(h_ll_.p, d_ll_alpha_.p, sizeof(double) * nr, cudaMemcpyDeviceToHost, stream_));
    CCS_CUDA(cudaMemcpyAsync(h_ll_.p + nr, d_ll_beta_.p, sizeof(double) * nr, cudaMemcpyDeviceToHost, stream_));
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
    stats.d2h_bytes += 16ll * nr;
    for (int r = 0; r < nr; ++r) {
        const bool filled = status_[r] == 0 || status_[r] == 1 || status_[r] == 3;
        if (ll_alpha) ll_alpha[r] = (filled && status_[r] != 3) ? h_ll_.p[r] : NAN;
        if (ll_beta) ll_beta[r] = (filled && status_[r] != 3) ? h_ll_.p[nr + r] : NAN;
        if (status) status[r] = status_[r];
    }
}

void ArrowEngine::dump_pair(int r, float* alpha, float* beta, int32_t* start, int32_t* aexp, int32_t* bexp) {
    const DevRead& rd = reads_[r];
    const int J = rd.J;
    if (J <= 0) return;
    std::vector<ColInfo> ci(J);
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
    if (alpha) CCS_CUDA(cudaMemcpy(alpha, d_alpha_.p + rd.col_off * 32, sizeof(float) * 32 * J, cudaMemcpyDeviceToHost));
    if (beta) CCS_CUDA(cudaMemcpy(beta, d_beta_.p + rd.col_off * 32, sizeof(float) * 32 * J, cudaMemcpyDeviceToHost));
    CCS_CUDA(cudaMemcpy(ci.data(), d_colinfo_.p + rd.col_off, sizeof(ColInfo) * J, cudaMemcpyDeviceToHost));
    if (bexp) CCS_CUDA(cudaMemcpy(bexp, d_bexp_.p + rd.col_off, sizeof(int32_t) * J, cudaMemcpyDeviceToHost));
    for (int j = 0; j < J; ++j) { if (start) start[j] = ci[j].start; if (aexp) aexp[j] = ci[j].cumexp; }
}

// ---------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------
// Ranges start at multiples of 16 in the flattened work list: the 16 octets of a scoring CTA then belong to one ZMW
// (the kernels skip the padding items past a range's end).
static inline int64_t pad16(int64_t n) { return (n + 15) & ~15ll; }

void ArrowEngine::score_ranges(const std::vector<ScoreRange>& ranges, int64_t n_items) {
    if (ranges.empty() || n_items == 0) return;
    d_ranges_.ensure(ranges.size());
    h_ranges_.ensure(ranges.size());
    std::memcpy(h_ranges_.p, ranges.data(), sizeof(ScoreRange) * ranges.size());
    d_delta_.ensure((size_t)(total_delta_rows_ + 2) * kDeltaStride, budget_);
    CCS_CUDA(cudaMemcpyAsync(d_ranges_.p, h_ranges_.p, sizeof(ScoreRange) * ranges.size(), cudaMemcpyHostToDevice, stream_));
    stats.h2d_bytes += (int64_t)sizeof(ScoreRange) * ranges.size();
    const ArrowBatchView V = view();
    // algorithmic bytes (SURVEY.md 8d B_score): five 128-byte band columns in + 64 bytes of partial sums out per
    // (covering read, position)
    int64_t sbytes = 0;
    if (timing_enabled)
        for (const ScoreRange& rg : ranges) {
            const ZmwState& zs = zstate_[rg.zmw];
            for (int r = zs.read_begin; r < zs.read_end; ++r)
                if (reads_[r].active)
                    sbytes += 704ll * std::max(0, std::min(rg.p_end, reads_[r].te) - std::max(rg.p_begin, reads_[r].ts));
        }
    stats.bytes_score += sbytes;
    span_begin(&stats.ms_score, sbytes, &stats.top_score_bytes, &stats.top_score_ms);
    launch_score(V, d_ranges_.p, (int)ranges.size(), n_items, d_delta_.p, stream_, generic_score, score_variant);
    span_end();
    CCS_CUDA(cudaGetLastError());
    ++stats.n_score;
    for (const ScoreRange& r : ranges) stats.score_items += r.p_end - r.p_begin;
    n_ranges_ = (int)ranges.size();
    n_range_items_ = n_items;
}

void ArrowEngine::score_all_positions() {
    std::vector<ScoreRange> ranges;
    int64_t first = 0;
    for (int z = 0; z < (int)zstate_.size(); ++z) {
        const ZmwState& zs = zstate_[z];
        if (zs.failed || zs.tpl.empty()) continue;
        ranges.push_back(ScoreRange{z, 0, (int32_t)zs.tpl.size(), 0, first});
        first += pad16((int64_t)zs.tpl.size());
    }
    score_ranges(ranges, first);
}

int64_t ArrowEngine::pick(std::vector<Candidate>& out) {
    out.clear();
    if (n_ranges_ == 0 || n_range_items_ == 0) return 0;
    const ArrowBatchView V = view();
    size_t cap = std::max<size_t>(d_cand_.cap, (size_t)std::min<int64_t>(n_range_items_, 1 << 20));
    for (int attempt = 0; attempt < 2; ++attempt) {
        d_cand_.ensure(cap);
        CCS_CUDA(cudaMemsetAsync(d_counter_.p, 0, sizeof(int32_t), stream_));
        span_begin(&stats.ms_pick);
        launch_pick(V, d_ranges_.p, n_ranges_, n_range_items_, d_delta_.p, d_cand_.p, (int)d_cand_.cap, d_counter_.p, stream_);
        span_end();
        // one synchronisation per round: the counter and a first slab of candidates come back together (late rounds have
        // a handful of candidates; the first round of a big chunk fetches the rest with a second copy)
        const size_t slab = std::min<size_t>(d_cand_.cap, 16384);
        h_cand_.ensure(slab + 1);
        CCS_CUDA(cudaMemcpyAsync(h_counter_.p, d_counter_.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        CCS_CUDA(cudaMemcpyAsync(h_cand_.p, d_cand_.p, sizeof(Candidate) * slab, cudaMemcpyDeviceToHost, stream_));
        CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
        resolve_spans();
        ++stats.n_pick;
        const int64_t n = h_counter_.p[0];
        if ((size_t)n <= d_cand_.cap) {
            if ((size_t)n > slab) {
                h_cand_.ensure((size_t)n + 1);      // (grow-only; the slab is copied again with the rest)
                CCS_CUDA(cudaMemcpyAsync(h_cand_.p, d_cand_.p, sizeof(Candidate) * n, cudaMemcpyDeviceToHost, stream_));
                CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
            }
            stats.d2h_bytes += (int64_t)sizeof(Candidate) * std::max<int64_t>(n, (int64_t)slab);
            out.assign(h_cand_.p, h_cand_.p + n);
            return n;
        }
        cap = (size_t)n + 1024;   // overflow: grow and retry once
    }
    throw CudaError("candidate buffer overflow");
}

void ArrowEngine::download_delta(int z, double* out) {
    const DevZmw& dz = zmws_[z];
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
    std::vector<double> tmp((size_t)dz.J * kDeltaStride);
    CCS_CUDA(cudaMemcpy(tmp.data(), d_delta_.p + dz.delta_off * kDeltaStride, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
    for (int p = 0; p < dz.J; ++p)
        for (int s = 0; s < 9; ++s)
            out[(size_t)p * 9 + s] = tmp[(size_t)p * kDeltaStride + s] + (s >= 5 ? tmp[(size_t)p * kDeltaStride + s + 4] : 0.0);
}

// ---------------------------------------------------------------------------------------------
// Polish(): iterate {score -> keep improving mutations -> greedy best with separation ->
// apply -> refill} until no mutation improves any ZMW of the batch
// (docs/how-does-ccs-work.md:96-101).
// ---------------------------------------------------------------------------------------------
void ArrowEngine::polish(const PolishParams& pp) {
    ab_tol_ = pp.ab_mismatch_tol;
    const int nz = (int)zstate_.size();
    score_mark_ = stats.n_score;
    CCS_CUDA(cudaEventRecord(evA_, stream_));   // inputs are resident: load() has been issued on this stream
    fill();
    zscore_filter(pp.min_zscore);
    auto check_usable = [&](int z) {
        ZmwState& zs = zstate_[z];
        int act = 0;
        for (int r = zs.read_begin; r < zs.read_end; ++r) act += reads_[r].active;
        if (act == 0 || act < pp.min_active_fraction * zs.n_mapped) { zs.failed = true; zs.done = true; }
    };
    for (int z = 0; z < nz; ++z) check_usable(z);
    if (!zmw_group_.empty())      // a draft with an unusable window is not polished at all
        for (int z = 0; z < nz;) {
            int e = z;
            bool bad = false;
            while (e < nz && zmw_group_[e] == zmw_group_[z]) { bad |= zstate_[e].failed; ++e; }
            if (bad) for (int k = z; k < e; ++k) zstate_[k].done = true;
            z = e;
        }

    std::vector<Candidate> cands;
    for (int it = 0; it < pp.max_iterations; ++it) {
        // ranges to score this round
        std::vector<ScoreRange> ranges;
        int64_t first = 0;
        for (int z = 0; z < nz; ++z) {
            ZmwState& zs = zstate_[z];
            if (zs.done) continue;
            const int J = (int)zs.tpl.size();
            zs.iterations = it + 1;
            // mutations are tested within test_margin bases of the window core only (the whole template when it is one window)
            const int tb = std::max(0, zs.core_b - pp.test_margin), te = std::min(J, zs.core_e + pp.test_margin);
            if (it == 0) {
                if (te > tb) {
                    ranges.push_back(ScoreRange{z, tb, te, 0, first});
                    first += pad16(te - tb);
                    zs.n_tested += count_canonical(zs.tpl, tb, te);
                }
            } else {
                // union of +-neighborhood around the last-applied sites
                std::vector<std::pair<int, int>> iv;
                for (int s : zs.sites) {
                    const int b = std::max(tb, s - pp.neighborhood), e = std::min(te, s + pp.neighborhood + 1);
                    if (e > b) iv.push_back({b, e});
                }
                std::sort(iv.begin(), iv.end());
                int cb = -1, ce = -1;
                auto flush = [&]() {
                    if (ce > cb) {
                        const int e = std::min(ce, J);
                        if (e > cb) {
                            ranges.push_back(ScoreRange{z, cb, e, 0, first});
                            first += pad16(e - cb);
                            zs.n_tested += count_canonical(zs.tpl, cb, e);
                        }
                    }
                };
                for (auto& v : iv) {
                    if (v.first > ce) { flush(); cb = v.first; ce = v.second; }
                    else ce = std::max(ce, v.second);
                }
                flush();
            }
        }
        if (ranges.empty()) break;
        ++stats.rounds;
        score_ranges(ranges, first);
        if (reuse_scores)      // these positions now carry delta-LLs of the current template
            for (const ScoreRange& rg : ranges) {
                std::vector<uint8_t>& st = zstate_[rg.zmw].stale;
                if (st.size() != zstate_[rg.zmw].tpl.size()) st.assign(zstate_[rg.zmw].tpl.size(), 1);
                std::fill(st.begin() + rg.p_begin, st.begin() + rg.p_end, (uint8_t)0);
            }
        { HostPhase hp("polish.score+pick (wait)"); pick(cands); }
        HostPhase hp_round("polish.round select+apply");
        // group candidates per ZMW
        std::vector<std::vector<HostMutation>> per(nz);
        for (const Candidate& c : cands) per[c.zmw].push_back(HostMutation{c.type, c.pos, c.base, c.score});
        std::atomic<int> applied_flag(0);
        parallel_for(nz, host_threads, [&](int z) {
            ZmwState& zs = zstate_[z];
            if (zs.done) return;
            auto& sc = per[z];
            if (sc.empty()) { zs.converged = true; zs.done = true; return; }
            // BestMutations: greedy by score, chosen sites >= separation apart (polish_host.h)
            const int J = (int)zs.tpl.size();
            std::vector<HostMutation> best = select_best_mutations(sc, J, pp.separation);
            std::vector<uint8_t> next = apply_to_template(zs.tpl, best);
            uint64_t h = tpl_hash(next);
            if (std::find(zs.seen.begin(), zs.seen.end(), h) != zs.seen.end()) {   // cycle guard
                best.assign(1, sc.front());
                next = apply_to_template(zs.tpl, best);
                h = tpl_hash(next);
            }
            zs.seen.push_back(h);
            zs.sites.clear();
            int off = 0;
            for (const auto& m : best) {
                zs.sites.push_back(m.pos + off);
                off += m.type == 1 ? 1 : (m.type == 2 ? -1 : 0);
            }
            if ((int)next.size() > tpl_cap_[z]) {
                // the template keeps growing past its reserved room (only seen on junk ZMWs whose reads do not
                // agree): stop refining it; reported as NON_CONVERGENT
                zs.done = true;
                return;
            }
            zs.n_applied += (int)best.size();
            zs.J_before = (int32_t)zs.tpl.size();
            zs.remap_sites = zs.sites;
            zs.remap_shifts.clear();
            { int acc = 0; for (const auto& m : best) { acc += m.type == 1 ? 1 : (m.type == 2 ? -1 : 0); zs.remap_shifts.push_back(acc); } }
            // span bookkeeping of the ZMW's reads (Integrator::ApplyMutations)
            for (int r = zs.read_begin; r < zs.read_end; ++r) {
                DevRead& rd = reads_[r];
                int ds = 0, de = 0;
                for (const auto& m : best) {
                    if (m.type == 1) { if (m.pos <= rd.ts) ++ds; if (m.pos < rd.te) ++de; }
                    else if (m.type == 2) { if (m.pos < rd.ts) --ds; if (m.pos < rd.te) --de; }
                }
                rd.ts += ds; rd.te += de;
            }
            {   // the window core's borders move like a read's tstart: an insertion at a border goes to its left
                int db = 0, de = 0;
                for (const auto& m : best) {
                    if (m.type == 1) { if (m.pos <= zs.core_b) ++db; if (m.pos <= zs.core_e) ++de; }
                    else if (m.type == 2) { if (m.pos < zs.core_b) --db; if (m.pos < zs.core_e) --de; }
                }
                zs.core_b += db; zs.core_e += de;
            }
            if (reuse_scores) {
                // carry the per-position "stale" marks over to the new coordinates (same shifts as remap_deltas) and
                // mark everything within qv_halo of this round's edits
                const int Jn = (int)next.size(), Jo = (int)zs.tpl.size();
                std::vector<uint8_t> ns((size_t)Jn, 1);
                size_t e = 0;
                int shift = 0;
                for (int q = 0; q < Jn; ++q) {
                    while (e < zs.remap_sites.size() && zs.remap_sites[e] <= q) shift = zs.remap_shifts[e++];
                    const int old = q - shift;
                    if (old >= 0 && old < Jo && zs.stale.size() == (size_t)Jo) ns[q] = zs.stale[old];
                }
                for (int sx : zs.sites)
                    std::fill(ns.begin() + std::max(0, sx - qv_halo), ns.begin() + std::min(Jn, sx + qv_halo + 1), (uint8_t)1);
                zs.stale.swap(ns);
            }
            zs.tpl.swap(next);
            zs.dirty = true;
            applied_flag.store(1, std::memory_order_relaxed);
        });
        const bool any_applied = applied_flag.load() != 0;
        hp_round.~HostPhase();
        new (&hp_round) HostPhase("polish.round tail");
        if (!any_applied) break;
        upload_templates_and_reads();
        if (reuse_scores) remap_deltas();
        fill();
        for (int z = 0; z < nz; ++z) if (!zstate_[z].done) check_usable(z);
    }
    { HostPhase hp("polish.qv (wait)"); consensus_qvs(); }
    CCS_CUDA(cudaEventRecord(evB_, stream_));
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
    CCS_CUDA(cudaEventSynchronize(evB_));
    float ms = 0;
    cudaEventElapsedTime(&ms, evA_, evB_);
    stats.ms_resident += ms;
    stats.n_zmws += nz;
}

// Re-index the stored delta rows of the ZMWs edited in this round (see arrow_remap_delta_kernel).
void ArrowEngine::remap_deltas() {
    std::vector<RemapJob> jobs;
    std::vector<int32_t> sites, shifts;
    int64_t scratch = 0;
    for (size_t z = 0; z < zstate_.size(); ++z) {
        ZmwState& zs = zstate_[z];
        if (!zs.dirty || zs.remap_sites.empty()) continue;
        RemapJob j;
        j.delta_off = zmws_[z].delta_off; j.scratch_off = scratch; j.J_old = zs.J_before; j.J_new = (int32_t)zs.tpl.size();
        j.site_off = (int32_t)sites.size(); j.n_sites = (int32_t)zs.remap_sites.size();
        sites.insert(sites.end(), zs.remap_sites.begin(), zs.remap_sites.end());
        shifts.insert(shifts.end(), zs.remap_shifts.begin(), zs.remap_shifts.end());
        scratch += j.J_new + 2;
        jobs.push_back(j);
        zs.remap_sites.clear(); zs.remap_shifts.clear();
    }
    if (jobs.empty() || !d_delta_.p) return;
    d_delta_scratch_.ensure((size_t)(scratch + 2) * kDeltaStride);
    d_remap_jobs_.ensure(jobs.size()); d_remap_sites_.ensure(sites.size() + 1); d_remap_shifts_.ensure(shifts.size() + 1);
    // staged through pinned buffers of the engine: no synchronisation here (the next round's remap comes after this
    // round's fill and pick synchronisations, so the buffers are free again by then)
    h_remap_jobs_.ensure(jobs.size()); h_remap_ints_.ensure(sites.size() + shifts.size() + 2);
    std::memcpy(h_remap_jobs_.p, jobs.data(), sizeof(RemapJob) * jobs.size());
    std::memcpy(h_remap_ints_.p, sites.data(), 4 * sites.size());
    std::memcpy(h_remap_ints_.p + sites.size(), shifts.data(), 4 * shifts.size());
    CCS_CUDA(cudaMemcpyAsync(d_remap_jobs_.p, h_remap_jobs_.p, sizeof(RemapJob) * jobs.size(), cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_remap_sites_.p, h_remap_ints_.p, 4 * sites.size(), cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_remap_shifts_.p, h_remap_ints_.p + sites.size(), 4 * shifts.size(), cudaMemcpyHostToDevice, stream_));
    launch_remap_delta(d_remap_jobs_.p, (int)jobs.size(), d_remap_sites_.p, d_remap_shifts_.p, d_delta_.p,
                       d_delta_scratch_.p, stream_);
}

int64_t ArrowEngine::count_canonical(const std::vector<uint8_t>& t, int b, int e) const {
    return count_canonical_mutations(t, b, e);
}

void ArrowEngine::consensus_qvs() {
    const int nz = (int)zstate_.size();
    std::vector<ScoreRange> ranges, rescoring;
    int64_t first = 0, first_rs = 0;
    for (int z = 0; z < nz; ++z) {
        const ZmwState& zs = zstate_[z];
        if (zs.failed || zs.tpl.empty()) continue;
        ranges.push_back(ScoreRange{z, 0, (int32_t)zs.tpl.size(), 0, first});
        first += pad16((int64_t)zs.tpl.size());
        // With reuse_scores only the positions whose stored delta-LLs predate an edit within qv_halo of them (or that
        // were never scored) are scored again; a ZMW that lost a read after scoring began is scored again in full
        const int J = (int)zs.tpl.size();
        // only the window core goes into the result: QVs outside it are never looked at
        const int qb = std::max(0, zs.core_b), qe = std::min(J, zs.core_e);
        if (!reuse_scores || zs.stale_scores || zs.stale.size() != (size_t)J) {
            if (qe > qb) {
                rescoring.push_back(ScoreRange{z, qb, qe, 0, first_rs});
                first_rs += pad16((int64_t)(qe - qb));
            }
        } else {
            int p = qb;
            while (p < qe) {
                if (!zs.stale[p]) { ++p; continue; }
                int e = p + 1, last = p;                       // merge runs separated by short clean gaps
                while (e < qe && e - last <= 8) { if (zs.stale[e]) last = e; ++e; }
                rescoring.push_back(ScoreRange{z, p, last + 1, 0, first_rs});
                first_rs += pad16((int64_t)(last + 1 - p));
                p = last + 2;                                  // ranges of one ZMW must not touch
            }
        }
    }
    for (int z = 0; z < nz; ++z) qv_[z].clear();
    if (ranges.empty()) return;
    if (!rescoring.empty()) score_ranges(rescoring, first_rs);
    // the QV kernel walks every position of every live ZMW
    d_ranges_.ensure(ranges.size());
    h_ranges_qv_.ensure(ranges.size());
    std::memcpy(h_ranges_qv_.p, ranges.data(), sizeof(ScoreRange) * ranges.size());
    CCS_CUDA(cudaMemcpyAsync(d_ranges_.p, h_ranges_qv_.p, sizeof(ScoreRange) * ranges.size(), cudaMemcpyHostToDevice, stream_));
    d_delta_.ensure((size_t)(total_delta_rows_ + 2) * kDeltaStride, budget_);
    d_qv_.ensure((size_t)total_delta_rows_ + 16, budget_);
    h_qv_.ensure((size_t)total_delta_rows_ + 16);
    const ArrowBatchView V = view();
    span_begin(&stats.ms_qv);
    launch_qv(V, d_delta_.p, d_qv_.p, first, d_ranges_.p, (int)ranges.size(), stream_);
    span_end();
    CCS_CUDA(cudaMemcpyAsync(h_qv_.p, d_qv_.p, (size_t)total_delta_rows_, cudaMemcpyDeviceToHost, stream_));
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
    resolve_spans();
    CCS_CUDA(cudaGetLastError());
    ++stats.n_qv;
    stats.d2h_bytes += total_delta_rows_;
    for (int z = 0; z < nz; ++z) {
        const ZmwState& zs = zstate_[z];
        if (zs.failed || zs.tpl.empty()) continue;
        const DevZmw& dz = zmws_[z];
        qv_[z].assign(h_qv_.p + dz.delta_off, h_qv_.p + dz.delta_off + dz.J);
    }
}

}  // namespace ccs
