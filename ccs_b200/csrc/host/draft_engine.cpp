// Draft Stage host engine -- see draft_engine.h.
#include "draft_engine.h"
#include "draft_host.h"
#include "parallel.h"
#include "../cuda/poa_launch.h"
#include "../../../include/ccsgpu.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace ccs {

namespace {

// Carves typed arrays out of one block, so that all descriptors of a pass go up in a single copy.
struct Carver {
    size_t off = 0;
    template <class T>
    size_t take(size_t n) {
        off = (off + 15) & ~size_t(15);
        const size_t o = off;
        off += n * sizeof(T);
        return o;
    }
};

template <class T> T* at(uint8_t* base, size_t off) { return reinterpret_cast<T*>(base + off); }

}  // namespace

DraftEngine::DraftEngine(int device, size_t scratch_budget_bytes) : device_(device), budget_(scratch_budget_bytes) {
    CCS_CUDA(cudaSetDevice(device_));
    {   // the Draft Stage kernels are serial chains over the graph (latency bound, few warps): high priority, so that
        // their CTAs get SM slots ahead of the other lanes' short-lived scoring CTAs
        int lo = 0, hi = 0;
        const char* e = std::getenv("CCS_B200_PRIO");
        if (!(e && e[0] == '0') && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess && hi < lo)
            CCS_CUDA(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, hi));
        else
            CCS_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    }
    CCS_CUDA(make_blocking_event(&ev_sync_));
    if (budget_ == 0) budget_ = 12ull << 30;
}

DraftEngine::~DraftEngine() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    for (cudaEvent_t e : ev_pool_) cudaEventDestroy(e);
    if (ev_sync_) cudaEventDestroy(ev_sync_);
    if (stream_) cudaStreamDestroy(stream_);
}

void DraftEngine::release_buffers() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    d_codes_.release(); d_desc_.release(); d_rev_.release(); d_moves_.release(); d_draft_.release(); d_meta_.release();
    d_pred0_.release(); d_predx_.release(); d_rank_.release(); d_order_.release(); d_lo_.release();
    d_hrows_.release(); d_scratch_.release(); d_draft_len_.release(); d_steps_.release(); d_results_.release();
    d_grid_.release(); d_col_.release();
}

void DraftEngine::span(double* acc, int64_t bytes, int64_t* top_bytes, double* top_ms) {
    while (ev_used_ + 2 > ev_pool_.size()) {
        cudaEvent_t e;
        CCS_CUDA(cudaEventCreate(&e));
        ev_pool_.push_back(e);
    }
    Span sp{ev_pool_[ev_used_], ev_pool_[ev_used_ + 1], acc, bytes, top_bytes, top_ms};
    ev_used_ += 2;
    CCS_CUDA(cudaEventRecord(sp.a, stream_));
    spans_.push_back(sp);
}

void DraftEngine::span_end() { CCS_CUDA(cudaEventRecord(spans_.back().b, stream_)); }

void DraftEngine::resolve_spans() {   // call after a stream synchronisation
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

void DraftEngine::run(const DraftInput& in, const DraftParams& dp, DraftOutput& out) {
    CCS_CUDA(cudaSetDevice(device_));
    const int nz = in.n_zmws, nr = in.n_reads;
    out.status.assign(nz, CCS_ZMW_EXCEPTION_THROWN);
    out.draft.assign(nz, {});
    out.maps.assign(nr, ReadMap());
    out.keep.assign(nr, 0);
    out.grid.clear();
    out.grid_off.assign(nr, -1);
    if (nz == 0) return;
    std::vector<int32_t> lens(nr);
    for (int r = 0; r < nr; ++r) lens[r] = (int32_t)(in.read_off[r + 1] - in.read_off[r]);

    // ---- the batch's read codes go up once and stay resident (the Polish Stage of this lane reads them too) ----
    const int64_t code_base = nr ? in.read_off[0] : 0;
    const int64_t code_total = nr ? in.read_off[nr] - code_base : 0;
    {
        HostPhase hp("draft.upload codes");
        h_codes_.ensure((size_t)code_total + 64);
        d_codes_.ensure((size_t)code_total + 64);
        const int64_t piece = 1 << 20;
        const int np = (int)((code_total + piece - 1) / piece);
        parallel_for(np, host_threads, [&](int k) {
            const int64_t b = k * piece, e = std::min<int64_t>(code_total, b + piece);
            std::memcpy(h_codes_.p + b, in.codes + code_base + b, (size_t)(e - b));
        });
        if (code_total) CCS_CUDA(cudaMemcpyAsync(d_codes_.p, h_codes_.p, (size_t)code_total, cudaMemcpyHostToDevice, stream_));
        d_rev_.ensure((size_t)nr + 64);
        h_rev_.ensure((size_t)nr + 64);
        CCS_CUDA(cudaMemsetAsync(d_rev_.p, 0, (size_t)nr + 64, stream_));
        stats.h2d_bytes += code_total;
    }

    // ---- a1: filtering and POA read selection (host: a sort of a dozen lengths per ZMW) ----------------------
    std::vector<Zw> work(nz);
    { HostPhase hp("draft.a1 filter");
    parallel_for(nz, host_threads, [&](int z) {
        const int r0 = in.zmw_read_off[z], r1 = in.zmw_read_off[z + 1];
        const int n = r1 - r0;
        if (n == 0) { out.status[z] = CCS_ZMW_NO_SUBREADS; return; }
        const float* s = in.snr + 4 * z;
        if (std::min(std::min(s[0], s[1]), std::min(s[2], s[3])) < dp.min_snr) { out.status[z] = CCS_ZMW_POOR_SNR; return; }
        const int nfull = filter_reads(lens.data() + r0, in.cx + r0, n, dp.top_passes, out.keep.data() + r0);
        if (nfull < dp.min_passes) { out.status[z] = CCS_ZMW_TOO_FEW_PASSES; return; }
        Zw& w = work[z];
        const int max_poa = std::max(1, std::min(dp.max_poa_reads, kPoaMaxReads));
        for (int r = r0; r < r1 && (int)w.poa_reads.size() < max_poa; ++r)
            if (out.keep[r] && (in.cx[r] & 3) == 3) w.poa_reads.push_back(r);
        // no full-length read (only reachable with --min-passes 0): no draft can be generated
        if (w.poa_reads.empty()) { out.status[z] = CCS_ZMW_DRAFT_FAILURE; return; }
        if (lens[w.poa_reads[0]] > kPoaMaxRefLen) { out.status[z] = CCS_ZMW_TOO_LONG; return; }
        w.alive = true;
    });
    }

    // ---- a2-a5 on the device, in chunks of ZMWs that fit the scratch budget -------------------------------------
    auto run_alive = [&]() {
        std::vector<int> zlist;
        int64_t bytes = 0;
        auto flush = [&]() {
            if (!zlist.empty()) poa_chunk(in, dp, out, lens, work, zlist);
            zlist.clear();
            bytes = 0;
        };
        for (int z = 0; z < nz; ++z) {
            const Zw& w = work[z];
            if (!w.alive) continue;
            int64_t cap = lens[w.poa_reads[0]] + 8, nmax = 0;
            for (size_t k = 1; k < w.poa_reads.size(); ++k) {
                cap += poa_new_vertex_bound(lens[w.poa_reads[k]]);
                nmax = std::max<int64_t>(nmax, lens[w.poa_reads[k]]);
            }
            const int64_t need = cap * 440 + nmax * 32;
            if (!zlist.empty() && bytes + need > (int64_t)budget_ * 7 / 10) flush();   // the buffers grow with 25 % slack
            zlist.push_back(z);
            bytes += need;
        }
        flush();
    };
    run_alive();      // draft generator 0: SparsePoa over the first full-length reads, seeded by the first one

    // ---- draft cascade (docs/faq/accuracy-vs-passes.md:41-46): ZMWs whose draft could not be generated or that too few
    //      subreads map back to get one more try with the robust generator -- seeded by the full-length read closest to
    //      the median length, over up to 2 * max_poa_reads - 1 reads in order of closeness
    bool any = false;
    for (int z = 0; z < nz; ++z) {
        Zw& w = work[z];
        w.alive = false;
        if (w.poa_reads.empty()) continue;
        if (out.status[z] != CCS_ZMW_DRAFT_FAILURE && out.status[z] != CCS_ZMW_TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT) continue;
        const int r0 = in.zmw_read_off[z], r1 = in.zmw_read_off[z + 1];
        std::vector<int32_t> full, sl;
        for (int r = r0; r < r1; ++r) if (out.keep[r] && (in.cx[r] & 3) == 3) { full.push_back(r); sl.push_back(lens[r]); }
        std::sort(sl.begin(), sl.end());
        const int med = sl.empty() ? 0 : sl[sl.size() / 2];
        full.erase(std::remove(full.begin(), full.end(), w.poa_reads[0]), full.end());   // not the seed that just failed
        std::stable_sort(full.begin(), full.end(), [&](int a, int b) { return std::abs(lens[a] - med) < std::abs(lens[b] - med); });
        const size_t cap = (size_t)std::max(1, std::min(2 * dp.max_poa_reads - 1, kPoaMaxReads));
        if (full.size() > cap) full.resize(cap);
        if (full.empty() || lens[full[0]] > kPoaMaxRefLen) continue;   // nothing new to try
        w.poa_reads = full;
        w.alive = true;
        out.draft[z].clear();
        for (int r = r0; r < r1; ++r) { out.maps[r] = ReadMap(); out.grid_off[r] = -1; }
        any = true;
    }
    if (any) run_alive();
}

// One chunk of ZMWs through SparsePoa (rounds of align -> traceback -> CommitAdd), FindConsensus and the mapping.
void DraftEngine::poa_chunk(const DraftInput& in, const DraftParams& dp, DraftOutput& out, const std::vector<int32_t>& lens,
                            std::vector<Zw>& work, const std::vector<int>& zlist) {
    const int ng = (int)zlist.size();
    const int64_t code_base = in.read_off[0];
    auto coff = [&](int r) { return in.read_off[r] - code_base; };
    // ---- layout ------------------------------------------------------------------------------------------------
    std::vector<int64_t> voff(ng + 1, 0), soff(ng + 1, 0), stoff(ng + 1, 0);
    int max_rounds = 0, max_ref = 0;
    int64_t n_vote_reads = 0;
    for (int g = 0; g < ng; ++g) {
        const Zw& w = work[zlist[g]];
        int64_t cap = lens[w.poa_reads[0]] + 8, nmax = 0;
        for (size_t k = 1; k < w.poa_reads.size(); ++k) {
            cap += poa_new_vertex_bound(lens[w.poa_reads[k]]);
            nmax = std::max<int64_t>(nmax, lens[w.poa_reads[k]]);
        }
        cap = (cap + 15) & ~15ll;                          // rows stay 16-byte aligned in every pool
        voff[g + 1] = voff[g] + cap;
        soff[g + 1] = soff[g] + 5 * nmax + 12 * cap + 16;     // CommitAdd: 5 n + V + 1; FindConsensus: 12 V
        stoff[g + 1] = stoff[g] + nmax;
        max_rounds = std::max(max_rounds, (int)w.poa_reads.size() - 1);
        max_ref = std::max(max_ref, lens[w.poa_reads[0]]);
        n_vote_reads += (int64_t)w.poa_reads.size() - 1;
    }
    const int64_t pool = voff[ng];
    std::vector<std::vector<int>> round_graphs(max_rounds);
    for (int g = 0; g < ng; ++g)
        for (int k = 1; k < (int)work[zlist[g]].poa_reads.size(); ++k) round_graphs[k - 1].push_back(g);

    Carver cv;
    const size_t o_hdr = cv.take<PoaGraphHdr>(ng), o_seed = cv.take<PoaTask>(ng), o_graphs = cv.take<int32_t>(ng);
    const size_t o_soff = cv.take<int64_t>(ng), o_jobs = cv.take<PoaVoteJob>(ng), o_vreads = cv.take<PoaVoteRead>((size_t)n_vote_reads + 1);
    std::vector<size_t> o_tasks(max_rounds);
    for (int k = 0; k < max_rounds; ++k) o_tasks[k] = cv.take<PoaTask>(round_graphs[k].size());
    const size_t desc_bytes = cv.off + 64;
    h_desc_.ensure(desc_bytes);
    d_desc_.ensure(desc_bytes);
    uint8_t* hb = h_desc_.p;
    { HostPhase hp("draft.a2 descriptors");
    int64_t vr = 0;
    for (int g = 0; g < ng; ++g) {
        const Zw& w = work[zlist[g]];
        const int sr = w.poa_reads[0];
        PoaGraphHdr& H = at<PoaGraphHdr>(hb, o_hdr)[g];
        std::memset(&H, 0, sizeof(H));
        H.voff = voff[g]; H.cap = (int32_t)(voff[g + 1] - voff[g]);
        PoaTask& S = at<PoaTask>(hb, o_seed)[g];
        std::memset(&S, 0, sizeof(S));
        S.codes_off = coff(sr); S.n = lens[sr]; S.graph = g; S.rev_idx = sr;
        at<int32_t>(hb, o_graphs)[g] = g;
        at<int64_t>(hb, o_soff)[g] = soff[g];
        PoaVoteJob& J = at<PoaVoteJob>(hb, o_jobs)[g];
        J.ref_off = coff(sr); J.ref_len = lens[sr]; J.ref_is_codes = 1; J.read_begin = (int32_t)vr;
        for (size_t k = 1; k < w.poa_reads.size(); ++k) {
            const int r = w.poa_reads[k];
            at<PoaVoteRead>(hb, o_vreads)[vr++] = PoaVoteRead{coff(r), lens[r], r};
        }
        J.read_end = (int32_t)vr;
    }
    for (int k = 0; k < max_rounds; ++k)
        for (size_t x = 0; x < round_graphs[k].size(); ++x) {
            const int g = round_graphs[k][x];
            const int r = work[zlist[g]].poa_reads[k + 1];
            PoaTask& T = at<PoaTask>(hb, o_tasks[k])[x];
            std::memset(&T, 0, sizeof(T));
            T.codes_off = coff(r); T.row_off = voff[g]; T.step_off = stoff[g]; T.n = lens[r]; T.graph = g;
            T.rev_idx = r; T.scratch_off = soff[g];
        }
    }
    d_meta_.ensure((size_t)pool + 16); d_pred0_.ensure((size_t)pool + 16); d_predx_.ensure((size_t)pool * 7 + 16);
    d_rank_.ensure((size_t)pool + 16); d_order_.ensure((size_t)pool * 2 + 16); d_col_.ensure((size_t)pool + 16);
    d_lo_.ensure((size_t)pool + 16);
    d_moves_.ensure((size_t)pool * kPoaBand + 16); d_hrows_.ensure((size_t)pool * kPoaBand + 16);
    d_scratch_.ensure((size_t)soff[ng] + 16); d_steps_.ensure((size_t)stoff[ng] + 16);
    d_results_.ensure((size_t)ng + 1); d_draft_.ensure((size_t)pool + 16); d_draft_len_.ensure((size_t)ng + 1);
    h_draft_.ensure((size_t)pool + 16); h_draft_len_.ensure((size_t)ng + 1);
    CCS_CUDA(cudaMemcpyAsync(d_desc_.p, h_desc_.p, desc_bytes, cudaMemcpyHostToDevice, stream_));
    stats.h2d_bytes += (int64_t)desc_bytes;
    uint8_t* db = d_desc_.p;
    PoaGraphView G;
    G.hdr = at<PoaGraphHdr>(db, o_hdr); G.meta = d_meta_.p; G.pred0 = d_pred0_.p; G.predx = d_predx_.p; G.rank = d_rank_.p;
    G.order[0] = d_order_.p; G.order[1] = d_order_.p + pool;
    G.col = d_col_.p;

    // ---- a3 (seeding): orientation of the POA reads against the seed; a2: seed chains + SparsePoa rounds ----------
    span(&stats.ms_graph);
    CCS_CUDA(launch_poa_kmer_vote(at<PoaVoteJob>(db, o_jobs), ng, max_ref, at<PoaVoteRead>(db, o_vreads), d_codes_.p,
                                  d_draft_.p, d_rev_.p, stream_));
    launch_poa_graph_init(G, at<PoaTask>(db, o_seed), ng, d_codes_.p, stream_);
    span_end();
    stats.n_graph_launches += 2;
    for (int k = 0; k < max_rounds; ++k) {
        const int nt = (int)round_graphs[k].size();
        if (nt == 0) continue;
        const PoaTask* tk = at<PoaTask>(db, o_tasks[k]);
        int64_t rbytes = 0;
        for (int g : round_graphs[k]) {
            // rows of round k: the graph has grown by an unknown (device-side) number of vertices; count the seed length
            const int64_t rows = lens[work[zlist[g]].poa_reads[0]];
            stats.rows += rows;
            rbytes += rows * (kPoaBand + 4 + 4 * kPoaBand) + lens[work[zlist[g]].poa_reads[k + 1]] + rows;
        }
        stats.bytes_align += rbytes;
        span(&stats.ms_align, rbytes, &stats.top_align_bytes, &stats.top_align_ms);
        launch_poa_align(tk, nt, G, d_draft_.p, d_codes_.p, d_rev_.p, d_lo_.p, d_moves_.p, d_hrows_.p,
                         d_steps_.p, d_results_.p, stream_);
        span_end();
        span(&stats.ms_graph);
        launch_poa_commit(G, tk, nt, d_codes_.p, d_rev_.p, d_steps_.p, d_results_.p, d_scratch_.p, stream_);
        span_end();
        stats.n_align_launches += 2; stats.n_graph_launches += 1; stats.n_tasks += nt;
    }
    // ---- a4: consensus -------------------------------------------------------------------------------------------
    span(&stats.ms_graph);
    launch_poa_consensus(G, at<int32_t>(db, o_graphs), ng, at<int64_t>(db, o_soff), d_scratch_.p, d_draft_.p,
                         d_draft_len_.p, stream_);
    span_end();
    stats.n_graph_launches += 1;
    CCS_CUDA(cudaMemcpyAsync(h_draft_len_.p, d_draft_len_.p, sizeof(int32_t) * ng, cudaMemcpyDeviceToHost, stream_));
    CCS_CUDA(cudaMemcpyAsync(h_draft_.p, d_draft_.p, (size_t)pool, cudaMemcpyDeviceToHost, stream_));
    stats.d2h_bytes += pool + 4ll * ng;
    { HostPhase hp("draft.a2-a4 gpu (wait)");
    CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
    CCS_CUDA(cudaGetLastError());
    resolve_spans();
    }
    // length gates
    std::vector<int> live;      // graph indices that go on to the mapping
    for (int g = 0; g < ng; ++g) {
        const int z = zlist[g];
        const int J = h_draft_len_.p[g];
        out.draft[z].assign(h_draft_.p + voff[g], h_draft_.p + voff[g] + J);
        if (J == 0) out.status[z] = CCS_ZMW_DRAFT_FAILURE;
        else if (J < dp.min_length) out.status[z] = CCS_ZMW_TOO_SHORT;
        else if (dp.max_length > 0 && J > dp.max_length) out.status[z] = CCS_ZMW_TOO_LONG;
        else { live.push_back(g); continue; }
        work[z].alive = false;
    }

    // ---- a5: subread -> draft mapping of every kept read (orientation vote against the draft, then the same
    //      aligner on the linear template), in passes that fit the scratch budget -------------------------------
    size_t li = 0;
    while (li < live.size()) {
        std::vector<int> task_read, task_g;
        int64_t rows = 0;
        size_t lj = li;
        int max_J = 0;
        for (; lj < live.size(); ++lj) {
            const int g = live[lj], z = zlist[g];
            const int J = h_draft_len_.p[g];
            int nk = 0;
            for (int r = in.zmw_read_off[z]; r < in.zmw_read_off[z + 1]; ++r) nk += out.keep[r];
            if (lj > li && (rows + (int64_t)nk * J) * (kPoaBand + 4) > (int64_t)budget_) break;
            for (int r = in.zmw_read_off[z]; r < in.zmw_read_off[z + 1]; ++r)
                if (out.keep[r]) { task_read.push_back(r); task_g.push_back(g); }
            rows += (int64_t)nk * J;
            max_J = std::max(max_J, J);
        }
        const int nt = (int)task_read.size(), nj = (int)(lj - li);
        int64_t mbytes = 0;
        Carver c2;
        const size_t o_j2 = c2.take<PoaVoteJob>(nj), o_v2 = c2.take<PoaVoteRead>((size_t)nt + 1), o_t2 = c2.take<PoaTask>((size_t)nt + 1);
        const size_t bytes2 = c2.off + 64;
        h_desc_.ensure(bytes2);
        d_desc_.ensure(bytes2);
        hb = h_desc_.p; db = d_desc_.p;
        int64_t grid_total = 0;
        {
            int64_t ro = 0;
            int k = 0;
            mbytes = 0;
            for (size_t x = li; x < lj; ++x) {
                const int g = live[x];
                const int J = h_draft_len_.p[g];
                PoaVoteJob& Jb = at<PoaVoteJob>(hb, o_j2)[x - li];
                Jb.ref_off = voff[g]; Jb.ref_len = J; Jb.ref_is_codes = 0; Jb.read_begin = k;
                for (; k < nt && task_g[k] == g; ++k) {
                    const int r = task_read[k];
                    at<PoaVoteRead>(hb, o_v2)[k] = PoaVoteRead{coff(r), lens[r], r};
                    PoaTask& T = at<PoaTask>(hb, o_t2)[k];
                    std::memset(&T, 0, sizeof(T));
                    T.codes_off = coff(r); T.row_off = ro; T.tpl_off = voff[g]; T.n = lens[r]; T.graph = -1; T.V = J;
                    T.rev_idx = r;
                    T.grid_off = grid_total;
                    grid_total += J / kWindowGrid + 1;
                    ro += J;
                    mbytes += (int64_t)J * (kPoaBand + 4) + lens[r] + J;
                }
                Jb.read_end = k;
            }
        }
        d_lo_.ensure((size_t)rows + 16); d_moves_.ensure((size_t)rows * kPoaBand + 16);
        d_results_.ensure((size_t)nt + 1); h_results_.ensure((size_t)nt + 1);
        d_grid_.ensure((size_t)grid_total + 16); h_grid_.ensure((size_t)grid_total + 16);
        CCS_CUDA(cudaMemsetAsync(d_grid_.p, 0xff, sizeof(int32_t) * (size_t)grid_total, stream_));
        CCS_CUDA(cudaMemcpyAsync(d_desc_.p, h_desc_.p, bytes2, cudaMemcpyHostToDevice, stream_));
        stats.h2d_bytes += (int64_t)bytes2;
        PoaGraphView G0 = G;     // linear tasks never touch the graph arrays
        span(&stats.ms_graph);
        CCS_CUDA(launch_poa_kmer_vote(at<PoaVoteJob>(db, o_j2), nj, max_J, at<PoaVoteRead>(db, o_v2), d_codes_.p, d_draft_.p,
                                      d_rev_.p, stream_));
        span_end();
        stats.bytes_map += mbytes;
        span(&stats.ms_map, mbytes, &stats.top_map_bytes, &stats.top_map_ms);
        launch_poa_align(at<PoaTask>(db, o_t2), nt, G0, d_draft_.p, d_codes_.p, d_rev_.p, d_lo_.p, d_moves_.p,
                         nullptr, nullptr, d_results_.p, stream_, d_grid_.p);
        span_end();
        stats.n_graph_launches += 1; stats.n_align_launches += 2; stats.n_tasks += nt; stats.rows += rows;
        CCS_CUDA(cudaMemcpyAsync(h_results_.p, d_results_.p, sizeof(PoaResult) * nt, cudaMemcpyDeviceToHost, stream_));
        CCS_CUDA(cudaMemcpyAsync(h_grid_.p, d_grid_.p, sizeof(int32_t) * (size_t)grid_total, cudaMemcpyDeviceToHost, stream_));
        stats.d2h_bytes += 4 * grid_total;
        CCS_CUDA(cudaMemcpyAsync(h_rev_.p, d_rev_.p, (size_t)in.n_reads, cudaMemcpyDeviceToHost, stream_));
        stats.d2h_bytes += (int64_t)sizeof(PoaResult) * nt + in.n_reads;
        { HostPhase hp("draft.a5 gpu map (wait)");
        CCS_CUDA(stream_sync_blocking(stream_, ev_sync_));
        CCS_CUDA(cudaGetLastError());
        resolve_spans();
        }
        const int64_t grid_base = (int64_t)out.grid.size();
        out.grid.insert(out.grid.end(), h_grid_.p, h_grid_.p + grid_total);
        for (int k = 0; k < nt; ++k) {
            const PoaResult& pr = h_results_.p[k];
            const int r = task_read[k];
            ReadMap& m = out.maps[r];
            out.grid_off[r] = grid_base + at<PoaTask>(hb, o_t2)[k].grid_off;
            m.score = pr.score;
            m.strand = h_rev_.p[r];
            if (pr.first_t < 0) continue;
            m.tstart = pr.first_t; m.tend = pr.last_t + 1;
            m.rstart = pr.first_i; m.rend = pr.last_i + 1;
            m.mapped = pr.score >= lens[r];
            if (m.strand) { const int rs = lens[r] - m.rend, re = lens[r] - m.rstart; m.rstart = rs; m.rend = re; }
            if (m.mapped && (m.tend - m.tstart < 2 || m.rend - m.rstart < 2)) m.mapped = 0;
        }
        for (size_t x = li; x < lj; ++x) {
            const int z = zlist[live[x]];
            int mapped_full = 0, mapped = 0, kept = 0;
            for (int r = in.zmw_read_off[z]; r < in.zmw_read_off[z + 1]; ++r) {
                if (!out.keep[r]) continue;
                ++kept;
                if (out.maps[r].mapped) { ++mapped; if ((in.cx[r] & 3) == 3) ++mapped_full; }
            }
            // fewer than --min-passes full-length reads, or not more than half of the subreads, map back to the draft
            // (docs/faq/accuracy-vs-passes.md:31-39)
            out.status[z] = (mapped_full < dp.min_passes || 2 * mapped <= kept) ? CCS_ZMW_TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT
                                                                                : CCS_ZMW_SUCCESS;
        }
        li = lj;
    }
}

}  // namespace ccs
