// Draft Stage host engine -- see draft_engine.h.
#include "draft_engine.h"
#include "poa_graph.h"
#include "draft_host.h"
#include "parallel.h"
#include "../cuda/poa_launch.h"
#include "../../../include/ccsgpu.h"
#include <algorithm>
#include <cstring>

namespace ccs {

DraftEngine::DraftEngine(int device, size_t scratch_budget_bytes) : device_(device), budget_(scratch_budget_bytes) {
    CCS_CUDA(cudaSetDevice(device_));
    CCS_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    CCS_CUDA(cudaEventCreate(&ev0_));
    CCS_CUDA(cudaEventCreate(&ev1_));
    if (budget_ == 0) budget_ = 12ull << 30;
}

DraftEngine::~DraftEngine() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (stream_) cudaStreamDestroy(stream_);
}

// One GPU pass over the task list staged in the pinned buffers (already chunked to the scratch budget).
void DraftEngine::align_tasks(int nt, bool any_dag, bool want_paths, int64_t rows, int64_t path_bytes, size_t n_vbase,
                              size_t n_poff, size_t n_preds, size_t n_reads) {
    if (nt == 0) return;
    CCS_CUDA(cudaSetDevice(device_));
    d_tasks_.ensure(nt); d_vbase_.ensure(n_vbase + 16); d_reads_.ensure(n_reads + 16);
    d_poff_.ensure(n_poff + 16); d_preds_.ensure(n_preds + 16);
    d_lo_.ensure((size_t)rows + 16); d_besti_.ensure((size_t)rows + 16); d_moves_.ensure((size_t)rows * kPoaBand + 16);
    if (any_dag) d_hrows_.ensure((size_t)rows * kPoaBand + 16);
    if (want_paths) { d_paths_.ensure((size_t)path_bytes + 16); h_paths_.ensure((size_t)path_bytes + 16); }
    d_results_.ensure(nt);
    h_results_.ensure(nt);
    CCS_CUDA(cudaMemcpyAsync(d_tasks_.p, h_tasks_.p, sizeof(PoaTask) * nt, cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_vbase_.p, h_vbase_.p, n_vbase, cudaMemcpyHostToDevice, stream_));
    CCS_CUDA(cudaMemcpyAsync(d_reads_.p, h_reads_.p, n_reads, cudaMemcpyHostToDevice, stream_));
    if (n_poff) CCS_CUDA(cudaMemcpyAsync(d_poff_.p, h_poff_.p, n_poff * 4, cudaMemcpyHostToDevice, stream_));
    if (n_preds) CCS_CUDA(cudaMemcpyAsync(d_preds_.p, h_preds_.p, n_preds * 4, cudaMemcpyHostToDevice, stream_));
    stats.h2d_bytes += (int64_t)(sizeof(PoaTask) * nt + n_vbase + n_reads + 4 * (n_poff + n_preds));
    CCS_CUDA(cudaEventRecord(ev0_, stream_));
    launch_poa_align(d_tasks_.p, nt, d_vbase_.p, d_poff_.p, d_preds_.p, d_reads_.p, d_lo_.p, d_besti_.p, d_moves_.p,
                     any_dag ? d_hrows_.p : nullptr, want_paths ? d_paths_.p : nullptr, d_results_.p, stream_);
    CCS_CUDA(cudaEventRecord(ev1_, stream_));
    CCS_CUDA(cudaMemcpyAsync(h_results_.p, d_results_.p, sizeof(PoaResult) * nt, cudaMemcpyDeviceToHost, stream_));
    if (want_paths) {
        CCS_CUDA(cudaMemcpyAsync(h_paths_.p, d_paths_.p, (size_t)path_bytes, cudaMemcpyDeviceToHost, stream_));
        stats.d2h_bytes += path_bytes;
    }
    CCS_CUDA(stream_sync_blocking(stream_));
    CCS_CUDA(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0_, ev1_);
    stats.ms_align += ms;
    stats.n_align_launches += 2;
    stats.n_tasks += nt;
    stats.rows += rows;
    stats.d2h_bytes += (int64_t)sizeof(PoaResult) * nt;
    stats.bytes_align += rows * (kPoaBand + 8 + (any_dag ? 4 * kPoaBand : 0)) + (int64_t)n_reads + (int64_t)n_vbase;
}

void DraftEngine::run(const DraftInput& in, const DraftParams& dp, DraftOutput& out) {
    const int nz = in.n_zmws, nr = in.n_reads;
    out.status.assign(nz, CCS_ZMW_EXCEPTION_THROWN);
    out.draft.assign(nz, {});
    out.maps.assign(nr, ReadMap());
    out.keep.assign(nr, 0);
    std::vector<int32_t> lens(nr);
    for (int r = 0; r < nr; ++r) lens[r] = (int32_t)(in.read_off[r + 1] - in.read_off[r]);

    // ---- a1: filtering, POA read selection, seed graph, orientation votes ------------------
    struct ZmwWork {
        std::vector<int32_t> poa_reads;   // global read indices, seed first
        std::vector<uint8_t> poa_rev;     // orientation of each (vs the seed)
        std::vector<uint8_t> seed;
        HostPoaGraph graph;
        std::vector<int32_t> order;       // export of the current round
        bool alive = false;
    };
    std::vector<ZmwWork> work(nz);
    { HostPhase hp("draft.a1 filter+seed+kmer");
    parallel_for(nz, host_threads, [&](int z) {
        const int r0 = in.zmw_read_off[z], r1 = in.zmw_read_off[z + 1];
        const int n = r1 - r0;
        if (n == 0) { out.status[z] = CCS_ZMW_NO_SUBREADS; return; }
        const float* s = in.snr + 4 * z;
        if (std::min(std::min(s[0], s[1]), std::min(s[2], s[3])) < dp.min_snr) { out.status[z] = CCS_ZMW_POOR_SNR; return; }
        const int nfull = filter_reads(lens.data() + r0, in.cx + r0, n, dp.top_passes, out.keep.data() + r0);
        if (nfull < dp.min_passes) { out.status[z] = CCS_ZMW_TOO_FEW_PASSES; return; }
        ZmwWork& w = work[z];
        for (int r = r0; r < r1 && (int)w.poa_reads.size() < dp.max_poa_reads; ++r)
            if (out.keep[r] && (in.cx[r] & 3) == 3) w.poa_reads.push_back(r);
        const int sr = w.poa_reads[0];
        w.seed.resize(lens[sr]);
        orient(in.codes + in.read_off[sr], lens[sr], false, w.seed.data());
        w.graph.init(w.seed.data(), lens[sr]);
        KmerSet ks;
        ks.build(w.seed.data(), lens[sr]);
        w.poa_rev.assign(w.poa_reads.size(), 0);
        for (size_t k = 1; k < w.poa_reads.size(); ++k) {
            int64_t f, c;
            ks.count(in.codes + in.read_off[w.poa_reads[k]], std::min(lens[w.poa_reads[k]], kPoaVoteBases), f, c);
            w.poa_rev[k] = c > f;
        }
        w.alive = true;
    });
    }

    // ---- a2: SparsePoa rounds: round k aligns the k-th POA read of every ZMW on the GPU ----
    std::vector<int> task_zmw;
    const int64_t row_bytes = kPoaBand * 5 + 8;   // moves + score row + lo/best per vertex
    for (int round = 1; round < dp.max_poa_reads; ++round) {
        int z = 0;
        while (z < nz) {
            // pass 1 (serial, cheap): lay out the chunk
            task_zmw.clear();
            int64_t rows = 0, path_bytes = 0, n_preds = 0, n_reads = 0, n_poff = 0;
            std::vector<PoaTask> tl;
            for (; z < nz; ++z) {
                ZmwWork& w = work[z];
                if (!w.alive || (int)w.poa_reads.size() <= round) continue;
                const int V = w.graph.size();
                const int rd = w.poa_reads[round];
                if (!tl.empty() && (rows + V) * row_bytes > (int64_t)budget_) break;
                PoaTask t;
                t.vert_off = rows; t.poff_off = n_poff; t.pred_base = n_preds; t.read_off = n_reads;
                t.row_off = rows; t.path_off = path_bytes; t.V = V; t.n = lens[rd]; t.linear = 0; t.pad_ = 0;
                rows += V; n_poff += V + 1; n_preds += w.graph.n_edges(); n_reads += lens[rd]; path_bytes += V + lens[rd];
                tl.push_back(t);
                task_zmw.push_back(z);
            }
            const int nt = (int)tl.size();
            if (nt == 0) break;
            h_tasks_.ensure(nt); h_vbase_.ensure((size_t)rows + 16); h_poff_.ensure((size_t)n_poff + 16);
            h_preds_.ensure((size_t)n_preds + 16); h_reads_.ensure((size_t)n_reads + 16);
            std::memcpy(h_tasks_.p, tl.data(), sizeof(PoaTask) * nt);
            // pass 2 (parallel): export graphs and orient reads straight into pinned memory
            { HostPhase hp("draft.a2 export");
            parallel_for(nt, host_threads, [&](int k) {
                ZmwWork& w = work[task_zmw[k]];
                const PoaTask& t = tl[k];
                w.graph.export_topo(w.order, h_vbase_.p + t.vert_off, h_poff_.p + t.poff_off, h_preds_.p + t.pred_base);
                const int rd = w.poa_reads[round];
                orient(in.codes + in.read_off[rd], lens[rd], w.poa_rev[round], h_reads_.p + t.read_off);
            });
            }
            { HostPhase hp("draft.a2 gpu align (wait)");
            align_tasks(nt, true, true, rows, path_bytes, (size_t)rows, (size_t)n_poff, (size_t)n_preds, (size_t)n_reads);
            }
            HostPhase hp2("draft.a2 commit");
            parallel_for(nt, host_threads, [&](int k) {
                ZmwWork& w = work[task_zmw[k]];
                const PoaTask& t = tl[k];
                const PoaResult& r = h_results_.p[k];
                if (r.score >= t.n && r.path_len > 0)   // placed: CommitAdd
                    w.graph.commit(h_paths_.p + t.path_off, r.path_len, r.end_t, r.end_i, w.order,
                                   h_poff_.p + t.poff_off, h_preds_.p + t.pred_base, h_reads_.p + t.read_off);
            });
        }
    }

    // ---- a4: consensus + length gates; a3: orientation of every kept read against the draft ----
    std::vector<std::vector<uint8_t>> rev_flag(nz);
    { HostPhase hp("draft.a4 consensus+kmer");
    parallel_for(nz, host_threads, [&](int z) {
        ZmwWork& w = work[z];
        if (!w.alive) return;
        const int n = w.graph.n_reads();
        const int min_cov = n < 5 ? 1 : (n + 1) / 2 - 1;
        w.graph.consensus(min_cov, out.draft[z]);
        const int J = (int)out.draft[z].size();
        if (J == 0) { out.status[z] = CCS_ZMW_DRAFT_FAILURE; w.alive = false; }
        else if (J < dp.min_length) { out.status[z] = CCS_ZMW_TOO_SHORT; w.alive = false; }
        else if (J > dp.max_length) { out.status[z] = CCS_ZMW_TOO_LONG; w.alive = false; }
        if (!w.alive) return;
        const int r0 = in.zmw_read_off[z], r1 = in.zmw_read_off[z + 1];
        KmerSet ks;
        ks.build(out.draft[z].data(), J);
        rev_flag[z].assign(r1 - r0, 0);
        for (int r = r0; r < r1; ++r) {
            if (!out.keep[r]) continue;
            int64_t f, c;
            ks.count(in.codes + in.read_off[r], std::min(lens[r], kPoaVoteBases), f, c);
            rev_flag[z][r - r0] = c > f;
        }
    });
    }

    // ---- a5: subread -> draft mapping of every kept read (linear graphs on the same kernel) ---
    {
        int z = 0;
        std::vector<int> task_read;
        std::vector<PoaTask> tl;
        while (z < nz) {
            tl.clear(); task_read.clear();
            int64_t rows = 0, n_vbase = 0, n_reads = 0;
            for (; z < nz; ++z) {
                ZmwWork& w = work[z];
                if (!w.alive) continue;
                const int r0 = in.zmw_read_off[z], r1 = in.zmw_read_off[z + 1];
                const int J = (int)out.draft[z].size();
                int nk = 0;
                for (int r = r0; r < r1; ++r) nk += out.keep[r];
                if (!tl.empty() && (rows + (int64_t)nk * J) * (kPoaBand + 8) > (int64_t)budget_) break;
                for (int r = r0; r < r1; ++r) {
                    if (!out.keep[r]) continue;
                    PoaTask t;
                    t.vert_off = n_vbase; t.poff_off = 0; t.pred_base = 0; t.read_off = n_reads;
                    t.row_off = rows; t.path_off = 0; t.V = J; t.n = lens[r]; t.linear = 1; t.pad_ = z;
                    rows += J; n_reads += lens[r];
                    tl.push_back(t);
                    task_read.push_back(r);
                }
                n_vbase += J;
            }
            const int nt = (int)tl.size();
            if (nt == 0) break;
            h_tasks_.ensure(nt); h_vbase_.ensure((size_t)n_vbase + 16); h_reads_.ensure((size_t)n_reads + 16);
            std::memcpy(h_tasks_.p, tl.data(), sizeof(PoaTask) * nt);
            { HostPhase hp("draft.a5 map pack");
            parallel_for(nt, host_threads, [&](int k) {
                const PoaTask& t = tl[k];
                const int r = task_read[k], zz = t.pad_;
                if (k == 0 || tl[k - 1].pad_ != zz)   // first task of the ZMW copies the draft
                    std::memcpy(h_vbase_.p + t.vert_off, out.draft[zz].data(), out.draft[zz].size());
                orient(in.codes + in.read_off[r], lens[r], rev_flag[zz][r - in.zmw_read_off[zz]], h_reads_.p + t.read_off);
            });
            }
            { HostPhase hp("draft.a5 gpu map (wait)");
            align_tasks(nt, false, false, rows, 0, (size_t)n_vbase, 0, 0, (size_t)n_reads);
            }
            for (int k = 0; k < nt; ++k) {
                const PoaResult& pr = h_results_.p[k];
                ReadMap& m = out.maps[task_read[k]];
                m.score = pr.score;
                if (pr.first_t < 0) continue;
                m.tstart = pr.first_t; m.tend = pr.last_t + 1;
                m.rstart = pr.first_i; m.rend = pr.last_i + 1;
                m.mapped = pr.score >= tl[k].n;
            }
        }
    }
    parallel_for(nz, host_threads, [&](int z) {
        ZmwWork& w = work[z];
        if (!w.alive) return;
        const int r0 = in.zmw_read_off[z], r1 = in.zmw_read_off[z + 1];
        int mapped_full = 0;
        for (int r = r0; r < r1; ++r) {
            if (!out.keep[r]) continue;
            ReadMap& m = out.maps[r];
            m.strand = rev_flag[z][r - r0];
            if (m.strand) { const int rs = lens[r] - m.rend, re = lens[r] - m.rstart; m.rstart = rs; m.rend = re; }
            if (m.mapped && (m.tend - m.tstart < 2 || m.rend - m.rstart < 2)) m.mapped = 0;
            if (m.mapped && (in.cx[r] & 3) == 3) ++mapped_full;
        }
        out.status[z] = mapped_full < dp.min_passes ? CCS_ZMW_TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT : CCS_ZMW_SUCCESS;
    });
}

}  // namespace ccs
