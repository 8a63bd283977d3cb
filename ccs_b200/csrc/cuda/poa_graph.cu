// Draft Stage graph kernels on the device-resident partial-order graph: seed chain, CommitAdd, FindConsensus and
// the k-mer orientation vote, one CTA per graph (PoaGraph::CommitAdd / FindConsensus, SdpRangeFinder seeding --
// SURVEY.md 8a rows a2-a4; /root/reference/docs/how-does-ccs-work.md:34-47).  The logic lives in
// poa_graph_ops.cuh (phases of index-parallel loops); these kernels bind it to a CTA.  All integer work:
// bit-exact against the oracle.  Bytes moved are a few tens of bytes per vertex per round -- the kernels are
// latency-, not bandwidth-, limited and run concurrently with the other lanes' DP kernels.
#include <cuda_runtime.h>
#include "poa_graph_ops.cuh"
#include "poa_launch.h"

namespace ccs {

namespace {

constexpr int kGraphThreads = 256;

struct CtaExec {
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return (int)blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ uint32_t cas(uint32_t* p, uint32_t c, uint32_t v) const { return atomicCAS(p, c, v); }
};

// seeds[g] = task describing the seed read of graph g (codes_off, n; orientation forward)
__global__ void __launch_bounds__(kGraphThreads) poa_graph_init_kernel(const PoaGraphView G, const PoaTask* __restrict__ seeds,
                                                                       const int n_graphs, const uint8_t* __restrict__ codes) {
    const int g = blockIdx.x;
    if (g >= n_graphs) return;
    CtaExec x;
    const PoaTask T = seeds[g];
    PoaReadAcc R{codes + T.codes_off, T.n, 0};
    poa_graph_init(x, G, T.graph, R);
}

__global__ void __launch_bounds__(kGraphThreads) poa_commit_kernel(const PoaGraphView G, const PoaTask* __restrict__ tasks,
                                                                   const int n_tasks, const uint8_t* __restrict__ codes,
                                                                   const uint8_t* __restrict__ rev,
                                                                   const PoaStep* __restrict__ steps,
                                                                   const PoaResult* __restrict__ results,
                                                                   int32_t* __restrict__ scratch) {
    __shared__ int32_t sm[kGraphThreads];
    const int k = blockIdx.x;
    if (k >= n_tasks) return;
    CtaExec x;
    const PoaTask T = tasks[k];
    const PoaResult r = results[k];
    if (!(r.score >= T.n && r.path_len > 0)) return;       // not placed: the read is not threaded (TryAddRead)
    PoaReadAcc R{codes + T.codes_off, T.n, (int)rev[T.rev_idx]};
    poa_graph_commit(x, G, T.graph, steps + T.step_off, r.path_len, R, scratch + T.scratch_off, sm);
}

// graphs[k] = graph slot; the consensus of graph g goes to draft[voff ..), its length to draft_len[g]
__global__ void __launch_bounds__(kGraphThreads) poa_consensus_kernel(const PoaGraphView G, const int32_t* __restrict__ graphs,
                                                                      const int n_graphs,
                                                                      const int64_t* __restrict__ scratch_off,
                                                                      int32_t* __restrict__ scratch, uint8_t* __restrict__ draft,
                                                                      int32_t* __restrict__ draft_len) {
    __shared__ int32_t sh[kPoaConsShared];
    const int k = blockIdx.x;
    if (k >= n_graphs) return;
    CtaExec x;
    const int g = graphs[k];
    poa_graph_consensus(x, G, g, scratch + scratch_off[g], draft + G.hdr[g].voff, draft_len + g, sh);
}

// One CTA per vote job: hash set of the reference's sampled 11-mers in shared memory, then one warp per read.
__global__ void __launch_bounds__(kGraphThreads) poa_kmer_vote_kernel(const PoaVoteJob* __restrict__ jobs, const int n_jobs,
                                                                      const PoaVoteRead* __restrict__ reads,
                                                                      const uint8_t* __restrict__ codes,
                                                                      const uint8_t* __restrict__ drafts,
                                                                      uint8_t* __restrict__ rev) {
    extern __shared__ uint32_t s_tab[];
    const int j = blockIdx.x;
    if (j >= n_jobs) return;
    CtaExec x;
    const PoaVoteJob J = jobs[j];
    const int cap = poa_kmer_table_cap(J.ref_len);
    for (int i = threadIdx.x; i < cap; i += blockDim.x) s_tab[i] = 0u;
    __syncthreads();
    PoaBaseAcc ref{J.ref_is_codes ? codes + J.ref_off : drafts + J.ref_off, J.ref_is_codes};
    poa_kmer_build(x, s_tab, cap, ref, J.ref_len);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = J.read_begin + warp; r < J.read_end; r += nwarps) {
        const PoaVoteRead rd = reads[r];
        const int n = min(rd.n, kPoaVoteBases);
        const int npos = n - (kPoaKmer - 1);               // k-mer end positions kPoaKmer-1 .. n-1
        int f = 0, c = 0;
        if (npos > 0) {
            const int per = (npos + 31) / 32;
            const int b = kPoaKmer - 1 + lane * per, e = min(b + per, n);
            poa_kmer_count(s_tab, cap, codes + rd.codes_off, b, e, f, c);
        }
        f = __reduce_add_sync(0xffffffffu, f);
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) rev[rd.rev_idx] = (c > f) ? 1 : 0;
    }
}

}  // namespace

void launch_poa_graph_init(const PoaGraphView& G, const PoaTask* seeds, int n_graphs, const uint8_t* codes, cudaStream_t stream) {
    if (n_graphs <= 0) return;
    poa_graph_init_kernel<<<n_graphs, kGraphThreads, 0, stream>>>(G, seeds, n_graphs, codes);
}

void launch_poa_commit(const PoaGraphView& G, const PoaTask* tasks, int n_tasks, const uint8_t* codes, const uint8_t* rev,
                       const PoaStep* steps, const PoaResult* results, int32_t* scratch, cudaStream_t stream) {
    if (n_tasks <= 0) return;
    poa_commit_kernel<<<n_tasks, kGraphThreads, 0, stream>>>(G, tasks, n_tasks, codes, rev, steps, results, scratch);
}

void launch_poa_consensus(const PoaGraphView& G, const int32_t* graphs, int n_graphs, const int64_t* scratch_off,
                          int32_t* scratch, uint8_t* draft, int32_t* draft_len, cudaStream_t stream) {
    if (n_graphs <= 0) return;
    poa_consensus_kernel<<<n_graphs, kGraphThreads, 0, stream>>>(G, graphs, n_graphs, scratch_off, scratch, draft, draft_len);
}

cudaError_t launch_poa_kmer_vote(const PoaVoteJob* jobs, int n_jobs, int max_ref_len, const PoaVoteRead* reads,
                                 const uint8_t* codes, const uint8_t* drafts, uint8_t* rev, cudaStream_t stream) {
    if (n_jobs <= 0) return cudaSuccess;
    int cap = 256;
    while (cap * 4 < max_ref_len) cap <<= 1;
    const size_t smem = (size_t)cap * sizeof(uint32_t);
    if (smem > 48 * 1024) {    // per device, cheap: no caching across the GPUs a process may drive
        cudaError_t e = cudaFuncSetAttribute(poa_kmer_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    poa_kmer_vote_kernel<<<n_jobs, kGraphThreads, smem, stream>>>(jobs, n_jobs, reads, codes, drafts, rev);
    return cudaSuccess;
}

}  // namespace ccs
