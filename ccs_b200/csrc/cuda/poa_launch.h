// Host-callable launchers of the Draft Stage kernels (poa_align.cu, poa_graph.cu).
#pragma once
#include <cuda_runtime.h>
#include "poa_device.h"

namespace ccs {

// Runs the DP (one warp per task) and the traceback on `stream`.  hrows == NULL: every task is linear (mapping);
// steps are written for DAG tasks only; grid (optional): linear tasks with grid_off >= 0 record the read position at
// every kWindowGrid-th template base of their path.
void launch_poa_align(const PoaTask* tasks, int n_tasks, const PoaGraphView& G, const uint8_t* drafts, const uint8_t* codes,
                      const uint8_t* rev_flags, int32_t* lo, uint8_t* moves, int32_t* hrows,
                      PoaStep* steps, PoaResult* results, cudaStream_t stream, int32_t* grid = nullptr);

// seeds[k]: {graph, codes_off, n} of the seed read of graph k
void launch_poa_graph_init(const PoaGraphView& G, const PoaTask* seeds, int n_graphs, const uint8_t* codes, cudaStream_t stream);

// CommitAdd of every placed task (score >= n) into its graph
void launch_poa_commit(const PoaGraphView& G, const PoaTask* tasks, int n_tasks, const uint8_t* codes, const uint8_t* rev,
                       const PoaStep* steps, const PoaResult* results, int32_t* scratch, cudaStream_t stream);

// FindConsensus of graphs[0..n): bases to draft[voff ..), length to draft_len[graph]
void launch_poa_consensus(const PoaGraphView& G, const int32_t* graphs, int n_graphs, const int64_t* scratch_off,
                          int32_t* scratch, uint8_t* draft, int32_t* draft_len, cudaStream_t stream);

// k-mer orientation votes: rev[read.rev_idx] = reverse complement shares more sampled 11-mers with the reference
cudaError_t launch_poa_kmer_vote(const PoaVoteJob* jobs, int n_jobs, int max_ref_len, const PoaVoteRead* reads,
                                 const uint8_t* codes, const uint8_t* drafts, uint8_t* rev, cudaStream_t stream);

}  // namespace ccs
