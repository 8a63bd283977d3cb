// Device-side layout of the Draft Stage's sequence-to-DAG aligner (SparsePoa / PoaGraph::TryAddRead,
// SURVEY.md 8a rows a2-a3; plain PODs shared by poa_align.cu and the C++ host engine).
//
// A graph is shipped in topological order: base[t], pred_off[t..t+1] into preds[] (predecessor
// ranks, ascending).  The DP is a banded local alignment with one 64-cell row per vertex:
// row t covers read prefixes i in [lo[t], lo[t]+64), lo[t] = max_pred(best_i[pred]) + 1 - 32
// (clamped) -- the band follows the best cell of the predecessor rows (DESIGN.md "Draft stage").
#pragma once
#include <cstdint>

namespace ccs {

constexpr int kPoaBand = 64;
constexpr int kPoaMatch = 3, kPoaMismatch = -5, kPoaIns = -4, kPoaDel = -4;
constexpr int kPoaMaxPred = 8;
constexpr int kPoaKmer = 11;
constexpr int kPoaVoteBases = 2048;   // orientation vote: k-mers of the read's first 2048 bases (spec)

struct PoaTask {
    int64_t vert_off;     // this graph's vertices in base[]
    int64_t poff_off;     // this graph's V+1 predecessor offsets in pred_off[] (unused when linear)
    int64_t pred_base;    // this graph's predecessor list in preds[] (unused when linear)
    int64_t read_off;     // oriented read bases (0..3)
    int64_t row_off;      // first row of this task in lo[] / besti[] / moves / hrows (rows = V)
    int64_t path_off;     // traceback output (moves, end -> start), capacity V + n
    int32_t V;            // vertices
    int32_t n;            // read length
    int32_t linear;       // 1: every vertex t has the single predecessor t-1 (mapping to a draft)
    int32_t pad_;
};
static_assert(sizeof(PoaTask) == 64, "PoaTask layout");

struct PoaResult {
    int32_t score;        // best local score
    int32_t end_t, end_i; // best cell (vertex rank, read prefix)
    int32_t path_len;     // number of moves written by the traceback
    int32_t first_t, first_i;   // first aligned (vertex rank, read position) of the path
    int32_t last_t, last_i;     // last aligned (vertex rank, read position)
};

}  // namespace ccs
