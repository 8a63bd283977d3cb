// Device-side layout of the Draft Stage (SparsePoa / PoaGraph; plain PODs shared by the
// CUDA kernels and the C++ host engine).
//
// The partial-order graph of every ZMW lives in HBM for the whole stage (DESIGN.md "Draft stage"): vertices keep a
// stable id (append-only), `order[t]` lists the ids in topological order and `rank[id]` is its inverse.  A vertex
// has up to kPoaMaxPred predecessors, kept sorted by rank: the first in pred0[id], the others in predx[id][0..6].
// The aligner writes one 64-cell DP row per vertex, indexed by vertex id: row t covers read prefixes
// i in [lo, lo+64), lo = max_pred(best_i[pred]) + 1 - 32 (clamped) -- the band follows the best cell of the
// predecessor rows.
#pragma once
#include <cstdint>

namespace ccs {

constexpr int kPoaBand = 64;
constexpr int kPoaMatch = 3, kPoaMismatch = -5, kPoaIns = -4, kPoaDel = -4;
constexpr int kPoaMaxPred = 8;
constexpr int kPoaMaxReads = 16;      // reads threaded into one graph (spans kept in the header)
constexpr int kPoaKmer = 11;
constexpr int kPoaVoteBases = 2048;   // orientation vote: k-mers of the read's first 2048 bases (spec)
constexpr int kPoaMaxRefLen = 131072; // longest k-mer vote reference (hash set in shared memory)
// band rule of the aligner (spec, DESIGN.md "Draft stage"): the rows of a block of 32 share one anchor row
constexpr int kPoaBlock = 32;
constexpr int kPoaAnchorMin = 20;     // an anchor row whose best score is lower carries no alignment yet ...
constexpr int kPoaBandDecay = 16;     // ... the band then moves this many cells back towards the read start
constexpr int kWindowGrid = 64;       // window borders of the Polish Stage lie on multiples of 64 draft bases (spec)

// A threaded read creates at most floor(2n/7) vertices: it is threaded only if score >= n, every new vertex costs
// at least 4 and a reused one earns 3, so 3(n - x) - 4x >= n.
inline int poa_new_vertex_bound(int n) { return (2 * n) / 7 + 1; }

// vertex word: base | in-degree << 2 | nReads << 8
struct PoaGraphHdr {
    int64_t voff;          // first vertex slot of this graph in the pools (meta, pred0, rank, order[2]; predx at 7 * voff)
    int32_t V;             // vertices
    int32_t cap;           // vertex capacity
    int32_t n_reads;       // threaded reads (the seed included)
    int32_t n_spans;
    int32_t order_sel;     // which of the two order buffers is current
    int32_t error;         // != 0: capacity exceeded (cannot happen by poa_new_vertex_bound)
    int32_t span_first[kPoaMaxReads], span_last[kPoaMaxReads];   // vertex ids
};

struct PoaGraphView {
    PoaGraphHdr* hdr;
    uint32_t* meta;
    int32_t* pred0;
    int32_t* predx;        // [slot][7]
    int32_t* rank;
    int32_t* order[2];
    int32_t* col;          // seed coordinate: the index of a seed vertex; a vertex added later takes the value of the old
                           // vertex it was placed behind (0 at the head) -- what the aligner's band moves by
};
// constant indices only: a dynamically indexed member array would be copied to local memory
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int32_t* poa_order(const PoaGraphView& G, int sel) { return sel ? G.order[1] : G.order[0]; }

struct PoaTask {
    int64_t codes_off;     // the read's emission codes in native orientation (base = code & 3)
    int64_t row_off;       // first DP row of this task in lo[] / besti[] / moves / hrows (row = vertex id, or rank if linear)
    int64_t step_off;      // traceback steps of this task (DAG tasks), capacity n
    int64_t tpl_off;       // linear tasks: the draft's bases
    int32_t n;             // read length
    int32_t graph;         // graph slot, or -1: linear template (every vertex t has the single predecessor t-1)
    int32_t V;             // linear tasks: template length (DAG tasks read the graph header)
    int32_t rev_idx;       // index of the read's orientation flag (0 forward, 1 reverse complement)
    int64_t scratch_off;   // DAG tasks: the graph's bookkeeping scratch (ints), >= 5 n + 12 cap + 16
    int64_t grid_off;      // linear tasks: where the traceback records the read position at every kWindowGrid-th template
                           // base (windowing, DESIGN.md "Windowing"); < 0: not recorded
};
static_assert(sizeof(PoaTask) == 64, "PoaTask layout");

struct PoaStep { int32_t vertex; int32_t readpos; };   // vertex id or -1 (insertion); end -> start order

struct PoaResult {
    int32_t score;        // best local score
    int32_t end_t, end_i; // best cell (vertex id / rank if linear, read prefix)
    int32_t path_len;     // DAG tasks: number of steps written (matches + insertions); linear: moves walked
    int32_t first_t, first_i;   // first aligned (vertex, read position) of the path
    int32_t last_t, last_i;     // last aligned (vertex, read position)
};

// k-mer orientation vote job: reads [read_begin, read_end) of the job list against one reference
struct PoaVoteJob {
    int64_t ref_off;       // reference bases: emission codes (ref_is_codes) or plain bases
    int32_t ref_len;
    int32_t ref_is_codes;
    int32_t read_begin, read_end;   // into the vote read list
};
struct PoaVoteRead { int64_t codes_off; int32_t n; int32_t rev_idx; };

}  // namespace ccs
