// poa_align: banded local alignment of a read against a partial-order graph, one warp per
// (graph, read) task, one 64-cell DP row per vertex in topological order
// (SparsePoa::OrientAndAddRead -> PoaGraph::TryAddRead, SURVEY.md 8a row a2 and Appendix B;
// "approximate draft consensus from a few subreads", /root/reference/docs/how-does-ccs-work.md:14,34-47).
//
//   C[v][i] = max(0, max_u H[u][i-1] + s(v, r_i), max_u H[u][i] + DEL)        u in pred(v)
//   H[v][i] = max(C[v][i], H[v][i-1] + INS)            -- in-row chain = warp prefix-max scan
//
// int32 max-plus arithmetic, bit-exact against the oracle.  The graph is device resident (poa_device.h): the
// kernel walks order[t], rows are stored by vertex id.  Each lane owns two adjacent cells; the previous row lives in
// registers (the common predecessor), the last 8 rows of a graph task in a shared-memory ring, older rows are re-read
// from global memory only when a vertex has a predecessor further back.  The band of every row of a 32-row block is
// placed from the block's anchor row (block-anchored band rule, DESIGN.md "Draft stage").  Reads are the uploaded
// emission codes, oriented on the fly.  Per vertex the kernel writes 64 B of traceback moves and its band start
// (+ 256 B of scores for DAGs): algorithmic bytes per task = V * (64 [+256] + 4) + n.
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include "poa_device.h"
#include "poa_launch.h"

namespace ccs {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerCta = 4;

constexpr int kRing = 8;            // recent DP rows of a graph task kept in shared memory (power of two)
constexpr int kNeg = -(1 << 29);

__device__ __forceinline__ int row_get(const int* __restrict__ row, const int idx) {
    return ((unsigned)idx < (unsigned)kPoaBand) ? row[idx] : 0;
}

// One warp per (graph | draft, read) task, one 64-cell row per vertex in topological order, two adjacent cells per lane.
// The band of every row of a 32-row block is known when the block starts (block-anchored band rule, DESIGN.md "Draft
// stage"): the lanes fetch the block's vertices and place their bands in one go, and nothing per row depends on a
// reduction over the previous row any more -- the per-row dependent chain is {neighbour cells, candidates, the in-row
// insertion scan}.  The previous row stays in REGISTERS: when the only predecessor is the previous vertex (every vertex
// of a draft, most vertices of a graph) and the band moved by 0 or 1 cells, its three operands are the lane's own two
// cells and one shuffle, and the read bases move along with the band (one new base per lane and row).  Other rows take
// the general path (predecessor rows from a shared-memory ring of the last 8 rows, else from global memory).  The best
// cell of the task is tracked per lane and reduced once at the end; only the anchor row of a block is reduced.
template <bool kDag>
__global__ void __launch_bounds__(kWarpsPerCta * 32) poa_align_kernel(
    const PoaTask* __restrict__ tasks, const int n_tasks, const PoaGraphView G, const uint8_t* __restrict__ drafts,
    const uint8_t* __restrict__ codes, const uint8_t* __restrict__ rev_flags, int32_t* __restrict__ lo_arr,
    uint8_t* __restrict__ moves, int32_t* __restrict__ hrows, PoaResult* __restrict__ results) {
    __shared__ int s_ring[kDag ? kWarpsPerCta : 1][kRing][kPoaBand];
    __shared__ int s_rlo[kDag ? kWarpsPerCta : 1][kRing];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task_id = blockIdx.x * kWarpsPerCta + warp;
    if (task_id >= n_tasks) return;
    const PoaTask T = tasks[task_id];
    int V = T.V;
    int64_t voff = 0;
    const int32_t* __restrict__ ord = nullptr;
    if (kDag) {
        const PoaGraphHdr& H = G.hdr[T.graph];
        V = H.V; voff = H.voff;
        ord = poa_order(G, H.order_sel) + voff;
    }
    V = __reduce_max_sync(kFull, V);      // same value in every lane; as a reduction result it is provably warp-uniform
    const uint32_t* __restrict__ meta = G.meta + voff;
    const int32_t* __restrict__ pred0 = G.pred0 + voff;
    const int32_t* __restrict__ predx = G.predx + 7 * voff;
    const int32_t* __restrict__ colv = G.col + voff;
    const int32_t* __restrict__ rankv = G.rank + voff;
    const uint8_t* __restrict__ tplb = drafts + T.tpl_off;
    const int n = T.n;
    const bool rev = rev_flags[T.rev_idx] != 0;
    // oriented base of read position q: (rbp[q * rdir] & 3) ^ rxor
    const uint8_t* __restrict__ rbp = codes + T.codes_off + (rev ? n - 1 : 0);
    const int rdir = rev ? -1 : 1, rxor = rev ? 3 : 0;
    auto read_base = [&](const int i) -> int {     // oriented base in front of read prefix i, 255 outside the read
        return (i >= 1 && i <= n) ? ((rbp[(i - 1) * rdir] & 3) ^ rxor) : 255;
    };
    int32_t* __restrict__ lo_r = lo_arr + T.row_off;
    uint8_t* __restrict__ mv_r = moves + T.row_off * kPoaBand;
    int32_t* __restrict__ h_r = (kDag && hrows) ? hrows + T.row_off * kPoaBand : nullptr;
    const int lo_max = max(0, n + 1 - kPoaBand);
    const int wslot = kDag ? warp : 0;

    int lbest = 0, lt = 0x7fffffff, li = 0, lid = -1;    // this lane's best cell: value, row (topological index), prefix, vertex
    // anchor of the current block: band start, best cell, best score and seed coordinate of the row in front of it
    // (block 0: the cell in front of the first row)
    int a_lo = 0, a_bi = 0, a_max = kPoaAnchorMin, a_col = -1;
    int pH0 = 0, pH1 = 0, prev_lo = 0;                   // previous row (registers)
    int rb0 = 255, rb1 = 255;                            // read bases in front of the lane's two prefixes

    for (int t0 = 0; t0 < V; t0 += kPoaBlock) {
        // ---- the block's vertices (one per lane) and their bands
        const int tt = min(t0 + lane, V - 1);
        // mw = base | in-degree << 2; md = how many rows back each predecessor lies (one byte each, 255 = further than
        // the ring): predecessor ids and ranks are looked up here, 32 vertices at a time, off the per-row dependent chain
        int mid = tt, mcol = tt;
        uint32_t mw, md0 = 0xffffff01u, md1 = 0xffffffffu;
        if (!kDag) mw = (uint32_t)tplb[tt] | ((tt > 0) ? 4u : 0u);
        else {
            mid = ord[tt];
            mw = meta[mid] & 63u;
            mcol = colv[mid];
            const int nin = (int)(mw >> 2);
            md0 = 0xffffffffu;
            for (int k = 0; k < nin; ++k) {
                const int p = (k == 0) ? pred0[mid] : predx[7 * (int64_t)mid + k - 1];
                const uint32_t d = (uint32_t)min(tt - rankv[p], 255);
                if (k < 4) md0 = (md0 & ~(0xffu << (8 * k))) | (d << (8 * k));
                else md1 = (md1 & ~(0xffu << (8 * (k - 4)))) | (d << (8 * (k - 4)));
            }
        }
        int mlo = (a_max >= kPoaAnchorMin) ? a_bi + (mcol - a_col) - kPoaBand / 2 : a_lo - kPoaBandDecay;
        mlo = min(max(mlo, 0), lo_max);
        if (t0 + lane < V) lo_r[mid] = mlo;
        if (kDag) __syncwarp();      // a later row of the block may look the band start of an earlier one up
        const int nrows = min(kPoaBlock, V - t0);

        for (int r = 0; r < nrows; ++r) {
            const int t = t0 + r;
            const int id = kDag ? __shfl_sync(kFull, mid, r) : t;
            const uint32_t vw = __shfl_sync(kFull, mw, r);
            const int vb = (int)(vw & 3u), npred = (int)(vw >> 2);
            const int lo = __shfl_sync(kFull, mlo, r);
            const uint32_t d0 = kDag ? __shfl_sync(kFull, md0, r) : 1u;
            const int dl = lo - prev_lo;
            const int i0 = lo + 2 * lane, i1 = i0 + 1;
            // interior row (warp-uniform): every cell is a real read prefix with a base in front of it -- no validity tests
            const bool interior = lo >= 1 && lo + kPoaBand <= n;
            // insertion chain H[c] = max(C[c], H[c-1] + INS) over the row's cells with i <= n, row outputs, best cell
            auto finish = [&](auto tag, const int c0, const int c1, unsigned m0, unsigned m1) {
                constexpr bool kIn = decltype(tag)::value;       // interior row: every cell valid
                const int ofs = -kPoaIns * (2 * lane);
                int x0 = c0 + ofs, x1 = c1 + ofs - kPoaIns;
                if (!kIn) {
                    if (i0 > n) x0 = kNeg;
                    if (i1 > n) x1 = kNeg;
                }
                x1 = max(x1, x0);
                int sc = x1;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1)      // a lane below `off` gets its own value back: max is idempotent
                    sc = max(sc, __shfl_up_sync(kFull, sc, off));
                int carry = __shfl_up_sync(kFull, sc, 1);
                if (lane == 0) carry = kNeg;
                x0 = max(x0, carry);
                x1 = max(x1, x0);
                int H0 = x0 - ofs, H1 = x1 - ofs + kPoaIns;
                if (!kIn) {
                    if (i0 > n) H0 = 0;
                    if (i1 > n) H1 = 0;
                }
                if (H0 > c0) m0 = 3u;
                if (H1 > c1) m1 = 3u;
                // row outputs (by vertex id)
                *reinterpret_cast<uchar2*>(mv_r + (size_t)id * kPoaBand + 2 * lane) = make_uchar2((unsigned char)m0, (unsigned char)m1);
                if (kDag) {
                    __syncwarp();   // everyone has read the ring
                    int* srow = s_ring[wslot][t & (kRing - 1)];
                    srow[2 * lane] = H0; srow[2 * lane + 1] = H1;
                    if (lane == 0) s_rlo[wslot][t & (kRing - 1)] = lo;
                    if (h_r != nullptr) *reinterpret_cast<int2*>(h_r + (size_t)id * kPoaBand + 2 * lane) = make_int2(H0, H1);
                    __syncwarp();   // row visible in shared memory before the next vertex reads it
                }
                // this lane's best cell so far: the first row that reaches its maximum, smaller prefix first within a row
                if (H0 > lbest) { lbest = H0; lt = t; li = i0; lid = id; }
                if (H1 > lbest) { lbest = H1; lt = t; li = i1; lid = id; }
                pH0 = H0; pH1 = H1; prev_lo = lo;
            };
            if (npred == 1 && (d0 & 255u) == 1u && dl == 1 && interior) {
                // hot path: the only predecessor is the previous row (registers), the band moved with the diagonal, every
                // cell is valid -- own two cells + the next lane's first, one new read base
                const int hm1 = pH0, h0 = pH1;
                int h1 = __shfl_down_sync(kFull, pH0, 1);
                if (lane == 31) h1 = 0;
                rb0 = rb1;
                rb1 = (rbp[(i1 - 1) * rdir] & 3) ^ rxor;
                const int bm0 = hm1 + ((rb0 == vb) ? kPoaMatch : kPoaMismatch), bm1 = h0 + ((rb1 == vb) ? kPoaMatch : kPoaMismatch);
                const int bd0 = h0 + kPoaDel, bd1 = h1 + kPoaDel;
                const int c0 = max(max(bm0, bd0), 0), c1 = max(max(bm1, bd1), 0);
                finish(std::true_type(), c0, c1, (c0 == 0) ? 0u : ((bm0 >= bd0) ? 1u : 2u), (c1 == 0) ? 0u : ((bm1 >= bd1) ? 1u : 2u));
            } else if (npred == 1 && (d0 & 255u) == 1u) {
                // the only predecessor is the previous row, held in registers as band cells 2*lane, 2*lane + 1
                int hm1, h0, h1;
                if (dl == 1) {            // band moved with the diagonal: own two cells + the next lane's first
                    hm1 = pH0; h0 = pH1;
                    h1 = __shfl_down_sync(kFull, pH0, 1);
                    if (lane == 31) h1 = 0;
                    rb0 = rb1; rb1 = read_base(i1);
                } else if (dl == 0) {     // band stayed: the previous lane's second cell + own two
                    hm1 = __shfl_up_sync(kFull, pH1, 1);
                    if (lane == 0) hm1 = 0;
                    h0 = pH0; h1 = pH1;
                } else {                  // any other move (a new anchor, an irregular seed coordinate): cells by index
                    const int a = 2 * lane + dl;
                    const int q0 = __shfl_sync(kFull, pH0, ((a - 1) >> 1) & 31), q1 = __shfl_sync(kFull, pH1, ((a - 1) >> 1) & 31);
                    const int r0 = __shfl_sync(kFull, pH0, (a >> 1) & 31), r1 = __shfl_sync(kFull, pH1, (a >> 1) & 31);
                    const int u0 = __shfl_sync(kFull, pH0, ((a + 1) >> 1) & 31), u1 = __shfl_sync(kFull, pH1, ((a + 1) >> 1) & 31);
                    hm1 = ((unsigned)(a - 1) < (unsigned)kPoaBand) ? (((a - 1) & 1) ? q1 : q0) : 0;
                    h0 = ((unsigned)a < (unsigned)kPoaBand) ? ((a & 1) ? r1 : r0) : 0;
                    h1 = ((unsigned)(a + 1) < (unsigned)kPoaBand) ? (((a + 1) & 1) ? u1 : u0) : 0;
                    rb0 = read_base(i0); rb1 = read_base(i1);
                }
                const int bm0 = (rb0 != 255) ? hm1 + ((rb0 == vb) ? kPoaMatch : kPoaMismatch) : 0;
                const int bm1 = (rb1 != 255) ? h0 + ((rb1 == vb) ? kPoaMatch : kPoaMismatch) : 0;
                const int bd0 = (i0 <= n) ? h0 + kPoaDel : 0;
                const int bd1 = (i1 <= n) ? h1 + kPoaDel : 0;
                // match wins ties against deletion; candidates must be > 0.  Predecessor index 0 = the only one.
                const int c0 = max(max(bm0, bd0), 0), c1 = max(max(bm1, bd1), 0);
                finish(std::false_type(), c0, c1, (c0 == 0) ? 0u : ((bm0 >= bd0) ? 1u : 2u), (c1 == 0) ? 0u : ((bm1 >= bd1) ? 1u : 2u));
            } else {
                rb0 = read_base(i0); rb1 = read_base(i1);
                const int sc0 = (rb0 == vb) ? kPoaMatch : kPoaMismatch;
                const int sc1 = (rb1 == vb) ? kPoaMatch : kPoaMismatch;
                int bm0 = 0, bm1 = 0, bd0 = 0, bd1 = 0;        // best match / deletion candidates (must be > 0 to count)
                int km0 = 0, km1 = 0, kd0 = 0, kd1 = 0;
                if (npred == 0) {
                    if (rb0 != 255 && sc0 > 0) { bm0 = sc0; km0 = 63; }
                    if (rb1 != 255 && sc1 > 0) { bm1 = sc1; km1 = 63; }
                }
                if (kDag) {
                    const uint32_t d1 = (npred > 4) ? __shfl_sync(kFull, md1, r) : 0xffffffffu;
                    for (int k = 0; k < npred; ++k) {
                        // a predecessor at most kRing rows back sits in the ring slot of its rank; others come from global memory
                        const int d = (int)(((k < 4 ? d0 : d1) >> (8 * (k & 3))) & 255u);
                        const bool near = d <= kRing;
                        const int slot = (t - d) & (kRing - 1);
                        int pr = 0;
                        if (!near) pr = (k == 0) ? pred0[id] : predx[7 * (int64_t)id + k - 1];
                        const int* __restrict__ row = near ? s_ring[wslot][slot] : (h_r + (size_t)pr * kPoaBand);
                        const int a = 2 * lane + lo - (near ? s_rlo[wslot][slot] : lo_r[pr]);
                        const int hm1 = row_get(row, a - 1), h0 = row_get(row, a), h1 = row_get(row, a + 1);
                        if (rb0 != 255) { const int c = hm1 + sc0; if (c > bm0) { bm0 = c; km0 = k; } }
                        if (rb1 != 255) { const int c = h0 + sc1; if (c > bm1) { bm1 = c; km1 = k; } }
                        if (i0 <= n) { const int c = h0 + kPoaDel; if (c > bd0) { bd0 = c; kd0 = k; } }
                        if (i1 <= n) { const int c = h1 + kPoaDel; if (c > bd1) { bd1 = c; kd1 = k; } }
                    }
                }
                // match wins ties against deletion (oracle evaluation order)
                int c0, c1;
                unsigned m0, m1;
                if (bm0 >= bd0) { c0 = bm0; m0 = bm0 > 0 ? (1u | (km0 << 2)) : 0u; } else { c0 = bd0; m0 = 2u | (kd0 << 2); }
                if (bm1 >= bd1) { c1 = bm1; m1 = bm1 > 0 ? (1u | (km1 << 2)) : 0u; } else { c1 = bd1; m1 = 2u | (kd1 << 2); }
                finish(std::false_type(), c0, c1, m0, m1);
            }
        }
        // ---- anchor of the next block = this block's last row: best cell = largest value, smallest prefix on ties
        // (key = value * 64 + (63 - cell); scores are >= 0 and < 2^25)
        {
            const int hv = max(pH0, pH1);
            const int cell = (pH0 >= pH1) ? 2 * lane : 2 * lane + 1;
            const int rkey = __reduce_max_sync(kFull, hv * kPoaBand + (kPoaBand - 1 - cell));
            static_assert(kPoaBand == 64, "key packing assumes 64 cells per row");
            a_max = rkey >> 6;
            a_lo = prev_lo;
            a_bi = (a_max > 0) ? prev_lo + (kPoaBand - 1 - (rkey & (kPoaBand - 1))) : prev_lo;
            a_col = __shfl_sync(kFull, mcol, nrows - 1);
        }
    }
    // ---- best cell of the task: largest value; the first row in topological order; the smallest prefix
    {
        const int gbest = __reduce_max_sync(kFull, lbest);
        const int bt = __reduce_min_sync(kFull, (lbest == gbest && gbest > 0) ? lt : 0x7fffffff);
        const int bi = __reduce_min_sync(kFull, (lbest == gbest && gbest > 0 && lt == bt) ? li : 0x7fffffff);
        const unsigned who = __ballot_sync(kFull, lbest == gbest && gbest > 0 && lt == bt && li == bi);
        const int gid = who ? __shfl_sync(kFull, lid, __ffs(who) - 1) : -1;
        if (lane == 0) {
            PoaResult r;
            r.score = gbest; r.end_t = gid; r.end_i = who ? bi : -1; r.path_len = 0;
            r.first_t = r.first_i = r.last_t = r.last_i = -1;
            results[task_id] = r;
        }
    }
}

// Traceback: one warp per task follows the stored moves from the best cell back to the local start.
// Linear tasks (mapping a read to the draft) only need the extents: the warp stages a window of the next 32 rows
// (their band starts and 64-byte move rows, one coalesced 2 KB read) in shared memory and walks it from there.
// DAG tasks write the path as steps {vertex id or -1, read position}, end -> start, deletions left out -- the
// input of the CommitAdd kernel; rows are stored by vertex id, so each step is a dependent global load.
__global__ void __launch_bounds__(kWarpsPerCta * 32) poa_traceback_kernel(
    const PoaTask* __restrict__ tasks, const int n_tasks, const PoaGraphView G, const int32_t* __restrict__ lo_arr,
    const uint8_t* __restrict__ moves, PoaStep* __restrict__ steps, PoaResult* __restrict__ results,
    int32_t* __restrict__ grid) {
    __shared__ __align__(16) uint8_t s_mv[kWarpsPerCta][32 * kPoaBand];
    __shared__ int s_lo[kWarpsPerCta][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task_id = blockIdx.x * kWarpsPerCta + warp;
    if (task_id >= n_tasks) return;
    const PoaTask T = tasks[task_id];
    PoaResult r = results[task_id];
    const int32_t* __restrict__ lo_r = lo_arr + T.row_off;
    const uint8_t* __restrict__ mv_r = moves + T.row_off * kPoaBand;
    int t = r.end_t, i = r.end_i, len = 0;
    if (T.graph >= 0) {
        // Rows are stored by vertex id and a path mostly runs down consecutive ids along a diagonal (a match whose only
        // candidate predecessor is the vertex before it).  A window of 32 consecutive ids -- band starts, first
        // predecessors, 64-byte move rows: one coalesced read -- is staged in shared memory and re-staged only when the
        // path leaves it.  Every iteration the lanes look at the next cells of the diagonal at once, (t - j, i - j), and
        // the walk advances over the whole run of plain matches in front of it; the cell that ends the run (a mismatch
        // placed on another predecessor, a deletion, an insertion, the local start) is decoded from lane 0's cell.
        __shared__ int s_p0[kWarpsPerCta][32];
        const int64_t voff = G.hdr[T.graph].voff;
        const int32_t* __restrict__ pred0 = G.pred0 + voff;
        const int32_t* __restrict__ predx = G.predx + 7 * voff;
        PoaStep* __restrict__ out = steps + T.step_off;
        int wbase = 0, wtop = -1;
        while (t >= 0) {
            if (t < wbase || t > wtop) {
                __syncwarp();
                wbase = max(t - 31, 0); wtop = t;
                const int row = wbase + lane;
                if (row <= wtop) {
                    s_lo[warp][lane] = lo_r[row];
                    s_p0[warp][lane] = pred0[row];
                    const uint4* src = reinterpret_cast<const uint4*>(mv_r + (size_t)row * kPoaBand);
                    uint4* dst = reinterpret_cast<uint4*>(&s_mv[warp][lane * kPoaBand]);
                    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
                }
                __syncwarp();
            }
            const int tj = t - lane, ij = i - lane;
            int pj = -2;
            unsigned mj = 0u;
            bool inb = false;
            if (tj >= wbase) {
                const int c = ij - s_lo[warp][tj - wbase];
                pj = s_p0[warp][tj - wbase];
                inb = (unsigned)c < (unsigned)kPoaBand;
                if (inb) mj = s_mv[warp][(tj - wbase) * kPoaBand + c];
            }
            const unsigned okm = __ballot_sync(kFull, inb && mj == 1u && pj == tj - 1);
            const int run = (okm == kFull) ? 32 : __ffs(~okm) - 1;      // plain matches in a row, lane 0 first
            if (run > 0) {
                if (lane < run && len + lane < T.n) out[len + lane] = PoaStep{tj, ij - 1};
                if (r.last_t < 0) { r.last_t = t; r.last_i = i - 1; }
                r.first_t = t - (run - 1); r.first_i = i - 1 - (run - 1);
                len += run; t -= run; i -= run;
                continue;
            }
            // one step from lane 0's cell
            const bool in0 = __shfl_sync(kFull, inb ? 1 : 0, 0) != 0;
            if (!in0) break;
            const unsigned m = __shfl_sync(kFull, mj, 0);
            const int p0 = __shfl_sync(kFull, pj, 0);
            const unsigned kind = m & 3u, k = m >> 2;
            if (kind == 0u) break;
            if (kind == 1u) {           // match / mismatch: read base i-1 on vertex t
                if (lane == 0 && len < T.n) out[len] = PoaStep{t, i - 1};
                ++len;
                if (r.last_t < 0) { r.last_t = t; r.last_i = i - 1; }
                r.first_t = t; r.first_i = i - 1;
                if (k == 63u) break;
                t = (k == 0u) ? p0 : predx[7 * (int64_t)t + k - 1];
                i -= 1;
            } else if (kind == 2u) {    // deletion: vertex skipped
                t = (k == 0u) ? p0 : predx[7 * (int64_t)t + k - 1];
            } else {                    // insertion: read base i-1 without a vertex
                if (lane == 0 && len < T.n) out[len] = PoaStep{-1, i - 1};
                ++len;
                i -= 1;
            }
        }
        if (lane == 0) { r.path_len = len; results[task_id] = r; }
        return;
    }
    // windowing: grid[t / 64] = number of oriented read bases placed before template position t on the path (the base
    // matched there, or the next one when t is deleted), for every t that is a multiple of kWindowGrid
    int32_t* __restrict__ gr = (grid != nullptr && T.grid_off >= 0) ? grid + T.grid_off : nullptr;
    bool done = t < 0;
    while (!done) {
        const int wbase = max(t - 31, 0);
        {   // stage rows wbase .. t
            const int row = wbase + lane;
            if (row <= t) {
                s_lo[warp][lane] = lo_r[row];
                const uint4* src = reinterpret_cast<const uint4*>(mv_r + (size_t)row * kPoaBand);
                uint4* dst = reinterpret_cast<uint4*>(&s_mv[warp][lane * kPoaBand]);
                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
            }
        }
        __syncwarp();
        while (t >= wbase) {
            // the lanes look at the next cells of the diagonal at once, (t - j, i - j): the walk advances over the whole
            // run of matches placed on the position before them (move byte 1; the first row of the draft carries 253)
            const int tj = t - lane, ij = i - lane;
            unsigned mj = 0u;
            bool inb = false;
            if (tj >= wbase) {
                const int c = ij - s_lo[warp][tj - wbase];
                inb = (unsigned)c < (unsigned)kPoaBand;
                if (inb) mj = s_mv[warp][(tj - wbase) * kPoaBand + c];
            }
            const unsigned okm = __ballot_sync(kFull, inb && mj == 1u);
            const int run = (okm == kFull) ? 32 : __ffs(~okm) - 1;
            if (run > 0) {
                if (gr != nullptr && lane < run && (tj & (kWindowGrid - 1)) == 0) gr[tj / kWindowGrid] = ij - 1;
                if (r.last_t < 0) { r.last_t = t; r.last_i = i - 1; }
                r.first_t = t - (run - 1); r.first_i = i - 1 - (run - 1);
                len += run; t -= run; i -= run;
                continue;
            }
            // one step from lane 0's cell
            if (__shfl_sync(kFull, inb ? 1 : 0, 0) == 0) { done = true; break; }
            const unsigned m = __shfl_sync(kFull, mj, 0);
            const unsigned kind = m & 3u, k = m >> 2;
            if (kind == 0u) { done = true; break; }
            ++len;
            if (gr != nullptr && lane == 0 && kind != 3u && (t & (kWindowGrid - 1)) == 0) gr[t / kWindowGrid] = (kind == 1u) ? i - 1 : i;
            if (kind == 1u) {           // match / mismatch: read base i-1 on template position t
                if (r.last_t < 0) { r.last_t = t; r.last_i = i - 1; }
                r.first_t = t; r.first_i = i - 1;
                if (k == 63u) { done = true; break; }
                t -= 1; i -= 1;
            } else if (kind == 2u) {    // deletion: template position skipped
                t -= 1;
            } else {                    // insertion: read base i-1 without a template position
                i -= 1;
            }
        }
        if (t < 0) done = true;
        __syncwarp();
    }
    if (lane == 0) { r.path_len = len; results[task_id] = r; }
}

}  // namespace

void launch_poa_align(const PoaTask* tasks, int n_tasks, const PoaGraphView& G, const uint8_t* drafts, const uint8_t* codes,
                      const uint8_t* rev_flags, int32_t* lo, uint8_t* moves, int32_t* hrows,
                      PoaStep* steps, PoaResult* results, cudaStream_t stream, int32_t* grid) {
    if (n_tasks <= 0) return;
    const int blocks = (n_tasks + kWarpsPerCta - 1) / kWarpsPerCta;
    if (hrows == nullptr)        // linear templates (subread -> draft mapping)
        poa_align_kernel<false><<<blocks, kWarpsPerCta * 32, 0, stream>>>(tasks, n_tasks, G, drafts, codes, rev_flags, lo, moves,
                                                                         nullptr, results);
    else                         // sequence-to-graph rounds
        poa_align_kernel<true><<<blocks, kWarpsPerCta * 32, 0, stream>>>(tasks, n_tasks, G, drafts, codes, rev_flags, lo, moves,
                                                                        hrows, results);
    poa_traceback_kernel<<<blocks, kWarpsPerCta * 32, 0, stream>>>(tasks, n_tasks, G, lo, moves, steps, results, grid);
}

}  // namespace ccs
