// poa_align: banded local alignment of a read against a partial-order graph, one warp per
// (graph, read) task, one 64-cell DP row per vertex in topological order
// (SparsePoa::OrientAndAddRead -> PoaGraph::TryAddRead, SURVEY.md 8a row a2 and Appendix B;
// "approximate draft consensus from a few subreads", /root/reference/docs/how-does-ccs-work.md:14,34-47).
//
//   C[v][i] = max(0, max_u H[u][i-1] + s(v, r_i), max_u H[u][i] + DEL)        u in pred(v)
//   H[v][i] = max(C[v][i], H[v][i-1] + INS)            -- in-row chain = warp prefix-max scan
//
// int32 max-plus arithmetic, bit-exact against the oracle.  The graph is device resident (poa_device.h): the
// kernel walks order[t], rows are stored by vertex id.  Each lane owns two adjacent cells; the previous row
// lives in shared memory (the common predecessor), older rows are re-read from global memory only when a
// vertex has a non-adjacent predecessor.  Reads are the uploaded emission codes, oriented on the fly.  Per
// vertex the kernel writes 64 B of traceback moves (+ 256 B of scores for DAGs): algorithmic bytes per task
// = V * (64 [+256] + 8) + n.
#include <cuda_runtime.h>
#include <cstdint>
#include "poa_device.h"
#include "poa_launch.h"

namespace ccs {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerCta = 4;
constexpr int kRing = 8;            // recent DP rows kept in shared memory (power of two)

__device__ __forceinline__ int row_get(const int* __restrict__ row, const int idx) {
    return ((unsigned)idx < (unsigned)kPoaBand) ? row[idx] : 0;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) poa_align_kernel(
    const PoaTask* __restrict__ tasks, const int n_tasks, const PoaGraphView G, const uint8_t* __restrict__ drafts,
    const uint8_t* __restrict__ codes, const uint8_t* __restrict__ rev_flags, int32_t* __restrict__ lo_arr,
    int32_t* __restrict__ besti_arr, uint8_t* __restrict__ moves, int32_t* __restrict__ hrows,
    PoaResult* __restrict__ results) {
    // The last kRing rows stay in shared memory with their vertex ids, band starts and best cells: the predecessors of a
    // branch vertex are almost always among them (a bubble opened by one read is a few vertices long), so the dependent
    // chain of a branch vertex does not wait on global memory either.
    __shared__ int s_ring[kWarpsPerCta][kRing][kPoaBand];
    __shared__ int s_rid[kWarpsPerCta][kRing], s_rlo[kWarpsPerCta][kRing], s_rbi[kWarpsPerCta][kRing];
    __shared__ int s_meta[kWarpsPerCta][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task_id = blockIdx.x * kWarpsPerCta + warp;
    if (task_id >= n_tasks) return;
    const PoaTask T = tasks[task_id];
    const bool linear = T.graph < 0;
    int V = T.V;
    int64_t voff = 0;
    const int32_t* __restrict__ ord = nullptr;
    if (!linear) {
        const PoaGraphHdr& H = G.hdr[T.graph];
        V = H.V; voff = H.voff;
        ord = poa_order(G, H.order_sel) + voff;
    }
    V = __reduce_max_sync(kFull, V);      // same value in every lane; as a reduction result it is provably warp-uniform
    const uint32_t* __restrict__ meta = G.meta + voff;
    const int32_t* __restrict__ pred0 = G.pred0 + voff;
    const int32_t* __restrict__ predx = G.predx + 7 * voff;
    const uint8_t* __restrict__ tplb = drafts + T.tpl_off;
    const uint8_t* __restrict__ rc = codes + T.codes_off;
    const int n = T.n;
    const bool rev = rev_flags[T.rev_idx] != 0;
    int32_t* __restrict__ lo_r = lo_arr + T.row_off;
    int32_t* __restrict__ bi_r = besti_arr + T.row_off;
    uint8_t* __restrict__ mv_r = moves + T.row_off * kPoaBand;
    int32_t* __restrict__ h_r = hrows ? hrows + T.row_off * kPoaBand : nullptr;
    const bool store_h = !linear && h_r != nullptr;
    const int lo_max = max(0, n + 1 - kPoaBand);
    int* srow = s_ring[warp][kRing - 1];      // "previous row" before the first vertex: zeros
    auto read_base = [&](const int i) -> int {     // oriented base of read position i (0-based)
        return rev ? 3 - (rc[n - 1 - i] & 3) : (rc[i] & 3);
    };

    int gbest = 0, gt = -1, gi = -1;
    int prev_lo = 0, prev_besti = 0, prev_id = -1;
    srow[2 * lane] = 0; srow[2 * lane + 1] = 0;
    if (lane < kRing) s_rid[warp][lane] = -2;
    int* smeta = s_meta[warp];          // per block of 32 vertices: id, base, in-degree, first predecessor
    __syncwarp();

    for (int t = 0; t < V; ++t) {
        // Vertex metadata does not depend on the DP state: every 32 vertices the lanes fetch one
        // vertex each so that the per-vertex dependent chain below never waits on global memory for them.
        if ((t & 31) == 0) {
            const int tt = min(t + lane, V - 1);
            int mid = tt, mb, mn, mf = tt - 1;
            if (linear) { mb = tplb[tt]; mn = (tt > 0) ? 1 : 0; }
            else {
                mid = ord[tt];
                const uint32_t m = meta[mid];
                mb = (int)(m & 3u); mn = (int)((m >> 2) & 15u); mf = pred0[mid];
            }
            __syncwarp();
            smeta[lane] = mid; smeta[32 + lane] = mb; smeta[64 + lane] = mn; smeta[96 + lane] = mf;
            __syncwarp();
        }
        const int id = smeta[t & 31];
        const int vb = smeta[32 + (t & 31)];
        const int npred = smeta[64 + (t & 31)];
        const int first_pred = smeta[96 + (t & 31)];
        int lo, c0, c1;
        unsigned m0, m1;
        int i0, i1;
        if (npred == 1 && first_pred == prev_id) {
            // fast path (every vertex of a linear template, almost every vertex of a POA graph): the only
            // predecessor is the previous row, which sits in shared memory
            lo = min(max(prev_besti + 1 - kPoaBand / 2, 0), lo_max);
            i0 = lo + 2 * lane; i1 = i0 + 1;
            const int a = 2 * lane + (lo - prev_lo);
            const int hm1 = row_get(srow, a - 1), h0 = row_get(srow, a), h1 = row_get(srow, a + 1);
            const bool v0 = (i0 >= 1 && i0 <= n), v1 = (i1 >= 1 && i1 <= n);
            const int rb0 = v0 ? read_base(i0 - 1) : 255, rb1 = v1 ? read_base(i1 - 1) : 255;
            const int bm0 = v0 ? hm1 + ((rb0 == vb) ? kPoaMatch : kPoaMismatch) : 0;
            const int bm1 = v1 ? h0 + ((rb1 == vb) ? kPoaMatch : kPoaMismatch) : 0;
            const int bd0 = (i0 <= n) ? h0 + kPoaDel : 0;
            const int bd1 = (i1 <= n) ? h1 + kPoaDel : 0;
            // match wins ties against deletion; candidates must be > 0.  Predecessor index 0 = the only one.
            c0 = max(max(bm0, bd0), 0); c1 = max(max(bm1, bd1), 0);
            m0 = (c0 == 0) ? 0u : ((bm0 >= bd0) ? 1u : 2u);
            m1 = (c1 == 0) ? 0u : ((bm1 >= bd1) ? 1u : 2u);
        } else {
        // band start from the predecessors' best cells
        lo = 0;
        if (npred > 0) {
            int m = 0;
            for (int k = 0; k < npred; ++k) {
                const int pr = (k == 0) ? first_pred : predx[7 * (int64_t)id + k - 1];
                const unsigned hit = __ballot_sync(kFull, s_rid[warp][lane & (kRing - 1)] == pr) & ((1u << kRing) - 1u);
                m = max(m, hit ? s_rbi[warp][__ffs(hit) - 1] : bi_r[pr]);
            }
            lo = min(max(m + 1 - kPoaBand / 2, 0), lo_max);
        }
        i0 = lo + 2 * lane; i1 = i0 + 1;
        const int rb0 = (i0 >= 1 && i0 <= n) ? read_base(i0 - 1) : 255;
        const int rb1 = (i1 >= 1 && i1 <= n) ? read_base(i1 - 1) : 255;
        const int sc0 = (rb0 == vb) ? kPoaMatch : kPoaMismatch;
        const int sc1 = (rb1 == vb) ? kPoaMatch : kPoaMismatch;
        int bm0 = 0, bm1 = 0, bd0 = 0, bd1 = 0;        // best match / deletion candidates (must be > 0 to count)
        int km0 = 0, km1 = 0, kd0 = 0, kd1 = 0;
        if (npred == 0) {
            if (rb0 != 255 && sc0 > 0) { bm0 = sc0; km0 = 63; }
            if (rb1 != 255 && sc1 > 0) { bm1 = sc1; km1 = 63; }
        }
        for (int k = 0; k < npred; ++k) {
            const int pr = (k == 0) ? first_pred : predx[7 * (int64_t)id + k - 1];
            const unsigned hit = __ballot_sync(kFull, s_rid[warp][lane & (kRing - 1)] == pr) & ((1u << kRing) - 1u);
            const int slot = __ffs(hit) - 1;
            const int* __restrict__ row = hit ? s_ring[warp][slot] : (h_r + (size_t)pr * kPoaBand);
            const int dl = lo - (hit ? s_rlo[warp][slot] : lo_r[pr]);
            const int a = 2 * lane + dl;
            const int hm1 = row_get(row, a - 1), h0 = row_get(row, a), h1 = row_get(row, a + 1);
            if (rb0 != 255) { const int c = hm1 + sc0; if (c > bm0) { bm0 = c; km0 = k; } }
            if (rb1 != 255) { const int c = h0 + sc1; if (c > bm1) { bm1 = c; km1 = k; } }
            if (i0 <= n) { const int c = h0 + kPoaDel; if (c > bd0) { bd0 = c; kd0 = k; } }
            if (i1 <= n) { const int c = h1 + kPoaDel; if (c > bd1) { bd1 = c; kd1 = k; } }
        }
        // match wins ties against deletion (oracle evaluation order)
        if (bm0 >= bd0) { c0 = bm0; m0 = bm0 > 0 ? (1u | (km0 << 2)) : 0u; } else { c0 = bd0; m0 = 2u | (kd0 << 2); }
        if (bm1 >= bd1) { c1 = bm1; m1 = bm1 > 0 ? (1u | (km1 << 2)) : 0u; } else { c1 = bd1; m1 = 2u | (kd1 << 2); }
        }
        __syncwarp();   // everyone has read the previous row
        // insertion chain: H[c] = max(C[c], H[c-1] + INS) over the row's cells with i <= n
        const int NEG = -(1 << 29);
        int x0 = (i0 <= n) ? c0 - kPoaIns * (2 * lane) : NEG;
        int x1 = (i1 <= n) ? c1 - kPoaIns * (2 * lane + 1) : NEG;
        x1 = max(x1, x0);
        int sc = x1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(kFull, sc, off);
            if (lane >= off) sc = max(sc, y);
        }
        int carry = __shfl_up_sync(kFull, sc, 1);
        if (lane == 0) carry = NEG;
        x0 = max(x0, carry);
        x1 = max(x1, x0);
        int H0 = (i0 <= n) ? x0 + kPoaIns * (2 * lane) : 0;
        int H1 = (i1 <= n) ? x1 + kPoaIns * (2 * lane + 1) : 0;
        if (H0 > c0) m0 = 3u;
        if (H1 > c1) m1 = 3u;
        // row outputs (by vertex id)
        *reinterpret_cast<uchar2*>(mv_r + (size_t)id * kPoaBand + 2 * lane) = make_uchar2((unsigned char)m0, (unsigned char)m1);
        srow = s_ring[warp][t & (kRing - 1)];
        srow[2 * lane] = H0; srow[2 * lane + 1] = H1;
        if (store_h) *reinterpret_cast<int2*>(h_r + (size_t)id * kPoaBand + 2 * lane) = make_int2(H0, H1);
        // best cell of the row: largest value, smallest read prefix on ties
        // one reduction: key = value * 64 + (63 - cell) -- the maximum key is the largest value and, among equals,
        // the smallest cell (scores are >= 0 and < 2^25)
        const int hv = max(H0, H1);
        const int cell = (H0 >= H1) ? 2 * lane : 2 * lane + 1;
        const int rkey = __reduce_max_sync(kFull, hv * kPoaBand + (kPoaBand - 1 - cell));
        static_assert(kPoaBand == 64, "key packing assumes 64 cells per row");
        const int rmax = rkey >> 6;
        const int wc = kPoaBand - 1 - (rkey & (kPoaBand - 1));
        const int besti = (rmax > 0) ? lo + wc : lo;
        if (lane == 0) {
            lo_r[id] = lo; bi_r[id] = besti;
            s_rid[warp][t & (kRing - 1)] = id; s_rlo[warp][t & (kRing - 1)] = lo; s_rbi[warp][t & (kRing - 1)] = besti;
        }
        if (rmax > gbest) { gbest = rmax; gt = id; gi = besti; }
        prev_lo = lo; prev_besti = besti; prev_id = id;
        __syncwarp();   // row visible in shared memory before the next vertex reads it
    }
    if (lane == 0) {
        PoaResult r;
        r.score = gbest; r.end_t = gt; r.end_i = gi; r.path_len = 0;
        r.first_t = r.first_i = r.last_t = r.last_i = -1;
        results[task_id] = r;
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
        // rows are stored by vertex id; a path mostly walks down consecutive ids (the seed chain, or the vertices one
        // earlier read added), so a window of 32 ids is re-staged only when the path jumps
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
            const int c = i - s_lo[warp][t - wbase];
            if ((unsigned)c >= (unsigned)kPoaBand) break;
            const unsigned m = s_mv[warp][(t - wbase) * kPoaBand + c];
            const unsigned kind = m & 3u, k = m >> 2;
            if (kind == 0u) break;
            if (kind == 1u) {           // match / mismatch: read base i-1 on vertex t
                if (lane == 0 && len < T.n) out[len] = PoaStep{t, i - 1};
                ++len;
                if (r.last_t < 0) { r.last_t = t; r.last_i = i - 1; }
                r.first_t = t; r.first_i = i - 1;
                if (k == 63u) break;
                t = (k == 0u) ? s_p0[warp][t - wbase] : predx[7 * (int64_t)t + k - 1];
                i -= 1;
            } else if (kind == 2u) {    // deletion: vertex skipped
                t = (k == 0u) ? s_p0[warp][t - wbase] : predx[7 * (int64_t)t + k - 1];
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
    int32_t* __restrict__ gr = (grid != nullptr && T.grid_off >= 0 && lane == 0) ? grid + T.grid_off : nullptr;
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
            const int c = i - s_lo[warp][t - wbase];
            if ((unsigned)c >= (unsigned)kPoaBand) { done = true; break; }
            const unsigned m = s_mv[warp][(t - wbase) * kPoaBand + c];
            const unsigned kind = m & 3u, k = m >> 2;
            if (kind == 0u) { done = true; break; }
            ++len;
            if (gr != nullptr && kind != 3u && (t & (kWindowGrid - 1)) == 0) gr[t / kWindowGrid] = (kind == 1u) ? i - 1 : i;
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
                      const uint8_t* rev_flags, int32_t* lo, int32_t* besti, uint8_t* moves, int32_t* hrows,
                      PoaStep* steps, PoaResult* results, cudaStream_t stream, int32_t* grid) {
    if (n_tasks <= 0) return;
    const int blocks = (n_tasks + kWarpsPerCta - 1) / kWarpsPerCta;
    poa_align_kernel<<<blocks, kWarpsPerCta * 32, 0, stream>>>(tasks, n_tasks, G, drafts, codes, rev_flags, lo, besti,
                                                               moves, hrows, results);
    poa_traceback_kernel<<<blocks, kWarpsPerCta * 32, 0, stream>>>(tasks, n_tasks, G, lo, moves, steps, results, grid);
}

}  // namespace ccs
