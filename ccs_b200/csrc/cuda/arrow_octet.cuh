// Warp-octet primitives of the Arrow recursion (sm_100a).
//
// One (read, template) pair is owned by an OCTET = 8 consecutive lanes of a warp; each lane
// keeps 4 consecutive band slots in registers, so an octet holds one 32-row band column and a
// warp advances four independent pairs per instruction.  Slot k (0..31) always holds DP row
// i with i mod 32 == k ("rotated band"): when the band slides down by d rows, d slots are
// recycled in place and no cell moves between lanes.
//
//   column step (forward):  a_i = C_i + g_i * a_{i-1},
//       C_i = mm[code_i] * prev_{i-1} + D * prev_i             (match + deletion from column j-1)
//       g_i = gg[code_i]                                       (branch / stick inside column j)
//   (mm, gg = the folded per-ZMW factors below; the generic evaluator of arrow_score.cu computes the same
//   products from the unfolded tables)
//   The in-column first-order recurrence is solved by a 4-cell serial scan per lane plus a
//   3-step Kogge-Stone ring scan over the octet (pairs (g, C) under (g2 g1, g2 C1 + C2)); the
//   band start is a slot with g = 0, which is what makes the ring scan exact.
//
// Restates Recursor::FillAlpha / FillBeta / ExtendAlpha of the reference design
// (behaviour /root/reference/docs/how-does-ccs-work.md:94-96).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "arrow_device.h"

namespace ccs {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kEmStride = 16;
constexpr int kCtxStartRow = 16;
constexpr int kCtxEndRow = 20;
constexpr int kC4Sentinel = 48;   // row codes are stored pre-multiplied by 4 (byte offset into a table row)

// table row lookup by pre-scaled code
__device__ __forceinline__ float ldtab(const float* __restrict__ row, const int c4) {
    return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(row) + c4);
}

__device__ __forceinline__ float shfl_oct(float x, int src_lane_in_octet) {
    return __shfl_sync(kFullMask, x, src_lane_in_octet, 8);
}
__device__ __forceinline__ int shfl_oct_i(int x, int src_lane_in_octet) {
    return __shfl_sync(kFullMask, x, src_lane_in_octet, 8);
}

// ---- folded per-ZMW factor tables in shared memory -------------------------------------------
// The model (DESIGN.md "Arrow model"): mm[ctx][code] = fl32(em_match * match), gg[ctx][code] =
// fl32(em_ins * (cognate ? branch : stick)) -- one fp32 product each, identical to the oracle's
// Tables::mm / Tables::gg.  Kernels build the rows they need once per CTA (a CTA works on one ZMW).
constexpr int kDSlot = 13;         // unused code slot: carries {deletion transition of the match context, 0}

__device__ __forceinline__ float2 lds_f2(const unsigned addr) {
    float2 r;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ float lds_f1(const unsigned addr) {
    float r;
    asm("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}

// {match factor of context cm, insertion factor of context ci} for one emission code
__device__ __forceinline__ float2 folded_entry(const float* __restrict__ em_match, const float* __restrict__ em_ins,
                                               const float4* __restrict__ tr, const int cm, const int ci, const int code) {
    const float4 tm = tr[cm], ti = tr[ci];
    if (code == kDSlot) return make_float2(tm.y, 0.f);
    const float mm = __fmul_rn(em_match[cm * kEmStride + code], tm.x);
    const bool cognate = (code & 3) == (ci & 3);
    const float gg = __fmul_rn(em_ins[ci * kEmStride + code], cognate ? ti.z : ti.w);
    return make_float2(mm, gg);
}

// A loop-invariant value the compiler must keep in a register: ptxas otherwise re-derives lane constants from %tid
// inside hot loops (S2R + integer ops every iteration).  An identity shuffle is opaque to it.  Call converged.
__device__ __forceinline__ int pinned(const int x) { return __shfl_sync(0xffffffffu, x, (int)(threadIdx.x & 31)); }

// All shuffles below use the full-warp mask: callers must keep the warp converged around them
// (idle octets run the same instruction stream on zeros).

// Forward column, step 1: per-cell terms.  v[] = previous column (scaled).  rel[q] =
// (slot - s_new) & 31 is the row's position in the new band, d = s_new - s_prev.
__device__ __forceinline__ void octet_forward_terms(const float v[4], const int g, const int d, const int rel[4],
                                                    const int code[4], const float M, const float D, const float B,
                                                    const float S, const float* __restrict__ emm_row,
                                                    const float* __restrict__ emi_row, const int cur_base4 /* (base of the context's current template base) << 2 */,
                                                    float A[4], float G[4]) {
    float up[4];   // previous column one row up: slot - 1
    up[0] = shfl_oct(v[3], (g + 7) & 7);
    up[1] = v[0]; up[2] = v[1]; up[3] = v[2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int rd = rel[q] + d;                 // position of this row in the previous band
        const float pv = (rd < 32) ? v[q] : 0.f;   // row inside previous band
        const float uv = (rd >= 1 && rd <= 32) ? up[q] : 0.f;
        const float em = ldtab(emm_row, code[q]);
        A[q] = fmaf(M, em * uv, D * pv);
        const float gi = ldtab(emi_row, code[q]) * (((code[q] & 12) == cur_base4) ? B : S);
        G[q] = (rel[q] == 0) ? 0.f : gi;           // band start: no in-band predecessor
    }
}

// Forward column, step 2: solve a_i = A_i + G_i * a_{i-1} around the ring; result in v[].
__device__ __forceinline__ void octet_forward_scan(float A[4], float G[4], const int g, float v[4]) {
    A[1] = fmaf(G[1], A[0], A[1]); G[1] *= G[0];
    A[2] = fmaf(G[2], A[1], A[2]); G[2] *= G[1];
    A[3] = fmaf(G[3], A[2], A[3]); G[3] *= G[2];
    float At = A[3], Gt = G[3];
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
        const float As = shfl_oct(At, (g + 8 - off) & 7);
        const float Gs = shfl_oct(Gt, (g + 8 - off) & 7);
        At = fmaf(Gt, As, At);
        Gt *= Gs;
    }
    const float x = shfl_oct(At, (g + 7) & 7);     // final value of slot 4g-1
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = fmaf(G[q], x, A[q]);
}

__device__ __forceinline__ float octet_max(float x) {
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) x = fmaxf(x, __shfl_xor_sync(kFullMask, x, off, 8));
    return x;
}

__device__ __forceinline__ float octet_sum(float x) {
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) x += __shfl_xor_sync(kFullMask, x, off, 8);
    return x;
}

}  // namespace ccs
