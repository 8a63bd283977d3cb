// arrow_fill_alpha / arrow_fill_beta: banded, column-scaled forward / backward recursion of
// every subread against its template slice (Recursor::FillAlpha / FillBeta / FillAlphaBeta,
// SURVEY.md 8a rows a8-a10; /root/reference/docs/how-does-ccs-work.md:94-96,
// /root/reference/docs/faq/revio.md:20-25 "filling out matrices at its core").
//
// Mapping: one warp octet per (read, template) pair, 16 pairs per 128-thread CTA, pairs sorted
// by length so the four octets of a warp finish together.  Per column an octet reads one
// template base + one 16-byte transition row (L1-resident), looks the emissions up in shared
// memory, and writes one 128-byte line of fp32 cells -- the kernel is HBM-write bound:
// algorithmic bytes per pair = 4*32*(J-1) cells + 8 (alpha) or 4 (beta) bytes of column info
// per column + I + J input bytes (DESIGN.md "Roofline").
//
// Row codes: a lane's four slots hold rows 32*lap + 4g + {0..3}; one aligned 32-bit word of the
// row-code array carries exactly those four codes, so a lane keeps the words of the current
// and next lap (plus one prefetched) and assembles its codes with a single byte-permute; a
// new word is loaded once every 32 columns (coalesced 32 B per octet).
#include "arrow_octet.cuh"
#include "arrow_launch.h"

namespace ccs {

namespace {

__device__ __forceinline__ void load_emissions(const ArrowBatchView& V, float* s_emm, float* s_emi) {
    for (int k = threadIdx.x; k < 36 * kEmStride; k += blockDim.x) s_emm[k] = V.em_match[k];
    for (int k = threadIdx.x; k < 17 * kEmStride; k += blockDim.x) s_emi[k] = V.em_ins[k];
    __syncthreads();
}

// codes of this lane's four slots for band start s: slots below the start slot are one lap ahead
__device__ __forceinline__ void lane_codes(const unsigned w_lo, const unsigned w_hi, const int s, const int g, int code[4]) {
    const int t = min(max((s & 31) - 4 * g, 0), 4);          // how many of the lane's slots are in the next lap
    const unsigned sel = 0x3210u | (0x4444u & ((1u << (4 * t)) - 1u));
    const unsigned cw = __byte_perm(w_lo, w_hi, sel);
    code[0] = (int)(cw & 0xffu);
    code[1] = (int)__byte_perm(cw, 0u, 0x4441u);
    code[2] = (int)__byte_perm(cw, 0u, 0x4442u);
    code[3] = (int)(cw >> 24);
}

__global__ void __launch_bounds__(128) arrow_fill_alpha_kernel(const ArrowBatchView V, const int32_t* __restrict__ order,
                                                               const int n_items) {
    __shared__ float s_emm[36 * kEmStride];
    __shared__ float s_emi[17 * kEmStride];
    load_emissions(V, s_emm, s_emi);

    const int item = blockIdx.x * 16 + (threadIdx.x >> 3);
    const int g = threadIdx.x & 7;
    int r = -1;
    if (item < n_items) r = order[item];
    DevRead rd;
    rd.J = 0; rd.I = 0; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0; rd.zmw = 0; rd.last_code = 0; rd.code_stride = 64;
    rd.active = 0;
    if (r >= 0) rd = V.reads[r];
    const bool valid = r >= 0 && rd.active && rd.J >= 2 && rd.I >= 2;
    const int J = valid ? rd.J : 0;
    const int Jc = valid ? rd.J : 2;      // clamp bound for harmless loads of idle octets
    const int I = rd.I;
    int Jmax = J;
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) Jmax = max(Jmax, __shfl_xor_sync(kFullMask, Jmax, off));

    const unsigned* __restrict__ rc32 = reinterpret_cast<const unsigned*>(V.rowcode + rd.code_off);
    const int wmax = (rd.code_stride >> 2) - 1;
    const uint8_t* __restrict__ tp = V.tpl + rd.tpl_off;
    const float4* __restrict__ tr = reinterpret_cast<const float4*>(V.trans) + (size_t)rd.zmw * 36;
    float4* __restrict__ acol = reinterpret_cast<float4*>(V.alpha) + (size_t)rd.col_off * 8 + g;
    ColInfo* __restrict__ cinfo = V.colinfo + rd.col_off;

    // column 0: alpha(0,0) = 1
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (g == 0) v[0] = 1.f;
    int s = 0, cum = 0, edge = 0, lap = 0;
    unsigned w0 = rc32[min(g, wmax)], w1 = rc32[min(8 + g, wmax)], w2 = rc32[min(16 + g, wmax)];
    if (valid) {
        acol[0] = make_float4(v[0], v[1], v[2], v[3]);
        if (g == 0) cinfo[0] = ColInfo{0, 0};
    }
    const int t0 = tp[0] & 3, t1 = tp[min(1, Jc - 1)] & 3;   // & 3: idle octets read whatever sits at the buffer start
    int t2 = tp[min(2, Jc - 1)] & 3;
    int cm = kCtxStartRow + t0;          // match/deletion context of column 1: pinned first move
    int ci = 4 * t0 + t1;                // insertion context of column 1
    float4 tr_m = tr[cm];
    float4 tr_i = tr[ci];
    float final_val = 0.f;
    int final_cum = 0;

    for (int j = 1; j < Jmax; ++j) {
        const bool alive = j < J;
        const int s_new = max(s, edge + 2 + kBandMargin - kBandW);
        const int d = s_new - s;
        if ((s_new >> 5) != lap) {       // octet-uniform, once every 32 columns
            lap = s_new >> 5;
            w0 = w1; w1 = w2;
            w2 = rc32[min(8 * (lap + 2) + g, wmax)];
        }
        int rel[4], code[4];
        lane_codes(w0, w1, s_new, g, code);
#pragma unroll
        for (int q = 0; q < 4; ++q) rel[q] = (4 * g + q - s_new) & 31;
        // prefetch next column's transition row and the template base after it
        const int ci_next = ((ci & 3) << 2) | t2;
        const float4 tr_next = tr[ci_next];
        const int t3 = tp[min(j + 2, Jc - 1)] & 3;

        octet_forward_column(v, g, d, rel, code, tr_m.x, tr_m.y, tr_i.z, tr_i.w, s_emm + cm * kEmStride,
                             s_emi + ci * kEmStride, (ci & 3) << 2);
        int edge_rel;
        bool dead;
        const int k = octet_scale_column(v, rel, edge_rel, dead);
        cum += k;
        edge = s_new + edge_rel - 1;
        s = s_new;
        if (alive) {
            acol[(size_t)j * 8] = make_float4(v[0], v[1], v[2], v[3]);
            if (g == 0) cinfo[j] = ColInfo{s_new, cum};
            if (j == J - 1) {   // alpha(I-1, J-1) lives in slot (I-1) mod 32 if it is inside the band
                const int slot = (I - 1) & 31;
                const int rrel = (slot - s_new) & 31;
                const bool inband = (s_new + rrel) == (I - 1);
                float x = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) x = (4 * g + q == slot && inband) ? v[q] : x;
                final_val = x;
                final_cum = cum;
            }
        }
        cm = ci; ci = ci_next; tr_m = tr_i; tr_i = tr_next; t2 = t3;
    }
    // gather the final cell from whichever lane owns it, then the pinned last match
    const float fv = octet_max(final_val);
    if (valid) {
        if (g == 0) {
            const int ctxl = 4 * (tp[J - 2] & 3) + (tp[J - 1] & 3);
            const double a = (double)fv * (double)s_emm[(kCtxEndRow + ctxl) * kEmStride + rd.last_code];
            const double base = (a > 0.0) ? log(a) + 0.6931471805599453094 * (double)final_cum : -INFINITY;
            V.base_ll[r] = base;
            V.ll_alpha[r] = base - (double)I * V.log_cw;
        }
    } else if (r >= 0 && g == 0) {
        V.base_ll[r] = -INFINITY;
        V.ll_alpha[r] = -INFINITY;
    }
}

__global__ void __launch_bounds__(128) arrow_fill_beta_kernel(const ArrowBatchView V, const int32_t* __restrict__ order,
                                                              const int n_items) {
    __shared__ float s_emm[36 * kEmStride];
    __shared__ float s_emi[17 * kEmStride];
    load_emissions(V, s_emm, s_emi);

    const int item = blockIdx.x * 16 + (threadIdx.x >> 3);
    const int g = threadIdx.x & 7;
    int r = -1;
    if (item < n_items) r = order[item];
    DevRead rd;
    rd.J = 0; rd.I = 0; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0; rd.zmw = 0; rd.last_code = 0; rd.first_code = 0;
    rd.code_stride = 64; rd.active = 0;
    if (r >= 0) rd = V.reads[r];
    const bool valid = r >= 0 && rd.active && rd.J >= 2 && rd.I >= 2;
    const int J = valid ? rd.J : 0;
    const int I = rd.I;
    int Jmax = J;
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) Jmax = max(Jmax, __shfl_xor_sync(kFullMask, Jmax, off));

    // second copy of the row codes, shifted by one row: word k holds the codes of rows 4k+1 .. 4k+4
    const unsigned* __restrict__ rc32 = reinterpret_cast<const unsigned*>(V.rowcode + rd.code_off + rd.code_stride);
    const int wmax = (rd.code_stride >> 2) - 1;
    const uint8_t* __restrict__ tp = V.tpl + rd.tpl_off;
    const float4* __restrict__ tr = reinterpret_cast<const float4*>(V.trans) + (size_t)rd.zmw * 36;
    float4* __restrict__ bcol = reinterpret_cast<float4*>(V.beta) + (size_t)rd.col_off * 8 + g;
    const ColInfo* __restrict__ cinfo = V.colinfo + rd.col_off;
    int32_t* __restrict__ bexp = V.beta_exp + rd.col_off;

    // The warp walks columns from (Jmax-1) down to 1; an octet is alive once j <= J-1.
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    int s_next = 0, cum = 0, lap = 0;
    int s_cur = 0;   // band start of column j (prefetched)
    int t_hi = 0, t_lo = 0, t_lo2 = 0;   // template bases j, j-1, j-2
    unsigned w0 = 0x30303030u, w1 = 0x30303030u, wm = 0x30303030u;   // laps L, L+1, L-1 (sentinels)
    float4 tr_c = make_float4(0.f, 0.f, 0.f, 0.f), tr_p = tr_c;
    float first_val = 0.f;
    int first_cum = 0;
    bool started = false;

    for (int j = Jmax - 1; j >= 1; --j) {
        const bool alive = j <= J - 1;
        if (alive && !started) {
            // prologue at j == J-1
            s_cur = cinfo[j].start;
            s_next = s_cur;
            t_hi = tp[j] & 3; t_lo = tp[j - 1] & 3; t_lo2 = tp[max(j - 2, 0)] & 3;
            tr_c = tr[4 * t_lo + t_hi];
            tr_p = tr[4 * t_lo2 + t_lo];
            lap = s_cur >> 5;
            w0 = rc32[min(8 * lap + g, wmax)];
            w1 = rc32[min(8 * (lap + 1) + g, wmax)];
            wm = rc32[min(8 * max(lap - 1, 0) + g, wmax)];
        }
        if (alive && (s_cur >> 5) != lap) {   // moved up into the previous lap
            lap = s_cur >> 5;
            w1 = w0; w0 = wm;
            wm = rc32[min(8 * max(lap - 1, 0) + g, wmax)];
        }
        const int ci = 4 * t_lo + t_hi;              // context of column j (= match context of j+1)
        const int d = s_next - s_cur;
        int rel[4], code1[4];
        lane_codes(w0, w1, s_cur, g, code1);
#pragma unroll
        for (int q = 0; q < 4; ++q) rel[q] = (4 * g + q - s_cur) & 31;
        // prefetch for column j-1
        const int s_prev = (alive && j >= 2) ? cinfo[j - 1].start : s_cur;
        const int t_lo3 = (alive && j >= 3) ? (tp[j - 3] & 3) : 0;
        const float4 tr_pp = alive ? tr[4 * t_lo3 + t_lo2] : tr_p;

        float A[4], G[4];
        octet_backward_terms(v, g, d, rel, code1, tr_c.x, tr_c.y, tr_c.z, tr_c.w, s_emm + ci * kEmStride,
                             s_emi + ci * kEmStride, (ci & 3) << 2, A, G);
        if (alive && !started) {
            // last column: beta(I-1, J-1) = pinned last match; rows above it by insertions
            const float endv = s_emm[(kCtxEndRow + ci) * kEmStride + rd.last_code];
#pragma unroll
            for (int q = 0; q < 4; ++q) A[q] = (s_cur + rel[q] == I - 1) ? endv : 0.f;
        }
        if (!alive) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { A[q] = 0.f; G[q] = 0.f; }
        }
        octet_backward_scan(A, G, g, v);
        int edge_rel;
        bool dead;
        const int k = octet_scale_column(v, rel, edge_rel, dead);
        if (alive) {
            cum += k;
            bcol[(size_t)j * 8] = make_float4(v[0], v[1], v[2], v[3]);
            if (g == 0) bexp[j] = cum;
            if (j == 1) {   // beta(1,1) lives in slot 1 if row 1 is inside the band
                const int rrel = (1 - s_cur) & 31;
                const bool inband = (s_cur + rrel) == 1;
                float x = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) x = (4 * g + q == 1 && inband) ? v[q] : x;
                first_val = x;
                first_cum = cum;
            }
            started = true;
            s_next = s_cur; s_cur = s_prev;
            t_hi = t_lo; t_lo = t_lo2; t_lo2 = t_lo3;
            tr_c = tr_p; tr_p = tr_pp;
        }
    }
    const float fv = octet_max(first_val);
    if (valid) {
        if (g == 0) {
            const double b = (double)fv * (double)s_emm[(kCtxStartRow + (tp[0] & 3)) * kEmStride + rd.first_code];
            const double lb = (b > 0.0) ? log(b) + 0.6931471805599453094 * (double)first_cum - (double)I * V.log_cw : -INFINITY;
            V.ll_beta[r] = lb;
            const double la = V.ll_alpha[r];
            int st = 0;  // CCS_READ_VALID
            if (!(la > -INFINITY) || !(lb > -INFINITY)) st = 3;                 // CCS_READ_DEAD
            else if (!(fabs(1.0 - la / lb) <= V.ab_tol)) st = 1;               // CCS_READ_ALPHA_BETA_MISMATCH
            V.status[r] = st;
        }
    } else if (r >= 0 && g == 0) {
        V.ll_beta[r] = -INFINITY;
        V.status[r] = (rd.active && (rd.J < 2 || rd.I < 2)) ? 2 : (rd.active ? 3 : 4);
    }
}


// ---------------------------------------------------------------------------------------------
// Generalised lane mapping: CPL cells per lane, LPP = 32 / CPL lanes per pair, CPL pairs per warp.
// More cells per lane amortise the per-column bookkeeping (band slide, prefetch, reductions, loop)
// and shorten the ring scan; fewer lanes per pair need more pairs in flight to fill the machine.
// The memory layout is the same for every CPL (slot = row mod 32 inside the 128-B column line).
// ---------------------------------------------------------------------------------------------
template <int LPP>
__device__ __forceinline__ float shfl_grp(float x, int src) { return __shfl_sync(kFullMask, x, src, LPP); }

template <int CPL>
struct LaneWords {   // row-code words of this lane for laps L, L+1 and (prefetched) L+2 / L-1
    unsigned w[3][CPL / 4];
};

template <int CPL>
__device__ __forceinline__ void lane_codes_n(const unsigned* lo, const unsigned* hi, const int s, const int g, int code[CPL]) {
#pragma unroll
    for (int k = 0; k < CPL / 4; ++k) {
        const int t = min(max((s & 31) - (CPL * g + 4 * k), 0), 4);
        const unsigned sel = 0x3210u | (0x4444u & ((1u << (4 * t)) - 1u));
        const unsigned cw = __byte_perm(lo[k], hi[k], sel);
        code[4 * k + 0] = (int)(cw & 0xffu);
        code[4 * k + 1] = (int)__byte_perm(cw, 0u, 0x4441u);
        code[4 * k + 2] = (int)__byte_perm(cw, 0u, 0x4442u);
        code[4 * k + 3] = (int)(cw >> 24);
    }
}

// column normalisation + leading edge over a group of LPP lanes
template <int CPL>
__device__ __forceinline__ int group_scale_column(float v[CPL], const int rel0, int& edge_rel) {
    constexpr int LPP = 32 / CPL;
    const float thr = __uint_as_float((unsigned)(127 + kEdgeLog2) << 23);
    float mx = 0.f;
    int er = 0;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        mx = fmaxf(mx, v[q]);
        const int rl = (rel0 + q) & 31;
        er = (v[q] >= thr) ? max(er, rl + 1) : er;
    }
    unsigned key = (__float_as_uint(mx) & 0xffff0000u) | (unsigned)er;
#pragma unroll
    for (int off = 1; off < LPP; off <<= 1) key = __vmaxu2(key, __shfl_xor_sync(kFullMask, key, off, LPP));
    edge_rel = (int)(key & 0xffffu);
    const int ebits = (int)((key >> 23) & 255u);
    const int k = ((key >> 16) == 0u) ? 0 : ebits - 127;
    const float sc = __uint_as_float((unsigned)(127 - k) << 23);
#pragma unroll
    for (int q = 0; q < CPL; ++q) v[q] *= sc;
    return k;
}

template <int LPP>
__device__ __forceinline__ float group_max(float x) {
#pragma unroll
    for (int off = 1; off < LPP; off <<= 1) x = fmaxf(x, __shfl_xor_sync(kFullMask, x, off, LPP));
    return x;
}

template <int CPL>
__global__ void __launch_bounds__(128) arrow_fill_alpha_n_kernel(const ArrowBatchView V, const int32_t* __restrict__ order,
                                                                 const int n_items) {
    constexpr int LPP = 32 / CPL;          // lanes per pair
    constexpr int PPC = 128 / LPP;         // pairs per CTA
    constexpr int NW = CPL / 4;            // row-code words per lane and lap
    __shared__ float s_emm[36 * kEmStride];
    __shared__ float s_emi[17 * kEmStride];
    load_emissions(V, s_emm, s_emi);

    const int item = blockIdx.x * PPC + (threadIdx.x / LPP);
    const int g = threadIdx.x % LPP;
    int r = -1;
    if (item < n_items) r = order[item];
    DevRead rd;
    rd.J = 0; rd.I = 0; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0; rd.zmw = 0; rd.last_code = 0; rd.code_stride = 256;
    rd.active = 0;
    if (r >= 0) rd = V.reads[r];
    const bool valid = r >= 0 && rd.active && rd.J >= 2 && rd.I >= 2;
    const int J = valid ? rd.J : 0;
    const int Jc = valid ? rd.J : 2;
    const int I = rd.I;
    int Jmax = J;
#pragma unroll
    for (int off = LPP; off < 32; off <<= 1) Jmax = max(Jmax, __shfl_xor_sync(kFullMask, Jmax, off));

    const unsigned* __restrict__ rc32 = reinterpret_cast<const unsigned*>(V.rowcode + rd.code_off);
    const int wmax = (rd.code_stride >> 2) - 1;
    const uint8_t* __restrict__ tp = V.tpl + rd.tpl_off;
    const float4* __restrict__ tr = reinterpret_cast<const float4*>(V.trans) + (size_t)rd.zmw * 36;
    float4* __restrict__ acol = reinterpret_cast<float4*>(V.alpha) + (size_t)rd.col_off * 8 + NW * g;
    ColInfo* __restrict__ cinfo = V.colinfo + rd.col_off;

    float v[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) v[q] = 0.f;
    if (g == 0) v[0] = 1.f;
    int s = 0, cum = 0, edge = 0, lap = 0;
    unsigned w0[NW], w1[NW], w2[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        w0[k] = rc32[min(NW * g + k, wmax)];
        w1[k] = rc32[min(8 + NW * g + k, wmax)];
        w2[k] = rc32[min(16 + NW * g + k, wmax)];
    }
    if (valid) {
#pragma unroll
        for (int k = 0; k < NW; ++k) acol[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        if (g == 0) cinfo[0] = ColInfo{0, 0};
    }
    const int t0 = tp[0] & 3, t1 = tp[min(1, Jc - 1)] & 3;   // & 3: idle octets read whatever sits at the buffer start
    int t2 = tp[min(2, Jc - 1)] & 3;
    int cm = kCtxStartRow + t0;
    int ci = 4 * t0 + t1;
    float4 tr_m = tr[cm];
    float4 tr_i = tr[ci];
    float final_val = 0.f;
    int final_cum = 0;

    for (int j = 1; j < Jmax; ++j) {
        const bool alive = j < J;
        const int s_new = max(s, edge + 2 + kBandMargin - kBandW);
        const int d = s_new - s;
        if ((s_new >> 5) != lap) {
            lap = s_new >> 5;
#pragma unroll
            for (int k = 0; k < NW; ++k) { w0[k] = w1[k]; w1[k] = w2[k]; w2[k] = rc32[min(8 * (lap + 2) + NW * g + k, wmax)]; }
        }
        int code[CPL];
        lane_codes_n<CPL>(w0, w1, s_new, g, code);
        const int rel0 = (CPL * g - s_new) & 31;
        const int ci_next = ((ci & 3) << 2) | t2;
        const float4 tr_next = tr[ci_next];
        const int t3 = tp[min(j + 2, Jc - 1)] & 3;

        const float* __restrict__ emm_row = s_emm + cm * kEmStride;
        const float* __restrict__ emi_row = s_emi + ci * kEmStride;
        const int cb4 = (ci & 3) << 2;
        const float M = tr_m.x, D = tr_m.y, B = tr_i.z, S = tr_i.w;
        // terms + serial scan inside the lane
        float A[CPL], G[CPL];
        float up = shfl_grp<LPP>(v[CPL - 1], (g + LPP - 1) % LPP);
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int rl = (rel0 + q) & 31;
            const int rdd = rl + d;
            const float old = v[q];
            const float pv = (rdd < 32) ? old : 0.f;
            const float uv = ((unsigned)(rdd - 1) < 32u) ? up : 0.f;
            const float a = fmaf(M, ldtab(emm_row, code[q]) * uv, D * pv);
            float gi = ldtab(emi_row, code[q]) * (((code[q] & 12) == cb4) ? B : S);
            gi = (rl == 0) ? 0.f : gi;
            if (q == 0) { A[0] = a; G[0] = gi; }
            else { A[q] = fmaf(gi, A[q - 1], a); G[q] = gi * G[q - 1]; }
            up = old;
        }
        float At = A[CPL - 1], Gt = G[CPL - 1];
#pragma unroll
        for (int off = 1; off < LPP; off <<= 1) {
            const float As = shfl_grp<LPP>(At, (g + LPP - off) % LPP);
            const float Gs = shfl_grp<LPP>(Gt, (g + LPP - off) % LPP);
            At = fmaf(Gt, As, At);
            Gt *= Gs;
        }
        const float x = shfl_grp<LPP>(At, (g + LPP - 1) % LPP);
#pragma unroll
        for (int q = 0; q < CPL; ++q) v[q] = fmaf(G[q], x, A[q]);

        int edge_rel;
        const int k = group_scale_column<CPL>(v, rel0, edge_rel);
        cum += k;
        edge = s_new + edge_rel - 1;
        s = s_new;
        if (alive) {
#pragma unroll
            for (int kk = 0; kk < NW; ++kk)
                acol[(size_t)j * 8 + kk] = make_float4(v[4 * kk], v[4 * kk + 1], v[4 * kk + 2], v[4 * kk + 3]);
            if (g == 0) cinfo[j] = ColInfo{s_new, cum};
            if (j == J - 1) {
                const int slot = (I - 1) & 31;
                const int rrel = (slot - s_new) & 31;
                const bool inband = (s_new + rrel) == (I - 1);
                float xx = 0.f;
#pragma unroll
                for (int q = 0; q < CPL; ++q) xx = (CPL * g + q == slot && inband) ? v[q] : xx;
                final_val = xx;
                final_cum = cum;
            }
        }
        cm = ci; ci = ci_next; tr_m = tr_i; tr_i = tr_next; t2 = t3;
    }
    const float fv = group_max<LPP>(final_val);
    if (valid) {
        if (g == 0) {
            const int ctxl = 4 * (tp[J - 2] & 3) + (tp[J - 1] & 3);
            const double a = (double)fv * (double)s_emm[(kCtxEndRow + ctxl) * kEmStride + rd.last_code];
            const double base = (a > 0.0) ? log(a) + 0.6931471805599453094 * (double)final_cum : -INFINITY;
            V.base_ll[r] = base;
            V.ll_alpha[r] = base - (double)I * V.log_cw;
        }
    } else if (r >= 0 && g == 0) {
        V.base_ll[r] = -INFINITY;
        V.ll_alpha[r] = -INFINITY;
    }
}

template <int CPL>
__global__ void __launch_bounds__(128) arrow_fill_beta_n_kernel(const ArrowBatchView V, const int32_t* __restrict__ order,
                                                                const int n_items) {
    constexpr int LPP = 32 / CPL;
    constexpr int PPC = 128 / LPP;
    constexpr int NW = CPL / 4;
    __shared__ float s_emm[36 * kEmStride];
    __shared__ float s_emi[17 * kEmStride];
    load_emissions(V, s_emm, s_emi);

    const int item = blockIdx.x * PPC + (threadIdx.x / LPP);
    const int g = threadIdx.x % LPP;
    int r = -1;
    if (item < n_items) r = order[item];
    DevRead rd;
    rd.J = 0; rd.I = 0; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0; rd.zmw = 0; rd.last_code = 0; rd.first_code = 0;
    rd.code_stride = 256; rd.active = 0;
    if (r >= 0) rd = V.reads[r];
    const bool valid = r >= 0 && rd.active && rd.J >= 2 && rd.I >= 2;
    const int J = valid ? rd.J : 0;
    const int I = rd.I;
    int Jmax = J;
#pragma unroll
    for (int off = LPP; off < 32; off <<= 1) Jmax = max(Jmax, __shfl_xor_sync(kFullMask, Jmax, off));

    const unsigned* __restrict__ rc32 = reinterpret_cast<const unsigned*>(V.rowcode + rd.code_off + rd.code_stride);
    const int wmax = (rd.code_stride >> 2) - 1;
    const uint8_t* __restrict__ tp = V.tpl + rd.tpl_off;
    const float4* __restrict__ tr = reinterpret_cast<const float4*>(V.trans) + (size_t)rd.zmw * 36;
    float4* __restrict__ bcol = reinterpret_cast<float4*>(V.beta) + (size_t)rd.col_off * 8 + NW * g;
    const ColInfo* __restrict__ cinfo = V.colinfo + rd.col_off;
    int32_t* __restrict__ bexp = V.beta_exp + rd.col_off;

    float v[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) v[q] = 0.f;
    int s_next = 0, cum = 0, lap = 0, s_cur = 0;
    int t_hi = 0, t_lo = 0, t_lo2 = 0;
    unsigned w0[NW], w1[NW], wm[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) { w0[k] = 0x30303030u; w1[k] = 0x30303030u; wm[k] = 0x30303030u; }
    float4 tr_c = make_float4(0.f, 0.f, 0.f, 0.f), tr_p = tr_c;
    float first_val = 0.f;
    int first_cum = 0;
    bool started = false;

    for (int j = Jmax - 1; j >= 1; --j) {
        const bool alive = j <= J - 1;
        if (alive && !started) {
            s_cur = cinfo[j].start;
            s_next = s_cur;
            t_hi = tp[j] & 3; t_lo = tp[j - 1] & 3; t_lo2 = tp[max(j - 2, 0)] & 3;
            tr_c = tr[4 * t_lo + t_hi];
            tr_p = tr[4 * t_lo2 + t_lo];
            lap = s_cur >> 5;
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                w0[k] = rc32[min(8 * lap + NW * g + k, wmax)];
                w1[k] = rc32[min(8 * (lap + 1) + NW * g + k, wmax)];
                wm[k] = rc32[min(8 * max(lap - 1, 0) + NW * g + k, wmax)];
            }
        }
        if (alive && (s_cur >> 5) != lap) {
            lap = s_cur >> 5;
#pragma unroll
            for (int k = 0; k < NW; ++k) { w1[k] = w0[k]; w0[k] = wm[k]; wm[k] = rc32[min(8 * max(lap - 1, 0) + NW * g + k, wmax)]; }
        }
        const int ci = 4 * t_lo + t_hi;
        const int d = s_next - s_cur;
        int code1[CPL];
        lane_codes_n<CPL>(w0, w1, s_cur, g, code1);
        const int rel0 = (CPL * g - s_cur) & 31;
        const int s_prev = (alive && j >= 2) ? cinfo[j - 1].start : s_cur;
        const int t_lo3 = (alive && j >= 3) ? (tp[j - 3] & 3) : 0;
        const float4 tr_pp = alive ? tr[4 * t_lo3 + t_lo2] : tr_p;

        const float* __restrict__ emm_row = s_emm + ci * kEmStride;
        const float* __restrict__ emi_row = s_emi + ci * kEmStride;
        const int cb4 = (ci & 3) << 2;
        const float M = tr_c.x, D = tr_c.y, B = tr_c.z, S = tr_c.w;
        const bool init_col = alive && !started;
        const float endv = s_emm[(kCtxEndRow + ci) * kEmStride + rd.last_code];
        float A[CPL], G[CPL];
        float dn = shfl_grp<LPP>(v[0], (g + 1) % LPP);
#pragma unroll
        for (int q = CPL - 1; q >= 0; --q) {
            const int rl = (rel0 + q) & 31;
            const float old = v[q];
            const float nx = (rl >= d) ? old : 0.f;
            const float nd = (rl >= d - 1 && rl <= d + 30) ? dn : 0.f;
            float a = fmaf(M, ldtab(emm_row, code1[q]) * nd, D * nx);
            if (init_col) a = (s_cur + rl == I - 1) ? endv : 0.f;
            float gi = ldtab(emi_row, code1[q]) * (((code1[q] & 12) == cb4) ? B : S);
            gi = (rl == 31) ? 0.f : gi;
            if (!alive) { a = 0.f; gi = 0.f; }
            if (q == CPL - 1) { A[q] = a; G[q] = gi; }
            else { A[q] = fmaf(gi, A[q + 1], a); G[q] = gi * G[q + 1]; }
            dn = old;
        }
        float At = A[0], Gt = G[0];
#pragma unroll
        for (int off = 1; off < LPP; off <<= 1) {
            const float As = shfl_grp<LPP>(At, (g + off) % LPP);
            const float Gs = shfl_grp<LPP>(Gt, (g + off) % LPP);
            At = fmaf(Gt, As, At);
            Gt *= Gs;
        }
        const float x = shfl_grp<LPP>(At, (g + 1) % LPP);
#pragma unroll
        for (int q = 0; q < CPL; ++q) v[q] = fmaf(G[q], x, A[q]);
        int edge_rel;
        const int k = group_scale_column<CPL>(v, rel0, edge_rel);
        if (alive) {
            cum += k;
#pragma unroll
            for (int kk = 0; kk < NW; ++kk)
                bcol[(size_t)j * 8 + kk] = make_float4(v[4 * kk], v[4 * kk + 1], v[4 * kk + 2], v[4 * kk + 3]);
            if (g == 0) bexp[j] = cum;
            if (j == 1) {
                const int rrel = (1 - s_cur) & 31;
                const bool inband = (s_cur + rrel) == 1;
                float xx = 0.f;
#pragma unroll
                for (int q = 0; q < CPL; ++q) xx = (CPL * g + q == 1 && inband) ? v[q] : xx;
                first_val = xx;
                first_cum = cum;
            }
            started = true;
            s_next = s_cur; s_cur = s_prev;
            t_hi = t_lo; t_lo = t_lo2; t_lo2 = t_lo3;
            tr_c = tr_p; tr_p = tr_pp;
        }
    }
    const float fv = group_max<LPP>(first_val);
    if (valid) {
        if (g == 0) {
            const double b = (double)fv * (double)s_emm[(kCtxStartRow + (tp[0] & 3)) * kEmStride + rd.first_code];
            const double lb = (b > 0.0) ? log(b) + 0.6931471805599453094 * (double)first_cum - (double)I * V.log_cw : -INFINITY;
            V.ll_beta[r] = lb;
            const double la = V.ll_alpha[r];
            int st = 0;
            if (!(la > -INFINITY) || !(lb > -INFINITY)) st = 3;
            else if (!(fabs(1.0 - la / lb) <= V.ab_tol)) st = 1;
            V.status[r] = st;
        }
    } else if (r >= 0 && g == 0) {
        V.ll_beta[r] = -INFINITY;
        V.status[r] = (rd.active && (rd.J < 2 || rd.I < 2)) ? 2 : (rd.active ? 3 : 4);
    }
}

}  // namespace

// cells_per_lane: 4 (octet kernels above), 8, 16 or 32
void launch_fill_alpha(const ArrowBatchView& V, const int32_t* order, int n_items, cudaStream_t stream, int cells_per_lane) {
    if (n_items <= 0) return;
    switch (cells_per_lane) {
        case 8: arrow_fill_alpha_n_kernel<8><<<(n_items + 31) / 32, 128, 0, stream>>>(V, order, n_items); break;
        case 16: arrow_fill_alpha_n_kernel<16><<<(n_items + 63) / 64, 128, 0, stream>>>(V, order, n_items); break;
        case 32: arrow_fill_alpha_n_kernel<32><<<(n_items + 127) / 128, 128, 0, stream>>>(V, order, n_items); break;
        default: arrow_fill_alpha_kernel<<<(n_items + 15) / 16, 128, 0, stream>>>(V, order, n_items);
    }
}

void launch_fill_beta(const ArrowBatchView& V, const int32_t* order, int n_items, cudaStream_t stream, int cells_per_lane) {
    if (n_items <= 0) return;
    switch (cells_per_lane) {
        case 8: arrow_fill_beta_n_kernel<8><<<(n_items + 31) / 32, 128, 0, stream>>>(V, order, n_items); break;
        case 16: arrow_fill_beta_n_kernel<16><<<(n_items + 63) / 64, 128, 0, stream>>>(V, order, n_items); break;
        case 32: arrow_fill_beta_n_kernel<32><<<(n_items + 127) / 128, 128, 0, stream>>>(V, order, n_items); break;
        default: arrow_fill_beta_kernel<<<(n_items + 15) / 16, 128, 0, stream>>>(V, order, n_items);
    }
}

}  // namespace ccs
