// arrow_fill_alpha / arrow_fill_beta: banded, column-scaled forward / backward recursion of
// every subread against its template slice (Recursor::FillAlpha / FillBeta / FillAlphaBeta,
// SURVEY.md 8a rows a8-a10; /root/reference/docs/how-does-ccs-work.md:94-96,
// /root/reference/docs/faq/revio.md:20-25 "filling out matrices at its core").
//
// Mapping: one CTA (128 threads) per group of <= 16 reads of ONE ZMW, one warp octet (8 lanes x 4
// cells) per (read, template) pair.  Per column an octet reads one template byte, looks its four
// cells' folded factors up in shared memory (one 8-byte load per cell) and writes one 128-byte line
// of fp32 cells -- the kernel is HBM-write bound by construction: algorithmic bytes per pair =
// 4*32*(J-1) cells + 8 (alpha) or 4 (beta) bytes of column info per column + I + J input bytes
// (DESIGN.md "Roofline").
//
// What keeps the per-column instruction count low (DESIGN.md "Kernels"):
//  * the band start only moves in steps of 4 rows (spec: quantised slide), so band bookkeeping is
//    per LANE, not per cell: the lane that leaves the band at the bottom re-enters at the top, one
//    predicate zeroes it, one predicate marks the band-start lane, and the leading-edge test looks
//    at three cells of the top lane only;
//  * the CTA's ZMW has its transition factors folded into the emission tables once per CTA:
//    alpha indexes one row per template TRINUCLEOTIDE (t[j-2], t[j-1], t[j]) holding
//    {match factor of context (t[j-2],t[j-1]), insertion factor of context (t[j-1],t[j])} per
//    code, plus the deletion transition in a spare code slot -- the template byte IS the row index;
//  * row codes: a lane's four slots are one aligned 32-bit word of the row-code array.
#include "arrow_octet.cuh"
#include "arrow_launch.h"

namespace ccs {

namespace {

constexpr int kT3Rows = 80;        // 64 trinucleotide rows + 16 pinned-first-move rows (64 + 4*t0 + t1)
constexpr int kT2Rows = 16;        // beta: one row per dinucleotide context

__device__ __forceinline__ unsigned vmax_oct(unsigned key) {
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) key = __vmaxu2(key, __shfl_xor_sync(kFullMask, key, off, 8));
    return key;
}

template <int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) arrow_fill_alpha_kernel(const ArrowBatchView V, const int32_t* __restrict__ order,
                                                               const int n_items) {
    __shared__ __align__(128) float2 s_t3[kT3Rows * kEmStride];
    {
        const int r0 = order[blockIdx.x * 16];                 // a group's first slot is always a real read
        const int zmw = (r0 >= 0) ? V.reads[r0].zmw : 0;
        const float4* __restrict__ tr = reinterpret_cast<const float4*>(V.trans) + (size_t)zmw * 36;
        for (int idx = threadIdx.x; idx < kT3Rows * kEmStride; idx += blockDim.x) {
            const int row = idx >> 4, code = idx & 15;
            const int cm = (row < 64) ? (row >> 2) : kCtxStartRow + ((row - 64) >> 2);
            const int ci = row & 15;
            s_t3[idx] = folded_entry(V.em_match, V.em_ins, tr, cm, ci, code);
        }
    }
    __syncthreads();

    const int item = blockIdx.x * 16 + (threadIdx.x >> 3);
    const int g = pinned(threadIdx.x & 7);
    int r = -1;
    if (item < n_items) r = order[item];
    DevRead rd;
    rd.J = 0; rd.I = 0; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0; rd.zmw = 0; rd.last_code = 0; rd.code_stride = 64;
    rd.active = 0;
    if (r >= 0) rd = V.reads[r];
    const bool valid = r >= 0 && rd.active && rd.J >= 2 && rd.I >= 2;
    const int J = pinned(valid ? rd.J : 0);
    const int I = rd.I;
    const int Jmax = __reduce_max_sync(kFullMask, J);     // warp-uniform trip count (a reduction: provably convergent loop)

    const unsigned* __restrict__ rc32 = reinterpret_cast<const unsigned*>(V.rowcode + rd.code_off);
    const int wmax = (rd.code_stride >> 2) - 1;
    const uint8_t* __restrict__ tp = V.tpl + rd.tpl_off;
    float4* __restrict__ acol = reinterpret_cast<float4*>(V.alpha) + (size_t)rd.col_off * 8 + g;
    ColInfo* __restrict__ cinfo = V.colinfo + rd.col_off;
    const unsigned t3_base = (unsigned)pinned((int)__cvta_generic_to_shared(s_t3));
    const int src_up = pinned((g + 7) & 7), src_up2 = pinned((g + 6) & 7), src_up4 = pinned((g + 4) & 7);
    const int g4 = pinned(4 * g);
    const float thr = __uint_as_float((unsigned)(127 + kEdgeLog2) << 23);
    const int oct_shift = pinned((int)(threadIdx.x & 24));

    // column 0: alpha(0,0) = 1
    float v0 = (g == 0) ? 1.f : 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    int s = 0, cum = 0, lap = 0;
    bool slid = false;                   // band start moved by 4 rows between the previous column and this one
    unsigned w0 = rc32[min(g, wmax)], w1 = rc32[min(8 + g, wmax)], w2 = rc32[min(16 + g, wmax)];
    if (valid) {
        acol[0] = make_float4(v0, v1, v2, v3);
        if (g == 0) cinfo[0] = ColInfo{0, 0};
    }
    // template bytes carry the trinucleotide index 16*t[j-2] + 4*t[j-1] + t[j]; column 1 uses the pinned-first-move rows
    int x_cur = valid ? 64 + (tp[1] & 15) : 0;
    int x_nxt = (2 < J) ? tp[2] : 0;

    auto column = [&](const int j, const bool rescale) {
        const bool alive = j < J;
        int x_pre = 0;
        if (j + 2 < J) x_pre = tp[j + 2];          // prefetch, two columns ahead
        const int ss = s & 31;
        const unsigned cw = (g4 < ss) ? w1 : w0;   // lanes below the start slot hold rows of the next lap
        const int r0 = (g4 - s) & 31;              // band-relative row of this lane's first cell
        const bool startl = r0 == 0, topl = r0 == 28;
        float up0 = __shfl_sync(kFullMask, v3, src_up, 8);   // alpha(row-1, j-1) of the lane's first cell
        if (startl && !slid) up0 = 0.f;            // row s-1 is outside the previous band
        if (topl && slid) { v0 = 0.f; v1 = 0.f; v2 = 0.f; v3 = 0.f; }   // this lane just re-entered at the top: new rows
        const unsigned row = t3_base + ((unsigned)x_cur << 7);
        const float2 e0 = lds_f2(row + ((cw & 0xffu) << 1));
        const float2 e1 = lds_f2(row + (__byte_perm(cw, 0u, 0x4441u) << 1));
        const float2 e2 = lds_f2(row + (__byte_perm(cw, 0u, 0x4442u) << 1));
        const float2 e3 = lds_f2(row + ((cw >> 24) << 1));
        const float D = lds_f1(row + kDSlot * 8);
        float A0 = fmaf(e0.x, up0, D * v0);
        float A1 = fmaf(e1.x, v0, D * v1);
        float A2 = fmaf(e2.x, v1, D * v2);
        float A3 = fmaf(e3.x, v2, D * v3);
        float G0 = startl ? 0.f : e0.y;            // band start: no in-band predecessor (makes the ring scan exact)
        float G1 = e1.y, G2 = e2.y, G3 = e3.y;
        // a_i = A_i + G_i * a_{i-1}: serial inside the lane, Kogge-Stone ring over the octet
        A1 = fmaf(G1, A0, A1); G1 *= G0;
        A2 = fmaf(G2, A1, A2); G2 *= G1;
        A3 = fmaf(G3, A2, A3); G3 *= G2;
        float At = A3, Gt = G3;
        {
            float As = __shfl_sync(kFullMask, At, src_up, 8), Gs = __shfl_sync(kFullMask, Gt, src_up, 8);
            At = fmaf(Gt, As, At); Gt *= Gs;
            As = __shfl_sync(kFullMask, At, src_up2, 8); Gs = __shfl_sync(kFullMask, Gt, src_up2, 8);
            At = fmaf(Gt, As, At); Gt *= Gs;
            As = __shfl_sync(kFullMask, At, src_up4, 8);
            At = fmaf(Gt, As, At);
        }
        const float xin = __shfl_sync(kFullMask, At, src_up, 8);
        v0 = fmaf(G0, xin, A0); v1 = fmaf(G1, xin, A1); v2 = fmaf(G2, xin, A2); v3 = fmaf(G3, xin, A3);

        // leading edge: does one of the band's last three rows reach 2^-60?  Only the top lane can say so; one ballot
        // tells the whole octet
        const float m3 = fmaxf(fmaxf(v1, v2), v3);
        const unsigned bal = __ballot_sync(kFullMask, topl && m3 >= thr);
        if (rescale) {                             // every 4th column is rescaled (spec); known at compile time here
            // the sign bit of a maximum is 0, so key >> 23 is its biased exponent e: scale by 2^(127-e) (an all-zero
            // or denormal column gives e = 0 -- the read is dead or about to be, and its LL ends up -inf either way)
            const unsigned key = vmax_oct(__float_as_uint(fmaxf(m3, v0)) & 0xffff0000u);
            const float sc = __uint_as_float(0x7f000000u - ((key >> 23) << 23));
            v0 *= sc; v1 *= sc; v2 *= sc; v3 *= sc;
            cum += (int)(key >> 23) - 127;
        }
        if (alive) {
            acol[(size_t)j * 8] = make_float4(v0, v1, v2, v3);
            if (g == 0) cinfo[j] = ColInfo{s, cum};
        }
        slid = ((bal >> oct_shift) & 0xffu) != 0u;
        s += slid ? 4 : 0;
        if ((s >> 5) != lap) {           // octet-uniform, once every 32 rows
            lap = s >> 5;
            w0 = w1; w1 = w2;
            w2 = rc32[min(8 * (lap + 2) + g, wmax)];
        }
        x_cur = x_nxt; x_nxt = x_pre;
    };
    // columns 1..3, then groups of four starting at a multiple of kScaleEvery (the first of each group is rescaled:
    // no per-column test, no rotation of the software-pipelined registers), then the tail
    static_assert(kScaleEvery == 4, "the column loop is unrolled in step with the rescale schedule");
    int j = 1;
    for (; j < 4 && j < Jmax; ++j) column(j, false);
    for (; j + 3 < Jmax; j += 4) { column(j, true); column(j + 1, false); column(j + 2, false); column(j + 3, false); }
    for (; j < Jmax; ++j) column(j, (j & 3) == 0);
    // alpha(I-1, J-1) lives in slot (I-1) mod 32 of the last column if it is inside the band; lane 0 of the octet reads
    // it back from the column line the octet just wrote (ordered by __syncwarp), then applies the pinned last match
    __syncwarp();
    if (valid) {
        if (g == 0) {
            const ColInfo cl = cinfo[J - 1];
            const int slot = (I - 1) & 31;
            const bool inband = (cl.start + ((slot - cl.start) & 31)) == (I - 1);
            const float fv = inband ? V.alpha[((size_t)rd.col_off + (J - 1)) * 32 + slot] : 0.f;
            const int final_cum = cl.cumexp;
            const int ctxl = tp[J - 1] & 15;
            const double a = (double)fv * (double)V.em_match[(kCtxEndRow + ctxl) * kEmStride + rd.last_code];
            const double base = (a > 0.0) ? log(a) + 0.6931471805599453094 * (double)final_cum : -INFINITY;
            V.base_ll[r] = base;
            V.ll_alpha[r] = base - (double)I * V.log_cw;
        }
    } else if (r >= 0 && g == 0) {
        V.base_ll[r] = -INFINITY;
        V.ll_alpha[r] = -INFINITY;
    }
}

template <int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) arrow_fill_beta_kernel(const ArrowBatchView V, const int32_t* __restrict__ order,
                                                              const int n_items) {
    __shared__ __align__(128) float2 s_t2[kT2Rows * kEmStride];
    {
        const int r0 = order[blockIdx.x * 16];
        const int zmw = (r0 >= 0) ? V.reads[r0].zmw : 0;
        const float4* __restrict__ tr = reinterpret_cast<const float4*>(V.trans) + (size_t)zmw * 36;
        for (int idx = threadIdx.x; idx < kT2Rows * kEmStride; idx += blockDim.x) {
            const int row = idx >> 4, code = idx & 15;
            s_t2[idx] = folded_entry(V.em_match, V.em_ins, tr, row, row, code);
        }
    }
    __syncthreads();

    const int item = blockIdx.x * 16 + (threadIdx.x >> 3);
    const int g = pinned(threadIdx.x & 7);
    int r = -1;
    if (item < n_items) r = order[item];
    DevRead rd;
    rd.J = 0; rd.I = 0; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0; rd.zmw = 0; rd.last_code = 0; rd.first_code = 0;
    rd.code_stride = 64; rd.active = 0;
    if (r >= 0) rd = V.reads[r];
    const bool valid = r >= 0 && rd.active && rd.J >= 2 && rd.I >= 2;
    const int J = pinned(valid ? rd.J : 0);
    const int I = rd.I;
    int Jmax = J;
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) Jmax = max(Jmax, __shfl_xor_sync(kFullMask, Jmax, off));

    // second copy of the row codes, shifted by one row: word k holds the codes of rows 4k+1 .. 4k+4
    const unsigned* __restrict__ rc32 = reinterpret_cast<const unsigned*>(V.rowcode + rd.code_off + rd.code_stride);
    const int wmax = (rd.code_stride >> 2) - 1;
    const uint8_t* __restrict__ tp = V.tpl + rd.tpl_off;
    float4* __restrict__ bcol = reinterpret_cast<float4*>(V.beta) + (size_t)rd.col_off * 8 + g;
    const ColInfo* __restrict__ cinfo = V.colinfo + rd.col_off;
    int32_t* __restrict__ bexp = V.beta_exp + rd.col_off;
    const unsigned t2_base = (unsigned)pinned((int)__cvta_generic_to_shared(s_t2));
    const int src_dn = pinned((g + 1) & 7), src_dn2 = pinned((g + 2) & 7), src_dn4 = pinned((g + 4) & 7);
    const int g4 = pinned(4 * g);

    // The warp walks columns from (Jmax-1) down to 1; an octet is alive once j <= J-1.
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    int cum = 0, lap = 0;
    int s_cur = 0, s_nxt = 0;            // band starts of columns j and j+1
    int s_p1 = 0, s_p2 = 0;              // prefetched band starts of columns j-1 and j-2
    int x_cur = 0, x_p1 = 0, x_p2 = 0;   // dinucleotide contexts of columns j, j-1, j-2
    unsigned w0 = 0x30303030u, w1 = 0x30303030u, wm = 0x30303030u;   // laps L, L+1, L-1 (sentinels)
    bool started = false;

#pragma unroll 2
    for (int j = Jmax - 1; j >= 1; --j) {
        const bool alive = j <= J - 1;
        const bool init_col = alive && !started;
        if (init_col) {
            // prologue at j == J-1
            s_cur = cinfo[j].start;
            s_nxt = s_cur;
            s_p1 = cinfo[max(j - 1, 0)].start;
            x_cur = tp[j] & 15;
            x_p1 = tp[max(j - 1, 0)] & 15;
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
        // prefetch for column j-2
        if (alive && j >= 3) { s_p2 = cinfo[j - 2].start; x_p2 = tp[j - 2] & 15; }

        const bool slid = s_nxt != s_cur;          // band start of column j+1 is 4 rows further
        const int ss = s_cur & 31;
        const unsigned cw = (g4 < ss) ? w1 : w0;
        const int r0 = (g4 - s_cur) & 31;
        const bool startl = r0 == 0, topl = r0 == 28;
        float dn3 = __shfl_sync(kFullMask, v0, src_dn, 8);   // beta(row+1, j+1) of the lane's last cell
        if (topl && !slid) dn3 = 0.f;              // row s+32 is outside the band of column j+1
        if (startl && slid) { v0 = 0.f; v1 = 0.f; v2 = 0.f; v3 = 0.f; }   // rows below the band of column j+1
        const unsigned row = t2_base + ((unsigned)x_cur << 7);
        const float2 e0 = lds_f2(row + ((cw & 0xffu) << 1));
        const float2 e1 = lds_f2(row + (__byte_perm(cw, 0u, 0x4441u) << 1));
        const float2 e2 = lds_f2(row + (__byte_perm(cw, 0u, 0x4442u) << 1));
        const float2 e3 = lds_f2(row + ((cw >> 24) << 1));
        const float D = lds_f1(row + kDSlot * 8);
        float A0 = fmaf(e0.x, v1, D * v0);
        float A1 = fmaf(e1.x, v2, D * v1);
        float A2 = fmaf(e2.x, v3, D * v2);
        float A3 = fmaf(e3.x, dn3, D * v3);
        float G0 = e0.y, G1 = e1.y, G2 = e2.y;
        float G3 = topl ? 0.f : e3.y;              // band end: no in-band successor
        if (init_col) {
            // last column: beta(I-1, J-1) = pinned last match; rows above it by insertions
            const float endv = V.em_match[(kCtxEndRow + x_cur) * kEmStride + rd.last_code];
            const int row0 = s_cur + r0;
            A0 = (row0 == I - 1) ? endv : 0.f;
            A1 = (row0 + 1 == I - 1) ? endv : 0.f;
            A2 = (row0 + 2 == I - 1) ? endv : 0.f;
            A3 = (row0 + 3 == I - 1) ? endv : 0.f;
        }
        // (an octet that has not started yet holds an all-zero column: its terms A are zero whatever the table row
        //  says, the scan returns zeros, and nothing is stored -- no masking needed)
        // b_i = A_i + G_i * b_{i+1}
        A2 = fmaf(G2, A3, A2); G2 *= G3;
        A1 = fmaf(G1, A2, A1); G1 *= G2;
        A0 = fmaf(G0, A1, A0); G0 *= G1;
        float At = A0, Gt = G0;
        {
            float As = __shfl_sync(kFullMask, At, src_dn, 8), Gs = __shfl_sync(kFullMask, Gt, src_dn, 8);
            At = fmaf(Gt, As, At); Gt *= Gs;
            As = __shfl_sync(kFullMask, At, src_dn2, 8); Gs = __shfl_sync(kFullMask, Gt, src_dn2, 8);
            At = fmaf(Gt, As, At); Gt *= Gs;
            As = __shfl_sync(kFullMask, At, src_dn4, 8);
            At = fmaf(Gt, As, At);
        }
        const float xin = __shfl_sync(kFullMask, At, src_dn, 8);
        v0 = fmaf(G0, xin, A0); v1 = fmaf(G1, xin, A1); v2 = fmaf(G2, xin, A2); v3 = fmaf(G3, xin, A3);

        int kcol = 0;
        if ((j & (kScaleEvery - 1)) == 0) {        // warp-uniform: every 4th column is rescaled (spec)
            const float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
            const unsigned key = vmax_oct(__float_as_uint(mx) & 0xffff0000u);
            const float sc = __uint_as_float(0x7f000000u - ((key >> 23) << 23));
            v0 *= sc; v1 *= sc; v2 *= sc; v3 *= sc;
            kcol = (int)(key >> 23) - 127;
        }
        if (alive) {
            cum += kcol;
            bcol[(size_t)j * 8] = make_float4(v0, v1, v2, v3);
            if (g == 0) bexp[j] = cum;
            started = true;
            s_nxt = s_cur; s_cur = s_p1; s_p1 = s_p2;
            x_cur = x_p1; x_p1 = x_p2;
        }
    }
    // the warp's last iteration is column 1 of every live octet: beta(1,1) is cell 1 of lane 0 if row 1 is in the band
    if (valid) {
        if (g == 0) {
            const bool inband = s_nxt <= 1;        // after the last column s_nxt holds the band start of column 1
            const float fv = inband ? v1 : 0.f;
            const int first_cum = cum;
            const double b = (double)fv * (double)V.em_match[(kCtxStartRow + (tp[0] & 3)) * kEmStride + rd.first_code];
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

}  // namespace

// order[n_items]: n_items is a multiple of 16; every aligned group of 16 entries holds reads of ONE ZMW
// (longest template first), padded with -1; a group's first entry is a real read.
void launch_fill_alpha(const ArrowBatchView& V, const int32_t* order, int n_items, cudaStream_t stream) {
    if (n_items <= 0) return;
    // occupancy target (CTAs of 128 threads per SM): CCS_B200_FILL_VARIANT = 0 (compiler's choice), 10, 12 -- for A/B runs
    static const int variant = [] { const char* e = std::getenv("CCS_B200_FILL_VARIANT"); return e ? std::atoi(e) : 0; }();
    const int grid = (n_items + 15) / 16;
    if (variant == 12) arrow_fill_alpha_kernel<12><<<grid, 128, 0, stream>>>(V, order, n_items);
    else if (variant == 10) arrow_fill_alpha_kernel<10><<<grid, 128, 0, stream>>>(V, order, n_items);
    else arrow_fill_alpha_kernel<0><<<grid, 128, 0, stream>>>(V, order, n_items);
}

void launch_fill_beta(const ArrowBatchView& V, const int32_t* order, int n_items, cudaStream_t stream) {
    if (n_items <= 0) return;
    static const int variant = [] { const char* e = std::getenv("CCS_B200_FILL_VARIANT"); return e ? std::atoi(e) : 0; }();
    const int grid = (n_items + 15) / 16;
    if (variant == 12) arrow_fill_beta_kernel<10><<<grid, 128, 0, stream>>>(V, order, n_items);
    else if (variant == 10) arrow_fill_beta_kernel<8><<<grid, 128, 0, stream>>>(V, order, n_items);
    else arrow_fill_beta_kernel<0><<<grid, 128, 0, stream>>>(V, order, n_items);
}

}  // namespace ccs
