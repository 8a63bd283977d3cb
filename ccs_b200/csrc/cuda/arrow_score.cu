// arrow_score: delta log-likelihood of every single-base mutation of a template position,
// summed over the ZMW's subreads in read order (Evaluator::LL(Mutation) = ExtendAlpha over
// the columns whose dinucleotide context changed + LinkAlphaBeta; Integrator::LL sums the
// reads -- SURVEY.md 8a rows a11-a14; /root/reference/docs/how-does-ccs-work.md:96-99
// "substituting one of the other three nucleotides, inserting one of the four nucleotides
// ..., or deleting the position").
//
// Mapping: one warp octet per (ZMW, template position); the octet loops over the ZMW's reads
// in index order and keeps the nine per-slot sums {SUB A,C,G,T, DEL, INS A,C,G,T} in fp64
// registers, so the reduction over reads is deterministic and needs no atomics.  Per
// (read, position) it loads two alpha columns + two beta columns (128 B each, L2/L1 shared
// with the neighbouring positions) and runs 8 virtual-template evaluations of <= 3 extension
// columns each: the kernel is issue-bound, not HBM-bound (DESIGN.md "Roofline").
#include "arrow_octet.cuh"
#include "arrow_launch.h"

namespace ccs {

namespace {

struct ReadCtx {
    const uint8_t* rc;
    const uint8_t* tp;
    const float4* tr;
    const float4* acol;
    const float4* bcol;
    const ColInfo* cinfo;
    const int32_t* bexp;
    int I, J;
    int last_code;
};

// virtual template base at index j of the mutated template; `word` packs template bases
// T[q-3 .. q+4] two bits each (bit 2*(x-q+3)).
__device__ __forceinline__ int tv_base(const unsigned word, const int type, const int q, const int base, const int j) {
    int x;   // index into the original template, or -1 for the new base
    if (type == 0) x = (j == q) ? -1 : j;
    else if (type == 1) x = (j < q) ? j : ((j == q) ? -1 : j - 1);
    else x = (j < q) ? j : j + 1;
    const int sh = 2 * (x - q + 3);
    const int t = (int)((word >> (sh & 31)) & 3u);
    return (x < 0) ? base : t;
}

// One mutation of one read.  All octets of the warp call this together (full-mask shuffles);
// `live` = this octet really wants the result.  Returns value v and exponent e with
// LL' = log(v) + ln2 * e (+ the read's constant counter-weight term).
__device__ __forceinline__ float eval_mutation(const ReadCtx& R, const int g, const float* s_emm, const float* s_emi,
                                               const unsigned word, const int type, const int q, const int base,
                                               const bool live, int& exp_out) {
    const int J = R.J, I = R.I;
    const int delta = (type == 1) ? 1 : ((type == 2) ? -1 : 0);
    const int Jp = J + delta;
    const int a = max(1, q);
    const int bp = (type == 2) ? q + 1 : q + 2;
    const int borig = bp - delta;
    const bool terminal = bp > Jp - 1;
    const int last = terminal ? Jp - 1 : bp - 1;
    int n_ext = live ? max(0, last - a + 1) : 0;
    int n_max = n_ext;
    n_max = max(n_max, __shfl_xor_sync(kFullMask, n_max, 8));
    n_max = max(n_max, __shfl_xor_sync(kFullMask, n_max, 16));

    // start column a-1
    const int ca = live ? min(a - 1, J - 1) : 0;
    const ColInfo ci0 = R.cinfo[ca];
    const float4 c0 = R.acol[(size_t)ca * 8 + g];
    float v[4] = {c0.x, c0.y, c0.z, c0.w};
    int S = ci0.start;
    int e = ci0.cumexp;
    const int code_max = I + kRowCodePad - 1;

    for (int k = 0; k < n_max; ++k) {
        const bool real = k < n_ext;
        const int jp = a + k;
        const int S_new = real ? max(S, R.cinfo[min(jp, J - 1)].start) : S;
        const int d = S_new - S;
        const int tm2 = tv_base(word, type, q, base, jp - 2);
        const int tm1 = tv_base(word, type, q, base, jp - 1);
        const int t00 = tv_base(word, type, q, base, jp);
        const int cm = (jp == 1) ? kCtxStartRow + tm1 : 4 * tm2 + tm1;
        const int ci = 4 * tm1 + t00;
        const float4 trm = R.tr[real ? cm : 0];
        const float4 tri = R.tr[real ? ci : 0];
        int rel[4], code[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            rel[x] = (4 * g + x - S_new) & 31;
            code[x] = real ? R.rc[min(S_new + rel[x], code_max)] : kC4Sentinel;
        }
        float A[4], G[4];
        octet_forward_terms(v, g, d, rel, code, trm.x, trm.y, tri.z, tri.w, s_emm + (real ? cm : 0) * kEmStride,
                            s_emi + (real ? ci : 0) * kEmStride, (ci & 3) << 2, A, G);
        float w[4];
        octet_forward_scan(A, G, g, w);
        if (real) {
#pragma unroll
            for (int x = 0; x < 4; ++x) v[x] = w[x];
            S = S_new;
        }
    }

    // terminal: alpha'(I-1, J'-1) * pinned last match
    float term_v;
    {
        const int slot = (I - 1) & 31;
        const int rrel = (slot - S) & 31;
        const bool inband = (S + rrel) == (I - 1);
        float x = 0.f;
#pragma unroll
        for (int y = 0; y < 4; ++y) x = (4 * g + y == slot && inband) ? v[y] : x;
        x = octet_max(x);
        const int ctxl = 4 * tv_base(word, type, q, base, Jp - 2) + tv_base(word, type, q, base, Jp - 1);
        term_v = x * s_emm[(kCtxEndRow + (terminal ? ctxl : 0)) * kEmStride + R.last_code];
    }
    // link with beta column borig
    float link_v;
    int link_e;
    {
        const bool lk = live && !terminal;
        const int cb = lk ? min(max(borig, 1), J - 1) : 1;
        const int sb = R.cinfo[cb].start;
        link_e = R.bexp[cb];
        const float4 b0 = R.bcol[(size_t)cb * 8 + g];
        const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
        float dn[4];
        dn[3] = shfl_oct(bv[0], (g + 1) & 7);
        dn[0] = bv[1]; dn[1] = bv[2]; dn[2] = bv[3];
        const int cmL = (bp == 1) ? kCtxStartRow + tv_base(word, type, q, base, 0)
                                  : 4 * tv_base(word, type, q, base, bp - 2) + tv_base(word, type, q, base, bp - 1);
        const float4 trl = R.tr[lk ? cmL : 0];
        const float* emm_row = s_emm + (lk ? cmL : 0) * kEmStride;
        float acc = 0.f;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int row = S + ((4 * g + x - S) & 31);
            const float bx = (row >= sb && row < sb + 32) ? bv[x] : 0.f;
            const float bd = (row + 1 >= sb && row + 1 < sb + 32) ? dn[x] : 0.f;
            const int c1 = lk ? R.rc[min(row + 1, code_max)] : kC4Sentinel;
            const float wv = fmaf(trl.x, ldtab(emm_row, c1) * bd, trl.y * bx);
            acc = fmaf(v[x], wv, acc);
        }
        link_v = octet_sum(acc);
    }
    exp_out = e + (terminal ? 0 : link_e);
    return terminal ? term_v : link_v;
}

__global__ void __launch_bounds__(128) arrow_score_generic_kernel(const ArrowBatchView V, const ScoreRange* __restrict__ ranges,
                                                          const int n_ranges, const long long n_items,
                                                          double* __restrict__ delta) {
    __shared__ float s_emm[36 * kEmStride];
    __shared__ float s_emi[17 * kEmStride];
    for (int k = threadIdx.x; k < 36 * kEmStride; k += blockDim.x) s_emm[k] = V.em_match[k];
    for (int k = threadIdx.x; k < 17 * kEmStride; k += blockDim.x) s_emi[k] = V.em_ins[k];
    __syncthreads();

    const long long item = (long long)blockIdx.x * 16 + (threadIdx.x >> 3);
    const int g = threadIdx.x & 7;
    const bool have = item < n_items;
    int z = 0, p = 0;
    if (have) {
        int lo = 0, hi = n_ranges - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ranges[mid].first <= item) lo = mid; else hi = mid - 1;
        }
        const ScoreRange rg = ranges[lo];
        z = rg.zmw;
        p = rg.p_begin + (int)(item - rg.first);
    }
    DevZmw zm;
    zm.read_begin = zm.read_end = 0; zm.fwd_off = 0; zm.J = 0; zm.delta_off = 0;
    if (have) zm = V.zmws[z];
    const int n_reads = zm.read_end - zm.read_begin;
    int n_max = n_reads;
    n_max = max(n_max, __shfl_xor_sync(kFullMask, n_max, 8));
    n_max = max(n_max, __shfl_xor_sync(kFullMask, n_max, 16));
    const int tbase = have ? (V.tpl[zm.fwd_off + p] & 3) : 0;

    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;

    for (int k = 0; k < n_max; ++k) {
        const bool has_read = k < n_reads;
        DevRead rd;
        rd.active = 0; rd.ts = rd.te = 0; rd.strand = 0; rd.I = 2; rd.J = 2; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0;
        rd.zmw = 0; rd.last_code = 0;
        int st = 1;
        const int r = zm.read_begin + k;
        if (has_read) { rd = V.reads[r]; st = V.status[r]; }
        const bool usable = has_read && rd.active && st == 0;
        if (!usable) {   // idle octets still execute the shared instruction stream: give them harmless operands
            rd.ts = rd.te = 0; rd.strand = 0; rd.I = 2; rd.J = 2; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0;
            rd.zmw = 0; rd.last_code = 0;
        }
        const bool cov_sd = usable && p >= rd.ts && p < rd.te;       // SUB / DEL
        const bool cov_in = usable && p > rd.ts && p < rd.te;        // INS (before p)
        if (!__any_sync(kFullMask, cov_sd)) continue;

        ReadCtx R;
        R.rc = V.rowcode + rd.code_off;
        R.tp = V.tpl + rd.tpl_off;
        R.tr = reinterpret_cast<const float4*>(V.trans) + (size_t)rd.zmw * 36;
        R.acol = reinterpret_cast<const float4*>(V.alpha) + (size_t)rd.col_off * 8;
        R.bcol = reinterpret_cast<const float4*>(V.beta) + (size_t)rd.col_off * 8;
        R.cinfo = V.colinfo + rd.col_off;
        R.bexp = V.beta_exp + rd.col_off;
        R.I = rd.I; R.J = rd.J; R.last_code = rd.last_code;
        const double base_ll = cov_sd ? V.base_ll[r] : 0.0;

        const int q_sd = cov_sd ? (rd.strand ? rd.te - 1 - p : p - rd.ts) : 0;
        const int q_in = cov_in ? (rd.strand ? rd.te - p : p - rd.ts) : 1;
        // template windows around q_sd and q_in: bases T[q-3 .. q+4], 2 bits each
        unsigned word_sd = 0, word_in = 0;
        {
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const int j1 = q_sd - 3 + x, j2 = q_in - 3 + x;
                const unsigned b1 = (cov_sd && j1 >= 0 && j1 < rd.J) ? (R.tp[j1] & 3u) : 0u;
                const unsigned b2 = (cov_in && j2 >= 0 && j2 < rd.J) ? (R.tp[j2] & 3u) : 0u;
                word_sd |= b1 << (2 * x);
                word_in |= b2 << (2 * x);
            }
        }
        // 3 substitutions
#pragma unroll 1
        for (int m = 0; m < 3; ++m) {
            const int bf = (tbase + 1 + m) & 3;                 // forward-strand base
            const int bl = rd.strand ? 3 - bf : bf;
            int e;
            const float val = eval_mutation(R, g, s_emm, s_emi, word_sd, 0, q_sd, bl, cov_sd, e);
            const double dll = (val > 0.f) ? (double)logf(val) + 0.6931471805599453094 * (double)e - base_ll : -INFINITY;
            if (cov_sd) {
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2) if (s2 == bf) acc[s2] += dll;
            }
        }
        {   // deletion
            int e;
            const float val = eval_mutation(R, g, s_emm, s_emi, word_sd, 2, q_sd, 0, cov_sd && rd.J >= 3, e);
            const double dll = (val > 0.f) ? (double)logf(val) + 0.6931471805599453094 * (double)e - base_ll : -INFINITY;
            if (cov_sd) acc[4] += (rd.J >= 3) ? dll : -INFINITY;
        }
        if (__any_sync(kFullMask, cov_in)) {
#pragma unroll 1
            for (int bf = 0; bf < 4; ++bf) {
                const int bl = rd.strand ? 3 - bf : bf;
                int e;
                const float val = eval_mutation(R, g, s_emm, s_emi, word_in, 1, q_in, bl, cov_in, e);
                const double dll = (val > 0.f) ? (double)logf(val) + 0.6931471805599453094 * (double)e - base_ll : -INFINITY;
                if (cov_in) {
#pragma unroll
                    for (int s2 = 0; s2 < 4; ++s2) if (s2 == bf) acc[5 + s2] += dll;
                }
            }
        }
    }
    if (have && g == 0) {
        double* out = delta + (size_t)(zm.delta_off + p) * kDeltaStride;
#pragma unroll
        for (int k = 0; k < 9; ++k) out[k] = acc[k];
#pragma unroll
        for (int k = 9; k < 13; ++k) out[k] = 0.0;
    }
}

// -------------------------------------------------------------------------------------------
// Fast path for interior positions: the eight mutations of a position share
//   * the match/deletion terms of the first extension column (context of t[q-1] unchanged),
//   * the first extension column X_b per new base b (SUB(q,b), INS(before q,b) and, for
//     b = t[q+1], DEL(q) all start with context (t[q-1], b)),
//   * the match/deletion terms of the second extension column per b.
// => 4 + 7 in-column scans and 8 links per (read, position) instead of 14 + 8 with every
//    operand reloaded.  Reverse-strand reads compute the insertion BEFORE their local q, which
//    is forward INS(p+1): those sums go to slots 9..12 of row p+1 (combined by the consumers).
// -------------------------------------------------------------------------------------------
struct FastOut { float sub[4]; float del; float ins[4]; int e_sd; int e_in; };

__device__ __forceinline__ float link_dot(const float y[4], const float bx[4], const float bd[4], const int codeL[4],
                                          const float4 trl, const float* __restrict__ emm_row) {
    float acc = 0.f;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const float wv = fmaf(trl.x, ldtab(emm_row, codeL[x]) * bd[x], trl.y * bx[x]);
        acc = fmaf(y[x], wv, acc);
    }
    return octet_sum(acc);
}

__device__ __forceinline__ void fast_eval(const ReadCtx& R, const int g, const float* s_emm, const float* s_emi,
                                          const int q_in, const bool live, FastOut& out) {
    const int J = R.J, I = R.I;
    const int q = live ? q_in : 2;
    const int jm2 = min(max(q - 2, 0), J - 1), jm1 = min(max(q - 1, 0), J - 1), j0 = min(q, J - 1);
    const int j1 = min(q + 1, J - 1), j2 = min(q + 2, J - 1);
    const int code_max = I + kRowCodePad - 1;
    const ColInfo cim1 = R.cinfo[jm1];
    const int s0 = cim1.start;
    const int sq0 = R.cinfo[j0].start, sq1 = R.cinfo[j1].start, sq2 = R.cinfo[j2].start;
    const int s1 = max(s0, sq0), s2 = max(s1, sq1);
    const float4 a0 = R.acol[(size_t)jm1 * 8 + g];
    const float4 b1v = R.bcol[(size_t)j1 * 8 + g];
    const float4 b2v = R.bcol[(size_t)j2 * 8 + g];
    out.e_sd = cim1.cumexp + R.bexp[j2];
    out.e_in = cim1.cumexp + R.bexp[j1];
    const int tm2 = R.tp[jm2] & 3, tm1 = R.tp[jm1] & 3, t0 = R.tp[j0] & 3, tp1 = R.tp[j1] & 3;   // & 3: idle octets read arbitrary bytes

    int rel1[4], rel2[4], codeU1[4], codeG1[4], codeU2[4], codeG2[4], codeL2[4], codeL1[4], codeLb1[4];
    float pvm2[4];
    const float v0[4] = {a0.x, a0.y, a0.z, a0.w};
    const int d1 = s1 - s0, d2 = s2 - s1;
    float up1[4];
    up1[0] = shfl_oct(v0[3], (g + 7) & 7); up1[1] = v0[0]; up1[2] = v0[1]; up1[3] = v0[2];
    // beta columns, masked onto the rows of the extension bands
    float bx2[4], bd2[4], bx1[4], bd1[4], bxD[4], bdD[4];
    const float be2[4] = {b2v.x, b2v.y, b2v.z, b2v.w}, be1[4] = {b1v.x, b1v.y, b1v.z, b1v.w};
    float dn2[4], dn1[4];
    dn2[3] = shfl_oct(be2[0], (g + 1) & 7); dn2[0] = be2[1]; dn2[1] = be2[2]; dn2[2] = be2[3];
    dn1[3] = shfl_oct(be1[0], (g + 1) & 7); dn1[0] = be1[1]; dn1[1] = be1[2]; dn1[2] = be1[3];
    float A1[4];
    const int cmq = 4 * tm2 + tm1;
    const float4 trq = R.tr[cmq];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int slot = 4 * g + x;
        rel1[x] = (slot - s1) & 31;
        rel2[x] = (slot - s2) & 31;
        const int row1 = s1 + rel1[x], row2 = s2 + rel2[x];
        // (idle octets run on whatever band starts sit in the first columns of the store: clamp both ways)
        const int c1 = R.rc[min(max(row1, 0), code_max)];
        const int c2 = R.rc[min(max(row2, 0), code_max)];
        const int c1n = R.rc[min(max(row1 + 1, 0), code_max)];
        const int c2n = R.rc[min(max(row2 + 1, 0), code_max)];
        const int rd1 = rel1[x] + d1, rd2 = rel2[x] + d2;
        const float pv = (rd1 < 32) ? v0[x] : 0.f;
        codeU1[x] = ((unsigned)(rd1 - 1) < 32u) ? c1 : kC4Sentinel;
        codeG1[x] = (rel1[x] == 0) ? kC4Sentinel : c1;
        codeU2[x] = ((unsigned)(rd2 - 1) < 32u) ? c2 : kC4Sentinel;
        codeG2[x] = (rel2[x] == 0) ? kC4Sentinel : c2;
        pvm2[x] = (rd2 < 32) ? 1.f : 0.f;
        A1[x] = fmaf(trq.x, ldtab(s_emm + cmq * kEmStride, codeU1[x]) * up1[x], trq.y * pv);
        // link operands: beta(i, c) / beta(i+1, c) at the rows of the extension band
        bx2[x] = (row2 >= sq2 && row2 < sq2 + 32) ? be2[x] : 0.f;
        bd2[x] = dn2[x];
        codeL2[x] = (row2 + 1 >= sq2 && row2 + 1 < sq2 + 32) ? c2n : kC4Sentinel;
        bx1[x] = (row2 >= sq1 && row2 < sq1 + 32) ? be1[x] : 0.f;
        bd1[x] = dn1[x];
        codeLb1[x] = (row2 + 1 >= sq1 && row2 + 1 < sq1 + 32) ? c2n : kC4Sentinel;
        bxD[x] = (row1 >= sq2 && row1 < sq2 + 32) ? be2[x] : 0.f;
        bdD[x] = dn2[x];
        codeL1[x] = (row1 + 1 >= sq2 && row1 + 1 < sq2 + 32) ? c1n : kC4Sentinel;
    }
    float Xdel[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int ci = 4 * tm1 + b;
        const float4 tri = R.tr[ci];
        const float* emi_row = s_emi + ci * kEmStride;
        const float* emm_row = s_emm + ci * kEmStride;
        float A[4], G[4], X[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            A[x] = A1[x];
            G[x] = ldtab(emi_row, codeG1[x]) * (((codeG1[x] & 12) == (b << 2)) ? tri.z : tri.w);
        }
        octet_forward_scan(A, G, g, X);
        if (b == tp1) {
#pragma unroll
            for (int x = 0; x < 4; ++x) Xdel[x] = X[x];
        }
        // second extension column: match/deletion terms shared by SUB(b) and INS(b)
        float up2[4], A2[4];
        up2[0] = shfl_oct(X[3], (g + 7) & 7); up2[1] = X[0]; up2[2] = X[1]; up2[3] = X[2];
#pragma unroll
        for (int x = 0; x < 4; ++x) A2[x] = fmaf(tri.x, ldtab(emm_row, codeU2[x]) * up2[x], tri.y * (X[x] * pvm2[x]));
        {   // SUB(q, b): insertion context (b, t[q+1]); link into beta column q+2
            const int c2 = 4 * b + tp1;
            const float4 tr2 = R.tr[c2];
            const float* e2 = s_emi + c2 * kEmStride;
            float Aa[4], Ga[4], Y[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                Aa[x] = A2[x];
                Ga[x] = ldtab(e2, codeG2[x]) * (((codeG2[x] & 12) == (tp1 << 2)) ? tr2.z : tr2.w);
            }
            octet_forward_scan(Aa, Ga, g, Y);
            const float v = link_dot(Y, bx2, bd2, codeL2, tr2, s_emm + c2 * kEmStride);
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k == b) out.sub[k] = v;
        }
        {   // INS(before q, b): insertion context (b, t[q]); link into beta column q+1
            const int c3 = 4 * b + t0;
            const float4 tr3 = R.tr[c3];
            const float* e3 = s_emi + c3 * kEmStride;
            float Aa[4], Ga[4], Z[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                Aa[x] = A2[x];
                Ga[x] = ldtab(e3, codeG2[x]) * (((codeG2[x] & 12) == (t0 << 2)) ? tr3.z : tr3.w);
            }
            octet_forward_scan(Aa, Ga, g, Z);
            const float v = link_dot(Z, bx1, bd1, codeLb1, tr3, s_emm + c3 * kEmStride);
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k == b) out.ins[k] = v;
        }
    }
    {   // DEL(q): X_{t[q+1]} links straight into beta column q+2 with context (t[q-1], t[q+1])
        const int cd = 4 * tm1 + tp1;
        out.del = link_dot(Xdel, bxD, bdD, codeL1, R.tr[cd], s_emm + cd * kEmStride);
    }
}

// Running product of the per-read link values of one mutation slot, kept as a mantissa in [1,2) times
// 2^pexp (pexp also absorbs the scale exponents of the stored columns): sum_r log(val_r) costs one FMUL and
// a few integer ops per read instead of a logf and fp64 adds; a non-positive value zeroes the product (-inf).
__device__ __forceinline__ void prod_mul(float& prod, int& pexp, float val, const int e) {
    int adj = e;
    if (val < 1.17549435e-38f) { val *= 1.8446744073709552e19f; adj -= 64; }   // subnormal -> normal
    const unsigned u = __float_as_uint(val);
    const bool nz = val > 0.f;
    float p = prod * __uint_as_float((u & 0x007fffffu) | 0x3f800000u);
    int ev = (int)(u >> 23) - 127 + adj;
    if (p >= 2.f) { p *= 0.5f; ev += 1; }
    prod = nz ? p : 0.f;
    pexp += nz ? ev : 0;
}

__device__ __forceinline__ double prod_dll(const float prod, const int pexp, const double base_sum) {
    return (prod > 0.f) ? (double)logf(prod) + 0.6931471805599453094 * (double)pexp - base_sum : -INFINITY;
}

__device__ __forceinline__ double dll_of(const float val, const int e, const double base_ll) {
    return (val > 0.f) ? (double)logf(val) + 0.6931471805599453094 * (double)e - base_ll : -INFINITY;
}

__global__ void __launch_bounds__(128) arrow_score_kernel(const ArrowBatchView V, const ScoreRange* __restrict__ ranges,
                                                          const int n_ranges, const long long n_items,
                                                          double* __restrict__ delta) {
    __shared__ float s_emm[36 * kEmStride];
    __shared__ float s_emi[17 * kEmStride];
    for (int k = threadIdx.x; k < 36 * kEmStride; k += blockDim.x) s_emm[k] = V.em_match[k];
    for (int k = threadIdx.x; k < 17 * kEmStride; k += blockDim.x) s_emi[k] = V.em_ins[k];
    __syncthreads();

    const long long item = (long long)blockIdx.x * 16 + (threadIdx.x >> 3);
    const int g = threadIdx.x & 7;
    const bool have = item < n_items;
    int z = 0, p = 0, p_begin = 0;
    if (have) {
        int lo = 0, hi = n_ranges - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ranges[mid].first <= item) lo = mid; else hi = mid - 1;
        }
        const ScoreRange rg = ranges[lo];
        z = rg.zmw;
        p = rg.p_begin + (int)(item - rg.first);
        p_begin = rg.p_begin;
    }
    DevZmw zm;
    zm.read_begin = zm.read_end = 0; zm.fwd_off = 0; zm.J = 0; zm.delta_off = 0;
    if (have) zm = V.zmws[z];
    const int n_reads = zm.read_end - zm.read_begin;
    int n_max = n_reads;
    n_max = max(n_max, __shfl_xor_sync(kFullMask, n_max, 8));
    n_max = max(n_max, __shfl_xor_sync(kFullMask, n_max, 16));
    const int tbase = have ? (V.tpl[zm.fwd_off + p] & 3) : 0;

    // per-slot running products {SUB A,C,G,T, DEL}, {INS A,C,G,T}, {INS' A,C,G,T} and, per group, the sum of the
    // contributing reads' base log-likelihoods
    float pr_sd[5], pr_ia[4], pr_ib[4];
    int px_sd[5], px_ia[4], px_ib[4];
#pragma unroll
    for (int k = 0; k < 5; ++k) { pr_sd[k] = 1.f; px_sd[k] = 0; }
#pragma unroll
    for (int k = 0; k < 4; ++k) { pr_ia[k] = 1.f; px_ia[k] = 0; pr_ib[k] = 1.f; px_ib[k] = 0; }
    double bs_sd = 0.0, bs_ia = 0.0, bs_ib = 0.0;

    for (int k = 0; k < n_max; ++k) {
        const bool has_read = k < n_reads;
        DevRead rd;
        rd.active = 0; rd.ts = rd.te = 0; rd.strand = 0; rd.I = 2; rd.J = 2; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0;
        rd.zmw = 0; rd.last_code = 0;
        int st = 1;
        const int r = zm.read_begin + k;
        if (has_read) { rd = V.reads[r]; st = V.status[r]; }
        const bool usable = has_read && rd.active && st == 0;
        if (!usable) {   // idle octets still execute the shared instruction stream: give them harmless operands
            rd.ts = rd.te = 0; rd.strand = 0; rd.I = 2; rd.J = 2; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0;
            rd.zmw = 0; rd.last_code = 0;
        }
        const bool cov_sd = usable && p >= rd.ts && p < rd.te;       // SUB / DEL
        const bool cov_in = usable && p > rd.ts && p < rd.te;        // INS (before p)
        if (!__any_sync(kFullMask, cov_sd)) continue;

        ReadCtx R;
        R.rc = V.rowcode + rd.code_off;
        R.tp = V.tpl + rd.tpl_off;
        R.tr = reinterpret_cast<const float4*>(V.trans) + (size_t)rd.zmw * 36;
        R.acol = reinterpret_cast<const float4*>(V.alpha) + (size_t)rd.col_off * 8;
        R.bcol = reinterpret_cast<const float4*>(V.beta) + (size_t)rd.col_off * 8;
        R.cinfo = V.colinfo + rd.col_off;
        R.bexp = V.beta_exp + rd.col_off;
        R.I = rd.I; R.J = rd.J; R.last_code = rd.last_code;
        const double base_ll = cov_sd ? V.base_ll[r] : 0.0;
        const int q_sd = cov_sd ? (rd.strand ? rd.te - 1 - p : p - rd.ts) : 0;
        const bool interior = cov_sd && q_sd >= 2 && q_sd <= rd.J - 4;
        const bool gen_sd = cov_sd && !interior;
        const bool prev_interior = rd.strand && p > p_begin && (q_sd + 1 >= 2) && (q_sd + 1 <= rd.J - 4);
        const bool gen_in = cov_in && (rd.strand ? !prev_interior : !interior);

        if (__any_sync(kFullMask, interior)) {
            FastOut fo;
#pragma unroll
            for (int b = 0; b < 4; ++b) { fo.sub[b] = 0.f; fo.ins[b] = 0.f; }
            fo.del = 0.f; fo.e_sd = 0; fo.e_in = 0;
            fast_eval(R, g, s_emm, s_emi, q_sd, interior, fo);
            if (interior) {
                bs_sd += base_ll;
                if (rd.strand) bs_ib += base_ll; else bs_ia += base_ll;
#pragma unroll
                for (int b = 0; b < 4; ++b) {                        // b = forward-strand base of the slot
                    const float vs = rd.strand ? fo.sub[3 - b] : fo.sub[b];
                    const float vi = rd.strand ? fo.ins[3 - b] : fo.ins[b];
                    if (b != tbase) prod_mul(pr_sd[b], px_sd[b], vs, fo.e_sd);
                    if (rd.strand) prod_mul(pr_ib[b], px_ib[b], vi, fo.e_in); else prod_mul(pr_ia[b], px_ia[b], vi, fo.e_in);
                }
                prod_mul(pr_sd[4], px_sd[4], fo.del, fo.e_sd);
            }
        }
        if (__any_sync(kFullMask, gen_sd) || __any_sync(kFullMask, gen_in)) {
            const int q_in = cov_in ? (rd.strand ? rd.te - p : p - rd.ts) : 1;
            unsigned word_sd = 0, word_in = 0;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const int j1 = q_sd - 3 + x, j2 = q_in - 3 + x;
                const unsigned b1 = (cov_sd && j1 >= 0 && j1 < rd.J) ? (R.tp[j1] & 3u) : 0u;
                const unsigned b2 = (cov_in && j2 >= 0 && j2 < rd.J) ? (R.tp[j2] & 3u) : 0u;
                word_sd |= b1 << (2 * x);
                word_in |= b2 << (2 * x);
            }
            if (gen_sd) bs_sd += base_ll;
            if (gen_in) bs_ia += base_ll;
#pragma unroll 1
            for (int m = 0; m < 8; ++m) {
                // m = 0..2 SUB, 3 DEL, 4..7 INS
                const bool is_ins = m >= 4;
                const int type = (m < 3) ? 0 : ((m == 3) ? 2 : 1);
                const int bf = (m < 3) ? ((tbase + 1 + m) & 3) : (is_ins ? m - 4 : 0);
                const int bl = rd.strand ? 3 - bf : bf;
                const bool lv = is_ins ? gen_in : (gen_sd && (m != 3 || rd.J >= 3));
                if (!__any_sync(kFullMask, is_ins ? gen_in : gen_sd)) continue;
                int e;
                float val = eval_mutation(R, g, s_emm, s_emi, is_ins ? word_in : word_sd, type,
                                          is_ins ? q_in : q_sd, bl, lv, e);
                if (m == 3 && rd.J < 3) val = 0.f;     // deleting one of two template bases leaves no template
                const bool take = is_ins ? gen_in : gen_sd;
                if (take) {
                    if (m < 3) {
#pragma unroll
                        for (int s2 = 0; s2 < 4; ++s2) if (s2 == bf) prod_mul(pr_sd[s2], px_sd[s2], val, e);
                    } else if (m == 3) {
                        prod_mul(pr_sd[4], px_sd[4], val, e);
                    } else {
#pragma unroll
                        for (int s2 = 0; s2 < 4; ++s2) if (s2 == bf) prod_mul(pr_ia[s2], px_ia[s2], val, e);
                    }
                }
            }
        }
    }
    if (have && g == 0) {
        double* out = delta + (size_t)(zm.delta_off + p) * kDeltaStride;
        // slots nobody contributed to keep delta-LL 0 (product 1, exponent 0, no base term)
#pragma unroll
        for (int k = 0; k < 5; ++k) out[k] = (k < 4 && k == tbase) ? 0.0 : prod_dll(pr_sd[k], px_sd[k], bs_sd);
#pragma unroll
        for (int k = 0; k < 4; ++k) out[5 + k] = prod_dll(pr_ia[k], px_ia[k], bs_ia);
        // reverse-strand insertions before the local position are forward INS(p+1)
        double* nxt = out + kDeltaStride;
#pragma unroll
        for (int k = 0; k < 4; ++k) nxt[9 + k] = prod_dll(pr_ib[k], px_ib[k], bs_ib);
        if (p == p_begin) {
#pragma unroll
            for (int k = 0; k < 4; ++k) out[9 + k] = 0.0;
        }
    }
}

// pick: canonical (de-duplicated) mutations with delta-LL > 0 inside the scored ranges
// (Polish(): "keep m with LL(m) > LL()"; dedup rule SURVEY.md A.7).  One thread per position.
__global__ void __launch_bounds__(256) arrow_pick_kernel(const ArrowBatchView V, const ScoreRange* __restrict__ ranges,
                                                         const int n_ranges, const long long n_items,
                                                         const double* __restrict__ delta, Candidate* __restrict__ out,
                                                         const int cap, int* __restrict__ counter) {
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    int lo = 0, hi = n_ranges - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ranges[mid].first <= item) lo = mid; else hi = mid - 1;
    }
    const ScoreRange rg = ranges[lo];
    const int z = rg.zmw;
    const int p = rg.p_begin + (int)(item - rg.first);
    const DevZmw zm = V.zmws[z];
    const uint8_t* t = V.tpl + zm.fwd_off;
    const int tb = t[p] & 3;                       // template bytes carry context bits above the base
    const int tprev = (p > 0) ? (t[p - 1] & 3) : -1;
    const double* d = delta + (size_t)(zm.delta_off + p) * kDeltaStride;
#pragma unroll 1
    for (int s = 0; s < 9; ++s) {
        const double v = (s >= 5) ? d[s] + d[s + 4] : d[s];
        if (!(v > 0.0)) continue;
        int type, base;
        if (s < 4) { type = 0; base = s; if (base == tb) continue; }
        else if (s == 4) { type = 2; base = 0; if (p > 0 && tb == tprev) continue; }
        else { type = 1; base = s - 5; if (!(p >= 1 && p <= zm.J - 1) || base == tprev) continue; }
        const int idx = atomicAdd(counter, 1);
        if (idx < cap) {
            Candidate c;
            c.score_hi = (float)v; c.zmw = z; c.pos = p; c.type = (int16_t)type; c.base = (int16_t)base; c.score = v;
            out[idx] = c;
        }
    }
}

// ConsensusQualities: QV_p = -10 log10(s/(1+s)), s = sum over the position's mutations of
// exp(delta-LL), clamped to [0,93] (docs/how-does-ccs-work.md:103-106; docs/faq/qv-binning.md:25-31).
__global__ void __launch_bounds__(256) arrow_qv_kernel(const ArrowBatchView V, const double* __restrict__ delta,
                                                       uint8_t* __restrict__ qv,
                                                       const long long n_items, const ScoreRange* __restrict__ ranges,
                                                       const int n_ranges) {
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    int lo = 0, hi = n_ranges - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ranges[mid].first <= item) lo = mid; else hi = mid - 1;
    }
    const ScoreRange rg = ranges[lo];
    const int z = rg.zmw;
    const int p = rg.p_begin + (int)(item - rg.first);
    const DevZmw zm = V.zmws[z];
    const int tb = V.tpl[zm.fwd_off + p] & 3;
    const double* d = delta + (size_t)(zm.delta_off + p) * kDeltaStride;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k < 4 && k == tb) continue;
        if (k >= 5 && !(p >= 1 && p <= zm.J - 1)) continue;
        s += exp((k >= 5) ? d[k] + d[k + 4] : d[k]);
    }
    double q = (s > 0.0) ? -10.0 * log10(s / (1.0 + s)) : 93.0;
    if (!(q < 93.0)) q = 93.0;
    if (q < 0.0) q = 0.0;
    qv[zm.delta_off + p] = (uint8_t)llrint(q);
}


// Re-index the delta rows of ZMWs whose template was just edited: row p' of the new template takes
// the row of the old position p = p' - shift(p'), shift = net length change of the edits before p'.
// Rows near an edit are re-scored in the next round anyway; rows far from every edit keep their
// delta-LLs (an edit more than `neighborhood` positions away does not change them beyond rounding),
// which is what lets ConsensusQualities reuse them instead of re-scoring the whole template.
__global__ void __launch_bounds__(256) arrow_remap_delta_kernel(const RemapJob* __restrict__ jobs, const int n_jobs,
                                                                const int32_t* __restrict__ sites,
                                                                const int32_t* __restrict__ shifts,
                                                                const double* __restrict__ src, double* __restrict__ dst,
                                                                const int to_scratch) {
    const int job = blockIdx.y;
    if (job >= n_jobs) return;
    const RemapJob jb = jobs[job];
    const int lane = threadIdx.x & 15;                       // 16 doubles per row
    for (int p = blockIdx.x * 16 + (threadIdx.x >> 4); p <= jb.J_new; p += gridDim.x * 16) {
        if (to_scratch) {
            // number of edits whose new-coordinate site is <= p
            int lo = 0, hi = jb.n_sites;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (sites[jb.site_off + mid] <= p) lo = mid + 1; else hi = mid; }
            const int sh = lo ? shifts[jb.site_off + lo - 1] : 0;
            const int q = p - sh;
            const double v = (q >= 0 && q <= jb.J_old) ? src[(size_t)(jb.delta_off + q) * kDeltaStride + lane] : 0.0;
            dst[(size_t)(jb.scratch_off + p) * kDeltaStride + lane] = v;
        } else {
            dst[(size_t)(jb.delta_off + p) * kDeltaStride + lane] = src[(size_t)(jb.scratch_off + p) * kDeltaStride + lane];
        }
    }
}

}  // namespace

void launch_score(const ArrowBatchView& V, const ScoreRange* ranges, int n_ranges, long long n_items, double* delta,
                  cudaStream_t stream, bool generic) {
    if (n_items <= 0) return;
    const long long blocks = (n_items + 15) / 16;
    if (generic) arrow_score_generic_kernel<<<(unsigned)blocks, 128, 0, stream>>>(V, ranges, n_ranges, n_items, delta);
    else arrow_score_kernel<<<(unsigned)blocks, 128, 0, stream>>>(V, ranges, n_ranges, n_items, delta);
}

void launch_pick(const ArrowBatchView& V, const ScoreRange* ranges, int n_ranges, long long n_items, const double* delta,
                 Candidate* out, int cap, int* counter, cudaStream_t stream) {
    if (n_items <= 0) return;
    const long long blocks = (n_items + 255) / 256;
    arrow_pick_kernel<<<(unsigned)blocks, 256, 0, stream>>>(V, ranges, n_ranges, n_items, delta, out, cap, counter);
}

void launch_qv(const ArrowBatchView& V, const double* delta, uint8_t* qv, long long n_items,
               const ScoreRange* ranges, int n_ranges, cudaStream_t stream) {
    if (n_items <= 0) return;
    const long long blocks = (n_items + 255) / 256;
    arrow_qv_kernel<<<(unsigned)blocks, 256, 0, stream>>>(V, delta, qv, n_items, ranges, n_ranges);
}

}  // namespace ccs

namespace ccs {
void launch_remap_delta(const RemapJob* jobs, int n_jobs, const int32_t* sites, const int32_t* shifts, double* delta,
                        double* scratch, cudaStream_t stream) {
    if (n_jobs <= 0) return;
    dim3 grid(64, n_jobs);
    arrow_remap_delta_kernel<<<grid, 256, 0, stream>>>(jobs, n_jobs, sites, shifts, delta, scratch, 1);
    arrow_remap_delta_kernel<<<grid, 256, 0, stream>>>(jobs, n_jobs, sites, shifts, scratch, delta, 0);
}
}  // namespace ccs
