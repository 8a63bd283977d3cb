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
            code[x] = real ? R.rc[min(S_new + rel[x], code_max)] : 12;
        }
        float A[4], G[4];
        octet_forward_terms(v, g, d, rel, code, trm.x, trm.y, tri.z, tri.w, s_emm + (real ? cm : 0) * kEmStride,
                            s_emi + (real ? ci : 0) * kEmStride, ci & 3, A, G);
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
            const int c1 = lk ? R.rc[min(row + 1, code_max)] : 12;
            const float wv = fmaf(trl.x, emm_row[c1] * bd, trl.y * bx);
            acc = fmaf(v[x], wv, acc);
        }
        link_v = octet_sum(acc);
    }
    exp_out = e + (terminal ? 0 : link_e);
    return terminal ? term_v : link_v;
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
    const int tbase = have ? V.tpl[zm.fwd_off + p] : 0;

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
                const unsigned b1 = (cov_sd && j1 >= 0 && j1 < rd.J) ? R.tp[j1] : 0u;
                const unsigned b2 = (cov_in && j2 >= 0 && j2 < rd.J) ? R.tp[j2] : 0u;
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
        double* out = delta + (size_t)(zm.delta_off + p) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) out[k] = acc[k];
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
    const int tb = t[p];
    const int tprev = (p > 0) ? t[p - 1] : -1;
    const double* d = delta + (size_t)(zm.delta_off + p) * 9;
#pragma unroll 1
    for (int s = 0; s < 9; ++s) {
        const double v = d[s];
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
    const int tb = V.tpl[zm.fwd_off + p];
    const double* d = delta + (size_t)(zm.delta_off + p) * 9;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k < 4 && k == tb) continue;
        if (k >= 5 && !(p >= 1 && p <= zm.J - 1)) continue;
        s += exp(d[k]);
    }
    double q = (s > 0.0) ? -10.0 * log10(s / (1.0 + s)) : 93.0;
    if (!(q < 93.0)) q = 93.0;
    if (q < 0.0) q = 0.0;
    qv[zm.delta_off + p] = (uint8_t)llrint(q);
}

}  // namespace

void launch_score(const ArrowBatchView& V, const ScoreRange* ranges, int n_ranges, long long n_items, double* delta,
                  cudaStream_t stream) {
    if (n_items <= 0) return;
    const long long blocks = (n_items + 15) / 16;
    arrow_score_kernel<<<(unsigned)blocks, 128, 0, stream>>>(V, ranges, n_ranges, n_items, delta);
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
