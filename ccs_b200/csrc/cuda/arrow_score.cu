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
    const unsigned* rcA;     // row-code words: word k = codes of rows 4k..4k+3
    const unsigned* rcB;     // shifted copy: word k = codes of rows 4k+1..4k+4
    int wmax;
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
    const int n_ext = live ? max(0, last - a + 1) : 0;
    const int n_max = __reduce_max_sync(kFullMask, n_ext);      // warp-uniform trip count (provably convergent loop)

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
    bool have = item < n_items;
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
        have = p < rg.p_end;             // ranges are padded to multiples of 16 items
    }
    DevZmw zm;
    zm.read_begin = zm.read_end = 0; zm.fwd_off = 0; zm.J = 0; zm.delta_off = 0;
    if (have) zm = V.zmws[z];
    const int n_reads = zm.read_end - zm.read_begin;
    const int n_max = __reduce_max_sync(kFullMask, n_reads);
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
// => 4 + 8 in-column scans and 9 links per (read, position).  Band starts are multiples of 4
// (quantised slide), so every band mask is a per-LANE predicate, a lane's four row codes are one
// aligned word of the row-code array, and the per-ZMW folded factors sit in shared memory in
// CODE-MAJOR order [code][context]: a cell's table address is (cell base + context offset), with
// the part of the context that depends on the unrolled base b as an immediate.
// Reverse-strand reads compute the insertion BEFORE their local q, which is forward INS(p+1):
// those sums go to slots 9..12 of row p+1 (combined by the consumers).
// -------------------------------------------------------------------------------------------
// warp-wide OR through a reduction: the result lives in a uniform register, so the compiler knows that branches on it
// are convergent (a vote's result is not treated that way, and every shuffle behind such a branch gets wrapped in a
// WARPSYNC / ENDCOLLECTIVE pair)
__device__ __forceinline__ bool warp_any(const bool x) { return __reduce_or_sync(kFullMask, (unsigned)x) != 0u; }


// ---- band staging through shared memory (sm_90+ bulk copies, mbarrier completion) ------------------------------
// A scoring CTA works on 16 consecutive template positions of one ZMW.  Per read it needs alpha columns q-1 and beta
// columns q+1, q+2 of those positions: ONE contiguous window of <= 19 band columns (plus their column info and beta
// exponents).  Instead of 15 dependent global loads per (read, position), one elected thread issues four
// cp.async.bulk copies per (CTA, read) into a double-buffered stage while the previous read is being scored; the
// octets then read their operands with LDS.  (north_star: "TMA or shared-memory staging of the DP band".)
constexpr int kStageCols = 20;                 // band columns per stage (19 needed + alignment slack)
struct __align__(128) ScoreStage {
    float4 a[kStageCols * 8];                  // alpha columns [jw, jw + nw)
    float4 b[kStageCols * 8];                  // beta columns  [jw, jw + nw)
    ColInfo ci[kStageCols + 4];                // column info from an even (16-byte aligned) index
    int be[kStageCols + 8];                    // beta exponents from a multiple-of-4 index
};
struct StageView { const float4* a; const float4* b; const ColInfo* ci; const int* be; int jw; int on; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, const unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, const unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, const unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();       // a lost copy must kill the kernel, never hang the device
    }
}

// The window of read `rd` a CTA with first position P0 needs: columns [jw, jw + nw); nw = 0: nothing to stage.
__device__ __forceinline__ void stage_window(const DevRead& rd, const int st, const int P0, int& jw, int& nw) {
    jw = 0; nw = 0;
    if (!rd.active || st != 0 || !(P0 < rd.te && P0 + 16 > rd.ts) || rd.J < 6) return;
    int lo, hi;                                 // local positions q of the CTA's first / last covered position
    if (!rd.strand) { lo = P0 - rd.ts; hi = lo + 15; } else { hi = rd.te - 1 - P0; lo = hi - 15; }
    jw = min(max(lo - 1, 0), rd.J - 1);
    nw = max(0, min(rd.J, hi + 3) - jw);
    nw = min(nw, kStageCols);
}

__device__ __forceinline__ void stage_issue(const ArrowBatchView& V, const DevRead& rd, const int jw, const int nw,
                                            ScoreStage* S, uint64_t* bar) {
    const long long c0 = rd.col_off + jw;
    const long long ci0 = c0 & ~1ll, ci1 = (c0 + nw + 1) & ~1ll;
    const long long be0 = c0 & ~3ll, be1 = (c0 + nw + 3) & ~3ll;
    const unsigned bytes_col = (unsigned)nw * 128u, bytes_ci = (unsigned)(ci1 - ci0) * 8u, bytes_be = (unsigned)(be1 - be0) * 4u;
    mbar_expect_tx(bar, 2u * bytes_col + bytes_ci + bytes_be);
    bulk_g2s(S->a, V.alpha + c0 * 32, bytes_col, bar);
    bulk_g2s(S->b, V.beta + c0 * 32, bytes_col, bar);
    bulk_g2s(S->ci, V.colinfo + ci0, bytes_ci, bar);
    bulk_g2s(S->be, V.beta_exp + be0, bytes_be, bar);
}

struct FastOut { float sub[4]; float del; float ins[4]; int e_sd; int e_in; };   // per-lane partial link sums

constexpr int kReadCache = 48;     // read descriptors of the CTA's ZMW kept in shared memory (more reads: global loads)

// Code-major table geometry: one code = 16 contexts x {mm, gg} = 32 words, padded to 33 so that the lanes of a warp
// (same context per octet, different codes) fall into different shared-memory banks.
constexpr int kTcCodeWords = 33;
constexpr int kTcCodeBytes = 4 * kTcCodeWords;

// byte address of code x of a row-code word (codes are stored x4) in the code-major table: tc + code * 132
__device__ __forceinline__ unsigned code_base(const unsigned tc, const unsigned w, const int x) {
    const unsigned byte = (x == 0) ? (w & 0xffu) : (x == 3) ? (w >> 24) : __byte_perm(w, 0u, 0x4440u + x);
    return byte * (kTcCodeBytes / 4) + tc;
}

// a_i = A_i + G_i * a_{i-1} around the octet ring (G of the band-start cell is 0)
__device__ __forceinline__ void ring_scan(float A[4], float G[4], const int src1, const int src2, const int src4, float v[4]) {
    A[1] = fmaf(G[1], A[0], A[1]); G[1] *= G[0];
    A[2] = fmaf(G[2], A[1], A[2]); G[2] *= G[1];
    A[3] = fmaf(G[3], A[2], A[3]); G[3] *= G[2];
    float At = A[3], Gt = G[3];
    float As = __shfl_sync(kFullMask, At, src1, 8), Gs = __shfl_sync(kFullMask, Gt, src1, 8);
    At = fmaf(Gt, As, At); Gt *= Gs;
    As = __shfl_sync(kFullMask, At, src2, 8); Gs = __shfl_sync(kFullMask, Gt, src2, 8);
    At = fmaf(Gt, As, At); Gt *= Gs;
    As = __shfl_sync(kFullMask, At, src4, 8);
    At = fmaf(Gt, As, At);
    const float x = __shfl_sync(kFullMask, At, src1, 8);
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = fmaf(G[q], x, A[q]);
}

// beta column masked onto the rows of an extension band: own = this lane's rows coincide, nxt = the next lane's first
// row is this lane's last row + 1
__device__ __forceinline__ void link_operands(const float4 b, const bool own, const bool nxt, const int src_dn,
                                              float bx[4], float dn[4]) {
    const float dn3 = __shfl_sync(kFullMask, b.x, src_dn, 8);
    bx[0] = own ? b.x : 0.f; bx[1] = own ? b.y : 0.f; bx[2] = own ? b.z : 0.f; bx[3] = own ? b.w : 0.f;
    dn[0] = bx[1]; dn[1] = bx[2]; dn[2] = bx[3]; dn[3] = nxt ? dn3 : 0.f;
}

// per-lane partial of sum_i y_i * (mm[ctx][code(i+1)] * beta(i+1) + D[ctx] * beta(i))
__device__ __forceinline__ float link_partial(const float y[4], const float bx[4], const float dn[4], const unsigned a0,
                                              const unsigned a1, const unsigned a2, const unsigned a3, const int imm,
                                              const float D) {
    float acc = y[0] * fmaf(lds_f1(a0 + imm), dn[0], D * bx[0]);
    acc = fmaf(y[1], fmaf(lds_f1(a1 + imm), dn[1], D * bx[1]), acc);
    acc = fmaf(y[2], fmaf(lds_f1(a2 + imm), dn[2], D * bx[2]), acc);
    acc = fmaf(y[3], fmaf(lds_f1(a3 + imm), dn[3], D * bx[3]), acc);
    return acc;
}

// tc = shared-memory byte address of the code-major folded table [16 codes][16 contexts] of {mm, gg}
template <int kUnrollB>
__device__ __forceinline__ void fast_eval(const ReadCtx& R, const StageView& SV, const int g4, const int src_up, const int src_up2,
                                          const int src_up4, const int src_dn, const unsigned tc, const int q_in,
                                          const bool live, FastOut& out) {
    const int J = R.J;
    const int q = live ? q_in : 2;
    const int jm1 = min(max(q - 1, 0), J - 1), j0 = min(q, J - 1), j1 = min(q + 1, J - 1), j2 = min(q + 2, J - 1);
    ColInfo cim1;
    int s1, s2, s3, be1, be2;
    float4 a0, b1, b2;
    if (SV.on) {                                // operands from the staged window (LDS); idle octets read its first columns
        const int w = live ? jm1 - SV.jw : 0;   // interior positions lie inside the window by construction
        cim1 = SV.ci[w];
        s1 = SV.ci[w + (j0 - jm1)].start; s2 = SV.ci[w + (j1 - jm1)].start; s3 = SV.ci[w + (j2 - jm1)].start;
        a0 = SV.a[w * 8];
        b1 = SV.b[(w + (j1 - jm1)) * 8];
        b2 = SV.b[(w + (j2 - jm1)) * 8];
        be1 = SV.be[w + (j1 - jm1)]; be2 = SV.be[w + (j2 - jm1)];
    } else {
        cim1 = R.cinfo[jm1];
        s1 = R.cinfo[j0].start; s2 = R.cinfo[j1].start; s3 = R.cinfo[j2].start;
        a0 = R.acol[(size_t)jm1 * 8];
        b1 = R.bcol[(size_t)j1 * 8];
        b2 = R.bcol[(size_t)j2 * 8];
        be1 = R.bexp[j1]; be2 = R.bexp[j2];
    }
    const int s0 = cim1.start;
    out.e_sd = cim1.cumexp + be2;
    out.e_in = cim1.cumexp + be1;
    // template bytes carry 16*t[j-2] + 4*t[j-1] + t[j]
    const int x0 = R.tp[j0], x1 = R.tp[j1];
    const int t0 = x0 & 3, tm1 = (x0 >> 2) & 3, cmq = (x0 >> 2) & 15, tp1 = x1 & 3;

    // lane geometry: first row of this lane in each band (band starts are multiples of 4)
    const int rb0 = s0 + ((g4 - s0) & 31), rb1 = s1 + ((g4 - s1) & 31), rb2 = s2 + ((g4 - s2) & 31);
    const int rb3 = s3 + ((g4 - s3) & 31);
    const int g4n = (g4 + 4) & 31;
    const int rb2n = s2 + ((g4n - s2) & 31), rb3n = s3 + ((g4n - s3) & 31);
    // row codes: word k of copy A = rows 4k..4k+3, of copy B = rows 4k+1..4k+4
    const unsigned wA1 = R.rcA[min(max(rb1 >> 2, 0), R.wmax)], wA2 = R.rcA[min(max(rb2 >> 2, 0), R.wmax)];
    const unsigned wB1 = R.rcB[min(max(rb1 >> 2, 0), R.wmax)], wB2 = R.rcB[min(max(rb2 >> 2, 0), R.wmax)];

    // ---- first extension column (band s1) from alpha column q-1 (band s0): shared match/deletion terms
    const bool start1 = rb1 == s1, start2 = rb2 == s2;
    float up = __shfl_sync(kFullMask, a0.w, src_up, 8);
    if (start1 && s1 == s0) up = 0.f;                        // row s1-1 is outside band s0
    const bool keep1 = rb1 == rb0;                           // false: the lane re-entered at the top, new rows
    const float v0[4] = {keep1 ? a0.x : 0.f, keep1 ? a0.y : 0.f, keep1 ? a0.z : 0.f, keep1 ? a0.w : 0.f};
    const unsigned c1[4] = {code_base(tc, wA1, 0), code_base(tc, wA1, 1), code_base(tc, wA1, 2), code_base(tc, wA1, 3)};
    const unsigned c2[4] = {code_base(tc, wA2, 0), code_base(tc, wA2, 1), code_base(tc, wA2, 2), code_base(tc, wA2, 3)};
    const unsigned n1[4] = {code_base(tc, wB1, 0), code_base(tc, wB1, 1), code_base(tc, wB1, 2), code_base(tc, wB1, 3)};
    const unsigned n2[4] = {code_base(tc, wB2, 0), code_base(tc, wB2, 1), code_base(tc, wB2, 2), code_base(tc, wB2, 3)};
    const unsigned dbase = tc + kDSlot * kTcCodeBytes;       // deletion transitions ride in code slot 13
    float A1[4];
    {
        const float Dq = lds_f1(dbase + cmq * 8);
        A1[0] = fmaf(lds_f1(c1[0] + cmq * 8), up, Dq * v0[0]);
        A1[1] = fmaf(lds_f1(c1[1] + cmq * 8), v0[0], Dq * v0[1]);
        A1[2] = fmaf(lds_f1(c1[2] + cmq * 8), v0[1], Dq * v0[2]);
        A1[3] = fmaf(lds_f1(c1[3] + cmq * 8), v0[2], Dq * v0[3]);
    }
    // ---- link operands
    float bxS[4], dnS[4], bxI[4], dnI[4], bxD[4], dnD[4];
    link_operands(b2, rb3 == rb2, rb3n == rb2 + 4, src_dn, bxS, dnS);     // band s2 -> beta column q+2 (band s3)
    link_operands(b1, true, rb2n == rb2 + 4, src_dn, bxI, dnI);           // band s2 -> beta column q+1 (band s2)
    link_operands(b2, rb3 == rb1, rb3n == rb1 + 4, src_dn, bxD, dnD);     // band s1 -> beta column q+2 (band s3)
    // per-cell table bases with the runtime part of the context folded in; the unrolled base b adds an immediate
    const int o_m1 = tm1 * 32, o_p1 = tp1 * 8, o_t0 = t0 * 8;
    const unsigned g1a[4] = {c1[0] + o_m1, c1[1] + o_m1, c1[2] + o_m1, c1[3] + o_m1};   // (tm1, b): + 8 b
    const unsigned m2a[4] = {c2[0] + o_m1, c2[1] + o_m1, c2[2] + o_m1, c2[3] + o_m1};   // (tm1, b): + 8 b
    const unsigned s2a[4] = {c2[0] + o_p1, c2[1] + o_p1, c2[2] + o_p1, c2[3] + o_p1};   // (b, tp1): + 32 b
    const unsigned i2a[4] = {c2[0] + o_t0, c2[1] + o_t0, c2[2] + o_t0, c2[3] + o_t0};   // (b, t0):  + 32 b
    const unsigned sla[4] = {n2[0] + o_p1, n2[1] + o_p1, n2[2] + o_p1, n2[3] + o_p1};   // link (b, tp1)
    const unsigned ila[4] = {n2[0] + o_t0, n2[1] + o_t0, n2[2] + o_t0, n2[3] + o_t0};   // link (b, t0)
    const unsigned d_m1 = dbase + o_m1, d_p1 = dbase + o_p1, d_t0 = dbase + o_t0;
    const bool keep2 = rb2 == rb1;
    const bool upok2 = !(start2 && s2 == s1);
    float Xdel[4] = {0.f, 0.f, 0.f, 0.f};
    float osub[4], oins[4];
#pragma unroll kUnrollB
    for (int b = 0; b < 4; ++b) {
        float A[4], G[4], X[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) { A[x] = A1[x]; G[x] = lds_f1(g1a[x] + 8 * b + 4); }
        if (start1) G[0] = 0.f;
        ring_scan(A, G, src_up, src_up2, src_up4, X);
        if (b == tp1) {
#pragma unroll
            for (int x = 0; x < 4; ++x) Xdel[x] = X[x];
        }
        // second extension column (band s2): match/deletion terms shared by SUB(b) and INS(b)
        float up2 = __shfl_sync(kFullMask, X[3], src_up, 8);
        if (!upok2) up2 = 0.f;
        const float Xm[4] = {keep2 ? X[0] : 0.f, keep2 ? X[1] : 0.f, keep2 ? X[2] : 0.f, keep2 ? X[3] : 0.f};
        const float Dm = lds_f1(d_m1 + 8 * b);
        float A2[4];
        A2[0] = fmaf(lds_f1(m2a[0] + 8 * b), up2, Dm * Xm[0]);
        A2[1] = fmaf(lds_f1(m2a[1] + 8 * b), Xm[0], Dm * Xm[1]);
        A2[2] = fmaf(lds_f1(m2a[2] + 8 * b), Xm[1], Dm * Xm[2]);
        A2[3] = fmaf(lds_f1(m2a[3] + 8 * b), Xm[2], Dm * Xm[3]);
        {   // SUB(q, b): insertion context (b, t[q+1]); link into beta column q+2
            float Aa[4], Ga[4], Y[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { Aa[x] = A2[x]; Ga[x] = lds_f1(s2a[x] + 32 * b + 4); }
            if (start2) Ga[0] = 0.f;
            ring_scan(Aa, Ga, src_up, src_up2, src_up4, Y);
            osub[b] = link_partial(Y, bxS, dnS, sla[0], sla[1], sla[2], sla[3], 32 * b, lds_f1(d_p1 + 32 * b));
        }
        {   // INS(before q, b): insertion context (b, t[q]); link into beta column q+1
            float Aa[4], Ga[4], Z[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { Aa[x] = A2[x]; Ga[x] = lds_f1(i2a[x] + 32 * b + 4); }
            if (start2) Ga[0] = 0.f;
            ring_scan(Aa, Ga, src_up, src_up2, src_up4, Z);
            oins[b] = link_partial(Z, bxI, dnI, ila[0], ila[1], ila[2], ila[3], 32 * b, lds_f1(d_t0 + 32 * b));
        }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) { out.sub[b] = osub[b]; out.ins[b] = oins[b]; }
    {   // DEL(q): X_{t[q+1]} links straight into beta column q+2 with context (t[q-1], t[q+1])
        const int od = o_m1 + o_p1;
        out.del = link_partial(Xdel, bxD, dnD, n1[0] + od, n1[1] + od, n1[2] + od, n1[3] + od, 0, lds_f1(dbase + od));
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

// Work items are (ZMW, position); `first` of every range is a multiple of 16, so the 16 octets of a CTA belong to one
// ZMW and share its folded factor table.  The octet loops over the ZMW's reads in index order.  The eight mutation
// slots of a position are spread over the octet's lanes: lane k < 4 owns SUB(forward base k) -- or DEL when k is the
// template base, whose substitution is the identity -- and lane 4 + k owns INS(forward base k); per read the lanes'
// partial link sums are transpose-reduced so that every lane ends up with the total of its own slot and keeps that
// slot's running product.  The order of the reduction is fixed: results are deterministic.
template <int kMinBlocks, int kUnrollB, bool kStage>
__global__ void __launch_bounds__(128, kMinBlocks) arrow_score_kernel(const ArrowBatchView V, const ScoreRange* __restrict__ ranges,
                                                          const int n_ranges, const long long n_items,
                                                          double* __restrict__ delta) {
    __shared__ float s_emm[36 * kEmStride];
    __shared__ float s_emi[17 * kEmStride];
    __shared__ __align__(16) float s_tc[16 * kTcCodeWords];  // code-major folded factors of the CTA's ZMW
    __shared__ DevRead s_rd[kReadCache];                     // the ZMW's read descriptors, statuses, base LLs
    __shared__ int s_st[kReadCache];
    __shared__ double s_bl[kReadCache];
    __shared__ ScoreStage s_stage[kStage ? 2 : 1];
    __shared__ __align__(8) uint64_t s_bar[2];
    if (kStage && threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int k = threadIdx.x; k < 36 * kEmStride; k += blockDim.x) s_emm[k] = V.em_match[k];
    for (int k = threadIdx.x; k < 17 * kEmStride; k += blockDim.x) s_emi[k] = V.em_ins[k];

    const long long item0 = (long long)blockIdx.x * 16;
    const long long item = item0 + (threadIdx.x >> 3);
    const int g = pinned(threadIdx.x & 7);
    ScoreRange rg;
    {
        int lo = 0, hi = n_ranges - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ranges[mid].first <= item0) lo = mid; else hi = mid - 1;
        }
        rg = ranges[lo];
    }
    const int z = rg.zmw;
    const int p = rg.p_begin + (int)(item - rg.first);
    const int p_begin = rg.p_begin;
    const bool have = item < n_items && p < rg.p_end;
    {
        const float4* __restrict__ tr = reinterpret_cast<const float4*>(V.trans) + (size_t)z * 36;
        for (int idx = threadIdx.x; idx < 256; idx += blockDim.x) {
            const int code = idx >> 4, ctx = idx & 15;
            const float2 e = folded_entry(V.em_match, V.em_ins, tr, ctx, ctx, code);
            s_tc[code * kTcCodeWords + 2 * ctx] = e.x;
            s_tc[code * kTcCodeWords + 2 * ctx + 1] = e.y;
        }
    }
    int nrc = 0;
    if (item0 < n_items) {
        const DevZmw zc = V.zmws[z];
        nrc = min(zc.read_end - zc.read_begin, kReadCache);
        const int* src = reinterpret_cast<const int*>(V.reads + zc.read_begin);
        for (int idx = threadIdx.x; idx < nrc * (int)(sizeof(DevRead) / 4); idx += blockDim.x) reinterpret_cast<int*>(s_rd)[idx] = src[idx];
        for (int idx = threadIdx.x; idx < nrc; idx += blockDim.x) {
            s_st[idx] = V.status[zc.read_begin + idx];
            s_bl[idx] = V.base_ll[zc.read_begin + idx];
        }
    }
    __syncthreads();
    const unsigned tc = (unsigned)pinned((int)__cvta_generic_to_shared(s_tc));
    const int g4 = pinned(4 * g);
    const int src_up = pinned((g + 7) & 7), src_up2 = pinned((g + 6) & 7), src_up4 = pinned((g + 4) & 7);
    const int src_dn = pinned((g + 1) & 7);

    DevZmw zm;
    zm.read_begin = zm.read_end = 0; zm.fwd_off = 0; zm.J = 0; zm.delta_off = 0;
    if (have) zm = V.zmws[z];
    const int n_reads = zm.read_end - zm.read_begin;
    // warp-uniform trip count through a reduction: the compiler then knows the read loop is convergent and does not
    // wrap every shuffle inside it in WARPSYNC / ENDCOLLECTIVE pairs
    int n_max = __reduce_max_sync(kFullMask, n_reads);
    const int tbase = have ? (V.tpl[zm.fwd_off + p] & 3) : 0;
    // staged variant: the read loop carries a CTA barrier, so every warp of the CTA runs the same number of iterations
    // (the CTA's ZMW has nrc cached reads; reads beyond the cache are scored from global memory)
    const int P0 = rg.p_begin + (int)(item0 - rg.first);          // first position of this CTA
    unsigned ph0 = 0, ph1 = 0;                                     // mbarrier phase parity of the two stages
    if (kStage) {
        int n_cta = 0;
        if (item0 < n_items) { const DevZmw zc = V.zmws[z]; n_cta = zc.read_end - zc.read_begin; }
        n_max = __reduce_max_sync(kFullMask, n_cta);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k0 = 0; k0 < 2; ++k0)
                if (k0 < nrc) {
                    int jw, nw;
                    stage_window(s_rd[k0], s_st[k0], P0, jw, nw);
                    if (nw > 0) stage_issue(V, s_rd[k0], jw, nw, &s_stage[k0], &s_bar[k0]);
                }
        }
    }

    // this lane's slot: running product(s) and the sums of the contributing reads' base log-likelihoods
    float prA = 1.f, prB = 1.f;          // A: SUB / DEL / INS from forward reads;  B: INS' (reverse-strand share of row p+1)
    int pxA = 0, pxB = 0;
    double bs_sd = 0.0, bs_ia = 0.0, bs_ib = 0.0;

    for (int k = 0; k < n_max; ++k) {
        const bool has_read = k < n_reads;
        DevRead rd;
        rd.active = 0; rd.ts = rd.te = 0; rd.strand = 0; rd.I = 2; rd.J = 2; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0;
        rd.zmw = 0; rd.last_code = 0; rd.code_stride = 16;
        int st = 1;
        const int r = zm.read_begin + k;
        if (has_read) {
            if (k < nrc) { rd = s_rd[k]; st = s_st[k]; } else { rd = V.reads[r]; st = V.status[r]; }
        }
        // start the next read's three column lines (and its column info) on their way from HBM to L2 while this one
        // is being scored: lanes 0..3 of the octet take one line each
        if (!kStage && have && k + 1 < nrc) {
            const DevRead& nx = s_rd[k + 1];
            if (nx.active && p >= nx.ts && p < nx.te) {
                const int qn = nx.strand ? nx.te - 1 - p : p - nx.ts;
                const int col = min(max((g == 0) ? qn - 1 : ((g == 1) ? qn + 1 : ((g == 2) ? qn + 2 : qn - 1)), 0), nx.J - 1);
                const void* a = (g == 0) ? (const void*)(V.alpha + ((size_t)nx.col_off + col) * 32)
                              : (g == 3) ? (const void*)(V.colinfo + nx.col_off + col)
                                         : (const void*)(V.beta + ((size_t)nx.col_off + col) * 32);
                if (g < 4) asm volatile("prefetch.global.L2 [%0];" :: "l"(a));
            }
        }
        const bool usable = has_read && rd.active && st == 0;
        if (!usable) {   // idle octets still execute the shared instruction stream: give them harmless operands
            rd.ts = rd.te = 0; rd.strand = 0; rd.I = 2; rd.J = 2; rd.code_off = 0; rd.col_off = 0; rd.tpl_off = 0;
            rd.zmw = 0; rd.last_code = 0; rd.code_stride = 16;
        }
        const bool cov_sd = usable && p >= rd.ts && p < rd.te;       // SUB / DEL
        const bool cov_in = usable && p > rd.ts && p < rd.te;        // INS (before p)
        StageView SV;
        SV.a = nullptr; SV.b = nullptr; SV.ci = nullptr; SV.be = nullptr; SV.jw = 0; SV.on = 0;
        if (kStage && k < nrc) {       // wait for this read's window (CTA-uniform decision from the cached descriptor)
            int jw, nw;
            stage_window(s_rd[k], s_st[k], P0, jw, nw);
            if (nw > 0) {
                const int bsel = k & 1;
                mbar_wait(&s_bar[bsel], bsel ? ph1 : ph0);
                if (bsel) ph1 ^= 1u; else ph0 ^= 1u;
                const long long c0 = s_rd[k].col_off + jw;
                SV.a = s_stage[bsel].a + g; SV.b = s_stage[bsel].b + g;
                SV.ci = s_stage[bsel].ci + (int)(c0 & 1ll); SV.be = s_stage[bsel].be + (int)(c0 & 3ll);
                SV.jw = jw; SV.on = 1;
            }
        }
        if (warp_any(cov_sd)) {

        ReadCtx R;
        R.rc = V.rowcode + rd.code_off;
        R.rcA = reinterpret_cast<const unsigned*>(V.rowcode + rd.code_off);
        R.rcB = reinterpret_cast<const unsigned*>(V.rowcode + rd.code_off + rd.code_stride);
        R.wmax = max((rd.code_stride >> 2) - 1, 0);
        R.tp = V.tpl + rd.tpl_off;
        R.tr = reinterpret_cast<const float4*>(V.trans) + (size_t)rd.zmw * 36;
        R.acol = reinterpret_cast<const float4*>(V.alpha) + (size_t)rd.col_off * 8 + g;
        R.bcol = reinterpret_cast<const float4*>(V.beta) + (size_t)rd.col_off * 8 + g;
        R.cinfo = V.colinfo + rd.col_off;
        R.bexp = V.beta_exp + rd.col_off;
        R.I = rd.I; R.J = rd.J; R.last_code = rd.last_code;
        const double base_ll = cov_sd ? ((k < nrc) ? s_bl[k] : V.base_ll[r]) : 0.0;
        const int q_sd = cov_sd ? (rd.strand ? rd.te - 1 - p : p - rd.ts) : 0;
        const bool interior = cov_sd && q_sd >= 2 && q_sd <= rd.J - 4;
        const bool gen_sd = cov_sd && !interior;
        const bool prev_interior = rd.strand && p > p_begin && (q_sd + 1 >= 2) && (q_sd + 1 <= rd.J - 4);
        const bool gen_in = cov_in && (rd.strand ? !prev_interior : !interior);

        if (warp_any(interior)) {
            FastOut fo;
            fast_eval<kUnrollB>(R, SV, g4, src_up, src_up2, src_up4, src_dn, tc, q_sd, interior, fo);
            // Q[k] = this lane's partial of the slot lane k owns (forward-strand base k & 3; reverse reads see 3 - base)
            const bool rv = rd.strand != 0;
            float Q[8];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float sb = rv ? fo.sub[3 - b] : fo.sub[b];
                Q[b] = (b == tbase) ? fo.del : sb;
                Q[4 + b] = rv ? fo.ins[3 - b] : fo.ins[b];
            }
            // transpose-reduce over the octet: after the three steps lane k holds the octet-wide sum of Q[k]
            const bool h4 = (g & 4) != 0, h2 = (g & 2) != 0, h1 = (g & 1) != 0;
            float R4[4], R2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float send = h4 ? Q[i] : Q[i + 4], keep = h4 ? Q[i + 4] : Q[i];
                R4[i] = keep + __shfl_xor_sync(kFullMask, send, 4, 8);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float send = h2 ? R4[i] : R4[i + 2], keep = h2 ? R4[i + 2] : R4[i];
                R2[i] = keep + __shfl_xor_sync(kFullMask, send, 2, 8);
            }
            const float send = h1 ? R2[0] : R2[1], keep = h1 ? R2[1] : R2[0];
            const float tot = keep + __shfl_xor_sync(kFullMask, send, 1, 8);
            if (interior) {
                bs_sd += base_ll;
                if (rv) bs_ib += base_ll; else bs_ia += base_ll;
                const int e = (g < 4) ? fo.e_sd : fo.e_in;
                if (g >= 4 && rv) prod_mul(prB, pxB, tot, e); else prod_mul(prA, pxA, tot, e);
            }
        }
        if (warp_any(gen_sd) || warp_any(gen_in)) {
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
            R.acol -= g; R.bcol -= g;    // the generic evaluator indexes the lane itself
#pragma unroll 1
            for (int m = 0; m < 8; ++m) {
                // m = 0..2 SUB, 3 DEL, 4..7 INS
                const bool is_ins = m >= 4;
                const int type = (m < 3) ? 0 : ((m == 3) ? 2 : 1);
                const int bf = (m < 3) ? ((tbase + 1 + m) & 3) : (is_ins ? m - 4 : 0);
                const int bl = rd.strand ? 3 - bf : bf;
                const bool lv = is_ins ? gen_in : (gen_sd && (m != 3 || rd.J >= 3));
                if (!warp_any(is_ins ? gen_in : gen_sd)) continue;
                int e;
                float val = eval_mutation(R, g, s_emm, s_emi, is_ins ? word_in : word_sd, type,
                                          is_ins ? q_in : q_sd, bl, lv, e);
                if (m == 3 && rd.J < 3) val = 0.f;     // deleting one of two template bases leaves no template
                const bool take = is_ins ? gen_in : gen_sd;
                const int owner = (m < 3) ? bf : ((m == 3) ? tbase : 4 + bf);
                if (take && g == owner) prod_mul(prA, pxA, val, e);
            }
        }
        }   // warp_any(cov_sd)
        if (kStage) {
            __syncthreads();               // every warp is done with stage k & 1
            if (threadIdx.x == 0 && k + 2 < nrc) {
                int jw, nw;
                stage_window(s_rd[k + 2], s_st[k + 2], P0, jw, nw);
                if (nw > 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before async writes
                    stage_issue(V, s_rd[k + 2], jw, nw, &s_stage[k & 1], &s_bar[k & 1]);
                }
            }
        }
    }
    if (have) {
        double* out = delta + (size_t)(zm.delta_off + p) * kDeltaStride;
        // slots nobody contributed to keep delta-LL 0 (product 1, exponent 0, no base term)
        if (g < 4) {
            const double v = prod_dll(prA, pxA, bs_sd);
            if (g == tbase) { out[4] = v; out[g] = 0.0; } else out[g] = v;
        } else {
            out[5 + (g - 4)] = prod_dll(prA, pxA, bs_ia);
            // reverse-strand insertions before the local position are forward INS(p+1); the last position of a range
            // keeps them to itself (row p_end belongs to nobody here, or to an earlier launch whose row must stay
            // self-consistent when stored delta-LLs are reused)
            if (p + 1 < rg.p_end) out[kDeltaStride + 9 + (g - 4)] = prod_dll(prB, pxB, bs_ib);
            if (p == p_begin) out[9 + (g - 4)] = 0.0;
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
    if (p >= rg.p_end) return;           // padding item (ranges start at multiples of 16)
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
    if (p >= rg.p_end) return;           // padding item (ranges start at multiples of 16)
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
                  cudaStream_t stream, bool generic, int variant) {
    if (n_items <= 0) return;
    const long long blocks = (n_items + 15) / 16;
    if (generic) arrow_score_generic_kernel<<<(unsigned)blocks, 128, 0, stream>>>(V, ranges, n_ranges, n_items, delta);
    else if (variant == 1) arrow_score_kernel<4, 4, false><<<(unsigned)blocks, 128, 0, stream>>>(V, ranges, n_ranges, n_items, delta);
    else if (variant == 2) arrow_score_kernel<5, 4, true><<<(unsigned)blocks, 128, 0, stream>>>(V, ranges, n_ranges, n_items, delta);
    else if (variant == 3) arrow_score_kernel<4, 4, true><<<(unsigned)blocks, 128, 0, stream>>>(V, ranges, n_ranges, n_items, delta);
    // default: 5 CTAs per SM (96 registers, 20 warps): measured 9 % faster than 4 CTAs at 128 registers despite the
    // extra spills -- the kernel is latency / issue bound and wants the warps (profiles/r2_score_variants.txt)
    else arrow_score_kernel<5, 4, false><<<(unsigned)blocks, 128, 0, stream>>>(V, ranges, n_ranges, n_items, delta);
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
