/* ccsgpu.h -- C ABI of the B200-native CCS consensus engine (libccsgpu.so).
 *
 * The reference (`ccs`, closed source; /root/reference is its documentation only) exposes
 * no plugin / FFI interface: it is one static binary.  Its own architecture has exactly
 * one seam on the hot path -- the Draft Stage and the Polish Stage, each with
 * interchangeable GPU / CPU-pool back ends (/root/reference/docs/img/ccs-impl.png;
 * "Arrow polishing is performed on GPU", /root/reference/docs/faq/revio.md:14,20-25).
 * This header reproduces that seam as plain C: pointers + sizes, caller-owned host
 * buffers, no exceptions across the boundary, no torch types.  INTEGRATION.md shows the
 * bindings (ctypes stub, C++ stage adaptor).
 *
 * Threading: one ccsgpu_ctx per GPU, one host thread per ctx (the reference's "-j" worker
 * threads feed a GPU stage through a queue, docs/faq/parallelize.md:17); distinct ctxs
 * share nothing.  All functions return CCS_OK (0) or a negative ccs_error.
 * There is NO CPU fallback: without a usable CUDA device every compute entry point
 * returns CCS_ERR_NO_DEVICE.
 */
#ifndef CCSGPU_H
#define CCSGPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ccs_error {
    CCS_OK = 0,
    CCS_ERR_CAPACITY = -1,   /* a caller buffer is too small (required size is written back) */
    CCS_ERR_CUDA = -2,       /* CUDA runtime failure; see ccsgpu_last_error() */
    CCS_ERR_ARG = -3,
    CCS_ERR_NO_DEVICE = -4,
    CCS_ERR_OOM = -5,        /* batch does not fit the ctx's device budget */
    CCS_ERR_IO = -6,
    CCS_ERR_CHEMISTRY = -7   /* unknown / missing chemistry: fatal in the reference too
                                (docs/faq/chemistry.md:7-10, docs/changelog.md:66) */
} ccs_error;

/* Per-ZMW outcome, in the order the reference reports them
 * (/root/reference/docs/faq/reports-aux-files.md:143-159). */
typedef enum ccs_zmw_status {
    CCS_ZMW_POOR_SNR = 0,
    CCS_ZMW_NO_SUBREADS,
    CCS_ZMW_TOO_FEW_PASSES,
    CCS_ZMW_LOW_PASS_SHORTCUT,
    CCS_ZMW_HETERODUPLEXES,
    CCS_ZMW_COVERAGE_DROPS,
    CCS_ZMW_INSUFFICIENT_SPANS,
    CCS_ZMW_TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT,
    CCS_ZMW_DRAFT_FAILURE,
    CCS_ZMW_TOO_LONG,
    CCS_ZMW_TOO_SHORT,
    CCS_ZMW_TOO_MANY_UNUSABLE,
    CCS_ZMW_EMPTY_WINDOW_DURING_POLISHING,
    CCS_ZMW_NON_CONVERGENT,
    CCS_ZMW_POOR_QUALITY,
    CCS_ZMW_EXCEPTION_THROWN,
    CCS_ZMW_SUCCESS
} ccs_zmw_status;

/* Per-read outcome of Recursor::FillAlphaBeta. */
typedef enum ccs_read_status {
    CCS_READ_VALID = 0,
    CCS_READ_ALPHA_BETA_MISMATCH = 1,
    CCS_READ_TEMPLATE_TOO_SMALL = 2,
    CCS_READ_DEAD = 3,          /* band lost the probability mass (LL = -inf) */
    CCS_READ_UNMAPPED = 4,      /* not placed on the draft / filtered before polishing */
    CCS_READ_POOR_ZSCORE = 5    /* log-likelihood against the draft too far below its expectation (Integrator::AddRead) */
} ccs_read_status;

/* ------------------------------------------------------------------------------------
 * Chemistry model (host only)
 * ---------------------------------------------------------------------------------- */
/* The model is an opaque POD blob of ccs_model_sizeof() bytes (ccs::ArrowModelParams). */
int  ccs_model_sizeof(void);
void ccs_model_synthetic(void* model_out);
/* Arrow model <-> JSON, the chemistry-bundle mechanism ($SMRT_CHEMISTRY_BUNDLE_DIR/arrow/<name>.json,
 * /root/reference/docs/faq/chemistry.md:28-56).  CCS_ERR_CHEMISTRY for an unsupported form / malformed file. */
int  ccs_model_save_json(const void* model, const char* path);
int  ccs_model_load_json(const char* path, void* model_out);

/* The synthetic-data generator (test / bench infrastructure) lives in its own library, libccssim.so:
 * include/ccssim.h. */

/* ------------------------------------------------------------------------------------
 * GPU context
 * ---------------------------------------------------------------------------------- */
typedef struct ccsgpu_ctx ccsgpu_ctx;

/* One context per GPU.  `model` = blob from ccs_model_synthetic()/ccs_model_load_json();
 * device_bytes_budget = 0 -> 90 % of free device memory.  Returns NULL on failure and
 * writes the ccs_error to *err (CCS_ERR_NO_DEVICE when no CUDA device: no CPU fallback). */
ccsgpu_ctx* ccsgpu_create(int device, const void* model, size_t device_bytes_budget, int* err);
void        ccsgpu_destroy(ccsgpu_ctx* ctx);
/* Number of concurrent lanes (engine sets with their own CUDA streams) a stage call spreads its batch
 * over; default 4 (env CCS_B200_LANES).  1 = strictly serial kernels (used for per-kernel timing). */
int         ccsgpu_set_lanes(ccsgpu_ctx* ctx, int n_lanes);
/* Message of the last failure on this ctx (or of the last failed ccsgpu_create if ctx == NULL). */
const char* ccsgpu_last_error(const ccsgpu_ctx* ctx);

/* ------------------------------------------------------------------------------------
 * Stage inputs / outputs (struct-of-arrays, caller-owned host memory)
 * ---------------------------------------------------------------------------------- */
/* A batch of ZMWs as the reader hands them to the stages (docs/img/ccs-impl.png "queue: ZMWs"). */
typedef struct ccs_batch {
    int32_t n_zmws, n_reads;
    const int32_t* zmw_read_off;  /* [n_zmws+1] ZMW -> reads */
    const int64_t* read_off;      /* [n_reads+1] read -> codes */
    const uint8_t* codes;         /* 1 B per base: 4*(min(pw,3)-1) + base(A,C,G,T = 0..3), native orientation
                                     (Recursor::EncodeRead; pw tag docs/faq/bam-output.md:20) */
    const float*   snr;           /* [n_zmws*4] `sn` tag order A,C,G,T (docs/faq/bam-output.md:28) */
    const uint8_t* cx;            /* [n_reads] local-context flags: 1 adapter before, 2 adapter after */
    const int32_t* hole;          /* [n_zmws] `zm` */
} ccs_batch;

/* Draft Stage output = Polish Stage input ("queue: ZMWs, Drafts, Windows"). */
typedef struct ccs_drafts {
    const int64_t* tpl_off;       /* [n_zmws+1] */
    const uint8_t* tpl;           /* draft templates, bases 0..3 */
    const uint8_t* strand;        /* [n_reads] 0 = same strand as the draft, 1 = reverse complement */
    const int32_t* tstart;        /* [n_reads] span of the read on the draft [tstart,tend); */
    const int32_t* tend;          /*           tend <= tstart: read not placed (excluded)   */
    const int32_t* rstart;        /* [n_reads] aligned part of the read [rstart,rend) in its native   */
    const int32_t* rend;          /*           orientation; both NULL: whole reads (ExtractMappedRead) */
} ccs_drafts;

typedef struct ccs_polish_cfg {
    int32_t max_iterations;       /* 40  */
    int32_t separation;           /* 10: minimum distance between mutations applied in one round */
    int32_t neighborhood;         /* 20: re-score +-neighborhood around applied mutations */
    int32_t min_length;           /* --min-length (docs/how-does-ccs-work.md:51) */
    int32_t max_length;           /* --max-length */
    double  min_rq;               /* --min-rq (docs/how-does-ccs-work.md:111-112) */
    double  ab_mismatch_tol;      /* 1e-3: |1 - LL_alpha/LL_beta| above this drops the read */
    double  min_active_fraction;  /* 0.5: fewer usable reads -> TOO_MANY_UNUSABLE */
    double  min_zscore;           /* -3.4: a read whose LL z-score against the draft is lower is dropped (POOR_ZSCORE) */
    /* Windowing (docs/how-does-ccs-work.md:57-61,108-110), ccsgpu_ccs only: a draft of at least 2 * window_size bases is
     * polished as floor(J / window_size) independent windows (cores of window_size bases, the last one running to the
     * end, each padded by window_overlap bases on both sides) whose polished cores are concatenated.  Both are rounded up
     * to multiples of 64; window_size 0 = one window per draft. */
    int32_t window_size;          /* 1024 */
    int32_t window_overlap;       /* 64 */
} ccs_polish_cfg;
void ccs_polish_cfg_default(ccs_polish_cfg* cfg);

typedef struct ccs_results {
    int64_t  seq_cap;             /* capacity of seq / qv in bases */
    int64_t* seq_off;             /* [n_zmws+1] out */
    uint8_t* seq;                 /* consensus bases 0..3 */
    uint8_t* qv;                  /* per-base QV 0..93 */
    float*   rq;                  /* [n_zmws] predicted accuracy = 1 - mean(10^(-QV/10)) */
    int32_t* status;              /* [n_zmws] ccs_zmw_status */
    int32_t* n_passes;            /* [n_zmws] `np`: full-length subreads used */
    int32_t* iterations;          /* [n_zmws] polish rounds run */
    int32_t* n_applied;           /* [n_zmws] mutations applied */
    int64_t* n_tested;            /* [n_zmws] mutations scored */
    double*  read_ll;             /* [n_reads] per-subread Arrow log-likelihood on the final template (NaN: not used) */
    int32_t* read_status;         /* [n_reads] ccs_read_status */
} ccs_results;

typedef struct ccs_draft_cfg {
    double  min_snr;              /* --min-snr    (docs/how-does-ccs-work.md:21) */
    int32_t min_passes;           /* --min-passes (docs/how-does-ccs-work.md:25) */
    int32_t top_passes;           /* --top-passes (docs/faq/accuracy-vs-passes.md:48-52) */
    int32_t max_poa_reads;        /* full-length subreads threaded into the POA ("a few subreads") */
    int32_t min_length;           /* --min-length / --max-length gate on the draft (docs/how-does-ccs-work.md:51) */
    int32_t max_length;
} ccs_draft_cfg;
void ccs_draft_cfg_default(ccs_draft_cfg* cfg);

/* Draft Stage output, caller-owned (mutable twin of ccs_drafts). */
typedef struct ccs_drafts_out {
    int64_t  tpl_cap;             /* capacity of tpl in bases; on CCS_ERR_CAPACITY the needed size is written back */
    int64_t* tpl_off;             /* [n_zmws+1] */
    uint8_t* tpl;
    uint8_t* strand;              /* [n_reads] */
    int32_t* tstart;              /* [n_reads] */
    int32_t* tend;                /* [n_reads] (0,0 when the read was filtered or not placed) */
    int32_t* rstart;              /* [n_reads] */
    int32_t* rend;                /* [n_reads] */
    int32_t* status;              /* [n_zmws] ccs_zmw_status; CCS_ZMW_SUCCESS = draft stage passed */
} ccs_drafts_out;

/* ------------------------------------------------------------------------------------
 * Stages
 * ---------------------------------------------------------------------------------- */
/* Draft Stage: FilterReads -> k-mer orientation votes -> SparsePoa (sequence-to-DAG alignment, traceback and
 * CommitAdd on the device-resident graph) -> FindConsensus -> subread-to-draft mapping, all kernels
 * (docs/how-does-ccs-work.md:19-55). */
int ccsgpu_draft(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_draft_cfg* cfg, ccs_drafts_out* out);

/* The whole per-ZMW hot path: Draft Stage + Polish Stage + final gates; per-ZMW status in the
 * reference's order (docs/faq/reports-aux-files.md:143-159). */
int ccsgpu_ccs(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_draft_cfg* dcfg, const ccs_polish_cfg* pcfg,
               ccs_results* out);

/* Polish Stage: Arrow refinement of every draft with its mapped subreads + per-base QVs
 * (Integrator + Polish + ConsensusQualities; docs/how-does-ccs-work.md:87-112).
 * Returns CCS_ERR_CAPACITY (and the needed size in out->seq_cap) if seq/qv are too small. */
int ccsgpu_polish(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_drafts* drafts, const ccs_polish_cfg* cfg,
                  ccs_results* out);

/* Test / bench hooks on the two hot kernels --------------------------------------------- */
/* Recursor::FillAlphaBeta for n independent (read, template) pairs; pair k uses snr[4k..].
 * Optional dump of pair `dump_pair` (>= 0): alpha/beta cells [J*32] (slot = row mod 32),
 * band starts [J], cumulative scale exponents of alpha (left to right) / beta (right to left). */
int ccsgpu_fill_alpha_beta(ccsgpu_ctx* ctx, int32_t n_pairs, const int64_t* tpl_off, const uint8_t* tpl,
                           const int64_t* read_off, const uint8_t* codes, const float* snr, double* ll_alpha,
                           double* ll_beta, int32_t* status, int32_t dump_pair, float* alpha_out, float* beta_out,
                           int32_t* start_out, int32_t* aexp_out, int32_t* bexp_out);
/* Integrator::LL(Mutation) - LL() for every slot of every draft position, no refinement:
 * delta[(tpl_off[z] + p) * 9 + slot], slots {SUB A,C,G,T, DEL, INS A,C,G,T}; read_ll[n_reads]. */
int ccsgpu_score_all(ccsgpu_ctx* ctx, const ccs_batch* in, const ccs_drafts* drafts, double* delta, double* read_ll,
                     int32_t* read_status);

/* Device-time accounting of the ctx since the last reset (CUDA events on the ctx stream). */
typedef struct ccs_stats {
    double  ms_fill_alpha, ms_fill_beta, ms_score, ms_pick, ms_qv, ms_h2d, ms_draft;
    int64_t launches_fill_alpha, launches_fill_beta, launches_score, launches_pick, launches_qv, launches_draft;
    int64_t bytes_fill_alpha, bytes_fill_beta;   /* algorithmic bytes (DESIGN.md "Roofline") */
    int64_t cells_fill, score_items, rounds;
    int64_t h2d_bytes, d2h_bytes;
    double  ms_poa_align;  /* poa_align + traceback kernels */
    int64_t launches_poa, poa_tasks, poa_rows, bytes_poa_align;
    double  ms_resident;   /* CUDA-event time of the stage with inputs already in HBM */
    double  ms_e2e;        /* host wall time of the stage calls: pack + H2D + kernels + D2H */
    /* with more than one lane the per-kernel ms above are sums over concurrently running lanes */
    int64_t n_zmws;        /* ZMWs processed */
    int64_t top_fill_alpha_bytes;   /* the largest single arrow_fill_alpha launch: algorithmic bytes */
    double  top_fill_alpha_ms;      /* and its duration (roofline numerator/denominator) */
    double  ms_poa_map;    /* poa_align + traceback launches of the subread -> draft mapping (linear templates) */
    double  ms_poa_graph;  /* graph kernels: seed chain, CommitAdd, FindConsensus, k-mer votes */
    int64_t launches_poa_graph, bytes_poa_map;
    int64_t bytes_score;   /* algorithmic bytes of arrow_score: 704 B per (covering read, position) (SURVEY.md 8d B_score) */
    int64_t launches_pack; /* arrow_pack_rowcodes launches */
    /* the largest single launch of each heavy kernel: algorithmic bytes and CUDA-event duration */
    int64_t top_fill_beta_bytes;  double top_fill_beta_ms;
    int64_t top_score_bytes;      double top_score_ms;
    int64_t top_poa_align_bytes;  double top_poa_align_ms;
    int64_t top_poa_map_bytes;    double top_poa_map_ms;
} ccs_stats;
int ccsgpu_get_stats(ccsgpu_ctx* ctx, ccs_stats* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* CCSGPU_H */
