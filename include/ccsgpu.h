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
    CCS_READ_UNMAPPED = 4       /* not placed on the draft / filtered before polishing */
} ccs_read_status;

/* ------------------------------------------------------------------------------------
 * Chemistry model and synthetic data (host only)
 * ---------------------------------------------------------------------------------- */
/* The model is an opaque POD blob of ccs_model_sizeof() bytes (ccs::ArrowModelParams). */
int  ccs_model_sizeof(void);
void ccs_model_synthetic(void* model_out);

typedef struct ccs_sim_config {
    int32_t insert_mean, insert_sd;
    int32_t passes_min, passes_max;
    int32_t partials;
    double  snr_mean[4];
    double  snr_sd;
    double  frac_low_snr;
    double  frac_few_passes;
    uint64_t seed;
} ccs_sim_config;

/* BASELINE.json configs 1..5 as simulator settings. */
void ccs_sim_get_config(int config_id, ccs_sim_config* out);
/* ZMW `index` of a config: truth template (bases 0..3), reads as emission codes
 * 4*(pw-1)+base in native orientation, cx flags, truth strand / span. */
int  ccs_sim_zmw(const void* model, const ccs_sim_config* cfg, int64_t index, float* snr, uint8_t* tpl,
                 int32_t tpl_cap, int32_t* tpl_len, uint8_t* codes, int64_t codes_cap, int32_t max_reads,
                 int32_t* n_reads, int64_t* read_off, uint8_t* cx, uint8_t* strand, int32_t* tstart, int32_t* tend);
/* Draft-like corruption of a template; map[j] = position of truth base j in `out` (len+1 entries). */
int  ccs_sim_corrupt(const uint8_t* tpl, int32_t len, double rate, uint64_t seed, uint8_t* out, int32_t out_cap,
                     int32_t* out_len, int32_t* map);

#ifdef __cplusplus
}
#endif
#endif /* CCSGPU_H */
