/* ccssim.h -- C ABI of libccssim.so: the synthetic Sequel-II-shape ZMW generator used by the tests, bench.py and
 * the demo inputs of the `ccs` command line.  Test / bench infrastructure, NOT part of the product library
 * (libccsgpu.so, include/ccsgpu.h): the reference arm of bench.py loads this library and the CPU oracle only.
 * Reads are sampled from the Arrow HMM itself, so the model is well specified. */
#ifndef CCSSIM_H
#define CCSSIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* The model is an opaque POD blob of ccs_model_sizeof() bytes (same as include/ccsgpu.h). */
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

/* Multi-threaded generation of ZMWs [first_index, first_index+n) into an opaque handle;
 * draft_error_rate >= 0 also makes a corrupted draft per ZMW (polish-only benches/tests).
 * sizes[4] = {n_reads, total_codes, total_truth_bases, total_draft_bases}. */
void* ccs_sim_batch_create(const void* model, const ccs_sim_config* cfg, int64_t first_index, int32_t n_zmws,
                           double draft_error_rate, int32_t n_threads);
void  ccs_sim_batch_free(void* handle);
void  ccs_sim_batch_sizes(const void* handle, int64_t* sizes);
void  ccs_sim_batch_copy(const void* handle, int32_t* zmw_read_off, int64_t* read_off, uint8_t* codes, float* snr,
                         uint8_t* cx, int32_t* hole, int64_t* truth_off, uint8_t* truth, uint8_t* strand,
                         int32_t* tstart, int32_t* tend, int64_t* draft_off, uint8_t* draft, int32_t* dstart,
                         int32_t* dend);

/* Writes ZMWs [first_index, first_index+n) of a config as a PacBio-style subreads.bam (hole = index+1;
 * tags zm qs qe cx sn pw RG) -- input for the `ccs` command line (ccs_b200/bin/ccs). */
int   ccs_sim_write_subreads_bam(const char* path, const char* movie, const void* model, const ccs_sim_config* cfg,
                                 int64_t first_index, int32_t n_zmws, int32_t with_chemistry);

#ifdef __cplusplus
}
#endif
#endif /* CCSSIM_H */
