/*
 * asmgpu_host.h -- flat C wrapper of include/asm_criteria.hpp (the C++ mirror of the reference's
 * MCMCConvergenceCriterion / TraceInfo / TreePSRF / TreeESS / TraceESS), so that tests and tools in any
 * language can drive the criteria exactly as AutoStopMCMC does:
 *     asmh_trace_info_create -> asmh_read_log_line / asmh_read_tree_log_line (one line per chain per check)
 *     asmh_*_create (inputs by the reference's names) -> asmh_criterion_setup -> asmh_criterion_converged
 * Return codes as in asmgpu.h; asmh_last_error() gives the message of the last failure on this thread.
 * IllegalArgumentException sites of the reference (TreePSRF.java:56-84,311-313; TraceInfo.java:82-85)
 * return ASM_E_INVALID.
 */
#ifndef ASMGPU_HOST_H_
#define ASMGPU_HOST_H_

#include "asmgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct asmh_trace_info asmh_trace_info; /* asmgpu::TraceInfo  (TraceInfo.java) */
typedef struct asmh_criterion asmh_criterion;   /* asmgpu::MCMCConvergenceCriterion (MCMCConvergenceCriterion.java) */

const char* asmh_last_error(void);

/* TraceInfo(chainCount) -- TraceInfo.java:31-41; owns one asm_ctx on `device` and the clade dictionary */
int32_t asmh_trace_info_create(asmh_trace_info** out, int32_t chain_count, int32_t device);
int32_t asmh_trace_info_destroy(asmh_trace_info* ti);
/* AutoStopMCMC.readLogLines (AutoStopMCMC.java:430-479) / readTreeLogLine (:486-538), one line at a time */
int32_t asmh_read_log_line(asmh_trace_info* ti, int32_t chain, const char* line, int32_t* appended);
int32_t asmh_read_tree_log_line(asmh_trace_info* ti, int32_t chain, const char* line, int32_t* appended);
/* One cell of a trace log as readLogLines converts it (AutoStopMCMC.java:460-466): Double.parseDouble's grammar, 0.0
 * where Java would throw NumberFormatException.  *is_number (may be NULL) says which of the two happened.  No GPU. */
double asmh_parse_log_cell(const char* token, int32_t* is_number);
/* Trees and log lines are staged on the host and uploaded, one copy per chain, when a criterion next asks for the
 * context (asmh_ctx, asmh_criterion_converged); asmh_flush does it now. */
int32_t asmh_flush(asmh_trace_info* ti);
/* direct feeding, for callers that hold parsed data */
int32_t asmh_set_taxon_count(asmh_trace_info* ti, int32_t n_taxa);
int32_t asmh_add_tree(asmh_trace_info* ti, int32_t chain, const int32_t* left, const int32_t* right, int32_t root);
int32_t asmh_add_log_line(asmh_trace_info* ti, int32_t chain, const double* values, int32_t n);
int32_t asmh_set_labels(asmh_trace_info* ti, const char* whitespace_separated); /* TraceInfo.setLabels :125-127 */
int32_t asmh_get_map(asmh_trace_info* ti, int32_t* map_out, int32_t cap, int32_t* n_out); /* getMap :52-68 */
int64_t asmh_tree_count(const asmh_trace_info* ti, int32_t chain);
int64_t asmh_log_count(const asmh_trace_info* ti, int32_t chain);
int64_t asmh_clade_count(const asmh_trace_info* ti);
asm_ctx* asmh_ctx(asmh_trace_info* ti);

/* criteria; inputs by the reference's names and in its units; initAndValidate runs inside create */
int32_t asmh_tree_psrf_create(asmh_criterion** out, double b, int32_t two_sided, int32_t check_ess, int32_t target_ess,
                              double smoothing, int32_t cache_limit, int32_t sample_size, int32_t method);
int32_t asmh_tree_ess_create(asmh_criterion** out, int32_t target_ess, double smoothing, int32_t cache_limit,
                             int32_t sample_size, int32_t method);
int32_t asmh_trace_ess_create(asmh_criterion** out, int32_t target_ess, const char* traces);
/* GelmanRubin (GelmanRubin.java:14) and CladeDifference (CladeDifference.java:16): input "threshold" */
int32_t asmh_gelman_rubin_create(asmh_criterion** out, double threshold);
int32_t asmh_clade_difference_create(asmh_criterion** out, double threshold);
int32_t asmh_criterion_last_value(asmh_criterion* c, double* out);
int32_t asmh_criterion_destroy(asmh_criterion* c);
/* setup(nChains, traceInfo) -- MCMCConvergenceCriterion.java:20 */
int32_t asmh_criterion_setup(asmh_criterion* c, int32_t n_chains, asmh_trace_info* ti);
/* converged(burnin, end) -- MCMCConvergenceCriterion.java:14 */
int32_t asmh_criterion_converged(asmh_criterion* c, const int32_t* burnin, int32_t end, int32_t* converged);
const char* asmh_criterion_header(asmh_criterion* c); /* header() :26 */
/* TreePSRF: delta (TreeESS.java:40), start0 (TreePSRF.java:28), grValues (:43-51) */
int32_t asmh_tree_psrf_state(asmh_criterion* c, int32_t* delta, int32_t* start0, int32_t* threw, double* gr_values,
                             int32_t cap, int32_t* n_gr);
/* TraceESS: currentESSs (TraceESS.java:24,62) and minESS (:63) of the last call */
int32_t asmh_trace_ess_current(asmh_criterion* c, double* ess_out, int32_t cap, int32_t* n_out, double* min_out);
const char* asmh_criterion_last_log(asmh_criterion* c);

/* BurnInDetector.burnIn(end) (BurnInDetector.java:21-36), the supplier of burnin[] for every converged() call
 * (AutoStopMCMC.java:397-409).  strategy 0 = Automatic (BurnInStrategy.java:5), 1 = Running with `percent` (0..99,
 * default 10); burnin_out has one entry per chain.  Host code, as in the reference. */
int32_t asmh_burn_in(asmh_trace_info* ti, int32_t strategy, int32_t percent, int32_t end, int32_t* burnin_out);

/* AutoStopMCMC.combinelogs (AutoStopMCMC.java:247-317) for one logger.  mode 0 = trace log (:287-312), 1 = tree
 * log (:260-285).  Writes `header` (may be NULL; stands for BEAST's Logger.init(), :254), then the samples
 * [burnin[i], last_reported) of every chain file renumbered 0, sample_delta, 2*sample_delta, ... (sample_delta =
 * (long)(logLines[0][0].get(1) - logLines[0][0].get(0)), :253), then `footer` (may be NULL), to out_path.
 * A chain file that ends early is the reference's NullPointerException: ASM_E_RANGE, out_path is not touched. */
int32_t asmh_combine_logs(int32_t mode, int32_t n_chains, const char* const* chain_paths, const int32_t* burnin,
                          int32_t last_reported, int64_t sample_delta, const char* header, const char* footer,
                          const char* out_path, int64_t* lines_out);

#ifdef __cplusplus
}
#endif
#endif
