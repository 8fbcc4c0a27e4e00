/*
 * asmgpu.h -- C ABI of libasmgpu.so, the B200 (sm_100a) implementation of ASM's
 * convergence-diagnostic hot path: TreePSRF's all-pairs tree distances with their within- and
 * between-chain reductions, and TraceESS's batched autocorrelation / effective sample size.
 *
 * This is the drop-in boundary.  The reference (rbouckaert/asm) is pure Java and has no FFI;
 * these are the entry points a Panama-FFM / JNI binding inside package asm.inference would
 * bind (see INTEGRATION.md for the stub), one per arithmetic site of the reference.  Citations
 * are file:line relative to the reference checkout.
 *
 * Conventions
 *   - plain C: pointers and sizes only; every function returns int32_t, 0 = ASM_OK, < 0 = ASM_E_*;
 *     nothing throws or aborts.  asm_last_error(ctx) gives the message of the last failure.
 *   - the caller owns every host buffer and may free it on return (the library copies).
 *   - one asm_ctx = one GPU (asm_create) or several GPUs of one box (asm_create_multi).  A context may be
 *     used from any thread, one call at a time (the reference calls criteria from the single
 *     LogWatcherThread, AutoStopMCMC.java:214, 397-409); the library selects its device on every call.
 *   - there is no CPU fallback: without a usable sm_100 device asm_create fails with
 *     ASM_E_NO_DEVICE and every compute entry point fails with it as well.
 *
 * Tree encoding: a tree is the set of its clades, a clade is the set of
 * leaf numbers under an internal node (CladeDifference.java:107-141, root included).  A global
 * dictionary numbers distinct clades 0..D-1 in first-seen order; a tree is a row of D bits,
 * stored as little-endian uint64 words, bit j of word w <-> clade id 64*w + j.  Robinson-Foulds
 * distance = popcount(row_a XOR row_b) = |a| + |b| - 2 a.b.
 */
#ifndef ASMGPU_H_
#define ASMGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASM_OK 0
#define ASM_E_INVALID (-1)    /* bad argument (IllegalArgumentException sites of the reference) */
#define ASM_E_NO_DEVICE (-2)  /* no CUDA device / not sm_100 / extension not usable          */
#define ASM_E_CUDA (-3)       /* a CUDA call or kernel failed; see asm_last_error            */
#define ASM_E_RANGE (-4)      /* index past the stored trees / samples (the Java would throw) */
#define ASM_E_NOMEM (-5)
#define ASM_E_UNSUPPORTED (-6)

/* distance-kernel selector ("method") */
#define ASM_RF_POPC 0 /* XOR + POPC tiles on the integer pipe, bulk-async (TMA) smem staging  */
#define ASM_RF_I8 1   /* int8 tcgen05.mma (UTCIMMA) contraction, TMA operands, TMEM accumulators */
#define ASM_RF_FP4 2  /* the same contraction on the block-scaled FP4 path (e2m1 indicators, unit scales, f32
                         accumulators: exact), twice the K per instruction; PSRF sums only (no asm_rf_block) */

#define ASM_MAX_CHAINS 64
#define ASM_MAX_LAG 2000 /* TraceESS.java:21 */

typedef struct asm_ctx asm_ctx;

/* ---- lifetime ------------------------------------------------------------------------------- */

const char* asm_version(void);
/* Number of CUDA devices visible, 0 if none (never fails). */
int32_t asm_device_count(void);
/* Create a context on CUDA device `device` for `n_chains` chains (TraceInfo(chainCount),
 * TraceInfo.java:31-41; criteria setup(nChains, traceInfo), MCMCConvergenceCriterion.java:20). */
int32_t asm_create(asm_ctx** out, int32_t device, int32_t n_chains);
int32_t asm_destroy(asm_ctx* ctx);
/* Host placement for the upload paths: bind the CALLING thread to the CPUs that are local to CUDA device `device`
 * (the local_cpulist of its PCI function in sysfs), so that buffers the thread allocates, fills and pins afterwards
 * sit on the GPU's own NUMA node -- with 8 GPUs on two sockets the 8 GB trace uploads of asm_ess_batch otherwise
 * cross the socket interconnect.  Returns the number of CPUs bound to, 0 where there is nothing to do (one NUMA node,
 * no sysfs entry, a virtual machine that hides the topology), < 0 on error.  The worker threads of asm_create_multi
 * call it for their own device; a one-process-per-GPU launcher calls it once per rank before it allocates
 * (AutoStopMCMC's chain threads, AutoStopMCMC.java:188-214, would call it where they create their log buffers). */
int32_t asm_bind_thread_near_device(int32_t device);

/* ---- more than one GPU ---------------------------------------------------------------------------------------
 * The reference calls its criteria from ONE thread of ONE process (LogWatcherThread, AutoStopMCMC.java:214,
 * 397-409), so the multi-GPU form is one context over several devices: asm_create_multi builds one per-device
 * context per entry of `devices` inside the calling process, joins them in an NCCL communicator
 * (ncclCommInitAll) and gives each a host thread.  EVERY entry point below takes such a context with unchanged
 * signature and meaning:
 *   - trees: asm_trees_set sends each row over PCIe once (device r uploads the r-th slice of the chain) and one
 *     in-place ncclAllGather over NVLink replicates it; asm_trees_append / asm_traces_append go to every device;
 *   - PSRF: tile pairs t of the symmetric pair matrix go to device t % N, the int64 partial tables are combined with
 *     ONE ncclAllReduce on the library's streams (exact in any order: sums of squares of small integers), the
 *     finish runs on the first device; results are bit-identical to one GPU;
 *   - ESS: traces (asm_ess_batch, asm_ess_traces_put) and selected columns (asm_ess_eval) are dealt to the devices
 *     in contiguous blocks; every device writes its slice of the caller's output arrays;
 *   - clade counts: every device counts its slice of [from, to), one ncclAllReduce(int32) adds the slices;
 *   - device pointers handed in (asm_trees_set_device, asm_psrf_sums_device, asm_ess_batch_device, ...) are on the
 *     FIRST device of the list.
 * NCCL (libnccl.so.2) is loaded on first use; one GPU needs none.  n_devices == 1 is allowed. */
int32_t asm_create_multi(asm_ctx** out, const int32_t* devices, int32_t n_devices, int32_t n_chains);
/* The same collectives for a launcher that runs one PROCESS per GPU (torchrun): rank 0 draws an id, the launcher's
 * own channel carries its ASM_COMM_ID_BYTES bytes to the other ranks, every rank joins with its plain asm_create
 * context.  From then on asm_trees_set, asm_psrf_sums*, asm_psrf_eval and asm_clade_counts on that context are
 * collective calls (every rank calls them with the same arguments and receives the complete result); ESS calls stay
 * local to the traces the rank was given and asm_comm_all_gather_f64 combines the per-rank result vectors. */
#define ASM_COMM_ID_BYTES 128
int32_t asm_comm_unique_id(uint8_t* id_out /* [ASM_COMM_ID_BYTES] */);
int32_t asm_comm_init_rank(asm_ctx* ctx, const uint8_t* id, int32_t rank, int32_t n_ranks);
/* rank / number of ranks of the context's communicator (0 / 1 without one; a multi-GPU context reports rank 0 of
 * n_devices) and the number of devices this context drives itself. */
int32_t asm_comm_info(const asm_ctx* ctx, int32_t* rank, int32_t* n_ranks, int32_t* n_local_devices);
/* NCCL_VERSION_CODE of the NCCL the library loaded, 0 if none could be loaded. */
int32_t asm_comm_nccl_version(void);
/* all[r * n_local + i] = rank r's local[i] (host buffers; the same n_local on every rank). */
int32_t asm_comm_all_gather_f64(asm_ctx* ctx, const double* local, int64_t n_local, double* all);
/* Message of the last failure on this context ("" if none).  ctx may be NULL: then the message
 * of the last failed asm_create on this thread. */
const char* asm_last_error(const asm_ctx* ctx);

/* ---- trees: replaces traceInfo.trees[chain] (TraceInfo.java:27, AutoStopMCMC.java:536) ------- */

/* Replace chain `chain`'s trees by n_trees bit rows, row-major [n_trees][words_per_tree]. */
int32_t asm_trees_set(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t words_per_tree, const uint64_t* bits);
/* Append n_new rows (the streaming runner adds one tree per chain per check).  words_per_tree
 * may exceed the stored width when the dictionary grew: older rows are zero-extended. */
int32_t asm_trees_append(asm_ctx* ctx, int32_t chain, int64_t n_new, int32_t words_per_tree, const uint64_t* bits);
/* The same two calls for callers that hold what the encoder emits -- the clade ids of every tree (CladeDifference's
 * clade identity, CladeDifference.java:107-141; asm_encoder_add_*) -- instead of finished bit rows: `ids` is row-major
 * [n_trees][ids_per_tree] uint16 dictionary ids, ASM_NO_CLADE in unused slots (a multifurcating tree has fewer clades);
 * n_clades is the dictionary size the ids were drawn from (sets the row width, at most 65,535).  The rows are built
 * on the GPU: a 100-taxon tree is 198 bytes of ids against 352 bytes of bits over PCIe.  The call returns when `ids`
 * has been read; an id that is neither ASM_NO_CLADE nor below n_clades is reported (ASM_E_RANGE) by the next call that
 * looks at the trees (asm_psrf_*, asm_rf_block, asm_clade_counts), which then fails.  On a multi-GPU context every
 * GPU takes the ids from the host itself (right for the runner's one tree per chain per check; for bulk loads over
 * several GPUs asm_trees_set sends every row over PCIe only once). */
#define ASM_NO_CLADE 0xFFFFu
int32_t asm_trees_set_ids(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t ids_per_tree, const uint16_t* ids,
                          int32_t n_clades);
int32_t asm_trees_append_ids(asm_ctx* ctx, int32_t chain, int64_t n_new, int32_t ids_per_tree, const uint16_t* ids,
                             int32_t n_clades);
int32_t asm_trees_count(const asm_ctx* ctx, int32_t chain, int64_t* n_trees, int32_t* words_per_tree);
/* Same as asm_trees_set, but `d_bits` is a device pointer on the context's GPU (used by the
 * benchmark's "inputs already resident in HBM" leg and by multi-GPU replicas). */
int32_t asm_trees_set_device(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t words_per_tree,
                             const uint64_t* d_bits);

/* ---- parity probes ---------------------------------------------------------------------------- */

/* out[(r-r0)*(c1-c0) + (c-c0)] = RF(tree[chainA][r], tree[chainB][c]) for r in [r0,r1),
 * c in [c0,c1): the value `m.distance(tree1, tree2)` supplies at TreeESS.java:179-180 (RF in
 * place of RNNIMetric).  `out` is a host buffer. */
int32_t asm_rf_block(asm_ctx* ctx, int32_t chainA, int64_t r0, int64_t r1, int32_t chainB, int64_t c0, int64_t c1,
                     int32_t method, uint16_t* out);
/* out[id] = number of trees in [from,to) of `chain` that contain clade id, id < 64*words:
 * the per-chain clade counts of CladeDifference.process (CladeDifference.java:147-159). */
int32_t asm_clade_counts(asm_ctx* ctx, int32_t chain, int64_t from, int64_t to, int32_t* out);

/* ---- PSRF: replaces calcPSRF + the row loops of converged2sided ------------------------------- */

/* Work split for multi-GPU runs: this context evaluates tiles t with t % shard_count ==
 * shard_index of the (symmetric) pair matrix; the partial S tables of all shards add up (exact
 * int64) to the full table.  {0, 1} = everything. */
typedef struct {
  int32_t shard_index;
  int32_t shard_count;
} asm_shard;

/* S[((a*n_rows) + r)*C + c] = sum over columns i of chain c, i = start(burnin[c]), +delta, ...
 * < end (burn-in rounded up to a multiple of delta, TreePSRF.java:296-301), of
 * distancePlusOne(a, x_r, c, i)^2 with x_r = row_start + r*delta, n_rows = ceil((end -
 * row_start)/delta): varIn (c == a) and varBetween (c != a) of calcPSRF (TreePSRF.java:275-286)
 * for every chain pair.  Exact integers.  S is a host buffer [C][n_rows][C]. */
int32_t asm_psrf_sums(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                      int32_t method, asm_shard shard, int64_t* S);
/* Same, S left on the device (d_S is a device pointer to [C][n_rows][C] int64) so that partial
 * tables can be combined with an NCCL all-reduce without a host round trip. */
int32_t asm_psrf_sums_device(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                             int32_t method, asm_shard shard, int64_t* d_S);
/* The same without the final host synchronisation: the work is enqueued on the context's stream
 * (asm_stream_handle), on which the caller orders its collective; asm_psrf_finish_device, enqueued on the same
 * stream, then follows the collective with no host round trip in between. */
int32_t asm_psrf_sums_device_async(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                                   int32_t method, asm_shard shard, int64_t* d_S);
/* The cudaStream_t (as void*) the context enqueues its PSRF work on. */
void* asm_stream_handle(asm_ctx* ctx);
/* From a complete S table on the device: psrf_rows[((a*n_rows)+r)*C + b] = S[a,r][a] != 0 ?
 * sqrt(S[a,r][b] / S[a,r][a]) : sqrt(S[a,r][b]) (TreePSRF.java:287-292) and psrf_mean[a*C + b] =
 * (sequential sum over r ascending) / n_rows (TreePSRF.java:264-271).  psrf_rows may be NULL.
 * Both outputs are host buffers. */
int32_t asm_psrf_finish_device(asm_ctx* ctx, const int64_t* d_S, int32_t n_rows, double* psrf_rows,
                               double* psrf_mean);
/* sum_out = values[0] + values[1] + ... + values[n-1] added left to right, every add rounded as the scalar loop
 * `for (i) s += values[i]` rounds it -- the summation TreePSRF.mean does (TreePSRF.java:264-271) and that
 * asm_psrf_finish_device applies to the row values -- evaluated in parallel on the device (the
 * roundings are replayed as an integer scan while the running sum stays inside one binade).  Any doubles are
 * accepted; negative and non-finite values take plain adds.  Host buffers. */
int32_t asm_seq_sum(asm_ctx* ctx, const double* values, int64_t n, double* sum_out);
/* asm_psrf_sums + asm_psrf_finish in one call on one GPU: everything converged2sided
 * (TreePSRF.java:141-158) needs from the device.  psrf_mean[0*C+1] is psrf1mean and
 * psrf_mean[1*C+0] is psrf2mean of the two-chain reference. */
int32_t asm_psrf_eval(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                      int32_t method, double* psrf_rows, double* psrf_mean);

/* ---- ESS: replaces TraceESS.ACT / calcESS ------------------------------------------------------ */

/* For each of n_traces independent traces (row t at traces + t*stride, n_samples doubles):
 * act_out[t] = ACT(trace, 0, n_samples) (TraceESS.java:130-179, max_lag = MAX_LAG = 2000) and
 * ess_out[t] = n_samples / act_out[t] (TraceESS.java:126-128).  Either output may be NULL.
 * All pointers are host pointers. */
int32_t asm_ess_batch(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces,
                      int32_t max_lag, double* act_out, double* ess_out);
/* Same with `d_traces` on the device; outputs are host buffers. */
int32_t asm_ess_batch_device(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride,
                             const double* d_traces, int32_t max_lag, double* act_out, double* ess_out);
/* Resident form of the batch (the benchmark's "inputs already in HBM" leg; a multi-GPU context deals the traces
 * to its devices here): asm_ess_traces_put uploads and keeps [n_traces][n_samples], asm_ess_traces_eval evaluates
 * what is resident, outputs as asm_ess_batch. */
int32_t asm_ess_traces_put(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces);
int32_t asm_ess_traces_count(const asm_ctx* ctx, int32_t* n_traces, int64_t* n_samples);
int32_t asm_ess_traces_eval(asm_ctx* ctx, int32_t max_lag, double* act_out, double* ess_out);
/* Streaming form.  Append n_new log lines of n_cols columns for `chain`, row-major
 * [n_new][n_cols] (readLogLines, AutoStopMCMC.java:459-466). */
int32_t asm_traces_append(asm_ctx* ctx, int32_t chain, int32_t n_cols, int64_t n_new, const double* values);
int32_t asm_traces_count(const asm_ctx* ctx, int32_t chain, int64_t* n_samples, int32_t* n_cols);
/* ess_out[j] = ESS of the concatenation over chains i of column cols[j], samples
 * [burnin[i], end) (TraceESS.java:53-62).  The min with Java NaN semantics (:63) and the
 * threshold (:68) stay with the caller. */
int32_t asm_ess_eval(asm_ctx* ctx, const int32_t* burnin, int32_t end, const int32_t* cols, int32_t n_cols,
                     double* ess_out);

/* Per chain and column, over samples [0, end): sums_out[chain*n_cols + col] = sequential sum of d and
 * sumsq_out[...] = sequential sum of d*d -- the accumulation loops of GelmanRubin.calcGRStat
 * (GelmanRubin.java:61-73), in their Java order.  Host buffers of n_chains*n_cols doubles. */
int32_t asm_trace_moments(asm_ctx* ctx, int32_t end, double* sums_out, double* sumsq_out);

/* ---- measurement hooks -------------------------------------------------------------------------- */

/* CUDA-event timing on the context's own stream (torch.cuda.Event only sees torch's stream).
 * asm_timer_mark records event `slot` (0..15); asm_timer_elapsed_ms synchronises on `to`. */
int32_t asm_timer_mark(asm_ctx* ctx, int32_t slot);
int32_t asm_timer_elapsed_ms(asm_ctx* ctx, int32_t from, int32_t to, float* ms);
/* Accumulated device time and launch count of the library's kernels since the last reset,
 * per kernel family (ASM_K_*), measured with events around every launch when profiling is on. */
#define ASM_K_GATHER 0
#define ASM_K_RF_POPC 1
#define ASM_K_RF_I8 2 /* both tensor-core kernels (ASM_RF_I8 and ASM_RF_FP4) */
#define ASM_K_PSRF_FINISH 3
#define ASM_K_CLADE_COUNTS 4
#define ASM_K_ESS_LAG 5
#define ASM_K_ESS_FINISH 6
#define ASM_K_MISC 7
#define ASM_K_COUNT 8
int32_t asm_profile_enable(asm_ctx* ctx, int32_t on);
int32_t asm_profile_reset(asm_ctx* ctx);
int32_t asm_profile_get(asm_ctx* ctx, int32_t family, double* total_ms, int64_t* launches);
/* Total kernel launches issued by this context since creation (always counted). */
int64_t asm_launch_count(const asm_ctx* ctx);
int32_t asm_sync(asm_ctx* ctx);
/* Write `bytes` of a device scratch buffer (> L2) to evict the L2 between timed iterations. */
int32_t asm_flush_l2(asm_ctx* ctx);
/* Number of (trace, lag) chains the last ESS call evaluated: n_traces * 2048 when every lag is computed, fewer
 * when the staged evaluation retired traces early (the integral loop of TraceESS.java:161-172 had stopped). */
int64_t asm_ess_last_lag_count(const asm_ctx* ctx);
/* on = 1 (default): staged lag evaluation; on = 0: every lag of every trace.  The results are bit-identical. */
int32_t asm_ess_set_adaptive(asm_ctx* ctx, int32_t on);
/* Layout facts of the last PSRF evaluation: padded row count, dictionary width in bits (padded),
 * tiles evaluated by this shard, tiles in the full problem. */
int32_t asm_psrf_last_layout(const asm_ctx* ctx, int64_t* padded_rows, int64_t* padded_bits, int64_t* tiles_done,
                             int64_t* tiles_total);

/* ---- host encoder (the step before the path): trees -> clades -> dictionary -> bit rows -------- */
/* CPU only; usable without a GPU.  Follows CladeDifference.traverse (CladeDifference.java:107-144)
 * for clade identity and the NEXUS/Newick conventions of readTreeLogLine (AutoStopMCMC.java:486-538). */

typedef struct asm_encoder asm_encoder;

int32_t asm_encoder_create(asm_encoder** out, int32_t n_taxa);
int32_t asm_encoder_destroy(asm_encoder* enc);
/* Trees as child arrays: left/right are [n_trees][2*n_taxa-1] (-1 for the leaves 0..n_taxa-1),
 * root is [n_trees].  clade_ids_out is [n_trees][n_taxa-1]: dictionary ids in post-order (children
 * before parent, left before right), new clades numbered in first-seen order. */
int32_t asm_encoder_add_topologies(asm_encoder* enc, int64_t n_trees, const int32_t* left, const int32_t* right,
                                   const int32_t* root, int32_t* clade_ids_out);
/* One Newick string whose leaf labels are integers (the NEXUS translate index); label_offset is
 * subtracted from each label (parser.offsetInput = 1, AutoStopMCMC.java:527).  Branch lengths and
 * [&...] annotations are skipped. */
int32_t asm_encoder_add_newick(asm_encoder* enc, const char* newick, int32_t label_offset, int32_t* clade_ids_out);
/* n_trees Newick strings at once: parsing and the clade walks use every host core (ASM_ENCODER_THREADS), the ids are
 * the ones one-by-one calls in the same order would hand out.  asm_encoder_add_topologies is parallel in the same
 * way (two phases per batch: look-ups against the frozen dictionary in parallel, insertions in first-seen order). */
int32_t asm_encoder_add_newick_batch(asm_encoder* enc, int64_t n_trees, const char* const* newicks, int32_t label_offset,
                                     int32_t* clade_ids_out);
int64_t asm_encoder_num_clades(const asm_encoder* enc);
/* Sorted leaf numbers of clade `id`; returns the clade size or ASM_E_*. */
int32_t asm_encoder_clade_leaves(const asm_encoder* enc, int64_t id, int32_t* leaves_out, int32_t cap);
/* Words per bit row for the current dictionary: ceil(D/64) rounded up to a multiple of 4 (256 bits). */
int32_t asm_encoder_words(const asm_encoder* enc);
/* bits_out[n_trees][words_per_tree] from clade_ids[n_trees][ids_per_tree]. */
int32_t asm_encode_bits(const int32_t* clade_ids, int64_t n_trees, int32_t ids_per_tree, int32_t words_per_tree,
                        uint64_t* bits_out);

#ifdef __cplusplus
}
#endif
#endif /* ASMGPU_H_ */
