/*
 * multi_gpu_check.c -- a plain C caller of include/asmgpu.h that drives several GPUs from ONE process, the way the
 * reference's single LogWatcherThread would through a Panama/JNI binding (AutoStopMCMC.java:397-409): no Python, no
 * torchrun.  It evaluates the same synthetic chains on one GPU and on N GPUs (asm_create_multi) and requires
 *   - the int64 table S of asm_psrf_sums to be IDENTICAL ("1 vs 2 vs 4 vs 8 must be identical int64"),
 *   - the PSRF means of asm_psrf_eval to be bit-identical,
 *   - the clade counts (sharded + ncclAllReduce int32) and the ESS values (traces dealt to the devices) to be identical.
 *
 * usage: multi_gpu_check N_GPUS [n_chains n_trees n_taxa [method [beta [n_traces n_samples]]]]
 *        method: 0 popcount, 1 int8 tcgen05, 2 fp4 tcgen05 (default)
 *        "multi_gpu_check 8 32 50000 500" is BASELINE.json's cfg5.
 * Prints one JSON line; exit code 0 = identical.
 * Inputs come from synthgen/libasmsynth.so (test infrastructure, not the product).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "asmgpu.h"
#include "asmsynth.h"

#define CHECK(call)                                                                          \
  do {                                                                                       \
    int32_t rc_ = (call);                                                                    \
    if (rc_ != ASM_OK) {                                                                     \
      fprintf(stderr, "%s:%d %s -> %d (%s | %s)\n", __FILE__, __LINE__, #call, (int)rc_,     \
              asm_last_error(ctx1), asm_last_error(ctxN));                                   \
      return 2;                                                                              \
    }                                                                                        \
  } while (0)

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char** argv) {
  asm_ctx *ctx1 = NULL, *ctxN = NULL;
  if (argc < 2) {
    fprintf(stderr, "usage: %s N_GPUS [n_chains n_trees n_taxa [method [beta [n_traces n_samples]]]]\n", argv[0]);
    return 2;
  }
  const int n_gpus = atoi(argv[1]);
  const int C = argc > 2 ? atoi(argv[2]) : 4;
  const int64_t n_trees = argc > 3 ? atoll(argv[3]) : 3000;
  const int n_taxa = argc > 4 ? atoi(argv[4]) : 60;
  const int method = argc > 5 ? atoi(argv[5]) : ASM_RF_FP4;
  const double beta = argc > 6 ? atof(argv[6]) : 3.0;
  const int n_traces = argc > 7 ? atoi(argv[7]) : 37;
  const int64_t n_samples = argc > 8 ? atoll(argv[8]) : 20000;
  if (n_gpus < 1 || n_gpus > 64 || asm_device_count() < n_gpus) {
    fprintf(stderr, "need %d CUDA devices, found %d\n", n_gpus, (int)asm_device_count());
    return 3;
  }

  /* ---- synthetic chains -> bit rows ---------------------------------------------------------------- */
  asm_encoder* enc = NULL;
  if (asm_encoder_create(&enc, n_taxa) != ASM_OK) return 2;
  int32_t** ids = (int32_t**)calloc((size_t)C, sizeof *ids);
  for (int c = 0; c < C; c++) {
    ids[c] = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_trees * (size_t)(n_taxa - 1));
    if (asm_synth_tree_chain(n_taxa, n_trees, 1, 20240100u + (uint64_t)c, 2, beta, NULL, NULL, NULL, enc, ids[c]) != ASM_OK)
      return 2;
  }
  const int words = asm_encoder_words(enc);
  const int64_t D = asm_encoder_num_clades(enc);
  uint64_t** bits = (uint64_t**)calloc((size_t)C, sizeof *bits);
  for (int c = 0; c < C; c++) {
    bits[c] = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n_trees * (size_t)words);
    if (asm_encode_bits(ids[c], n_trees, n_taxa - 1, words, bits[c]) != ASM_OK) return 2;
    free(ids[c]);
  }

  int32_t devices[64];
  for (int i = 0; i < n_gpus; i++) devices[i] = i;
  CHECK(asm_create(&ctx1, 0, C));
  CHECK(asm_create_multi(&ctxN, devices, n_gpus, C));
  for (int c = 0; c < C; c++) {
    CHECK(asm_trees_set(ctx1, c, n_trees, words, bits[c]));
    CHECK(asm_trees_set(ctxN, c, n_trees, words, bits[c]));
  }

  /* ---- PSRF: S and the means, ragged burn-in ---------------------------------------------------------- */
  int32_t burnin[ASM_MAX_CHAINS];
  for (int c = 0; c < C; c++) burnin[c] = (int32_t)((c * 37) % 101);
  const int32_t end = (int32_t)n_trees, delta = 1, row_start = 0;
  const size_t n_S = (size_t)C * (size_t)n_trees * (size_t)C;
  int64_t* S1 = (int64_t*)malloc(sizeof(int64_t) * n_S);
  int64_t* SN = (int64_t*)malloc(sizeof(int64_t) * n_S);
  const asm_shard all = {0, 1};
  CHECK(asm_psrf_sums(ctx1, burnin, row_start, end, delta, method, all, S1));
  CHECK(asm_psrf_sums(ctxN, burnin, row_start, end, delta, method, all, SN));
  const int s_same = memcmp(S1, SN, sizeof(int64_t) * n_S) == 0;
  uint64_t checksum = 0; /* sum of the table mod 2^64: the S_checksum of the bench lines */
  for (size_t i = 0; i < n_S; i++) checksum += (uint64_t)SN[i];

  double mean1[ASM_MAX_CHAINS * ASM_MAX_CHAINS], meanN[ASM_MAX_CHAINS * ASM_MAX_CHAINS];
  float ms1 = 0.f, msN = 0.f;
  for (int rep = 0; rep < 2; rep++) { /* second pass timed (buffers allocated, kernels loaded) */
    CHECK(asm_timer_mark(ctx1, 0));
    CHECK(asm_psrf_eval(ctx1, burnin, row_start, end, delta, method, NULL, mean1));
    CHECK(asm_timer_mark(ctx1, 1));
    CHECK(asm_timer_elapsed_ms(ctx1, 0, 1, &ms1));
    CHECK(asm_timer_mark(ctxN, 0));
    CHECK(asm_psrf_eval(ctxN, burnin, row_start, end, delta, method, NULL, meanN));
    CHECK(asm_timer_mark(ctxN, 1));
    CHECK(asm_timer_elapsed_ms(ctxN, 0, 1, &msN));
  }
  const int mean_same = memcmp(mean1, meanN, sizeof(double) * (size_t)C * (size_t)C) == 0;

  /* ---- clade counts: sharded + all-reduce ------------------------------------------------------------- */
  int32_t* cc1 = (int32_t*)malloc(sizeof(int32_t) * (size_t)words * 64);
  int32_t* ccN = (int32_t*)malloc(sizeof(int32_t) * (size_t)words * 64);
  int cc_same = 1;
  for (int c = 0; c < C && c < 3; c++) {
    CHECK(asm_clade_counts(ctx1, c, 3, n_trees - 1, cc1));
    CHECK(asm_clade_counts(ctxN, c, 3, n_trees - 1, ccN));
    cc_same &= memcmp(cc1, ccN, sizeof(int32_t) * (size_t)words * 64) == 0;
  }

  /* ---- ESS: traces dealt to the devices ---------------------------------------------------------------- */
  double* x = (double*)malloc(sizeof(double) * (size_t)n_traces * (size_t)n_samples);
  static const double phis[4] = {0.0, 0.5, 0.9, 0.99}, mus[3] = {0.0, 1.0, -1e4};
  for (int j = 0; j < n_traces; j++)
    if (asm_synth_ar1(x + (size_t)j * (size_t)n_samples, n_samples, mus[j % 3], phis[j % 4], 777u + (uint64_t)j) != ASM_OK) return 2;
  double* e1 = (double*)malloc(sizeof(double) * 4 * (size_t)n_traces);
  double *a1 = e1 + n_traces, *eN = e1 + 2 * n_traces, *aN = e1 + 3 * n_traces;
  CHECK(asm_ess_batch(ctx1, n_traces, n_samples, n_samples, x, ASM_MAX_LAG, a1, e1));
  CHECK(asm_ess_batch(ctxN, n_traces, n_samples, n_samples, x, ASM_MAX_LAG, aN, eN));
  int ess_same = memcmp(e1, eN, sizeof(double) * (size_t)n_traces) == 0 && memcmp(a1, aN, sizeof(double) * (size_t)n_traces) == 0;
  CHECK(asm_ess_traces_put(ctxN, n_traces, n_samples, n_samples, x));
  memset(eN, 0, sizeof(double) * 2 * (size_t)n_traces);
  const double t0 = now_s();
  CHECK(asm_ess_traces_eval(ctxN, ASM_MAX_LAG, aN, eN));
  const double ess_resident_ms = 1e3 * (now_s() - t0);
  ess_same &= memcmp(e1, eN, sizeof(double) * (size_t)n_traces) == 0 && memcmp(a1, aN, sizeof(double) * (size_t)n_traces) == 0;

  int32_t rank = -1, n_ranks = -1, n_local = -1;
  asm_comm_info(ctxN, &rank, &n_ranks, &n_local);
  const double T = (double)C * (double)n_trees;
  printf("{\"check\": \"multi_gpu_check\", \"n_gpus\": %d, \"n_ranks\": %d, \"n_local_devices\": %d, \"nccl_version\": %d, "
         "\"shape\": \"%d chains x %lld trees x %d taxa\", \"clades_D\": %lld, \"method\": %d, "
         "\"S_identical_to_1gpu\": %s, \"S_checksum\": \"%016llx\", \"psrf_means_identical\": %s, "
         "\"clade_counts_identical\": %s, \"ess_identical\": %s, "
         "\"psrf_eval_ms_1gpu\": %.3f, \"psrf_eval_ms_ngpu\": %.3f, \"tree_pairs_per_s_ngpu\": %.4g, "
         "\"ess_traces\": \"%d x %lld\", \"ess_resident_ms_ngpu\": %.3f, \"psrf_mean_01\": %.17g}\n",
         n_gpus, (int)n_ranks, (int)n_local, (int)asm_comm_nccl_version(), C, (long long)n_trees, n_taxa, (long long)D, method,
         s_same ? "true" : "false", (unsigned long long)checksum, mean_same ? "true" : "false", cc_same ? "true" : "false",
         ess_same ? "true" : "false", (double)ms1, (double)msN, T * T / (1e-3 * (double)msN), n_traces, (long long)n_samples,
         ess_resident_ms, C > 1 ? meanN[1] : 0.0);
  asm_destroy(ctx1);
  asm_destroy(ctxN);
  asm_encoder_destroy(enc);
  return (s_same && mean_same && cc_same && ess_same) ? 0 : 1;
}
