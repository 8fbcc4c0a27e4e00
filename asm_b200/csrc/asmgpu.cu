// asmgpu.cu -- context, tree / trace storage and the C-ABI glue of libasmgpu.so.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <vector>

#include "asm_internal.h"

static thread_local std::string g_create_err;

int32_t asm_fail(asm_ctx* ctx, int32_t code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->err = buf;
  else
    g_create_err = buf;
  return code;
}

int32_t asm_reserve(asm_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.bytes) return ASM_OK;
  if (b.p) {
    ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ASM_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
  }
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return asm_fail(ctx, ASM_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  b.bytes = want;
  return ASM_OK;
}

LaunchScope::LaunchScope(asm_ctx* c, int f, cudaStream_t s) : ctx(c), family(f), st(s ? s : c->stream) {
  ctx->launches++;
  if (ctx->profile) cudaEventRecord(ctx->pe0, st);
}
LaunchScope::~LaunchScope() {
  if (ctx->profile) {
    cudaEventRecord(ctx->pe1, st);
    cudaEventSynchronize(ctx->pe1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->pe0, ctx->pe1);
    ctx->prof_ms[family] += ms;
    ctx->prof_n[family] += 1;
  }
}

// Re-stride rows from `src_words` to `dst_words` (zero-extending), used when the dictionary grew.
__global__ void restride_kernel(const uint64_t* __restrict__ src, int src_words, uint64_t* __restrict__ dst,
                                int dst_words, int64_t n_rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n_rows * dst_words;
  for (; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / dst_words;
    int w = (int)(i - r * dst_words);
    dst[i] = w < src_words ? src[r * src_words + w] : 0ULL;
  }
}

__global__ void flush_kernel(uint4* p, size_t n, uint32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(v, v, v, v);
}

static int32_t select_device(const asm_ctx* ctx) {
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return asm_fail(const_cast<asm_ctx*>(ctx), ASM_E_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
  return ASM_OK;
}
#define ASM_ENTER(ctx)                         \
  do {                                         \
    if (!(ctx)) return ASM_E_INVALID;          \
    int32_t rc_ = select_device(ctx);          \
    if (rc_ != ASM_OK) return rc_;             \
  } while (0)

static int32_t ensure_tree_capacity(asm_ctx* ctx, int chain, int64_t need_rows, int32_t need_words) {
  // widen every chain if the dictionary grew
  if (need_words > ctx->words) {
    int32_t nw = need_words;
    for (int c = 0; c < ctx->n_chains; c++) {
      ChainTrees& t = ctx->trees[c];
      if (!t.d_bits || t.cap == 0) continue;
      uint64_t* nb = nullptr;
      cudaError_t e = cudaMalloc(&nb, sizeof(uint64_t) * (size_t)t.cap * nw);
      if (e != cudaSuccess) return asm_fail(ctx, ASM_E_NOMEM, "cudaMalloc widen failed: %s", cudaGetErrorString(e));
      if (t.n > 0) {
        LaunchScope ls(ctx, ASM_K_MISC);
        restride_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(t.d_bits, ctx->words, nb, nw, t.n);
      }
      ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      ASM_CUDA(ctx, cudaFree(t.d_bits));
      t.d_bits = nb;
    }
    ctx->words = nw;
  }
  ChainTrees& t = ctx->trees[chain];
  if (need_rows > t.cap) {
    int64_t nc = t.cap ? t.cap : 1024;
    while (nc < need_rows) nc *= 2;
    uint64_t* nb = nullptr;
    cudaError_t e = cudaMalloc(&nb, sizeof(uint64_t) * (size_t)nc * ctx->words);
    if (e != cudaSuccess) return asm_fail(ctx, ASM_E_NOMEM, "cudaMalloc trees failed: %s", cudaGetErrorString(e));
    if (t.n > 0)
      ASM_CUDA(ctx, cudaMemcpyAsync(nb, t.d_bits, sizeof(uint64_t) * (size_t)t.n * ctx->words, cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (t.d_bits) ASM_CUDA(ctx, cudaFree(t.d_bits));
    t.d_bits = nb;
    t.cap = nc;
  }
  return ASM_OK;
}

static int32_t trees_put(asm_ctx* ctx, int32_t chain, int64_t at, int64_t n_new, int32_t words, const uint64_t* bits,
                         cudaMemcpyKind kind) {
  if (chain < 0 || chain >= ctx->n_chains || n_new < 0 || words <= 0 || (n_new > 0 && !bits))
    return asm_fail(ctx, ASM_E_INVALID, "trees: bad argument (chain=%d n=%lld words=%d)", chain, (long long)n_new, words);
  int32_t rc = ensure_tree_capacity(ctx, chain, at + n_new, words);
  if (rc != ASM_OK) return rc;
  ChainTrees& t = ctx->trees[chain];
  if (n_new > 0) {
    if (words == ctx->words) {
      ASM_CUDA(ctx, cudaMemcpyAsync(t.d_bits + (size_t)at * ctx->words, bits, sizeof(uint64_t) * (size_t)n_new * words,
                                    kind, ctx->stream));
    } else {  // narrower than the stored width: 2D copy + zero fill of the tail columns
      ASM_CUDA(ctx, cudaMemsetAsync(t.d_bits + (size_t)at * ctx->words, 0, sizeof(uint64_t) * (size_t)n_new * ctx->words,
                                    ctx->stream));
      ASM_CUDA(ctx, cudaMemcpy2DAsync(t.d_bits + (size_t)at * ctx->words, sizeof(uint64_t) * ctx->words, bits,
                                      sizeof(uint64_t) * words, sizeof(uint64_t) * words, (size_t)n_new, kind,
                                      ctx->stream));
    }
    ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller may free `bits` on return
  }
  t.n = at + n_new;
  return ASM_OK;
}

extern "C" {

const char* asm_version(void) { return "asm-b200 0.1 (sm_100a)"; }

int32_t asm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int32_t asm_create(asm_ctx** out, int32_t device, int32_t n_chains) {
  if (!out) return ASM_E_INVALID;
  *out = nullptr;
  if (n_chains < 1 || n_chains > ASM_MAX_CHAINS)
    return asm_fail(nullptr, ASM_E_INVALID, "n_chains must be in [1,%d], not %d", ASM_MAX_CHAINS, n_chains);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return asm_fail(nullptr, ASM_E_NO_DEVICE, "no CUDA device (%s); libasmgpu has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= n) return asm_fail(nullptr, ASM_E_INVALID, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return asm_fail(nullptr, ASM_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return asm_fail(nullptr, ASM_E_NO_DEVICE, "device %d is sm_%d%d; libasmgpu is built for sm_100a only", device,
                    prop.major, prop.minor);
  auto* ctx = new (std::nothrow) asm_ctx();
  if (!ctx) return ASM_E_NOMEM;
  ctx->device = device;
  ctx->n_chains = n_chains;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
  ctx->trees.resize((size_t)n_chains);
  ctx->traces.resize((size_t)n_chains);
  {
    const char* e = getenv("ASM_ESS_ADAPTIVE");
    ctx->ess_adaptive = !(e && e[0] == '0');
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete ctx;
    return asm_fail(nullptr, ASM_E_CUDA, "stream create: %s", cudaGetErrorString(e));
  }
  cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->stream_copy, cudaStreamNonBlocking);
  for (int k = 0; k < asm_ctx::kMaxGroups; k++) {
    cudaStreamCreateWithFlags(&ctx->group_stream[k], cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->group_stream2[k], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->group_event[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->group_fork[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->group_join[k], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  for (auto& ev : ctx->timer) cudaEventCreate(&ev);
  cudaEventCreate(&ctx->pe0);
  cudaEventCreate(&ctx->pe1);
  *out = ctx;
  return ASM_OK;
}

int32_t asm_destroy(asm_ctx* ctx) {
  if (!ctx) return ASM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& t : ctx->trees)
    if (t.d_bits) cudaFree(t.d_bits);
  for (auto& t : ctx->traces)
    if (t.d_vals) cudaFree(t.d_vals);
  DevBuf* bufs[] = {&ctx->opnd_bits, &ctx->opnd_i8, &ctx->opnd_f4, &ctx->nbits, &ctx->S_pad, &ctx->blk_meta,
                    &ctx->S_compact, &ctx->psrf_rows, &ctx->psrf_mean, &ctx->misc, &ctx->flush, &ctx->ess_tr,
                    &ctx->ess_sls, &ctx->ess_out, &ctx->tile_counter, &ctx->rf_out, &ctx->walk_dev};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (auto& ev : ctx->timer)
    if (ev) cudaEventDestroy(ev);
  if (ctx->pe0) cudaEventDestroy(ctx->pe0);
  if (ctx->pe1) cudaEventDestroy(ctx->pe1);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream_copy) cudaStreamDestroy(ctx->stream_copy);
  for (int k = 0; k < asm_ctx::kMaxGroups; k++) {
    if (ctx->group_stream[k]) cudaStreamDestroy(ctx->group_stream[k]);
    if (ctx->group_stream2[k]) cudaStreamDestroy(ctx->group_stream2[k]);
    if (ctx->group_event[k]) cudaEventDestroy(ctx->group_event[k]);
    if (ctx->group_fork[k]) cudaEventDestroy(ctx->group_fork[k]);
    if (ctx->group_join[k]) cudaEventDestroy(ctx->group_join[k]);
  }
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return ASM_OK;
}

const char* asm_last_error(const asm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int32_t asm_trees_set(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t words_per_tree, const uint64_t* bits) {
  ASM_ENTER(ctx);
  return trees_put(ctx, chain, 0, n_trees, words_per_tree, bits, cudaMemcpyHostToDevice);
}

int32_t asm_trees_set_device(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t words_per_tree,
                             const uint64_t* d_bits) {
  ASM_ENTER(ctx);
  return trees_put(ctx, chain, 0, n_trees, words_per_tree, d_bits, cudaMemcpyDeviceToDevice);
}

int32_t asm_trees_append(asm_ctx* ctx, int32_t chain, int64_t n_new, int32_t words_per_tree, const uint64_t* bits) {
  ASM_ENTER(ctx);
  if (chain < 0 || chain >= ctx->n_chains) return asm_fail(ctx, ASM_E_INVALID, "chain %d out of range", chain);
  return trees_put(ctx, chain, ctx->trees[chain].n, n_new, words_per_tree, bits, cudaMemcpyHostToDevice);
}

int32_t asm_trees_count(const asm_ctx* ctx, int32_t chain, int64_t* n_trees, int32_t* words_per_tree) {
  if (!ctx || chain < 0 || chain >= ctx->n_chains) return ASM_E_INVALID;
  if (n_trees) *n_trees = ctx->trees[chain].n;
  if (words_per_tree) *words_per_tree = ctx->words;
  return ASM_OK;
}

int32_t asm_rf_block(asm_ctx* ctx, int32_t chainA, int64_t r0, int64_t r1, int32_t chainB, int64_t c0, int64_t c1,
                     int32_t method, uint16_t* out) {
  ASM_ENTER(ctx);
  if (chainA < 0 || chainA >= ctx->n_chains || chainB < 0 || chainB >= ctx->n_chains || !out || r0 < 0 || c0 < 0 ||
      r1 < r0 || c1 < c0)
    return asm_fail(ctx, ASM_E_INVALID, "rf_block: bad argument");
  if (r1 > ctx->trees[chainA].n || c1 > ctx->trees[chainB].n)
    return asm_fail(ctx, ASM_E_RANGE, "rf_block: range past the stored trees");
  if (r1 == r0 || c1 == c0) return ASM_OK;
  if (method == ASM_RF_POPC) return rf_popc_block(ctx, chainA, r0, r1, chainB, c0, c1, out);
  if (method == ASM_RF_I8) return rf_i8_block(ctx, chainA, r0, r1, chainB, c0, c1, out);
  if (method == ASM_RF_FP4) return asm_fail(ctx, ASM_E_UNSUPPORTED, "rf_block: the fp4 kernel has no distance dump; use ASM_RF_I8 or ASM_RF_POPC");
  return asm_fail(ctx, ASM_E_INVALID, "rf_block: unknown method %d", method);
}

int32_t asm_clade_counts(asm_ctx* ctx, int32_t chain, int64_t from, int64_t to, int32_t* out) {
  ASM_ENTER(ctx);
  if (chain < 0 || chain >= ctx->n_chains || !out || from < 0 || to < from)
    return asm_fail(ctx, ASM_E_INVALID, "clade_counts: bad argument");
  if (to > ctx->trees[chain].n) return asm_fail(ctx, ASM_E_RANGE, "clade_counts: range past the stored trees");
  return clade_counts(ctx, chain, from, to, out);
}

static int32_t psrf_sums_impl(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                              int32_t method, asm_shard shard, int64_t* d_S_out) {
  if (shard.shard_count < 1 || shard.shard_index < 0 || shard.shard_index >= shard.shard_count)
    return asm_fail(ctx, ASM_E_INVALID, "psrf: bad shard {%d,%d}", shard.shard_index, shard.shard_count);
  if (method != ASM_RF_POPC && method != ASM_RF_I8 && method != ASM_RF_FP4)
    return asm_fail(ctx, ASM_E_INVALID, "psrf: unknown method %d", method);
  int32_t rc = psrf_build_layout(ctx, burnin, row_start, end, delta, method == ASM_RF_FP4 ? 240 : kBlockRows);
  if (rc != ASM_OK) return rc;
  if (ctx->layout.n_rows == 0) return ASM_OK;
  if (method == ASM_RF_FP4) {
    if ((rc = psrf_gather_f4(ctx)) != ASM_OK) return rc;
    if ((rc = rf_f4_sums(ctx, shard)) != ASM_OK) return rc;
  } else if (method == ASM_RF_I8) {
    if ((rc = psrf_gather_i8(ctx)) != ASM_OK) return rc;
    if ((rc = rf_i8_sums(ctx, shard)) != ASM_OK) return rc;
  } else {
    if ((rc = psrf_gather_bits(ctx)) != ASM_OK) return rc;
    if ((rc = rf_popc_sums(ctx, shard)) != ASM_OK) return rc;
  }
  return psrf_compact(ctx, shard, d_S_out);
}

int32_t asm_psrf_sums_device(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                             int32_t method, asm_shard shard, int64_t* d_S) {
  ASM_ENTER(ctx);
  if (!burnin || !d_S) return asm_fail(ctx, ASM_E_INVALID, "psrf_sums: null argument");
  int32_t rc = psrf_sums_impl(ctx, burnin, row_start, end, delta, method, shard, d_S);
  if (rc != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

int32_t asm_psrf_sums_device_async(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                                   int32_t method, asm_shard shard, int64_t* d_S) {
  ASM_ENTER(ctx);
  if (!burnin || !d_S) return asm_fail(ctx, ASM_E_INVALID, "psrf_sums: null argument");
  return psrf_sums_impl(ctx, burnin, row_start, end, delta, method, shard, d_S);
}

void* asm_stream_handle(asm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int32_t asm_psrf_sums(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                      int32_t method, asm_shard shard, int64_t* S) {
  ASM_ENTER(ctx);
  if (!burnin || !S) return asm_fail(ctx, ASM_E_INVALID, "psrf_sums: null argument");
  int32_t rc = psrf_build_layout(ctx, burnin, row_start, end, delta);  // to size the output
  if (rc != ASM_OK) return rc;
  size_t bytes = sizeof(int64_t) * (size_t)ctx->n_chains * (size_t)ctx->layout.n_rows * (size_t)ctx->n_chains;
  if (bytes == 0) return ASM_OK;
  if ((rc = asm_reserve(ctx, ctx->S_compact, bytes)) != ASM_OK) return rc;
  rc = psrf_sums_impl(ctx, burnin, row_start, end, delta, method, shard, (int64_t*)ctx->S_compact.p);
  if (rc != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemcpyAsync(S, ctx->S_compact.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

int32_t asm_psrf_finish_device(asm_ctx* ctx, const int64_t* d_S, int32_t n_rows, double* psrf_rows, double* psrf_mean) {
  ASM_ENTER(ctx);
  if (!d_S || !psrf_mean || n_rows < 0) return asm_fail(ctx, ASM_E_INVALID, "psrf_finish: bad argument");
  return psrf_finish(ctx, d_S, n_rows, psrf_rows, psrf_mean);
}

int32_t asm_psrf_eval(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                      int32_t method, double* psrf_rows, double* psrf_mean) {
  ASM_ENTER(ctx);
  if (!burnin || !psrf_mean) return asm_fail(ctx, ASM_E_INVALID, "psrf_eval: null argument");
  int32_t rc = psrf_build_layout(ctx, burnin, row_start, end, delta);
  if (rc != ASM_OK) return rc;
  const int C = ctx->n_chains;
  size_t bytes = sizeof(int64_t) * (size_t)C * (size_t)ctx->layout.n_rows * (size_t)C;
  if ((rc = asm_reserve(ctx, ctx->S_compact, bytes ? bytes : 8)) != ASM_OK) return rc;
  asm_shard all{0, 1};
  if ((rc = psrf_sums_impl(ctx, burnin, row_start, end, delta, method, all, (int64_t*)ctx->S_compact.p)) != ASM_OK)
    return rc;
  return psrf_finish(ctx, (const int64_t*)ctx->S_compact.p, ctx->layout.n_rows, psrf_rows, psrf_mean);
}

int32_t asm_ess_batch_device(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* d_traces,
                             int32_t max_lag, double* act_out, double* ess_out) {
  ASM_ENTER(ctx);
  if (n_traces < 0 || n_samples < 0 || stride < n_samples || (n_traces > 0 && !d_traces) || max_lag < 1 ||
      max_lag > ASM_MAX_LAG)
    return asm_fail(ctx, ASM_E_INVALID, "ess_batch: bad argument");
  if (n_traces == 0) return ASM_OK;
  return ess_batch_device(ctx, n_traces, n_samples, stride, d_traces, max_lag, act_out, ess_out);
}

int32_t asm_ess_batch(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces,
                      int32_t max_lag, double* act_out, double* ess_out) {
  ASM_ENTER(ctx);
  if (n_traces < 0 || n_samples < 0 || stride < n_samples || (n_traces > 0 && !traces) || max_lag < 1 ||
      max_lag > ASM_MAX_LAG)
    return asm_fail(ctx, ASM_E_INVALID, "ess_batch: bad argument");
  if (n_traces == 0) return ASM_OK;
  // upload compactly: [n_traces][n_pad] with n_pad even so that every row is 16-byte aligned
  int64_t n_pad = (n_samples + 1) & ~int64_t(1);
  if (n_pad == 0) n_pad = 2;
  int32_t rc = asm_reserve(ctx, ctx->ess_tr, sizeof(double) * (size_t)n_traces * (size_t)n_pad);
  if (rc != ASM_OK) return rc;
  return ess_batch_host(ctx, n_traces, n_samples, stride, traces, n_pad, (double*)ctx->ess_tr.p, max_lag, act_out, ess_out);
}

int32_t asm_traces_append(asm_ctx* ctx, int32_t chain, int32_t n_cols, int64_t n_new, const double* values) {
  ASM_ENTER(ctx);
  if (chain < 0 || chain >= ctx->n_chains || n_cols < 1 || n_new < 0 || (n_new > 0 && !values))
    return asm_fail(ctx, ASM_E_INVALID, "traces_append: bad argument");
  ChainTraces& t = ctx->traces[chain];
  if (t.n_cols == 0) t.n_cols = n_cols;
  if (t.n_cols != n_cols)
    return asm_fail(ctx, ASM_E_INVALID, "traces_append: chain %d has %d columns, not %d", chain, t.n_cols, n_cols);
  if (t.n + n_new > t.cap) {
    int64_t nc = t.cap ? t.cap : 4096;
    while (nc < t.n + n_new) nc *= 2;
    double* nb = nullptr;
    cudaError_t e = cudaMalloc(&nb, sizeof(double) * (size_t)nc * n_cols);
    if (e != cudaSuccess) return asm_fail(ctx, ASM_E_NOMEM, "cudaMalloc traces: %s", cudaGetErrorString(e));
    if (t.n > 0)
      ASM_CUDA(ctx, cudaMemcpyAsync(nb, t.d_vals, sizeof(double) * (size_t)t.n * n_cols, cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (t.d_vals) ASM_CUDA(ctx, cudaFree(t.d_vals));
    t.d_vals = nb;
    t.cap = nc;
  }
  if (n_new > 0) {
    ASM_CUDA(ctx, cudaMemcpyAsync(t.d_vals + (size_t)t.n * n_cols, values, sizeof(double) * (size_t)n_new * n_cols,
                                  cudaMemcpyHostToDevice, ctx->stream));
    ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  t.n += n_new;
  return ASM_OK;
}

int32_t asm_traces_count(const asm_ctx* ctx, int32_t chain, int64_t* n_samples, int32_t* n_cols) {
  if (!ctx || chain < 0 || chain >= ctx->n_chains) return ASM_E_INVALID;
  if (n_samples) *n_samples = ctx->traces[chain].n;
  if (n_cols) *n_cols = ctx->traces[chain].n_cols;
  return ASM_OK;
}

int32_t asm_ess_eval(asm_ctx* ctx, const int32_t* burnin, int32_t end, const int32_t* cols, int32_t n_cols,
                     double* ess_out) {
  ASM_ENTER(ctx);
  if (!burnin || !cols || n_cols < 0 || !ess_out || end < 0) return asm_fail(ctx, ASM_E_INVALID, "ess_eval: bad argument");
  for (int c = 0; c < ctx->n_chains; c++) {
    if (end > ctx->traces[c].n) return asm_fail(ctx, ASM_E_RANGE, "ess_eval: end %d past chain %d's %lld samples", end, c, (long long)ctx->traces[c].n);
    if (burnin[c] < 0) return asm_fail(ctx, ASM_E_RANGE, "ess_eval: negative burn-in");
    for (int j = 0; j < n_cols; j++)
      if (cols[j] < 0 || cols[j] >= ctx->traces[c].n_cols)
        return asm_fail(ctx, ASM_E_RANGE, "ess_eval: column %d out of range", cols[j]);
  }
  if (n_cols == 0) return ASM_OK;
  return ess_eval_stream(ctx, burnin, end, cols, n_cols, ess_out);
}

int32_t asm_trace_moments(asm_ctx* ctx, int32_t end, double* sums_out, double* sumsq_out) {
  ASM_ENTER(ctx);
  if (!sums_out || !sumsq_out || end < 0) return asm_fail(ctx, ASM_E_INVALID, "trace_moments: bad argument");
  const int n_cols = ctx->traces[0].n_cols;
  for (int c = 0; c < ctx->n_chains; c++) {
    if (ctx->traces[c].n_cols != n_cols) return asm_fail(ctx, ASM_E_INVALID, "trace_moments: chains differ in column count");
    if (end > ctx->traces[c].n)
      return asm_fail(ctx, ASM_E_RANGE, "trace_moments: end %d past chain %d's %lld samples", end, c, (long long)ctx->traces[c].n);
  }
  if (n_cols == 0) return ASM_OK;
  return trace_moments(ctx, end, sums_out, sumsq_out);
}

int32_t asm_timer_mark(asm_ctx* ctx, int32_t slot) {
  ASM_ENTER(ctx);
  if (slot < 0 || slot >= 16) return asm_fail(ctx, ASM_E_INVALID, "timer slot %d", slot);
  ASM_CUDA(ctx, cudaEventRecord(ctx->timer[slot], ctx->stream));
  return ASM_OK;
}

int32_t asm_timer_elapsed_ms(asm_ctx* ctx, int32_t from, int32_t to, float* ms) {
  ASM_ENTER(ctx);
  if (from < 0 || from >= 16 || to < 0 || to >= 16 || !ms) return asm_fail(ctx, ASM_E_INVALID, "timer slots");
  ASM_CUDA(ctx, cudaEventSynchronize(ctx->timer[to]));
  ASM_CUDA(ctx, cudaEventElapsedTime(ms, ctx->timer[from], ctx->timer[to]));
  return ASM_OK;
}

int32_t asm_profile_enable(asm_ctx* ctx, int32_t on) {
  if (!ctx) return ASM_E_INVALID;
  ctx->profile = on != 0;
  return ASM_OK;
}
int32_t asm_profile_reset(asm_ctx* ctx) {
  if (!ctx) return ASM_E_INVALID;
  for (int i = 0; i < ASM_K_COUNT; i++) {
    ctx->prof_ms[i] = 0;
    ctx->prof_n[i] = 0;
  }
  return ASM_OK;
}
int32_t asm_profile_get(asm_ctx* ctx, int32_t family, double* total_ms, int64_t* launches) {
  if (!ctx || family < 0 || family >= ASM_K_COUNT) return ASM_E_INVALID;
  if (total_ms) *total_ms = ctx->prof_ms[family];
  if (launches) *launches = ctx->prof_n[family];
  return ASM_OK;
}
int64_t asm_launch_count(const asm_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t asm_sync(asm_ctx* ctx) {
  ASM_ENTER(ctx);
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

int32_t asm_flush_l2(asm_ctx* ctx) {
  ASM_ENTER(ctx);
  const size_t bytes = 256u << 20;  // 2x the 126 MB L2
  int32_t rc = asm_reserve(ctx, ctx->flush, bytes);
  if (rc != ASM_OK) return rc;
  static uint32_t v = 0;
  {
    LaunchScope ls(ctx, ASM_K_MISC);
    flush_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint4*)ctx->flush.p, bytes / 16, ++v);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

int64_t asm_ess_last_lag_count(const asm_ctx* ctx) { return ctx ? ctx->ess_lag_slots : 0; }
int32_t asm_ess_set_adaptive(asm_ctx* ctx, int32_t on) {
  if (!ctx) return ASM_E_INVALID;
  ctx->ess_adaptive = on != 0;
  return ASM_OK;
}

int32_t asm_psrf_last_layout(const asm_ctx* ctx, int64_t* padded_rows, int64_t* padded_bits, int64_t* tiles_done,
                             int64_t* tiles_total) {
  if (!ctx) return ASM_E_INVALID;
  if (padded_rows) *padded_rows = ctx->layout.Tp;
  if (padded_bits) *padded_bits = ctx->layout.bits_used;
  if (tiles_done) *tiles_done = ctx->layout.tiles_done;
  if (tiles_total) *tiles_total = ctx->layout.tiles_total;
  return ASM_OK;
}

}  // extern "C"
