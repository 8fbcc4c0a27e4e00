// asmgpu.cu -- context, tree / trace storage and the C-ABI glue of libasmgpu.so.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <new>
#include <vector>

#include <sched.h>

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


// Clade ids -> bit rows, one warp per tree: the row is assembled in shared memory (ids of one tree hit the same few words)
// and written out whole, so the stored rows never see an atomic.
__global__ void __launch_bounds__(256) ids_to_bits_kernel(const uint16_t* __restrict__ ids, int ids_per_tree, int64_t n_trees,
                                                           uint64_t* __restrict__ rows, int words, int n_clades,
                                                           int* __restrict__ bad) {
  extern __shared__ unsigned long long row_s[];  // [warps per block][words]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  unsigned long long* row = row_s + (size_t)warp * words;
  for (int64_t t = (int64_t)blockIdx.x * wpb + warp; t < n_trees; t += (int64_t)gridDim.x * wpb) {
    for (int w = lane; w < words; w += 32) row[w] = 0ull;
    __syncwarp();
    const uint16_t* src = ids + (size_t)t * ids_per_tree;
    for (int k = lane; k < ids_per_tree; k += 32) {
      const unsigned id = src[k];
      if (id == ASM_NO_CLADE) continue;
      if ((int)id >= n_clades) {
        atomicExch(bad, 1);
        continue;
      }
      atomicOr(&row[id >> 6], 1ull << (id & 63));
    }
    __syncwarp();
    for (int w = lane; w < words; w += 32) rows[(size_t)t * words + w] = row[w];
    __syncwarp();
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

// Rows [at, at + n_new) of `chain` <- bits (host or device per `kind`), zero-extended to the stored width.  Enqueued
// on the context's stream; the caller synchronises.
static int32_t trees_upload_rows(asm_ctx* ctx, int32_t chain, int64_t at, int64_t n_new, int32_t words, const uint64_t* bits,
                                 cudaMemcpyKind kind) {
  ChainTrees& t = ctx->trees[chain];
  if (n_new <= 0) return ASM_OK;
  if (words == ctx->words) {
    ASM_CUDA(ctx, cudaMemcpyAsync(t.d_bits + (size_t)at * ctx->words, bits, sizeof(uint64_t) * (size_t)n_new * words, kind,
                                  ctx->stream));
  } else {  // narrower than the stored width: 2D copy + zero fill of the tail columns
    ASM_CUDA(ctx, cudaMemsetAsync(t.d_bits + (size_t)at * ctx->words, 0, sizeof(uint64_t) * (size_t)n_new * ctx->words,
                                  ctx->stream));
    ASM_CUDA(ctx, cudaMemcpy2DAsync(t.d_bits + (size_t)at * ctx->words, sizeof(uint64_t) * ctx->words, bits,
                                    sizeof(uint64_t) * words, sizeof(uint64_t) * words, (size_t)n_new, kind, ctx->stream));
  }
  return ASM_OK;
}

static int32_t trees_check(asm_ctx* ctx, int32_t chain, int64_t n_new, int32_t words, const void* bits) {
  if (chain < 0 || chain >= ctx->n_chains || n_new < 0 || words <= 0 || (n_new > 0 && !bits))
    return asm_fail(ctx, ASM_E_INVALID, "trees: bad argument (chain=%d n=%lld words=%d)", chain, (long long)n_new, words);
  return ASM_OK;
}

static int32_t trees_put(asm_ctx* ctx, int32_t chain, int64_t at, int64_t n_new, int32_t words, const uint64_t* bits,
                         cudaMemcpyKind kind) {
  int32_t rc = trees_check(ctx, chain, n_new, words, bits);
  if (rc != ASM_OK) return rc;
  if ((rc = ensure_tree_capacity(ctx, chain, at + n_new, words)) != ASM_OK) return rc;
  if ((rc = trees_upload_rows(ctx, chain, at, n_new, words, bits, kind)) != ASM_OK) return rc;
  if (n_new > 0) ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller may free or overwrite `bits` on return
  ctx->trees[chain].n = at + n_new;
  return ASM_OK;
}

// Rows [at, at + n_new) of `chain` from clade ids on the host (asm_trees_set_ids / asm_trees_append_ids).
static int32_t trees_put_ids(asm_ctx* ctx, int32_t chain, int64_t at, int64_t n_new, int32_t ids_per_tree, const uint16_t* ids,
                             int32_t n_clades) {
  if (chain < 0 || chain >= ctx->n_chains || n_new < 0 || ids_per_tree < 1 || n_clades < 1 || n_clades > 0xFFFF || (n_new > 0 && !ids))
    return asm_fail(ctx, ASM_E_INVALID, "trees (ids): bad argument (chain=%d n=%lld ids_per_tree=%d n_clades=%d)", chain,
                    (long long)n_new, ids_per_tree, n_clades);
  const int32_t words = std::max(4, (n_clades + 255) / 256 * 4);  // whole 256-clade units, as the host encoder lays rows out
  int32_t rc = ensure_tree_capacity(ctx, chain, at + n_new, words);
  if (rc != ASM_OK) return rc;
  if (n_new > 0) {
    const size_t bytes = sizeof(uint16_t) * (size_t)n_new * (size_t)ids_per_tree;
    // Two staging buffers and the copy stream: the copy of one chain's ids runs while the previous chain's rows are
    // still being built (a check uploads chain after chain).  ev_ids_used[slot] keeps a copy from overwriting ids the
    // kernel two calls back has not read yet.
    const int slot = (ctx->ids_slot ^= 1);
    DevBuf& stage = ctx->ids_stage[slot];
    if (stage.bytes < bytes) ASM_CUDA(ctx, cudaEventSynchronize(ctx->ev_ids_used[slot]));  // about to be freed
    if ((rc = asm_reserve(ctx, stage, bytes)) != ASM_OK) return rc;
    uint16_t* d_ids = (uint16_t*)stage.p;
    if (!ctx->ids_flag.p) {  // set by the kernel when an id is outside the dictionary; read with the next evaluation
      if ((rc = asm_reserve(ctx, ctx->ids_flag, sizeof(int))) != ASM_OK) return rc;
      ASM_CUDA(ctx, cudaMemsetAsync(ctx->ids_flag.p, 0, sizeof(int), ctx->stream));
    }
    ASM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_copy, ctx->ev_ids_used[slot], 0));
    ASM_CUDA(ctx, cudaMemcpyAsync(d_ids, ids, bytes, cudaMemcpyHostToDevice, ctx->stream_copy));
    ASM_CUDA(ctx, cudaEventRecord(ctx->ev_ids, ctx->stream_copy));
    ASM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ids, 0));
    const int wpb = 8;
    const size_t smem = sizeof(uint64_t) * (size_t)wpb * (size_t)ctx->words;
    if (smem > 48 * 1024 && !(ctx->func_attr_done & kAttrIdsToBits)) {
      ASM_CUDA(ctx, cudaFuncSetAttribute(ids_to_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin));
      ctx->func_attr_done |= kAttrIdsToBits;
    }
    {
      LaunchScope ls(ctx, ASM_K_MISC);
      const int blocks = (int)std::min<int64_t>((n_new + wpb - 1) / wpb, (int64_t)ctx->sm_count * 8);
      ids_to_bits_kernel<<<blocks, 32 * wpb, smem, ctx->stream>>>(d_ids, ids_per_tree, n_new,
                                                                 ctx->trees[chain].d_bits + (size_t)at * ctx->words, ctx->words,
                                                                 n_clades, (int*)ctx->ids_flag.p);
    }
    ASM_CUDA(ctx, cudaGetLastError());
    ASM_CUDA(ctx, cudaEventRecord(ctx->ev_ids_used[slot], ctx->stream));
    ctx->ids_check_pending = true;
    // the caller may free `ids` on return: wait for the copy; the expansion runs behind it in stream order, and so does
    // whatever reads the rows next
    ASM_CUDA(ctx, cudaEventSynchronize(ctx->ev_ids));
  }
  ctx->trees[chain].n = at + n_new;
  return ASM_OK;
}

// Did an asm_trees_*_ids call since the last check carry an id outside its dictionary?  (The flag is read here, by the
// next call that looks at the trees, so that the upload itself waits for nothing but its copy.)
int32_t trees_ids_verify(asm_ctx* ctx) {
  if (!ctx->ids_check_pending) return ASM_OK;
  int flag = 0;
  ASM_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->ids_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->ids_check_pending = false;
  if (flag) {
    ASM_CUDA(ctx, cudaMemsetAsync(ctx->ids_flag.p, 0, sizeof(int), ctx->stream));
    return asm_fail(ctx, ASM_E_RANGE, "trees (ids): an uploaded clade id was not below its n_clades");
  }
  return ASM_OK;
}

// Replace a chain from host rows on a context that is one rank of N: every row crosses PCIe once per job, not
// once per GPU -- rank r uploads rows [r*per, (r+1)*per) and one in-place all-gather over NVLink replicates them.
static int32_t trees_set_sharded(asm_ctx* ctx, int32_t chain, int64_t n, int32_t words, const uint64_t* bits) {
  int32_t rc = trees_check(ctx, chain, n, words, bits);
  if (rc != ASM_OK) return rc;
  const int64_t per = (n + ctx->world - 1) / ctx->world;
  if ((rc = ensure_tree_capacity(ctx, chain, per * ctx->world, words)) != ASM_OK) return rc;
  const int64_t lo = std::min<int64_t>((int64_t)ctx->rank * per, n), hi = std::min<int64_t>(lo + per, n);
  if ((rc = trees_upload_rows(ctx, chain, lo, hi - lo, words, bits ? bits + (size_t)lo * words : nullptr,
                              cudaMemcpyHostToDevice)) != ASM_OK)
    return rc;
  if ((rc = comm_all_gather_u64_inplace(ctx, ctx->trees[chain].d_bits, (size_t)per * ctx->words)) != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->trees[chain].n = n;
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

int32_t asm_bind_thread_near_device(int32_t device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return ASM_E_INVALID;
  }
  char bdf[32] = {0};
  if (cudaDeviceGetPCIBusId(bdf, (int)sizeof bdf, device) != cudaSuccess) {
    cudaGetLastError();
    return ASM_E_CUDA;
  }
  for (char* c = bdf; *c; c++) *c = (char)tolower((unsigned char)*c);
  auto slurp = [](const std::string& path) {
    std::string out;
    if (FILE* f = fopen(path.c_str(), "r")) {
      char buf[4096];
      const size_t k = fread(buf, 1, sizeof buf - 1, f);
      buf[k] = 0;
      out = buf;
      fclose(f);
    }
    return out;
  };
  const std::string dir = std::string("/sys/bus/pci/devices/") + bdf + "/";
  const std::string node = slurp(dir + "numa_node");
  if (node.empty() || atoi(node.c_str()) < 0) return 0;            // the platform does not say (e.g. a VM): nothing to do
  const std::string online = slurp("/sys/devices/system/node/online");
  if (online.find_first_of(",-") == std::string::npos) return 0;   // a single NUMA node
  const std::string list = slurp(dir + "local_cpulist");
  cpu_set_t set;
  CPU_ZERO(&set);
  int count = 0;
  for (const char* p = list.c_str(); *p;) {  // "0-15,32-47"
    if (!isdigit((unsigned char)*p)) {
      p++;
      continue;
    }
    char* end = nullptr;
    long a = strtol(p, &end, 10), b = a;
    if (*end == '-') b = strtol(end + 1, &end, 10);
    for (long c = a; c <= b && c < CPU_SETSIZE; c++) {
      CPU_SET((int)c, &set);
      count++;
    }
    p = end;
  }
  if (count == 0) return 0;
  if (sched_setaffinity(0, sizeof set, &set) != 0) return 0;       // not permitted (cgroup cpuset): leave the thread where it is
  return count;
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
  cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_ids, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_ids_used[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_ids_used[1], cudaEventDisableTiming);
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
  if (ctx->parent) return asm_fail(ctx, ASM_E_INVALID, "asm_destroy: a member of a multi-GPU context is destroyed with its parent");
  if (!ctx->members.empty() || ctx->pool) {
    group_destroy(ctx);
    delete ctx;
    return ASM_OK;
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx);
  for (auto& t : ctx->trees)
    if (t.d_bits) cudaFree(t.d_bits);
  for (auto& t : ctx->traces)
    if (t.d_vals) cudaFree(t.d_vals);
  DevBuf* bufs[] = {&ctx->opnd_bits, &ctx->opnd_i8, &ctx->opnd_f4, &ctx->nbits, &ctx->S_pad, &ctx->blk_meta,
                    &ctx->S_compact, &ctx->psrf_rows, &ctx->psrf_vals, &ctx->psrf_mean, &ctx->misc, &ctx->flush, &ctx->ess_tr,
                    &ctx->ess_sls, &ctx->ess_out, &ctx->tile_counter, &ctx->rf_out, &ctx->walk_dev, &ctx->ess_res, &ctx->f4_blk_dev, &ctx->f4_walk_dev, &ctx->top_word_dev, &ctx->ess_park, &ctx->ids_stage[0], &ctx->ids_stage[1], &ctx->ids_flag};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (auto& ev : ctx->timer)
    if (ev) cudaEventDestroy(ev);
  if (ctx->pe0) cudaEventDestroy(ctx->pe0);
  if (ctx->pe1) cudaEventDestroy(ctx->pe1);
  if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
  if (ctx->ev_ids) cudaEventDestroy(ctx->ev_ids);
  for (auto& ev : ctx->ev_ids_used)
    if (ev) cudaEventDestroy(ev);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
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

// A group parent forwards to its members' threads; `is_group` guards every entry point below.
static inline bool is_group(const asm_ctx* ctx) { return ctx && !ctx->members.empty(); }
#define ASM_NULLCHECK(ctx) \
  do {                     \
    if (!(ctx)) return ASM_E_INVALID; \
  } while (0)

int32_t asm_trees_set(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t words_per_tree, const uint64_t* bits) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))
    return group_run(ctx, [=](asm_ctx* m, int) { return asm_trees_set(m, chain, n_trees, words_per_tree, bits); });
  ASM_ENTER(ctx);
  if (ctx->world > 1) return trees_set_sharded(ctx, chain, n_trees, words_per_tree, bits);
  return trees_put(ctx, chain, 0, n_trees, words_per_tree, bits, cudaMemcpyHostToDevice);
}

// d_bits lives on this context's GPU (in a group: on the first member's).  A member of a group receives the rows
// of member 0 by broadcast; a rank of a multi-process communicator owns its own pointer and copies locally.
static int32_t trees_set_device_rank(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t words, const uint64_t* d_bits) {
  if (!ctx->parent || ctx->world == 1) return trees_put(ctx, chain, 0, n_trees, words, d_bits, cudaMemcpyDeviceToDevice);
  int32_t rc = trees_check(ctx, chain, n_trees, words, ctx->rank == 0 ? (const void*)d_bits : (const void*)ctx);
  if (rc != ASM_OK) return rc;
  if ((rc = ensure_tree_capacity(ctx, chain, n_trees, words)) != ASM_OK) return rc;
  if (ctx->rank == 0 && (rc = trees_upload_rows(ctx, chain, 0, n_trees, words, d_bits, cudaMemcpyDeviceToDevice)) != ASM_OK)
    return rc;
  if ((rc = comm_broadcast_u64(ctx, ctx->trees[chain].d_bits, (size_t)n_trees * ctx->words, 0)) != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->trees[chain].n = n_trees;
  return ASM_OK;
}

int32_t asm_trees_set_device(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t words_per_tree,
                             const uint64_t* d_bits) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))
    return group_run(ctx, [=](asm_ctx* m, int i) {
      return asm_trees_set_device(m, chain, n_trees, words_per_tree, i == 0 ? d_bits : nullptr);
    });
  ASM_ENTER(ctx);
  return trees_set_device_rank(ctx, chain, n_trees, words_per_tree, d_bits);
}

int32_t asm_trees_append(asm_ctx* ctx, int32_t chain, int64_t n_new, int32_t words_per_tree, const uint64_t* bits) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))  // a tree per chain per check: every member takes the few new rows from the host itself
    return group_run(ctx, [=](asm_ctx* m, int) { return asm_trees_append(m, chain, n_new, words_per_tree, bits); });
  ASM_ENTER(ctx);
  if (chain < 0 || chain >= ctx->n_chains) return asm_fail(ctx, ASM_E_INVALID, "chain %d out of range", chain);
  return trees_put(ctx, chain, ctx->trees[chain].n, n_new, words_per_tree, bits, cudaMemcpyHostToDevice);
}

int32_t asm_trees_set_ids(asm_ctx* ctx, int32_t chain, int64_t n_trees, int32_t ids_per_tree, const uint16_t* ids,
                          int32_t n_clades) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))  // ids are small: every member takes them from the host itself and builds its own rows
    return group_run(ctx, [=](asm_ctx* m, int) { return asm_trees_set_ids(m, chain, n_trees, ids_per_tree, ids, n_clades); });
  ASM_ENTER(ctx);
  return trees_put_ids(ctx, chain, 0, n_trees, ids_per_tree, ids, n_clades);
}

int32_t asm_trees_append_ids(asm_ctx* ctx, int32_t chain, int64_t n_new, int32_t ids_per_tree, const uint16_t* ids,
                             int32_t n_clades) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))
    return group_run(ctx, [=](asm_ctx* m, int) { return asm_trees_append_ids(m, chain, n_new, ids_per_tree, ids, n_clades); });
  ASM_ENTER(ctx);
  if (chain < 0 || chain >= ctx->n_chains) return asm_fail(ctx, ASM_E_INVALID, "chain %d out of range", chain);
  return trees_put_ids(ctx, chain, ctx->trees[chain].n, n_new, ids_per_tree, ids, n_clades);
}

int32_t asm_trees_count(const asm_ctx* ctx, int32_t chain, int64_t* n_trees, int32_t* words_per_tree) {
  if (is_group(ctx)) ctx = ctx->members[0];
  if (!ctx || chain < 0 || chain >= ctx->n_chains) return ASM_E_INVALID;
  if (n_trees) *n_trees = ctx->trees[chain].n;
  if (words_per_tree) *words_per_tree = ctx->words;
  return ASM_OK;
}

int32_t asm_rf_block(asm_ctx* ctx, int32_t chainA, int64_t r0, int64_t r1, int32_t chainB, int64_t c0, int64_t c1,
                     int32_t method, uint16_t* out) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))  // a parity probe: every member holds every tree, the first one answers
    return group_run_one(ctx, 0, [=](asm_ctx* m) { return asm_rf_block(m, chainA, r0, r1, chainB, c0, c1, method, out); });
  ASM_ENTER(ctx);
  if (chainA < 0 || chainA >= ctx->n_chains || chainB < 0 || chainB >= ctx->n_chains || !out || r0 < 0 || c0 < 0 ||
      r1 < r0 || c1 < c0)
    return asm_fail(ctx, ASM_E_INVALID, "rf_block: bad argument");
  if (r1 > ctx->trees[chainA].n || c1 > ctx->trees[chainB].n)
    return asm_fail(ctx, ASM_E_RANGE, "rf_block: range past the stored trees");
  if (r1 == r0 || c1 == c0) return ASM_OK;
  {
    const int32_t rcv = trees_ids_verify(ctx);
    if (rcv != ASM_OK) return rcv;
  }
  if (method == ASM_RF_POPC) return rf_popc_block(ctx, chainA, r0, r1, chainB, c0, c1, out);
  if (method == ASM_RF_I8) return rf_i8_block(ctx, chainA, r0, r1, chainB, c0, c1, out);
  if (method == ASM_RF_FP4) return asm_fail(ctx, ASM_E_UNSUPPORTED, "rf_block: the fp4 kernel has no distance dump; use ASM_RF_I8 or ASM_RF_POPC");
  return asm_fail(ctx, ASM_E_INVALID, "rf_block: unknown method %d", method);
}

static int32_t clade_counts_rank(asm_ctx* ctx, int32_t chain, int64_t from, int64_t to, int32_t* out, bool want_out) {
  if (chain < 0 || chain >= ctx->n_chains || (want_out && !out) || from < 0 || to < from)
    return asm_fail(ctx, ASM_E_INVALID, "clade_counts: bad argument");
  if (to > ctx->trees[chain].n) return asm_fail(ctx, ASM_E_RANGE, "clade_counts: range past the stored trees");
  {
    const int32_t rcv = trees_ids_verify(ctx);
    if (rcv != ASM_OK) return rcv;
  }
  return clade_counts(ctx, chain, from, to, want_out ? out : nullptr);
}

int32_t asm_clade_counts(asm_ctx* ctx, int32_t chain, int64_t from, int64_t to, int32_t* out) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx)) {
    if (!out) return asm_fail(ctx, ASM_E_INVALID, "clade_counts: bad argument");
    return group_run(ctx, [=](asm_ctx* m, int i) {
      int32_t rc = select_device(m);
      return rc != ASM_OK ? rc : clade_counts_rank(m, chain, from, to, out, i == 0);
    });
  }
  ASM_ENTER(ctx);
  return clade_counts_rank(ctx, chain, from, to, out, true);
}

// Partial sums of this rank's share of the tile pairs, then (with a communicator) one int64 all-reduce on the same
// stream: exact in any order, so every rank ends with the complete table.
static int32_t psrf_sums_impl(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                              int32_t method, asm_shard shard, int64_t* d_S_out, bool verify_ids = true) {
  if (shard.shard_count < 1 || shard.shard_index < 0 || shard.shard_index >= shard.shard_count)
    return asm_fail(ctx, ASM_E_INVALID, "psrf: bad shard {%d,%d}", shard.shard_index, shard.shard_count);
  if (method != ASM_RF_POPC && method != ASM_RF_I8 && method != ASM_RF_FP4)
    return asm_fail(ctx, ASM_E_INVALID, "psrf: unknown method %d", method);
  // the caller's split composes with the communicator's: slot t of the caller's shard goes to rank t % world
  const asm_shard eff{shard.shard_index * ctx->world + ctx->rank, shard.shard_count * ctx->world};
  int32_t rc;
  if (verify_ids && (rc = trees_ids_verify(ctx)) != ASM_OK) return rc;  // asm_psrf_eval checks with its final copy instead
  rc = psrf_build_layout(ctx, burnin, row_start, end, delta, method == ASM_RF_FP4 ? 240 : kBlockRows);
  if (rc != ASM_OK) return rc;
  if (ctx->layout.n_rows == 0) return ASM_OK;
  if (method == ASM_RF_FP4) {
    if ((rc = psrf_gather_f4(ctx)) != ASM_OK) return rc;
    if ((rc = rf_f4_sums(ctx, eff)) != ASM_OK) return rc;
  } else if (method == ASM_RF_I8) {
    if ((rc = psrf_gather_i8(ctx)) != ASM_OK) return rc;
    if ((rc = rf_i8_sums(ctx, eff)) != ASM_OK) return rc;
  } else {
    if ((rc = psrf_gather_bits(ctx)) != ASM_OK) return rc;
    if ((rc = rf_popc_sums(ctx, eff)) != ASM_OK) return rc;
  }
  if ((rc = psrf_compact(ctx, eff, d_S_out)) != ASM_OK) return rc;
  return comm_all_reduce_i64(ctx, d_S_out, (size_t)ctx->n_chains * (size_t)ctx->layout.n_rows * (size_t)ctx->n_chains);
}

// d_S == NULL: the table stays in the context's own buffer (members of a group other than the first)
static int32_t psrf_sums_device_rank(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                                     int32_t method, asm_shard shard, int64_t* d_S, bool sync, bool verify_ids = true) {
  if (!burnin) return asm_fail(ctx, ASM_E_INVALID, "psrf_sums: null argument");
  int32_t rc;
  if (!d_S) {
    if ((rc = psrf_build_layout(ctx, burnin, row_start, end, delta)) != ASM_OK) return rc;
    const size_t bytes = sizeof(int64_t) * (size_t)ctx->n_chains * (size_t)ctx->layout.n_rows * (size_t)ctx->n_chains;
    if ((rc = asm_reserve(ctx, ctx->S_compact, bytes ? bytes : 8)) != ASM_OK) return rc;
    d_S = (int64_t*)ctx->S_compact.p;
  }
  if ((rc = psrf_sums_impl(ctx, burnin, row_start, end, delta, method, shard, d_S, verify_ids)) != ASM_OK) return rc;
  if (sync) ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

int32_t asm_psrf_sums_device(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                             int32_t method, asm_shard shard, int64_t* d_S) {
  ASM_NULLCHECK(ctx);
  if (!d_S) return asm_fail(ctx, ASM_E_INVALID, "psrf_sums: null argument");
  if (is_group(ctx))
    return group_run(ctx, [=](asm_ctx* m, int i) {
      int32_t rc = select_device(m);
      return rc != ASM_OK ? rc : psrf_sums_device_rank(m, burnin, row_start, end, delta, method, shard, i == 0 ? d_S : nullptr, true);
    });
  ASM_ENTER(ctx);
  return psrf_sums_device_rank(ctx, burnin, row_start, end, delta, method, shard, d_S, true);
}

int32_t asm_psrf_sums_device_async(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                                   int32_t method, asm_shard shard, int64_t* d_S) {
  ASM_NULLCHECK(ctx);
  if (!d_S) return asm_fail(ctx, ASM_E_INVALID, "psrf_sums: null argument");
  if (is_group(ctx))
    return group_run(ctx, [=](asm_ctx* m, int i) {
      int32_t rc = select_device(m);
      return rc != ASM_OK ? rc : psrf_sums_device_rank(m, burnin, row_start, end, delta, method, shard, i == 0 ? d_S : nullptr, false);
    });
  ASM_ENTER(ctx);
  return psrf_sums_device_rank(ctx, burnin, row_start, end, delta, method, shard, d_S, false);
}

void* asm_stream_handle(asm_ctx* ctx) {
  if (is_group(ctx)) ctx = ctx->members[0];
  return ctx ? (void*)ctx->stream : nullptr;
}

// S == NULL (members of a group other than the first): compute and reduce, copy nothing out
static int32_t psrf_sums_rank(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                              int32_t method, asm_shard shard, int64_t* S) {
  if (!burnin) return asm_fail(ctx, ASM_E_INVALID, "psrf_sums: null argument");
  int32_t rc = psrf_sums_device_rank(ctx, burnin, row_start, end, delta, method, shard, nullptr, false);
  if (rc != ASM_OK) return rc;
  const size_t bytes = sizeof(int64_t) * (size_t)ctx->n_chains * (size_t)ctx->layout.n_rows * (size_t)ctx->n_chains;
  if (S && bytes) ASM_CUDA(ctx, cudaMemcpyAsync(S, ctx->S_compact.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

int32_t asm_psrf_sums(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                      int32_t method, asm_shard shard, int64_t* S) {
  ASM_NULLCHECK(ctx);
  if (!S) return asm_fail(ctx, ASM_E_INVALID, "psrf_sums: null argument");
  if (is_group(ctx))
    return group_run(ctx, [=](asm_ctx* m, int i) {
      int32_t rc = select_device(m);
      return rc != ASM_OK ? rc : psrf_sums_rank(m, burnin, row_start, end, delta, method, shard, i == 0 ? S : nullptr);
    });
  ASM_ENTER(ctx);
  return psrf_sums_rank(ctx, burnin, row_start, end, delta, method, shard, S);
}

int32_t asm_psrf_finish_device(asm_ctx* ctx, const int64_t* d_S, int32_t n_rows, double* psrf_rows, double* psrf_mean) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))  // d_S is on the first member's GPU
    return group_run_one(ctx, 0, [=](asm_ctx* m) { return asm_psrf_finish_device(m, d_S, n_rows, psrf_rows, psrf_mean); });
  ASM_ENTER(ctx);
  if (!d_S || !psrf_mean || n_rows < 0) return asm_fail(ctx, ASM_E_INVALID, "psrf_finish: bad argument");
  return psrf_finish(ctx, d_S, n_rows, psrf_rows, psrf_mean);
}

int32_t asm_seq_sum(asm_ctx* ctx, const double* values, int64_t n, double* sum_out) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx)) return group_run_one(ctx, 0, [=](asm_ctx* m) { return asm_seq_sum(m, values, n, sum_out); });
  ASM_ENTER(ctx);
  return seq_sum(ctx, values, n, sum_out);
}

// psrf_mean == NULL (members of a group other than the first): take part in the sums and the all-reduce only
static int32_t psrf_eval_rank(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                              int32_t method, double* psrf_rows, double* psrf_mean) {
  if (!burnin) return asm_fail(ctx, ASM_E_INVALID, "psrf_eval: null argument");
  asm_shard all{0, 1};
  int32_t rc = psrf_sums_device_rank(ctx, burnin, row_start, end, delta, method, all, nullptr, false, /*verify_ids=*/false);
  if (rc != ASM_OK) return rc;
  if (!psrf_mean) {
    ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASM_OK;
  }
  return psrf_finish(ctx, (const int64_t*)ctx->S_compact.p, ctx->layout.n_rows, psrf_rows, psrf_mean);
}

int32_t asm_psrf_eval(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                      int32_t method, double* psrf_rows, double* psrf_mean) {
  ASM_NULLCHECK(ctx);
  if (!burnin || !psrf_mean) return asm_fail(ctx, ASM_E_INVALID, "psrf_eval: null argument");
  if (is_group(ctx))
    return group_run(ctx, [=](asm_ctx* m, int i) {
      int32_t rc = select_device(m);
      return rc != ASM_OK ? rc : psrf_eval_rank(m, burnin, row_start, end, delta, method, i == 0 ? psrf_rows : nullptr,
                                                i == 0 ? psrf_mean : nullptr);
    });
  ASM_ENTER(ctx);
  return psrf_eval_rank(ctx, burnin, row_start, end, delta, method, psrf_rows, psrf_mean);
}

// ---- ESS ------------------------------------------------------------------------------------------------------
// A group deals the traces to its members in contiguous blocks (the synthetic shapes cycle the autocorrelation
// with the trace index, so blocks carry the same mix of short and long integral loops); every member writes its
// own slice of the caller's output arrays.
static inline void block_of(int64_t n, int world, int rank, int64_t* lo, int64_t* hi) {
  const int64_t per = (n + world - 1) / world;
  *lo = std::min<int64_t>((int64_t)rank * per, n);
  *hi = std::min<int64_t>(*lo + per, n);
}

static bool ess_args_bad(int32_t n_traces, int64_t n_samples, int64_t stride, const void* traces, int32_t max_lag) {
  return n_traces < 0 || n_samples < 0 || stride < n_samples || (n_traces > 0 && !traces) || max_lag < 1 || max_lag > ASM_MAX_LAG;
}

int32_t asm_ess_batch_device(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* d_traces,
                             int32_t max_lag, double* act_out, double* ess_out) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))  // the pointer is on the first member's GPU: that member evaluates (asm_ess_traces_put shards)
    return group_run_one(ctx, 0, [=](asm_ctx* m) {
      return asm_ess_batch_device(m, n_traces, n_samples, stride, d_traces, max_lag, act_out, ess_out);
    });
  ASM_ENTER(ctx);
  if (ess_args_bad(n_traces, n_samples, stride, d_traces, max_lag)) return asm_fail(ctx, ASM_E_INVALID, "ess_batch: bad argument");
  if (n_traces == 0) return ASM_OK;
  return ess_batch_device(ctx, n_traces, n_samples, stride, d_traces, max_lag, act_out, ess_out);
}

static int32_t ess_batch_host_rank(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces,
                                   int32_t max_lag, double* act_out, double* ess_out) {
  if (n_traces == 0) return ASM_OK;
  // upload compactly: [n_traces][n_pad] with n_pad even so that every row is 16-byte aligned
  int64_t n_pad = (n_samples + 1) & ~int64_t(1);
  if (n_pad == 0) n_pad = 2;
  int32_t rc = asm_reserve(ctx, ctx->ess_tr, sizeof(double) * (size_t)n_traces * (size_t)n_pad);
  if (rc != ASM_OK) return rc;
  return ess_batch_host(ctx, n_traces, n_samples, stride, traces, n_pad, (double*)ctx->ess_tr.p, max_lag, act_out, ess_out);
}

int32_t asm_ess_batch(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces,
                      int32_t max_lag, double* act_out, double* ess_out) {
  ASM_NULLCHECK(ctx);
  if (ess_args_bad(n_traces, n_samples, stride, traces, max_lag)) return asm_fail(ctx, ASM_E_INVALID, "ess_batch: bad argument");
  if (is_group(ctx)) {
    const int world = (int)ctx->members.size();
    return group_run(ctx, [=](asm_ctx* m, int i) {
      int64_t lo, hi;
      block_of(n_traces, world, i, &lo, &hi);
      int32_t rc = select_device(m);
      if (rc != ASM_OK || hi == lo) return rc;
      return ess_batch_host_rank(m, (int32_t)(hi - lo), n_samples, stride, traces + (size_t)lo * stride, max_lag,
                                 act_out ? act_out + lo : nullptr, ess_out ? ess_out + lo : nullptr);
    });
  }
  ASM_ENTER(ctx);
  return ess_batch_host_rank(ctx, n_traces, n_samples, stride, traces, max_lag, act_out, ess_out);
}

// Resident form of the batch: put uploads (a group: each member its block) and keeps the traces in HBM, eval runs
// the kernels on what is resident.  A rank of a multi-process communicator holds the traces it was given; the
// caller combines the per-rank vectors with asm_comm_all_gather_f64.
static int32_t ess_traces_put_rank(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces) {
  int64_t n_pad = (n_samples + 1) & ~int64_t(1);
  if (n_pad == 0) n_pad = 2;
  ctx->ess_res_n = 0;
  if (n_traces == 0) return ASM_OK;
  int32_t rc = asm_reserve(ctx, ctx->ess_res, sizeof(double) * (size_t)n_traces * (size_t)n_pad);
  if (rc != ASM_OK) return rc;
  if (n_samples > 0)
    ASM_CUDA(ctx, cudaMemcpy2DAsync(ctx->ess_res.p, sizeof(double) * (size_t)n_pad, traces, sizeof(double) * (size_t)stride,
                                    sizeof(double) * (size_t)n_samples, (size_t)n_traces, cudaMemcpyHostToDevice, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->ess_res_n = n_traces;
  ctx->ess_res_samples = n_samples;
  ctx->ess_res_stride = n_pad;
  return ASM_OK;
}

int32_t asm_ess_traces_put(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces) {
  ASM_NULLCHECK(ctx);
  if (ess_args_bad(n_traces, n_samples, stride, traces, 1)) return asm_fail(ctx, ASM_E_INVALID, "ess_traces_put: bad argument");
  if (is_group(ctx)) {
    const int world = (int)ctx->members.size();
    return group_run(ctx, [=](asm_ctx* m, int i) {
      int64_t lo, hi;
      block_of(n_traces, world, i, &lo, &hi);
      int32_t rc = select_device(m);
      return rc != ASM_OK ? rc : ess_traces_put_rank(m, (int32_t)(hi - lo), n_samples, stride, traces + (size_t)lo * stride);
    });
  }
  ASM_ENTER(ctx);
  return ess_traces_put_rank(ctx, n_traces, n_samples, stride, traces);
}

int32_t asm_ess_traces_count(const asm_ctx* ctx, int32_t* n_traces, int64_t* n_samples) {
  if (!ctx) return ASM_E_INVALID;
  int32_t n = 0;
  int64_t s = 0;
  if (is_group(ctx)) {
    for (const asm_ctx* m : ctx->members) n += m->ess_res_n, s = std::max(s, m->ess_res_samples);
  } else {
    n = ctx->ess_res_n, s = ctx->ess_res_samples;
  }
  if (n_traces) *n_traces = n;
  if (n_samples) *n_samples = s;
  return ASM_OK;
}

int32_t asm_ess_traces_eval(asm_ctx* ctx, int32_t max_lag, double* act_out, double* ess_out) {
  ASM_NULLCHECK(ctx);
  if (max_lag < 1 || max_lag > ASM_MAX_LAG) return asm_fail(ctx, ASM_E_INVALID, "ess_traces_eval: bad max_lag %d", max_lag);
  if (is_group(ctx)) {
    std::vector<int64_t> off(ctx->members.size() + 1, 0);
    for (size_t i = 0; i < ctx->members.size(); i++) off[i + 1] = off[i] + ctx->members[i]->ess_res_n;
    return group_run(ctx, [=, &off](asm_ctx* m, int i) {
      int32_t rc = select_device(m);
      if (rc != ASM_OK || m->ess_res_n == 0) return rc;
      return ess_batch_device(m, m->ess_res_n, m->ess_res_samples, m->ess_res_stride, (const double*)m->ess_res.p, max_lag,
                              act_out ? act_out + off[(size_t)i] : nullptr, ess_out ? ess_out + off[(size_t)i] : nullptr);
    });
  }
  ASM_ENTER(ctx);
  if (ctx->ess_res_n == 0) return ASM_OK;
  return ess_batch_device(ctx, ctx->ess_res_n, ctx->ess_res_samples, ctx->ess_res_stride, (const double*)ctx->ess_res.p,
                          max_lag, act_out, ess_out);
}

int32_t asm_traces_append(asm_ctx* ctx, int32_t chain, int32_t n_cols, int64_t n_new, const double* values) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx))  // one log line per chain per check: replicated, so that the columns of asm_ess_eval can be dealt out
    return group_run(ctx, [=](asm_ctx* m, int) { return asm_traces_append(m, chain, n_cols, n_new, values); });
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
  if (is_group(ctx)) ctx = ctx->members[0];
  if (!ctx || chain < 0 || chain >= ctx->n_chains) return ASM_E_INVALID;
  if (n_samples) *n_samples = ctx->traces[chain].n;
  if (n_cols) *n_cols = ctx->traces[chain].n_cols;
  return ASM_OK;
}

int32_t asm_ess_eval(asm_ctx* ctx, const int32_t* burnin, int32_t end, const int32_t* cols, int32_t n_cols,
                     double* ess_out) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx)) {  // every member holds every log line: the selected columns are dealt out in blocks
    if (!burnin || !cols || n_cols < 0 || !ess_out || end < 0) return asm_fail(ctx, ASM_E_INVALID, "ess_eval: bad argument");
    const int world = (int)ctx->members.size();
    return group_run(ctx, [=](asm_ctx* m, int i) -> int32_t {
      int64_t lo, hi;
      block_of(n_cols, world, i, &lo, &hi);
      if (hi == lo) return ASM_OK;
      return asm_ess_eval(m, burnin, end, cols + lo, (int32_t)(hi - lo), ess_out + lo);
    });
  }
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
  ASM_NULLCHECK(ctx);
  if (is_group(ctx)) return group_run_one(ctx, 0, [=](asm_ctx* m) { return asm_trace_moments(m, end, sums_out, sumsq_out); });
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

// ---- measurement hooks (a group: every member marks / flushes / synchronises; times are the max over members) ----
int32_t asm_timer_mark(asm_ctx* ctx, int32_t slot) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx)) return group_run(ctx, [=](asm_ctx* m, int) { return asm_timer_mark(m, slot); });
  ASM_ENTER(ctx);
  if (slot < 0 || slot >= 16) return asm_fail(ctx, ASM_E_INVALID, "timer slot %d", slot);
  ASM_CUDA(ctx, cudaEventRecord(ctx->timer[slot], ctx->stream));
  return ASM_OK;
}

int32_t asm_timer_elapsed_ms(asm_ctx* ctx, int32_t from, int32_t to, float* ms) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx)) {
    if (!ms) return asm_fail(ctx, ASM_E_INVALID, "timer slots");
    std::vector<float> each(ctx->members.size(), 0.f);
    float* e = each.data();
    int32_t rc = group_run(ctx, [=](asm_ctx* m, int i) { return asm_timer_elapsed_ms(m, from, to, e + i); });
    *ms = *std::max_element(each.begin(), each.end());
    return rc;
  }
  ASM_ENTER(ctx);
  if (from < 0 || from >= 16 || to < 0 || to >= 16 || !ms) return asm_fail(ctx, ASM_E_INVALID, "timer slots");
  ASM_CUDA(ctx, cudaEventSynchronize(ctx->timer[to]));
  ASM_CUDA(ctx, cudaEventElapsedTime(ms, ctx->timer[from], ctx->timer[to]));
  return ASM_OK;
}

int32_t asm_profile_enable(asm_ctx* ctx, int32_t on) {
  if (!ctx) return ASM_E_INVALID;
  for (asm_ctx* m : ctx->members) m->profile = on != 0;
  ctx->profile = on != 0;
  return ASM_OK;
}
int32_t asm_profile_reset(asm_ctx* ctx) {
  if (!ctx) return ASM_E_INVALID;
  for (asm_ctx* m : ctx->members) asm_profile_reset(m);
  for (int i = 0; i < ASM_K_COUNT; i++) {
    ctx->prof_ms[i] = 0;
    ctx->prof_n[i] = 0;
  }
  return ASM_OK;
}
int32_t asm_profile_get(asm_ctx* ctx, int32_t family, double* total_ms, int64_t* launches) {
  if (!ctx || family < 0 || family >= ASM_K_COUNT) return ASM_E_INVALID;
  double ms = ctx->prof_ms[family];
  int64_t n = ctx->prof_n[family];
  for (const asm_ctx* m : ctx->members) {  // the slowest member's time, every member's launches
    ms = std::max(ms, m->prof_ms[family]);
    n += m->prof_n[family];
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  return ASM_OK;
}
int64_t asm_launch_count(const asm_ctx* ctx) {
  if (!ctx) return 0;
  int64_t n = ctx->launches;
  for (const asm_ctx* m : ctx->members) n += m->launches;
  return n;
}

int32_t asm_sync(asm_ctx* ctx) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx)) return group_run(ctx, [](asm_ctx* m, int) { return asm_sync(m); });
  ASM_ENTER(ctx);
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

int32_t asm_flush_l2(asm_ctx* ctx) {
  ASM_NULLCHECK(ctx);
  if (is_group(ctx)) return group_run(ctx, [](asm_ctx* m, int) { return asm_flush_l2(m); });
  ASM_ENTER(ctx);
  const size_t bytes = 256u << 20;  // 2x the 126 MB L2
  int32_t rc = asm_reserve(ctx, ctx->flush, bytes);
  if (rc != ASM_OK) return rc;
  static std::atomic<uint32_t> v{0};
  {
    LaunchScope ls(ctx, ASM_K_MISC);
    flush_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint4*)ctx->flush.p, bytes / 16, ++v);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

int64_t asm_ess_last_lag_count(const asm_ctx* ctx) {
  if (!ctx) return 0;
  int64_t n = ctx->ess_lag_slots;
  for (const asm_ctx* m : ctx->members) n += m->ess_lag_slots;
  return n;
}
int32_t asm_ess_set_adaptive(asm_ctx* ctx, int32_t on) {
  if (!ctx) return ASM_E_INVALID;
  for (asm_ctx* m : ctx->members) m->ess_adaptive = on != 0;
  ctx->ess_adaptive = on != 0;
  return ASM_OK;
}

int32_t asm_psrf_last_layout(const asm_ctx* ctx, int64_t* padded_rows, int64_t* padded_bits, int64_t* tiles_done,
                             int64_t* tiles_total) {
  if (!ctx) return ASM_E_INVALID;
  const asm_ctx* c0 = is_group(ctx) ? ctx->members[0] : ctx;
  if (padded_rows) *padded_rows = c0->layout.Tp;
  if (padded_bits) {
    *padded_bits = c0->layout.bits_used;
    if (c0->layout.bits_used < 0) {  // fp4: the K extent was found on the device by the gather (rf_f4.cu); fetch it now
      int top = 0;
      if (select_device(c0) != ASM_OK || !c0->top_word_dev.p || cudaStreamSynchronize(c0->stream) != cudaSuccess ||
          cudaMemcpy(&top, c0->top_word_dev.p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
        return ASM_E_CUDA;
      *padded_bits = (int64_t)(top + 1) * 64;
    }
  }
  if (tiles_total) *tiles_total = c0->layout.tiles_total;
  if (tiles_done) {
    *tiles_done = c0->layout.tiles_done;
    for (size_t i = 1; i < ctx->members.size(); i++) *tiles_done += ctx->members[i]->layout.tiles_done;
  }
  return ASM_OK;
}

}  // extern "C"
