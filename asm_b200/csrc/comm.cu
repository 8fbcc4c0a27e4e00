// comm.cu -- multi-GPU plumbing of libasmgpu.so: NCCL communicators owned by the library, and the one-process
// group (asm_create_multi) whose members each have a host thread.
//
// The reference calls its criteria from ONE thread of ONE JVM (AutoStopMCMC.java:214, 397-409), so a drop-in
// cannot ask the caller to become N processes: asm_create_multi builds N per-device contexts inside the calling
// process (ncclCommInitAll), gives each a worker thread, and every entry point of include/asmgpu.h forwards to
// them with unchanged signatures.  A launcher that does run one process per GPU (torchrun) gets the same
// collectives through asm_comm_unique_id / asm_comm_init_rank.
//
// What crosses NVLink: the int64 partial sums S (ncclAllReduce, exact in any order), the
// int32 clade counts (ncclAllReduce), the bit rows when a chain is replaced (each rank uploads 1/N of the rows
// over PCIe, one in-place ncclAllGather replicates them), and 8-byte ESS values.
//
// NCCL is loaded with dlopen on first use: a single-GPU user needs no NCCL at all, and inside a process that has
// already mapped a libnccl.so.2 (PyTorch ships one) the same copy is reused instead of a second one.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <thread>

#include "asm_internal.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  std::string err;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return handle != nullptr && err.empty(); }
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("ASM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.err = std::string("NCCL not found (dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?") + "); set ASM_NCCL_LIB";
      return;
    }
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(api.handle, name);
      if (!p && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return api;
}

#define ASM_NCCL(ctx, call)                                                                                  \
  do {                                                                                                       \
    ncclResult_t r_ = (call);                                                                                \
    if (r_ != ncclSuccess)                                                                                   \
      return asm_fail((ctx), ASM_E_CUDA, "%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r_)); \
  } while (0)

int32_t need_nccl(asm_ctx* ctx) {
  NcclApi& a = nccl();
  if (!a.ok()) return asm_fail(ctx, ASM_E_UNSUPPORTED, "%s", a.err.c_str());
  return ASM_OK;
}

}  // namespace

// ---- worker threads of a one-process group -------------------------------------------------------------------
struct WorkerPool {
  struct W {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int32_t()> job;
    bool pending = false, done = true, quit = false;
    int32_t rc = ASM_OK;
  };
  std::vector<std::unique_ptr<W>> w;

  void start(int n, const std::vector<asm_ctx*>& members) {
    for (int i = 0; i < n; i++) {
      w.emplace_back(new W());
      W* me = w.back().get();
      const int dev = members[(size_t)i]->device;
      me->th = std::thread([me, dev] {
        cudaSetDevice(dev);
        asm_bind_thread_near_device(dev);  // the staging this thread allocates and the copies it drives stay on the GPU's node
        std::unique_lock<std::mutex> lk(me->mu);
        for (;;) {
          me->cv.wait(lk, [me] { return me->pending || me->quit; });
          if (me->quit) return;
          std::function<int32_t()> job = std::move(me->job);
          me->pending = false;
          lk.unlock();
          const int32_t rc = job();
          lk.lock();
          me->rc = rc;
          me->done = true;
          me->cv.notify_all();
        }
      });
    }
  }
  void post(int i, std::function<int32_t()> job) {
    W* me = w[(size_t)i].get();
    std::lock_guard<std::mutex> lk(me->mu);
    me->job = std::move(job);
    me->pending = true;
    me->done = false;
    me->cv.notify_all();
  }
  int32_t wait(int i) {
    W* me = w[(size_t)i].get();
    std::unique_lock<std::mutex> lk(me->mu);
    me->cv.wait(lk, [me] { return me->done; });
    return me->rc;
  }
  void stop() {
    for (auto& p : w) {
      {
        std::lock_guard<std::mutex> lk(p->mu);
        p->quit = true;
        p->cv.notify_all();
      }
      if (p->th.joinable()) p->th.join();
    }
    w.clear();
  }
};

int32_t group_run(asm_ctx* parent, const std::function<int32_t(asm_ctx*, int)>& fn) {
  const int n = (int)parent->members.size();
  for (int i = 0; i < n; i++) {
    asm_ctx* m = parent->members[(size_t)i];
    parent->pool->post(i, [&fn, m, i] { return fn(m, i); });
  }
  int32_t first = ASM_OK;
  for (int i = 0; i < n; i++) {
    const int32_t rc = parent->pool->wait(i);
    if (rc != ASM_OK && first == ASM_OK) {
      first = rc;
      parent->err = "gpu " + std::to_string(parent->members[(size_t)i]->device) + ": " + parent->members[(size_t)i]->err;
    }
  }
  return first;
}

int32_t group_run_one(asm_ctx* parent, int index, const std::function<int32_t(asm_ctx*)>& fn) {
  asm_ctx* m = parent->members[(size_t)index];
  parent->pool->post(index, [&fn, m] { return fn(m); });
  const int32_t rc = parent->pool->wait(index);
  if (rc != ASM_OK) parent->err = "gpu " + std::to_string(m->device) + ": " + m->err;
  return rc;
}

void comm_destroy(asm_ctx* ctx) {
  if (ctx->comm && nccl().ok()) nccl().CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->rank = 0;
  ctx->world = 1;
}

void group_destroy(asm_ctx* parent) {
  if (parent->pool) {
    // every member tears itself down on its own thread (its device is current there)
    group_run(parent, [](asm_ctx* m, int) {
      m->parent = nullptr;
      return asm_destroy(m);
    });
    parent->pool->stop();
    delete parent->pool;
    parent->pool = nullptr;
  }
  parent->members.clear();
}

// ---- collectives on the context's stream ----------------------------------------------------------------------
int32_t comm_all_reduce_i64(asm_ctx* ctx, int64_t* d_buf, size_t count) {
  if (ctx->world == 1 || count == 0) return ASM_OK;
  ASM_NCCL(ctx, nccl().AllReduce(d_buf, d_buf, count, ncclInt64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return ASM_OK;
}
int32_t comm_all_reduce_i32(asm_ctx* ctx, int32_t* d_buf, size_t count) {
  if (ctx->world == 1 || count == 0) return ASM_OK;
  ASM_NCCL(ctx, nccl().AllReduce(d_buf, d_buf, count, ncclInt32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return ASM_OK;
}
int32_t comm_all_gather_u64_inplace(asm_ctx* ctx, uint64_t* d_buf, size_t count_per_rank) {
  if (ctx->world == 1 || count_per_rank == 0) return ASM_OK;
  ASM_NCCL(ctx, nccl().AllGather(d_buf + (size_t)ctx->rank * count_per_rank, d_buf, count_per_rank, ncclUint64,
                                 (ncclComm_t)ctx->comm, ctx->stream));
  return ASM_OK;
}
int32_t comm_broadcast_u64(asm_ctx* ctx, uint64_t* d_buf, size_t count, int root) {
  if (ctx->world == 1 || count == 0) return ASM_OK;
  ASM_NCCL(ctx, nccl().Broadcast(d_buf, d_buf, count, ncclUint64, root, (ncclComm_t)ctx->comm, ctx->stream));
  return ASM_OK;
}
int32_t comm_all_gather_f64(asm_ctx* ctx, const double* d_send, double* d_recv, size_t count_per_rank) {
  if (count_per_rank == 0) return ASM_OK;
  if (ctx->world == 1) {
    ASM_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, sizeof(double) * count_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
    return ASM_OK;
  }
  ASM_NCCL(ctx, nccl().AllGather(d_send, d_recv, count_per_rank, ncclFloat64, (ncclComm_t)ctx->comm, ctx->stream));
  return ASM_OK;
}

// ---- C ABI ----------------------------------------------------------------------------------------------------
extern "C" {

int32_t asm_create_multi(asm_ctx** out, const int32_t* devices, int32_t n_devices, int32_t n_chains) {
  if (!out) return ASM_E_INVALID;
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 64)
    return asm_fail(nullptr, ASM_E_INVALID, "asm_create_multi: n_devices must be in [1,64], not %d", n_devices);
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) return asm_fail(nullptr, ASM_E_INVALID, "asm_create_multi: device %d listed twice", devices[i]);
  if (n_devices > 1 && !nccl().ok()) return asm_fail(nullptr, ASM_E_UNSUPPORTED, "%s", nccl().err.c_str());
  auto* parent = new (std::nothrow) asm_ctx();
  if (!parent) return ASM_E_NOMEM;
  parent->device = devices[0];
  parent->n_chains = n_chains;
  parent->world = n_devices;
  auto fail = [&](int32_t rc) {
    for (asm_ctx* m : parent->members) asm_destroy(m);
    delete parent;
    return rc;
  };
  for (int i = 0; i < n_devices; i++) {
    asm_ctx* m = nullptr;
    const int32_t rc = asm_create(&m, devices[i], n_chains);  // leaves its message in asm_last_error(NULL)
    if (rc != ASM_OK) return fail(rc);
    m->rank = i;
    m->world = n_devices;
    parent->members.push_back(m);
  }
  parent->sm_count = parent->members[0]->sm_count;
  if (n_devices > 1) {
    std::vector<ncclComm_t> comms((size_t)n_devices);
    std::vector<int> devs(devices, devices + n_devices);
    const ncclResult_t r = nccl().CommInitAll(comms.data(), n_devices, devs.data());
    if (r != ncclSuccess)
      return fail(asm_fail(nullptr, ASM_E_CUDA, "ncclCommInitAll(%d devices): %s", n_devices, nccl().GetErrorString(r)));
    for (int i = 0; i < n_devices; i++) parent->members[(size_t)i]->comm = comms[(size_t)i];
  }
  for (asm_ctx* m : parent->members) m->parent = parent;
  parent->pool = new WorkerPool();
  parent->pool->start(n_devices, parent->members);
  *out = parent;
  return ASM_OK;
}

int32_t asm_comm_unique_id(uint8_t* id_out) {
  if (!id_out) return ASM_E_INVALID;
  if (!nccl().ok()) return asm_fail(nullptr, ASM_E_UNSUPPORTED, "%s", nccl().err.c_str());
  static_assert(sizeof(ncclUniqueId) == ASM_COMM_ID_BYTES, "ASM_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
  ncclUniqueId id;
  const ncclResult_t r = nccl().GetUniqueId(&id);
  if (r != ncclSuccess) return asm_fail(nullptr, ASM_E_CUDA, "ncclGetUniqueId: %s", nccl().GetErrorString(r));
  std::memcpy(id_out, &id, sizeof id);
  return ASM_OK;
}

int32_t asm_comm_init_rank(asm_ctx* ctx, const uint8_t* id, int32_t rank, int32_t n_ranks) {
  if (!ctx) return ASM_E_INVALID;
  if (!ctx->members.empty() || ctx->parent || ctx->comm)
    return asm_fail(ctx, ASM_E_INVALID, "asm_comm_init_rank: the context already belongs to a communicator or group");
  if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return asm_fail(ctx, ASM_E_INVALID, "asm_comm_init_rank: bad rank %d of %d", rank, n_ranks);
  if (n_ranks == 1) return ASM_OK;
  int32_t rc = need_nccl(ctx);
  if (rc != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof uid);
  ncclComm_t c = nullptr;
  ASM_NCCL(ctx, nccl().CommInitRank(&c, n_ranks, uid, rank));
  ctx->comm = c;
  ctx->rank = rank;
  ctx->world = n_ranks;
  return ASM_OK;
}

int32_t asm_comm_info(const asm_ctx* ctx, int32_t* rank, int32_t* n_ranks, int32_t* n_local_devices) {
  if (!ctx) return ASM_E_INVALID;
  const bool group = !ctx->members.empty();
  if (rank) *rank = group ? 0 : ctx->rank;
  if (n_ranks) *n_ranks = ctx->world;
  if (n_local_devices) *n_local_devices = group ? (int32_t)ctx->members.size() : 1;
  return ASM_OK;
}

int32_t asm_comm_nccl_version(void) {
  int v = 0;
  if (!nccl().ok() || nccl().GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

int32_t asm_comm_all_gather_f64(asm_ctx* ctx, const double* local, int64_t n_local, double* all) {
  if (!ctx || n_local < 0 || (n_local > 0 && (!local || !all))) return ASM_E_INVALID;
  if (!ctx->members.empty()) {  // a one-process group holds everything already
    if (n_local > 0 && all != local) std::memmove(all, local, sizeof(double) * (size_t)n_local);
    return ASM_OK;
  }
  if (n_local == 0) return ASM_OK;
  ASM_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)n_local;
  int32_t rc = asm_reserve(ctx, ctx->misc, sizeof(double) * n * (size_t)(ctx->world + 1));
  if (rc != ASM_OK) return rc;
  double* d_send = (double*)ctx->misc.p;
  double* d_recv = d_send + n;
  ASM_CUDA(ctx, cudaMemcpyAsync(d_send, local, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = comm_all_gather_f64(ctx, d_send, d_recv, n)) != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemcpyAsync(all, d_recv, sizeof(double) * n * (size_t)ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

}  // extern "C"
