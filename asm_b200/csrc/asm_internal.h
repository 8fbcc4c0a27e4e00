// asm_internal.h -- shared between the translation units of libasmgpu.so.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/asmgpu.h"

// ---- layout constants shared by host and kernels -------------------------------------------------
constexpr int kTileRows = 128;     // rows of one distance tile (both kernels)
constexpr int kBlockRows = 256;    // chain segments are padded to this many rows
constexpr int kBitsPad = 1024;     // dictionary width is padded to this many bits (16 words = one
                                   // 128-byte K chunk of the popcount kernel; 8 K-chunks of the int8 one)
constexpr int kMaxSegments = 2 * ASM_MAX_CHAINS;

// flags of a 128-row tile / 256-row block of the padded, chain-major operand matrix
constexpr uint32_t kFlagCol = 1;   // its trees are columns (post-burn-in) of their chain
constexpr uint32_t kFlagRows = 2;  // it contains at least one row (tree index >= row_start)

struct Segment {        // a run of thinned trees of one chain in the padded operand
  int64_t dst_off;      // first padded row
  int64_t n;            // live rows (the rest up to the next segment is zero padding)
  int64_t src_first;    // tree index of the first live row
  int32_t chain;
  int32_t is_col;       // 1: B segment (columns), 0: A segment (burn-in trees that are rows only)
};

struct PsrfLayout {
  int32_t C = 0, delta = 1, row_start = 0, end = 0;
  int32_t n_rows = 0;             // output rows per chain
  int32_t n_seg = 0;
  Segment seg[kMaxSegments];
  int64_t Tp = 0;                 // padded rows, multiple of 256
  int32_t Wp = 0;                 // padded words per row, multiple of 16 (popcount operand)
  int32_t pitch8 = 0;             // bytes per row of the int8 operand, multiple of 128
  int32_t pitch4 = 0;             // bytes per row of the fp4 operand (two clades per byte), multiple of 128
  int32_t block_real = 256;       // live rows per 256-row block (always 256 since round 2: segments are dense)
  int32_t col_block = 256;        // trees per column block: 256, or 240 for the fp4 kernel (N = 240 tiles over the SAME
                                  // dense rows: a column block starts at any row, TMA boxes need no alignment)
  // padded position of output row r of chain a: r < rowsA[a] ? segA_off[a] + rowA0[a] + r
  //                                                          : segB_off[a] + rowB0[a] + r - rowsA[a]
  int64_t segA_off[ASM_MAX_CHAINS], segB_off[ASM_MAX_CHAINS];
  int64_t rowA0[ASM_MAX_CHAINS], rowB0[ASM_MAX_CHAINS];
  int32_t rowsA[ASM_MAX_CHAINS];
  // zero rows that the kernels see beside the column trees of chain c: padC[c] when those trees are the columns of a
  // tile (column blocks of col_block trees), padR[c] when they are its rows (256-row blocks) and the sums reach the
  // other operand through the symmetric half.  Removed in closed form by the compaction kernel.
  int64_t padC[ASM_MAX_CHAINS], padR[ASM_MAX_CHAINS];
  int64_t tiles_total = 0, tiles_done = 0;
  int64_t bits_used = 0;          // operand width (bits) of the kernel that ran last
};

struct ChainTrees {
  uint64_t* d_bits = nullptr;  // [cap][ctx->words]
  int64_t n = 0, cap = 0;
};

struct ChainTraces {
  double* d_vals = nullptr;  // row-major [cap][n_cols]
  int64_t n = 0, cap = 0;
  int32_t n_cols = 0;
};

constexpr int kF4CounterOff = 8;

struct DevBuf {  // grow-only device scratch
  void* p = nullptr;
  size_t bytes = 0;
};

struct asm_ctx {
  int device = 0;
  int n_chains = 0;
  int sm_count = 0;
  int smem_optin = 0;  // cudaDevAttrMaxSharedMemoryPerBlockOptin
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream (ESS sum chain)
  cudaStream_t stream_copy = nullptr;  // H2D copies overlapped with compute (asm_ess_batch)
  static constexpr int kMaxGroups = 8;
  cudaStream_t group_stream[kMaxGroups] = {}, group_stream2[kMaxGroups] = {};   // ESS chunk groups
  cudaEvent_t group_event[kMaxGroups] = {}, group_fork[kMaxGroups] = {}, group_join[kMaxGroups] = {};
  void* h_pinned = nullptr;            // small page-locked scratch for asynchronous counters
  DevBuf top_word_dev;                 // fp4 path scalars: [0] highest non-zero word of the gathered rows (gather_f4_kernel ->
                                       // rf_f4x2_kernel), [kF4CounterOff] the contraction's work-slot counter
  DevBuf ids_stage[2], ids_flag;       // clade ids on their way to bit rows (asm_trees_set_ids), double-buffered; "an id was out of range"
  int ids_slot = 0;
  cudaEvent_t ev_ids_used[2] = {};     // the expansion kernel has read ids_stage[slot]
  bool ids_check_pending = false;      // ids_flag has not been looked at since the last ids upload (trees_ids_verify)
  cudaEvent_t ev_ids = nullptr;        // the copy of the caller's ids has run
  void* h_stage = nullptr;             // page-locked staging of the per-step block tables
  size_t h_stage_bytes = 0;
  cudaEvent_t ev_stage = nullptr;      // the copy out of h_stage has run
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int32_t words = 0;  // stored row width of every chain (uint64 words)
  std::vector<ChainTrees> trees;
  std::vector<ChainTraces> traces;
  // scratch
  DevBuf opnd_bits, opnd_i8, opnd_f4, nbits, S_pad, blk_meta, S_compact, psrf_rows, psrf_vals, psrf_mean, misc, flush, ess_tr,
      ess_sls, ess_out, tile_counter, rf_out;
  PsrfLayout layout;
  std::vector<int2> meta_host;  // staging for tile metadata
  std::vector<uint32_t> walk_host;  // Morton tile walk of the int8 kernel (rebuilt when the block count changes)
  DevBuf walk_dev;
  long long walk_nb = -1;
  // fp4 kernel: row blocks (256 trees) and column blocks (240 trees) over the same dense operand rows
  std::vector<int2> f4_blk_host;   // [nR row blocks][nC column blocks]: {first padded row, chain | flags << 8}
  DevBuf f4_blk_dev;
  int32_t f4_nR = 0, f4_nC = 0;
  std::vector<uint32_t> f4_walk_host;
  std::vector<long long> f4_walk_key;  // first padded rows of the blocks the cached walk was built for
  DevBuf f4_walk_dev;
  uint32_t func_attr_done = 0;  // bit per kernel whose max-dynamic-smem attribute is set on this device
  // measurement
  cudaEvent_t timer[16] = {};
  bool profile = false;
  cudaEvent_t pe0 = nullptr, pe1 = nullptr;
  double prof_ms[ASM_K_COUNT] = {};
  int64_t prof_n[ASM_K_COUNT] = {};
  int64_t launches = 0;
  bool ess_adaptive = true;     // staged lag evaluation (ess.cu)
  long long ess_lag_slots = 0;  // (trace, lag) chains evaluated by the last ESS call
  // resident batch traces (asm_ess_traces_put): [ess_res_n][ess_res_stride] doubles
  DevBuf ess_res, ess_park;
  int32_t ess_res_n = 0;
  int64_t ess_res_samples = 0, ess_res_stride = 0;
  // ---- multi-GPU (comm.cu) ------------------------------------------------------------------------------
  // A context is a single GPU (rank 0 of 1), one rank of an NCCL communicator shared with other processes
  // (asm_comm_init_rank), a member of a one-process group (asm_create_multi), or the parent of such a group
  // (members non-empty; the parent owns no device state and forwards every call to its members' threads).
  void* comm = nullptr;  // ncclComm_t
  int rank = 0, world = 1;
  asm_ctx* parent = nullptr;
  std::vector<asm_ctx*> members;
  struct WorkerPool* pool = nullptr;
  std::string err;
};

constexpr uint32_t kAttrPopc = 1, kAttrI8 = 2, kAttrI8Dump = 4, kAttrI8x2 = 8, kAttrEss = 16, kAttrF4x2 = 32, kAttrEssTicket = 64, kAttrIdsToBits = 128;

// ---- error plumbing ---------------------------------------------------------------------------------
int32_t asm_fail(asm_ctx* ctx, int32_t code, const char* fmt, ...);
#define ASM_CUDA(ctx, call)                                                                          \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return asm_fail((ctx), ASM_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

int32_t asm_reserve(asm_ctx* ctx, DevBuf& b, size_t bytes);

// profile bracket around one launch
struct LaunchScope {
  asm_ctx* ctx;
  int family;
  cudaStream_t st;
  LaunchScope(asm_ctx* c, int f, cudaStream_t s = nullptr);
  ~LaunchScope();
};

// ---- kernels (launchers live next to their kernels) ---------------------------------------------------
// psrf_layout.cu
int32_t psrf_build_layout(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                          int32_t col_block = kBlockRows);
int32_t psrf_gather_bits(asm_ctx* ctx);  // layout -> opnd_bits [Tp][Wp], nbits [Tp], blk_meta
int32_t psrf_gather_i8(asm_ctx* ctx);    // layout -> opnd_i8 [Tp][pitch8] directly, nbits [Tp], blk_meta
int32_t psrf_gather_f4(asm_ctx* ctx);    // layout (col_block = 240) -> opnd_f4 [Tp][pitch4] e2m1 nibbles, nbits, f4 block tables
int32_t psrf_check_max_bits(asm_ctx* ctx, int32_t limit);  // ASM_E_UNSUPPORTED if a gathered row has more set bits
int32_t psrf_compact(asm_ctx* ctx, asm_shard shard, int64_t* d_S_out);  // S_pad -> [C][n_rows][C], pad-corrected
int32_t psrf_finish(asm_ctx* ctx, const int64_t* d_S, int32_t n_rows, double* psrf_rows_host, double* psrf_mean_host);
int32_t seq_sum(asm_ctx* ctx, const double* values_host, int64_t n, double* sum_host);  // left-to-right sum, seq_sum.cuh
int32_t trees_ids_verify(asm_ctx* ctx);  // asmgpu.cu: ASM_E_RANGE if an ids upload since the last check named a clade outside its dictionary
// rf_popc.cu
int32_t rf_popc_sums(asm_ctx* ctx, asm_shard shard);
int32_t rf_popc_block(asm_ctx* ctx, int32_t chainA, int64_t r0, int64_t r1, int32_t chainB, int64_t c0, int64_t c1,
                      uint16_t* out_host);
// rf_i8.cu / rf_f4.cu (tc_common.cuh)
int32_t rf_i8_sums(asm_ctx* ctx, asm_shard shard);
int32_t rf_f4_sums(asm_ctx* ctx, asm_shard shard);  // block-scaled FP4 variant (layout with 240 live rows per block)
int32_t rf_i8_block(asm_ctx* ctx, int32_t chainA, int64_t r0, int64_t r1, int32_t chainB, int64_t c0, int64_t c1,
                    uint16_t* out_host);
// clade.cu (with a communicator: this rank counts its share of [from, to), an int32 all-reduce adds the shares;
// out_host may be NULL)
int32_t clade_counts(asm_ctx* ctx, int32_t chain, int64_t from, int64_t to, int32_t* out_host);
// comm.cu -- collectives on the context's own stream (NCCL, loaded on first use); no-ops when world == 1
int32_t comm_all_reduce_i64(asm_ctx* ctx, int64_t* d_buf, size_t count);
int32_t comm_all_reduce_i32(asm_ctx* ctx, int32_t* d_buf, size_t count);
int32_t comm_all_gather_u64_inplace(asm_ctx* ctx, uint64_t* d_buf, size_t count_per_rank);  // rank r's part at d_buf + r*count
int32_t comm_broadcast_u64(asm_ctx* ctx, uint64_t* d_buf, size_t count, int root);
int32_t comm_all_gather_f64(asm_ctx* ctx, const double* d_send, double* d_recv, size_t count_per_rank);
void comm_destroy(asm_ctx* ctx);
// run fn(member, index) on every member's thread and wait; returns the first failure (message copied to the parent)
int32_t group_run(asm_ctx* parent, const std::function<int32_t(asm_ctx*, int)>& fn);
int32_t group_run_one(asm_ctx* parent, int index, const std::function<int32_t(asm_ctx*)>& fn);
void group_destroy(asm_ctx* parent);
// moments.cu
int32_t trace_moments(asm_ctx* ctx, int32_t end, double* sums_host, double* sumsq_host);
// ess.cu
int32_t ess_batch_device(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* d_traces,
                         int32_t max_lag, double* act_out_host, double* ess_out_host);
int32_t ess_batch_host(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces,
                       int64_t n_pad, double* d_tr, int32_t max_lag, double* act_out_host, double* ess_out_host);
int32_t ess_eval_stream(asm_ctx* ctx, const int32_t* burnin, int32_t end, const int32_t* cols, int32_t n_cols,
                        double* ess_out_host);
