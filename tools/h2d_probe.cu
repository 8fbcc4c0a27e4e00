// h2d_probe.cu -- how long consecutive pinned host->device copies of one chain's bit rows (3.5 MB at cfg2) take, one
// after the other, each waited for with an event: with and without an L2-sized write before them, from one source
// buffer or from four.  (The end-to-end bench re-uploads four chains per step; this isolates the copies.)
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/h2d_probe tools/h2d_probe.cu
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));              \
      std::exit(1);                                                              \
    }                                                                            \
  } while (0)

__global__ void fill(uint4* p, size_t n, unsigned v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(v, v, v, v);
}

static double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  const size_t bytes = argc > 1 ? (size_t)atol(argv[1]) : 3520000;
  const int n_copies = 8, reps = 30;
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaEvent_t ev;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  std::vector<void*> h(n_copies), d(n_copies);
  for (int i = 0; i < n_copies; i++) {
    CK(cudaMallocHost(&h[i], bytes));
    std::memset(h[i], i + 1, bytes);
    CK(cudaMalloc(&d[i], bytes));
  }
  void* flush;
  const size_t fb = 256u << 20;
  CK(cudaMalloc(&flush, fb));
  for (int with_flush = 0; with_flush < 2; with_flush++)
    for (int one_src = 0; one_src < 2; one_src++)
      for (int spin = 0; spin < 2; spin++) {  // spin: poll the event instead of blocking on it
        std::vector<std::vector<double>> t(n_copies);
        for (int r = 0; r < reps + 3; r++) {
          if (with_flush) fill<<<148 * 8, 256, 0, st>>>((uint4*)flush, fb / 16, (unsigned)r);
          CK(cudaStreamSynchronize(st));
          for (int i = 0; i < n_copies; i++) {
            const double t0 = now_us();
            CK(cudaMemcpyAsync(d[i], h[one_src ? 0 : i], bytes, cudaMemcpyHostToDevice, st));
            CK(cudaEventRecord(ev, st));
            if (spin) {
              while (cudaEventQuery(ev) == cudaErrorNotReady) {
              }
            } else {
              CK(cudaEventSynchronize(ev));
            }
            if (r >= 3) t[i].push_back(now_us() - t0);
          }
        }
        std::printf("{\"bytes\": %zu, \"l2_flush_before\": %d, \"one_source\": %d, \"spin\": %d, \"copy_us\": [", bytes, with_flush,
                    one_src, spin);
        for (int i = 0; i < n_copies; i++) {
          std::vector<double>& v = t[i];
          std::sort(v.begin(), v.end());
          std::printf("%s%.1f", i ? ", " : "", v[v.size() / 2]);
        }
        std::printf("]}\n");
      }
  return 0;
}
