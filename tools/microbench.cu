// microbench.cu -- issue-rate microbenchmarks for the roofline denominators of the non-HBM-bound
// kernels ("must be measured"): POPC, LOP3, IMAD, REDUX, DADD/DMUL/DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CHAINS 8
template <int OP>
__global__ void __launch_bounds__(256) k(uint64_t* out, int iters, uint32_t seed, double dseed) {
  uint32_t x[CHAINS];
  double d[CHAINS];
  for (int c = 0; c < CHAINS; c++) { x[c] = seed + threadIdx.x * 7919u + c * 104729u; d[c] = dseed + c + threadIdx.x * 1e-3; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      if (OP == 0) x[c] = __popc(x[c]) + seed;                         // POPC (+IADD)
      if (OP == 1) x[c] = (x[c] ^ seed) & (x[c] >> 1 | seed);          // LOP3 (+SHF)
      if (OP == 2) x[c] = x[c] * seed + 12345u;                        // IMAD
      if (OP == 3) x[c] = __reduce_add_sync(0xffffffffu, x[c]) + seed; // REDUX
      if (OP == 4) d[c] = __dadd_rn(d[c], dseed);                      // DADD
      if (OP == 5) d[c] = __dmul_rn(d[c], dseed);                      // DMUL
      if (OP == 6) d[c] = __fma_rn(d[c], dseed, dseed);                // DFMA
      if (OP == 7) d[c] = __dadd_rn(d[c], __dmul_rn(d[(c + 1) % CHAINS], dseed));  // DMUL+DADD pair
      if (OP == 8) x[c] = __popc(x[c] ^ seed) + __popc(x[(c + 1) % CHAINS]);       // 2 POPC + LOP3 + IADD
    }
  }
  uint64_t s = 0;
  for (int c = 0; c < CHAINS; c++) s += x[c] + (uint64_t)__double_as_longlong(d[c]);
  if (s == 0x1234567ULL) out[0] = s;
}

template <int OP>
double run(const char* name, double ops_per_iter_chain, int sms) {
  uint64_t* d; cudaMalloc(&d, 8);
  int iters = 20000, blocks = sms * 8;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<blocks, 256>>>(d, 100, 3u, 1.0000001);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<OP><<<blocks, 256>>>(d, iters, 3u, 1.0000001);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)blocks * 256 * iters * CHAINS * ops_per_iter_chain;
  double rate = ops / (ms * 1e-3);
  printf("  \"%s\": {\"ops_per_s\": %.4e, \"ms\": %.3f, \"per_sm_per_ns\": %.3f},\n", name, rate, ms, rate / sms / 1e9);
  cudaFree(d);
  return rate;
}


// Dependent-issue latency of FP64 adds: one warp, C independent accumulator chains, each step one DADD per chain
// (MUL = 1: the addend is a fresh DMUL, as in the ESS lag kernel).  Reports clocks per step of one warp.
template <int C, int MUL>
__global__ void lat_k(double* out, long long* clk, int iters, double dseed) {
  double a[C], m[C];
  for (int c = 0; c < C; c++) { a[c] = dseed + c; m[c] = dseed * (c + 2) + threadIdx.x; }
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < C; c++) {
      if (MUL) { a[c] = __dadd_rn(a[c], m[c]); m[c] = __dmul_rn(m[c], dseed); } else a[c] = __dadd_rn(a[c], m[c]);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < C; c++) s += a[c] + m[c];
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  if (s == 1.2345) out[0] = s;
}
template <int C, int MUL>
void run_lat(const char* name, int warps) {
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
  int iters = 100000;
  lat_k<C, MUL><<<1, 32 * warps>>>(d, c, 100, 1.0000001);
  lat_k<C, MUL><<<1, 32 * warps>>>(d, c, iters, 1.0000001);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("  \"%s\": {\"chains\": %d, \"warps_in_cta\": %d, \"clk_per_step\": %.2f},\n", name, C, warps, (double)h / iters);
  cudaFree(d); cudaFree(c);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, p.multiProcessorCount, p.clockRate);
  int sms = p.multiProcessorCount;
  run<0>("popc", 1, sms);
  run<1>("lop3", 2, sms);
  run<2>("imad", 1, sms);
  run<3>("redux_warp_instr_x32", 1, sms);
  run<4>("dadd", 1, sms);
  run<5>("dmul", 1, sms);
  run<6>("dfma", 1, sms);
  run<7>("dmul_dadd_pair_instrs", 2, sms);
  run<8>("popc2_mix", 2, sms);
  run_lat<1, 0>("dadd_dep_1chain", 1);
  run_lat<2, 0>("dadd_dep_2chain", 1);
  run_lat<4, 0>("dadd_dep_4chain", 1);
  run_lat<8, 0>("dadd_dep_8chain", 1);
  run_lat<1, 1>("dmul_dadd_dep_1chain", 1);
  run_lat<2, 1>("dmul_dadd_dep_2chain", 1);
  run_lat<4, 1>("dmul_dadd_dep_4chain", 1);
  run_lat<8, 1>("dmul_dadd_dep_8chain", 1);
  run_lat<4, 1>("dmul_dadd_dep_4chain_4warps", 4);
  run_lat<4, 1>("dmul_dadd_dep_4chain_8warps", 8);
  run_lat<8, 1>("dmul_dadd_dep_8chain_4warps", 4);
  run_lat<8, 1>("dmul_dadd_dep_8chain_8warps", 8);
  printf("  \"note\": \"lane-ops per second; per_sm_per_ns ~ lane-ops per clock per SM at ~1 GHz (divide by the clock in GHz)\"\n}\n");
  return 0;
}
