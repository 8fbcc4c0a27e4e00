// fp4_probe.cu -- is the block-scaled FP4 tensor-core path exact on 0/1 clade indicators, and how fast is it
// next to kind::i8?  One CTA, cta_group::1, M=128, N=240.  Operands: 0/1 bits as e2m1 nibbles (1.0 = 0x2),
// unit scale factors (UE8M0 0x7F in every byte of the scale-factor columns), f32 accumulators: the dot product
// of two bit rows is an integer < 2^24, so the result is exact.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fp4_probe tools/fp4_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kM = 128, kN = 240, kNB = 256, kKBytes = 128;  // one K block: 128 bytes = 256 fp4 or 128 int8
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mma_i8(uint32_t c, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(c), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_mxf4(uint32_t c, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(c), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}
// byte (row r, byte kb of the 128-byte K block) of a K-major tile with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128(int r, int kb) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((kb >> 4) ^ (r & 7)) << 4) | (kb & 15))); }

// mode 0: kind::i8 (bits as bytes, K = 128 per block), mode 1: kind::mxf4 (bits as nibbles, K = 256 per block)
__global__ void __launch_bounds__(128) probe(const uint8_t* __restrict__ a_bits, const uint8_t* __restrict__ b_bits, int mode,
                                             int reps, uint32_t* __restrict__ out, long long* __restrict__ clk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sa = smem;                 // 128 x 128 B
  unsigned char* sb = smem + kM * kKBytes;  // 256 x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kelems = mode ? 256 : 128;
  // fill the tiles
  for (int idx = threadIdx.x; idx < (kM + kNB) * kKBytes; idx += blockDim.x) {
    const int r = idx / kKBytes, kb = idx % kKBytes;
    const bool isb = r >= kM;
    const int rr = isb ? r - kM : r;
    const uint8_t* src = (isb ? b_bits : a_bits) + (size_t)rr * 256;  // one byte per bit, 256 bits per row
    uint8_t v;
    if (mode) v = (uint8_t)((src[2 * kb] ? 0x2 : 0) | (src[2 * kb + 1] ? 0x20 : 0));
    else v = src[kb];
    (isb ? sb : sa)[sw128(rr, kb)] = v;
  }
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  // unit scale factors: 0x7F in every byte of columns 480..511, all 128 lanes
  {
    const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16) + 480u;
    const uint32_t v = 0x7F7F7F7Fu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(t), "r"(v) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint64_t ad = make_desc(smem_u32(sa)), bd = make_desc(smem_u32(sb));
    // kind::i8: c_format S32 = 2 @ [4,6), a/b S8 = 1 @ [7,10)/[10,13), N>>3 @ [17,23), M>>4 @ [24,29)
    const uint32_t idesc_i8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    // kind::mxf4 (InstrDescriptorBlockScaled): a/b E2M1 = 1 @ [7,10)/[10,13), N>>3 @ [17,23), scale UE8M0 = 1 @ 23,
    // M>>4 @ [24,29), scale-factor ids 0, K = 64
    const uint32_t idesc_f4 = (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | (1u << 23) | ((uint32_t)(kM >> 4) << 24);
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; rep++) {
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const uint32_t acc = (rep | kk) != 0 ? 1u : 0u;
        if (mode) mma_mxf4(tmem, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc_f4, acc, tmem + 480u, tmem + 496u);
        else mma_i8(tmem, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc_i8, acc);
      }
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    clk[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read the 240 accumulator columns of this warp's 32 lanes
  for (int c0 = 0; c0 < kN; c0 += 16) {
    uint32_t r[16];
    const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(t));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; j++) out[(size_t)(warp * 32 + lane) * kN + c0 + j] = r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  std::vector<uint8_t> a(kM * 256), b(kNB * 256);
  srand(7);
  for (auto& v : a) v = (rand() % 3) == 0;
  for (auto& v : b) v = (rand() % 3) == 0;
  uint8_t *da, *db;
  uint32_t* dout;
  long long* dclk;
  cudaMalloc(&da, a.size()); cudaMalloc(&db, b.size()); cudaMalloc(&dout, kM * kN * 4); cudaMalloc(&dclk, 8);
  cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
  const size_t smem = (kM + kNB) * kKBytes + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<uint32_t> out(kM * kN);
  printf("{\n");
  for (int mode = 0; mode < 2; mode++) {
    const int kelems = mode ? 256 : 128;
    probe<<<1, 128, smem>>>(da, db, mode, 1, dout, dclk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" \"error_mode%d\": \"%s\"\n}\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    long long bad = 0, first = -1;
    for (int r = 0; r < kM; r++)
      for (int c = 0; c < kN; c++) {
        int want = 0;
        for (int k = 0; k < kelems; k++) want += a[r * 256 + k] & b[c * 256 + k];
        float gotf; uint32_t u = out[r * kN + c];
        memcpy(&gotf, &u, 4);
        const long long got = mode ? (long long)gotf : (long long)(int32_t)u;
        if (got != want) { if (first < 0) first = r * kN + c; bad++; }
      }
    const int reps = 4000;
    probe<<<1, 128, smem>>>(da, db, mode, reps, dout, dclk);
    cudaDeviceSynchronize();
    long long clk;
    cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
    const double per_mma = (double)clk / (reps * 4.0);
    printf(" \"%s\": {\"mismatches\": %lld, \"first_bad\": %lld, \"out00\": %u, \"clk_per_mma\": %.1f, \"k_per_mma\": %d, \"flop_per_clk_per_sm\": %.0f},\n",
           mode ? "mxf4" : "i8", bad, first, out[0], per_mma, mode ? 64 : 32, 2.0 * kM * kN * (mode ? 64 : 32) / per_mma);
  }
  printf(" \"shape\": \"M=128 N=240 cta_group::1, one K block, accumulators in TMEM\"\n}\n");
  return 0;
}
