// Microbenchmark: cycles per tcgen05.mma (SS operands, SWIZZLE_NONE K-major canonical layout) for
// kind::tf32 (K=8) and kind::f16/bf16 (K=16), M=128, several N; one CTA per SM, one issuing thread.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../nequip_b200/csrc/nqb_tc.cuh"

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// MODE 0: tf32, 1: bf16.  lbo/sbo: descriptor strides in bytes.  spread: operand address varies per MMA.
// pattern: 0 = alternate two accumulators, 1 = ONE accumulator (dependent chain), 2 = kernel pattern
// (4 MMAs into acc0 then 8 into acc1), 3 = four accumulators round robin.  noise: the other three warps
// hammer shared memory with 16-byte stores while the MMAs run.
__global__ void __launch_bounds__(128, 1) k2(long long* out, int nmma, int N, int pattern, int noise) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t tbase;
  __shared__ volatile int stop;
  const int nonzero = noise & 2, commits = noise & 4;
  noise &= 1;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    ((float*)sm)[i] = nonzero ? (float)((i * 2654435761u) >> 8 & 0xffff) * 1.0e-4f - 3.0f : 0.f;
  if (threadIdx.x == 0) for (int q = 0; q < 4; ++q) mbar_init(&bar2[q], 1);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 64 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const uint32_t off = (uint32_t)((i & 3) * 256);
      const uint64_t da = make_desc(a0 + off, 128, 1024), db = make_desc(b0 + off, 128, 1024);
      uint32_t acc;
      if (pattern == 0) acc = i & 1;
      else if (pattern == 1) acc = 0;
      else if (pattern == 2) acc = ((i % 12) < 4) ? 0 : 1;
      else if (pattern == 3) acc = i & 3;
      else if (pattern == 4) acc = (i & 1) * 2;                       // columns 0 / 256 alternating
      else if (pattern == 5) acc = ((i % 12) < 4) ? 0 : 2;            // 4x col 0 then 8x col 256
      else if (pattern == 6) acc = ((i % 3) == 0) ? 0 : 2;            // hh, x, x interleaved, cols 0 / 256
      else acc = ((i % 3) == 0) ? 0 : 1;                              // hh, x, x interleaved, cols 0 / 128
      umma_tf32(tbase + acc * 128, da, db, idesc, 1);
      if (commits && (i % 12) == 11) umma_commit(&bar2[(i / 12) & 3]);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if (noise && threadIdx.x >= 32) {
    float4* p = reinterpret_cast<float4*>(sm + 128 * 1024) + (threadIdx.x - 32);
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    while (!stop) {
#pragma unroll
      for (int r = 0; r < 16; ++r) p[r * 96] = v;
      v.x += 1.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

void run2(const char* name, int pattern, int noise, long long* d_out, int nsm) {
  const int nmma = 2048, N = 128;
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k2<<<nsm, 128, 200 * 1024>>>(d_out, nmma, N, pattern, noise);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: ERROR %s\n", name, cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, d_out, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nsm; ++i) avg += (double)h[i]; avg /= nsm;
  printf("%-52s %7.1f cycles/MMA\n", name, avg / nmma);
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int nmma, int N, uint32_t lbo, uint32_t sbo, int spread) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) ((float*)sm)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = (MODE == 0) ? make_idesc(128, N) : idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 64 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const uint32_t off = spread ? (uint32_t)((i & 3) * 256) : 0u;
      const uint64_t da = make_desc(a0 + off, lbo, sbo), db = make_desc(b0 + off, lbo, sbo);
      const uint32_t d = tbase + (uint32_t)((i & 1) * 256);
      if (MODE == 0) umma_tf32(d, da, db, idesc, i > 1);
      else umma_bf16(d, da, db, idesc, i > 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

template <int MODE> void run(const char* name, int N, uint32_t lbo, uint32_t sbo, int spread, long long* d_out, int nsm) {
  const int nmma = 2048;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  k<MODE><<<nsm, 128, 160 * 1024>>>(d_out, nmma, N, lbo, sbo, spread);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: ERROR %s\n", name, cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, d_out, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nsm; ++i) avg += (double)h[i]; avg /= nsm;
  const int K = MODE == 0 ? 8 : 16;
  printf("%-44s N=%3d  %7.1f cycles/MMA   %8.1f flop/clk/SM\n", name, N, avg / nmma, 2.0 * 128 * N * K * nmma / avg);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int nsm = p.multiProcessorCount;
  long long* d_out; cudaMalloc(&d_out, 256 * sizeof(long long));
  printf("%s, %d SMs (all SMs issue concurrently)\n", p.name, nsm);
  for (int N : {32, 64, 128, 256}) run<0>("tf32 SS no-swizzle LBO=128 SBO=1024 (chunk K=32)", N, 128, 1024, 1, d_out, nsm);
  for (int N : {32, 128}) run<0>("tf32 SS no-swizzle LBO=128 SBO=4096 (K=128)", N, 128, 4096, 1, d_out, nsm);
  for (int N : {128}) run<0>("tf32 SS no-swizzle same operands", N, 128, 1024, 0, d_out, nsm);
  for (int N : {32, 64, 128, 256}) run<1>("bf16 SS no-swizzle LBO=128 SBO=1024", N, 128, 1024, 1, d_out, nsm);
  run<0>("tf32 on ONE SM only", 128, 128, 1024, 1, d_out, 1);
  run2("tf32 N=128 two accumulators alternating", 0, 0, d_out, nsm);
  run2("tf32 N=128 ONE accumulator (dependent chain)", 1, 0, d_out, nsm);
  run2("tf32 N=128 kernel pattern 4x acc0 + 8x acc1", 2, 0, d_out, nsm);
  run2("tf32 N=128 four accumulators round robin", 3, 0, d_out, nsm);
  run2("tf32 N=128 alternating + smem store noise (3 warps)", 0, 1, d_out, nsm);
  run2("tf32 N=128 kernel pattern + smem store noise", 2, 1, d_out, nsm);
  run2("tf32 N=128 cols 0/256 alternating", 4, 0, d_out, nsm);
  run2("tf32 N=128 4x col0 then 8x col256", 5, 0, d_out, nsm);
  run2("tf32 N=128 hh,x,x interleaved cols 0/256", 6, 0, d_out, nsm);
  run2("tf32 N=128 hh,x,x interleaved cols 0/128", 7, 0, d_out, nsm);
  run2("tf32 N=128 hh,x,x interleaved cols 0/256 + commits", 6, 4, d_out, nsm);
  run2("tf32 N=128 kernel pattern, NONZERO operands", 2, 2, d_out, nsm);
  run2("tf32 N=128 kernel pattern, commit every 12", 2, 4, d_out, nsm);
  run2("tf32 N=128 kernel pattern, nonzero + commits + noise", 2, 7, d_out, nsm);
  run2("tf32 N=128 alternating, NONZERO operands", 0, 2, d_out, nsm);
  return 0;
}
