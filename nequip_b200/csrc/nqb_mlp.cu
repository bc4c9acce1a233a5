// Radial-MLP hidden layer (K = 8) on CUDA cores, sm_100a.
//
// Reference op (paths under /root/reference):
//   edge_weight = ScalarMLPFunction(edge_embedding)            nequip/nn/mlp.py:80-195, 262-268
//   built with depth 1 by InteractionBlock                     nequip/nn/interaction_block.py:119-127, 196
//     h   = silu(emb @ (W1 * a1))          [E, NB] x [NB, H]   (NB = 8 Bessel functions, H = 128)   <- here
//     w   =      h   @ (W2 * a2)           [E, H ] x [H, W]    nqb_gemm.cu (unfused) / the fused TP kernels
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nqb.h"
#include "nqb_tc.cuh"

namespace {

constexpr int H = 128;       // hidden width == K of the second-layer GEMM
constexpr int NB = 8;        // Bessel functions

// ---------------------------------------------------------------------------------------------
// first radial layer on CUDA cores (K = 8): h = silu(emb @ W1s)  and its backward
//   gemb[e, k] = sum_m gh[e, m] * silu'(pre[e, m]) * W1s[k, m],  pre recomputed from emb
// (feeds / follows the grouped tensor-core GEMM of the second layer, nqb_gemm.cu)
// ---------------------------------------------------------------------------------------------
// Persistent warps: lane = 4 hidden units whose 8 x 4 first-layer weights live in registers for the whole
// kernel; a warp walks over edges (grid-stride), reads the 8 basis values of the edge (one broadcast
// 32-byte load) and writes the edge's 128 activations as one 512-byte row.  (The first version re-staged
// the 4 KB weight matrix per 8 edges -- as many bytes as it wrote.)
__device__ __forceinline__ float sigmoid_fast(float p) { return __fdividef(1.0f, 1.0f + expf(-p)); }

__global__ void __launch_bounds__(256) k_hidden_fwd(const float* __restrict__ emb, const float* __restrict__ W1s,
                                                    int64_t E, float* __restrict__ h, float* __restrict__ h_lo) {
  const int lane = threadIdx.x & 31, m0 = lane * 4;
  float w[NB][4];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(W1s + k * H + m0));
    w[k][0] = t.x; w[k][1] = t.y; w[k][2] = t.z; w[k][3] = t.w;
  }
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < E; e += nwarps) {
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(emb + e * NB));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(emb + e * NB + 4));
    const float x[NB] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float p = 0.f;
#pragma unroll
      for (int k = 0; k < NB; ++k) p = fmaf(x[k], w[k][q], p);
      o[q] = p * sigmoid_fast(p);
    }
    __stcs(reinterpret_cast<float4*>(h + e * H + m0), make_float4(o[0], o[1], o[2], o[3]));
    if (h_lo) __stcs(reinterpret_cast<float4*>(h_lo + e * H + m0), make_float4(tf32_lo(o[0]), tf32_lo(o[1]), tf32_lo(o[2]), tf32_lo(o[3])));
  }
}

__global__ void __launch_bounds__(256) k_hidden_bwd(const float* __restrict__ emb, const float* __restrict__ W1s,
                                                    const float* __restrict__ gh, int64_t E, float* __restrict__ gemb) {
  const int lane = threadIdx.x & 31, m0 = lane * 4;
  float w[NB][4];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(W1s + k * H + m0));
    w[k][0] = t.x; w[k][1] = t.y; w[k][2] = t.z; w[k][3] = t.w;
  }
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < E; e += nwarps) {  // warp = edge
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(emb + e * NB));
  const float4 x1 = __ldg(reinterpret_cast<const float4*>(emb + e * NB + 4));
  const float x[NB] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  const float4 g4 = __ldcs(reinterpret_cast<const float4*>(gh + e * H + m0));
  const float g[4] = {g4.x, g4.y, g4.z, g4.w};
  float acc[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) acc[k] = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float p = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) p = fmaf(x[k], w[k][q], p);
    const float sg = sigmoid_fast(p);
    const float gp = g[q] * (sg * (1.0f + p * (1.0f - sg)));
#pragma unroll
    for (int k = 0; k < NB; ++k) acc[k] = fmaf(gp, w[k][q], acc[k]);
  }
  // reduce the 8 partial sums over the 32 lanes (halving butterfly: 4 + 2 + 1 + 2 shuffles)
#pragma unroll
  for (int o = 16, c = NB; o >= 1; o >>= 1) {
    if (c > 1) {
      c >>= 1;
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < c) {
          const float mine = up ? acc[j + c] : acc[j];
          const float theirs = up ? acc[j] : acc[j + c];
          acc[j] = mine + __shfl_xor_sync(0xffffffffu, theirs, o);
        }
      }
    } else {
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
    }
  }
  // after the three halving steps lane bits (16, 8, 4) select the component: k = 4*b16 + 2*b8 + b4
  if ((lane & 3) == 0) {
    const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    gemb[e * NB + k] = acc[0];
  }
  }  // edge loop
}

}  // namespace

extern "C" int nqb_set_error(const char* msg);  // defined in nqb_runtime.cu
extern "C" void nqb_count_launch(void);

// persistent grid: 8 CTAs of 256 threads per SM (or fewer when there is less work)
static unsigned hidden_grid(int64_t threads) {
  static int sms_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (sms_dev[dev] == 0) {
    cudaDeviceGetAttribute(&sms_dev[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms_dev[dev] <= 0) sms_dev[dev] = 148;
  }
  const int sms = sms_dev[dev];
  const int64_t need = (threads + 255) / 256, cap = (int64_t)sms * 8;
  return (unsigned)(need < cap ? need : cap);
}

extern "C" int nqb_mlp_hidden_fwd(const float* emb, const float* W1s, int64_t E, int num_bessel, int hidden, float* h,
                                  float* h_lo, nqb_stream_t st) {
  if (num_bessel != NB || hidden != H) return nqb_set_error("nqb_mlp_hidden_fwd: only num_bessel=8, hidden=128 is built");
  if (E < 0) return nqb_set_error("nqb_mlp_hidden_fwd: negative size");
  if (E == 0) return 0;
  if (!emb || !W1s || !h) return nqb_set_error("nqb_mlp_hidden_fwd: null pointer");
  const int64_t threads = E * 32;
  k_hidden_fwd<<<hidden_grid(threads), 256, 0, (cudaStream_t)st>>>(emb, W1s, E, h, h_lo);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}

extern "C" int nqb_mlp_hidden_bwd(const float* emb, const float* W1s, const float* grad_h, int64_t E, int num_bessel,
                                  int hidden, float* grad_emb, nqb_stream_t st) {
  if (num_bessel != NB || hidden != H) return nqb_set_error("nqb_mlp_hidden_bwd: only num_bessel=8, hidden=128 is built");
  if (E < 0) return nqb_set_error("nqb_mlp_hidden_bwd: negative size");
  if (E == 0) return 0;
  if (!emb || !W1s || !grad_h || !grad_emb) return nqb_set_error("nqb_mlp_hidden_bwd: null pointer");
  const int64_t threads = E * 32;
  k_hidden_bwd<<<hidden_grid(threads), 256, 0, (cudaStream_t)st>>>(emb, W1s, grad_h, E, grad_emb);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}
