// Radial-MLP kernels on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Reference op (paths under /root/reference):
//   edge_weight = ScalarMLPFunction(edge_embedding)            nequip/nn/mlp.py:80-195, 262-268
//   built with depth 1 by InteractionBlock                     nequip/nn/interaction_block.py:119-127, 196
//     h   = silu(emb @ (W1 * a1))          [E, NB] x [NB, H]   (NB = 8 Bessel functions, H = 128)
//     w   =      h   @ (W2 * a2)           [E, H ] x [H, W]    (W = tp.weight_numel, up to 2176)
// The second product is the genuine dense GEMM of the hot path (241 GFLOP per call for the
// Li3PO4 layer 2) and its [E, W] output is 92% of the TP kernel's bytes.  The reference runs it
// as an fp32 SGEMM (TF32 off, nequip/utils/global_state.py:141-149); to stay inside the 1e-5
// parity budget on tensor cores both operands are split  a = hi + lo  with hi = rn_tf32(a),
// lo = a - hi  and three kind::tf32 MMAs are accumulated in fp32 TMEM:
//     D = A_hi B_hi + A_lo B_hi + A_hi B_lo        (missing A_lo B_lo ~ 2^-22 relative)
//
// Forward kernel (k_mlp_fwd): persistent CTAs, one 128-edge tile at a time.
//   warps 0-3  compute h for their edge row straight from the 8 Bessel values (K = 8 is done on
//              CUDA cores), split hi/lo, store it in the canonical K-major core-matrix layout;
//              later they are the epilogue (tcgen05.ld -> st.global)
//   warp  4    one lane streams 32-column weight tiles (pre-split, pre-laid-out) with
//              cp.async.bulk + mbarrier complete_tx
//   warp  5    one lane issues tcgen05.mma (M=128, N=32, K=8) x 16 k-steps x 3 terms per tile,
//              tcgen05.commit frees the weight stage and publishes the accumulator
// Backward kernel (k_mlp_bwd): grad_h = grad_w @ (W2 a2)^T  streamed over K = W in 32-wide chunks
//   (grad_w is split hi/lo on its way through registers), then the epilogue applies silu' and
//   the 128 -> 8 product with W1 on CUDA cores and writes grad_emb [E, 8].
//
// Operand layouts (SWIZZLE_NONE, K-major "interleave" canonical form; units of 16 B):
//   core matrix = 8 rows x 16 B (4 tf32), contiguous 128 B;
//   LBO = byte distance between the two core matrices adjacent in K  (128 B here)
//   SBO = byte distance between 8-row groups along M/N
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nqb.h"
#include "nqb_tc.cuh"

namespace {

constexpr int H = 128;       // hidden width == K of the forward GEMM
constexpr int NB = 8;        // Bessel functions
constexpr int TILE_M = 128;  // edges per tile
constexpr int FWD_N = 32;    // weight columns per forward MMA tile
constexpr int FWD_STAGES = 2;
constexpr int BWD_KC = 32;   // K chunk of the backward GEMM
constexpr int BWD_STAGES = 3;

// ---------------------------------------------------------------------------------------------
// weight preparation (once per model): scaled, split hi/lo, laid out per MMA tile
// ---------------------------------------------------------------------------------------------
// forward tiles: for column tile j (FWD_N columns):  [hi | lo], each [FWD_N x H] K-major canonical
__global__ void k_prep_fwd(const float* __restrict__ W2, float alpha, int W, float* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over W * H
  if (idx >= W * H) return;
  int n = idx / H, k = idx % H;
  float v = W2[(int64_t)k * W + n] * alpha;
  float hi = tf32_rn(v), lo = v - hi;
  int j = n / FWD_N, nn = n % FWD_N;
  float* tile = out + (int64_t)j * (2 * FWD_N * H);
  int off = canon_off(nn, k, H / 4);
  tile[off] = hi;
  tile[FWD_N * H + off] = lo;
}
// backward chunks: for K chunk c (BWD_KC weight columns): [hi | lo], each [H x BWD_KC] K-major canonical
// (B operand of grad_h = grad_w @ W2s^T has N = hidden index, K = weight column)
__global__ void k_prep_bwd(const float* __restrict__ W2, float alpha, int W, float* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over H * W
  if (idx >= W * H) return;
  int m = idx / W, n = idx % W;  // hidden m, weight column n
  float v = W2[(int64_t)m * W + n] * alpha;
  float hi = tf32_rn(v), lo = v - hi;
  int c = n / BWD_KC, kk = n % BWD_KC;
  float* tile = out + (int64_t)c * (2 * H * BWD_KC);
  int off = canon_off(m, kk, BWD_KC / 4);
  tile[off] = hi;
  tile[H * BWD_KC + off] = lo;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct FwdSmem {
  float a_hi[TILE_M * H];                       // 64 KB
  float a_lo[TILE_M * H];                       // 64 KB
  float b[FWD_STAGES][2 * FWD_N * H];           // 2 x 32 KB  (hi | lo)
  float w1[NB * H];                             // 4 KB (scaled first-layer weights)
  uint64_t a_full, a_empty;
  uint64_t b_full[FWD_STAGES], b_empty[FWD_STAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1)
k_mlp_fwd(const float* __restrict__ emb, const float* __restrict__ W1s, const float* __restrict__ Bprep,
          int64_t E, int W, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntile_n = W / FWD_N;
  const int64_t ntile_m = (E + TILE_M - 1) / TILE_M;

  for (int i = tid; i < NB * H; i += blockDim.x) S.w1[i] = W1s[i];
  if (tid == 0) {
    mbar_init(&S.a_full, 128);
    mbar_init(&S.a_empty, 1);
    for (int s = 0; s < FWD_STAGES; ++s) { mbar_init(&S.b_full[s], 1); mbar_init(&S.b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&S.acc_full[s], 1); mbar_init(&S.acc_empty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&S.tmem_base, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp < 4) {
    // ===================== h producer + epilogue (thread == edge row == TMEM lane) ======================
    const int row = tid;  // 0..127
    uint32_t it_tile = 0;     // number of M tiles done by this CTA
    uint32_t acc_iter = 0;    // global N-tile counter (accumulator ring)
    for (int64_t tm = blockIdx.x; tm < ntile_m; tm += gridDim.x, ++it_tile) {
      const int64_t e = tm * TILE_M + row;
      float x[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) x[k] = (e < E) ? __ldg(emb + e * NB + k) : 0.f;
      // A buffers free? (all MMAs of the previous tile retired)
      if (it_tile > 0) mbar_wait(&S.a_empty, (it_tile - 1) & 1);
      for (int kg = 0; kg < H / 4; ++kg) {
        float hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = kg * 4 + q;
          float p = 0.f;
#pragma unroll
          for (int k = 0; k < NB; ++k) p = fmaf(x[k], S.w1[k * H + m], p);
          const float hv = silu_f(p);
          hi[q] = tf32_rn(hv);
          lo[q] = hv - hi[q];
        }
        const int off = (row >> 3) * (H / 4 * 32) + kg * 32 + (row & 7) * 4;
        *reinterpret_cast<float4*>(&S.a_hi[off]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(&S.a_lo[off]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&S.a_full);
      // epilogue over the N tiles of this M tile
      for (int j = 0; j < ntile_n; ++j, ++acc_iter) {
        const uint32_t slot = acc_iter & 1;
        mbar_wait(&S.acc_full[slot], (acc_iter >> 1) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + slot * FWD_N, v);
        tc_fence_before();
        mbar_arrive(&S.acc_empty[slot]);
        if (e < E) {
          float4* dst = reinterpret_cast<float4*>(out + e * (int64_t)W + (int64_t)j * FWD_N);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
  } else if (warp == 4) {
    // ===================== weight-tile loader ==========================================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tm = blockIdx.x; tm < ntile_m; tm += gridDim.x) {
        for (int j = 0; j < ntile_n; ++j, ++it) {
          const uint32_t s = it % FWD_STAGES, ph = (it / FWD_STAGES) & 1;
          if (it >= FWD_STAGES) mbar_wait(&S.b_empty[s], ph ^ 1);
          constexpr uint32_t bytes = 2 * FWD_N * H * sizeof(float);
          mbar_expect_tx(&S.b_full[s], bytes);
          bulk_g2s(S.b[s], Bprep + (int64_t)j * (2 * FWD_N * H), bytes, &S.b_full[s]);
        }
      }
    }
  } else {
    // ===================== MMA issuer ===================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE_M, FWD_N);
      constexpr uint32_t SBO = (H / 4) * 128;  // bytes between 8-row groups
      constexpr uint32_t LBO = 128;            // bytes between K-adjacent core matrices
      const uint32_t a_hi = smem_u32(S.a_hi), a_lo = smem_u32(S.a_lo);
      uint32_t it = 0, it_tile = 0;
      for (int64_t tm = blockIdx.x; tm < ntile_m; tm += gridDim.x, ++it_tile) {
        mbar_wait(&S.a_full, it_tile & 1);
        tc_fence_after();
        for (int j = 0; j < ntile_n; ++j, ++it) {
          const uint32_t s = it % FWD_STAGES, ph = (it / FWD_STAGES) & 1;
          const uint32_t slot = it & 1;
          mbar_wait(&S.b_full[s], ph);
          if (it >= 2) mbar_wait(&S.acc_empty[slot], ((it >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t b_hi = smem_u32(S.b[s]), b_lo = b_hi + FWD_N * H * sizeof(float);
          const uint32_t d = tmem + slot * FWD_N;
          uint32_t acc = 0;
#pragma unroll 1
          for (int term = 0; term < 3; ++term) {
            const uint32_t a0 = (term == 1) ? a_lo : a_hi;
            const uint32_t b0 = (term == 2) ? b_lo : b_hi;
#pragma unroll
            for (int ks = 0; ks < H / 8; ++ks) {
              umma_tf32(d, make_desc(a0 + ks * 256, LBO, SBO), make_desc(b0 + ks * 256, LBO, SBO), idesc, acc);
              acc = 1;
            }
          }
          umma_commit(&S.b_empty[s]);      // weight stage can be refilled once these MMAs retire
          umma_commit(&S.acc_full[slot]);  // accumulator ready for the epilogue
        }
        umma_commit(&S.a_empty);           // A tile may be overwritten
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------------------------
// backward:  grad_emb[e, :] = ((grad_w[e, :] @ W2s^T) * silu'(pre[e, :])) @ W1s^T
// ---------------------------------------------------------------------------------------------
struct BwdSmem {
  float a[BWD_STAGES][2 * TILE_M * BWD_KC];  // 3 x 32 KB (hi | lo) grad_w chunk
  float b[BWD_STAGES][2 * H * BWD_KC];       // 3 x 32 KB (hi | lo) weight chunk
  float w1[NB * H];
  uint64_t a_full[BWD_STAGES], b_full[BWD_STAGES], empty[BWD_STAGES];
  uint64_t acc_full, acc_empty;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1)
k_mlp_bwd(const float* __restrict__ emb, const float* __restrict__ W1s, const float* __restrict__ Bprep,
          const float* __restrict__ gw, int64_t E, int W, float* __restrict__ gemb) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  BwdSmem& S = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = W / BWD_KC;
  const int64_t ntile_m = (E + TILE_M - 1) / TILE_M;

  for (int i = tid; i < NB * H; i += blockDim.x) S.w1[i] = W1s[i];
  if (tid == 0) {
    for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&S.a_full[s], 128); mbar_init(&S.b_full[s], 1); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.acc_full, 1);
    mbar_init(&S.acc_empty, 128);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&S.tmem_base, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp < 4) {
    // ============ grad_w chunk producer (split hi/lo through registers) + epilogue ======================
    uint32_t it = 0, it_tile = 0;
    const int r8 = lane & 7, kq = lane >> 3;  // lane -> (row within group of 8, 16-byte k group 0..3)
    for (int64_t tm = blockIdx.x; tm < ntile_m; tm += gridDim.x, ++it_tile) {
      const int64_t e0 = tm * TILE_M;
      for (int c = 0; c < nchunk; ++c, ++it) {
        const uint32_t s = it % BWD_STAGES, ph = (it / BWD_STAGES) & 1;
        // each warp owns rows [32*warp, 32*warp+32): 4 groups of 8 rows, 8 k-groups (two passes of 4)
        float4 v[8];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int64_t e = e0 + warp * 32 + g * 8 + r8;
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            const int kg = hp * 4 + kq;
            v[g * 2 + hp] = (e < E) ? __ldg(reinterpret_cast<const float4*>(gw + e * (int64_t)W + (int64_t)c * BWD_KC + kg * 4))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (it >= BWD_STAGES) mbar_wait(&S.empty[s], ph ^ 1);
        float* ahi = S.a[s];
        float* alo = S.a[s] + TILE_M * BWD_KC;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            const int row = warp * 32 + g * 8 + r8, kg = hp * 4 + kq;
            const float4 a = v[g * 2 + hp];
            float4 hi = make_float4(tf32_rn(a.x), tf32_rn(a.y), tf32_rn(a.z), tf32_rn(a.w));
            float4 lo = make_float4(a.x - hi.x, a.y - hi.y, a.z - hi.z, a.w - hi.w);
            const int off = (row >> 3) * (BWD_KC / 4 * 32) + kg * 32 + (row & 7) * 4;
            *reinterpret_cast<float4*>(ahi + off) = hi;
            *reinterpret_cast<float4*>(alo + off) = lo;
          }
        }
        fence_proxy_async();
        mbar_arrive(&S.a_full[s]);
      }
      // ---- epilogue for this M tile: thread == edge row
      const int row = tid;
      const int64_t e = e0 + row;
      float x[NB], ge[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) { x[k] = (e < E) ? __ldg(emb + e * NB + k) : 0.f; ge[k] = 0.f; }
      mbar_wait(&S.acc_full, it_tile & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < H / 32; ++cb) {
        float gh[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, gh);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int m = cb * 32 + q;
          float p = 0.f;
#pragma unroll
          for (int k = 0; k < NB; ++k) p = fmaf(x[k], S.w1[k * H + m], p);
          const float sg = 1.0f / (1.0f + expf(-p));
          const float ds = sg * (1.0f + p * (1.0f - sg));  // d silu / d p
          const float gp = gh[q] * ds;
#pragma unroll
          for (int k = 0; k < NB; ++k) ge[k] = fmaf(gp, S.w1[k * H + m], ge[k]);
        }
      }
      tc_fence_before();
      mbar_arrive(&S.acc_empty);
      if (e < E) {
        float4* dst = reinterpret_cast<float4*>(gemb + e * NB);
        dst[0] = make_float4(ge[0], ge[1], ge[2], ge[3]);
        dst[1] = make_float4(ge[4], ge[5], ge[6], ge[7]);
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tm = blockIdx.x; tm < ntile_m; tm += gridDim.x) {
        for (int c = 0; c < nchunk; ++c, ++it) {
          const uint32_t s = it % BWD_STAGES, ph = (it / BWD_STAGES) & 1;
          if (it >= BWD_STAGES) mbar_wait(&S.empty[s], ph ^ 1);
          constexpr uint32_t bytes = 2 * H * BWD_KC * sizeof(float);
          mbar_expect_tx(&S.b_full[s], bytes);
          bulk_g2s(S.b[s], Bprep + (int64_t)c * (2 * H * BWD_KC), bytes, &S.b_full[s]);
        }
      }
    }
  } else {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE_M, H);
      constexpr uint32_t SBO = (BWD_KC / 4) * 128, LBO = 128;
      uint32_t it = 0, it_tile = 0;
      for (int64_t tm = blockIdx.x; tm < ntile_m; tm += gridDim.x, ++it_tile) {
        if (it_tile > 0) mbar_wait(&S.acc_empty, (it_tile - 1) & 1);
        uint32_t acc = 0;
        for (int c = 0; c < nchunk; ++c, ++it) {
          const uint32_t s = it % BWD_STAGES, ph = (it / BWD_STAGES) & 1;
          mbar_wait(&S.a_full[s], ph);
          mbar_wait(&S.b_full[s], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(S.a[s]), a_lo = a_hi + TILE_M * BWD_KC * sizeof(float);
          const uint32_t b_hi = smem_u32(S.b[s]), b_lo = b_hi + H * BWD_KC * sizeof(float);
#pragma unroll 1
          for (int term = 0; term < 3; ++term) {
            const uint32_t a0 = (term == 1) ? a_lo : a_hi;
            const uint32_t b0 = (term == 2) ? b_lo : b_hi;
#pragma unroll
            for (int ks = 0; ks < BWD_KC / 8; ++ks) {
              umma_tf32(tmem, make_desc(a0 + ks * 256, LBO, SBO), make_desc(b0 + ks * 256, LBO, SBO), idesc, acc);
              acc = 1;
            }
          }
          umma_commit(&S.empty[s]);
        }
        umma_commit(&S.acc_full);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 128);
}

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

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int nqb_set_error(const char* msg);  // defined in nqb_runtime.cu
extern "C" void nqb_count_launch(void);

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

extern "C" size_t nqb_mlp_prepared_bytes(int W) { return (size_t)2 * H * (size_t)W * sizeof(float); }

extern "C" int nqb_mlp_prepare(const float* W2, float alpha2, int hidden, int W, float* prep_fwd, float* prep_bwd,
                               nqb_stream_t st) {
  if (hidden != H) return nqb_set_error("nqb_mlp_prepare: hidden width must be 128");
  if (W <= 0 || W % 32 != 0) return nqb_set_error("nqb_mlp_prepare: weight_numel must be a positive multiple of 32");
  if (!W2 || !prep_fwd || !prep_bwd) return nqb_set_error("nqb_mlp_prepare: null pointer");
  int n = W * H;
  k_prep_fwd<<<(n + 255) / 256, 256, 0, (cudaStream_t)st>>>(W2, alpha2, W, prep_fwd);
  k_prep_bwd<<<(n + 255) / 256, 256, 0, (cudaStream_t)st>>>(W2, alpha2, W, prep_bwd);
  nqb_count_launch();
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}

extern "C" int nqb_mlp_fwd(const float* emb, const float* W1s, const float* prep_fwd, int64_t E, int num_bessel,
                           int hidden, int W, float* out, nqb_stream_t st) {
  if (num_bessel != NB || hidden != H) return nqb_set_error("nqb_mlp_fwd: only num_bessel=8, hidden=128 is built");
  if (W <= 0 || W % FWD_N != 0) return nqb_set_error("nqb_mlp_fwd: weight_numel must be a multiple of 32");
  if (E < 0) return nqb_set_error("nqb_mlp_fwd: negative size");
  if (E == 0) return 0;
  if (!emb || !W1s || !prep_fwd || !out) return nqb_set_error("nqb_mlp_fwd: null pointer");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem) + 1024);
    if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
    attr_set = true;
  }
  int64_t tiles = (E + TILE_M - 1) / TILE_M;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  k_mlp_fwd<<<grid, 192, sizeof(FwdSmem) + 1024, (cudaStream_t)st>>>(emb, W1s, prep_fwd, E, W, out);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}

extern "C" int nqb_mlp_bwd(const float* emb, const float* W1s, const float* prep_bwd, const float* grad_w, int64_t E,
                           int num_bessel, int hidden, int W, float* grad_emb, nqb_stream_t st) {
  if (num_bessel != NB || hidden != H) return nqb_set_error("nqb_mlp_bwd: only num_bessel=8, hidden=128 is built");
  if (W <= 0 || W % BWD_KC != 0) return nqb_set_error("nqb_mlp_bwd: weight_numel must be a multiple of 32");
  if (E < 0) return nqb_set_error("nqb_mlp_bwd: negative size");
  if (E == 0) return 0;
  if (!emb || !W1s || !prep_bwd || !grad_w || !grad_emb) return nqb_set_error("nqb_mlp_bwd: null pointer");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem) + 1024);
    if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
    attr_set = true;
  }
  int64_t tiles = (E + TILE_M - 1) / TILE_M;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  k_mlp_bwd<<<grid, 192, sizeof(BwdSmem) + 1024, (cudaStream_t)st>>>(emb, W1s, prep_bwd, grad_w, E, W, grad_emb);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}

// persistent grid: 8 CTAs of 256 threads per SM (or fewer when there is less work)
static unsigned hidden_grid(int64_t threads) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
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
