// EXPERIMENTAL (round-2 candidate, not on the product path yet; tests/test_gemm_t_gpu.py is opt-in):
// fp32-accurate GEMM for K <= 128 with the WEIGHTS as the TMEM-resident A operand of tcgen05.mma, sm_100a.
//
//   C[M, N] (ldc) = A[M, K] (lda, fp32 row-major) @ B[K, N]            -- the radial-MLP last layer,
//                                                                          nequip/nn/mlp.py:262-268
// computed transposed:  D^T[n, m] = sum_k W^T[n, k] * A^T[k, m]  per (128-column N-tile, 64-row M-tile):
//   * MMA "A" operand = W^T tile [128 n x K] (tf32 hi and lo parts), written ONCE per N-tile into tensor
//     memory (256 of the 512 columns) -- the MMAs then read only the edge operand from shared memory
//     (2 KB per M128 x N64 x K8 instruction = 64 B/clk instead of the 128 B/clk that the operand-in-smem form of
//     nqb_gemm.cu needs; profiles/r01_gemm_roles.txt shows that form starves every other shared-memory user);
//   * MMA "B" operand = the A rows (edges), K-major canonical core-matrix layout in shared memory: the fp32
//     tile itself (hardware truncation = high part) and its low part, exactly as in nqb_gemm.cu;
//   * D^T accumulators (hi*hi and cross terms, 64 columns each, double buffered) in the other 256 columns;
//   * the epilogue thread owns one OUTPUT COLUMN n (TMEM lane) and 64 rows m: its stores C[m, n0 + lane] are
//     128-byte coalesced per warp without any staging transpose.
// 3xTF32 split and accumulation as in nqb_gemm.cu (K <= 128: 16 accumulate steps on hi*hi, 32 on the cross
// terms, one segment).
//
// Roles (512 threads): warps 0-7 two producer groups (tiles dealt round-robin; cp.async -> raw ring of 3 x 32 KB,
// low-part pass into a ring of 2 x 32 KB), warps 8-11 epilogue (+ weight upload), warp 12 MMA issue.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nqb.h"
#include "nqb_tc.cuh"

namespace {

constexpr int TE = 64;     // rows of A (edges) per tile = MMA N
constexpr int TW = 128;    // output columns per N-tile = MMA M = TMEM lanes
constexpr int KMAX = 128;  // resident K
constexpr int RAWT = 3, NLOT = 2;
constexpr int NT_THREADS = 512;
constexpr int TILE_FLOATS = TE * KMAX;  // 32 KB

struct SmemT {
  float araw[RAWT][TILE_FLOATS];
  float alo[NLOT][TILE_FLOATS];
  uint64_t a_full[RAWT], a_done[RAWT];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t w_full;
  uint32_t tmem_base;
};

// tcgen05.mma with the A operand in tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 32 columns: thread i of the warp writes row (lane base + i), 32 consecutive columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// work split: T N-tiles, G CTAs.  T <= G: CTA b owns N-tile b % T and the M-tiles r, r + R, ... (R = G / T);
// otherwise CTA b owns the N-tiles b, b + G, ... and all M-tiles.
struct SchedT {
  int q, q_step, T;
  int64_t m_start, m_step;
  __device__ SchedT(int b, int G, int T_) : T(T_) {
    if (T_ <= G) {
      const int R = G / T_;
      q = (b < T_ * R) ? (b % T_) : T_;
      q_step = T_;
      m_start = b / T_;
      m_step = R;
    } else {
      q = b; q_step = G; m_start = 0; m_step = 1;
    }
  }
};

__global__ void __launch_bounds__(NT_THREADS, 1)
k_gemm3x_t(const float* __restrict__ A, int64_t lda, const float* __restrict__ Wp, float* __restrict__ C, int64_t ldc,
           int64_t M, int K, int N) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemT& S = *reinterpret_cast<SmemT*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t mtiles = (M + TE - 1) / TE;
  const int ntiles = (N + TW - 1) / TW;
  const int ksteps = (K + 7) / 8;  // MMA k-steps (operands are zero beyond K)

  if (tid == 0) {
    for (int s = 0; s < RAWT; ++s) { mbar_init(&S.a_full[s], 128); mbar_init(&S.a_done[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&S.acc_full[s], 1); mbar_init(&S.acc_empty[s], 128); }
    mbar_init(&S.w_full, 128);
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(&S.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const SchedT sch(blockIdx.x, gridDim.x, ntiles);
  // TMEM columns: [0,128) W hi, [128,256) W lo, then per buffer b: [256 + 128 b, +64) hi*hi, [+64, +128) cross terms
  const bool has_work = sch.q < sch.T && sch.m_start < mtiles;

  if (warp < 8) {
    // =========================== producers: group g owns the tiles i = g (mod 2) ==========================
    const int grp = warp >> 2, gw = warp & 3;
    const int r8 = lane & 7, kq = lane >> 3;
    const int kgroups = ksteps * 2;  // 16-byte k-groups that the MMAs read
    // flat tile sequence (q outer, M-tiles inner); cursor = next own tile
    int cq = sch.q;
    int64_t cmt = sch.m_start;
    bool cvalid = has_work;
    uint32_t ci = 0;  // index of the cursor's tile in the sequence
    auto advance = [&]() {
      cmt += sch.m_step;
      if (cmt >= mtiles) { cmt = sch.m_start; cq += sch.q_step; cvalid = cq < sch.T; }
      ++ci;
    };
    for (int j = 0; j < grp && cvalid; ++j) advance();
    auto issue = [&]() {  // copy the cursor's tile (if any) into raw stage ci % RAWT, move to the next own tile
      if (cvalid) {
        float* dst = S.araw[ci % RAWT];
        const int64_t m0 = cmt * TE;
#pragma unroll 1
        for (int kb = gw; kb * 4 < kgroups; kb += 4) {
          const int kg = kb * 4 + kq;
          const int k = kg * 4;
#pragma unroll
          for (int rg = 0; rg < TE / 8; ++rg) {
            const int64_t m = m0 + rg * 8 + r8;
            const bool in = m < M && k < K;
            cp_async16(dst + rg * (KMAX / 4 * 32) + kg * 32 + r8 * 4, A + (in ? m * lda + k : 0), in ? 16u : 0u);
          }
        }
        advance();
        if (cvalid) advance();
      }
      cp_async_commit();
    };
    uint32_t i = grp;            // tile being finished
    bool ivalid = cvalid;
    issue();
    while (ivalid) {
      cp_async_wait<0>();        // my pieces of tile i have landed
      // low parts of tile i: stage i % NLOT was last read by tile i - NLOT (my previous tile)
      if (i >= NLOT) mbar_wait(&S.a_done[(i - NLOT) % RAWT], ((i - NLOT) / RAWT) & 1);
      const float* raw = S.araw[i % RAWT];
      float* lo = S.alo[i % NLOT];
#pragma unroll 1
      for (int kb = gw; kb * 4 < kgroups; kb += 4) {
        const int kg = kb * 4 + kq;
        float4 a[TE / 8];
#pragma unroll
        for (int rg = 0; rg < TE / 8; ++rg) a[rg] = *reinterpret_cast<const float4*>(raw + rg * (KMAX / 4 * 32) + kg * 32 + r8 * 4);
#pragma unroll
        for (int rg = 0; rg < TE / 8; ++rg)
          *reinterpret_cast<float4*>(lo + rg * (KMAX / 4 * 32) + kg * 32 + r8 * 4) =
              make_float4(tf32_lo(a[rg].x), tf32_lo(a[rg].y), tf32_lo(a[rg].z), tf32_lo(a[rg].w));
      }
      fence_proxy_async();
      mbar_arrive(&S.a_full[i % RAWT]);
      // only now prefetch my next tile (waiting for its stage first would delay tile i behind the MMAs of
      // tile i - 1): its raw stage was last read by tile inext - RAWT
      const bool nvalid = cvalid;
      const uint32_t inext = ci;
      if (nvalid && inext >= RAWT) mbar_wait(&S.a_done[(inext - RAWT) % RAWT], ((inext - RAWT) / RAWT) & 1);
      issue();
      i = inext;
      ivalid = nvalid;
    }
    cp_async_wait<0>();
  } else if (warp < 12) {
    // =========================== epilogue (+ weight upload), thread = output column =====================
    const int ew = warp - 8;  // TMEM lane quadrant == warp % 4
    const uint32_t tlane = tmem + ((uint32_t)(ew * 32) << 16);
    uint32_t i = 0, wq = 0;
    for (int q = sch.q; q < sch.T && sch.m_start < mtiles; q += sch.q_step, ++wq) {
      // all MMAs of the previous N-tile have retired (its last accumulators were drained below)
      const float* wrow = Wp + ((int64_t)q * 2 * TW + ew * 32 + lane) * KMAX;
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {        // hi rows, then lo rows
#pragma unroll 1
        for (int c = 0; c < KMAX / 32; ++c) {
          float v[32];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(wrow + (int64_t)part * TW * KMAX + c * 32 + u * 4));
            v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
          }
          tmem_st32(tlane + part * 128 + c * 32, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&S.w_full);
      const int n = q * TW + ew * 32 + lane;
      const bool n_ok = n < N;
      for (int64_t mt = sch.m_start; mt < mtiles; mt += sch.m_step, ++i) {
        const uint32_t buf = i & 1;
        mbar_wait(&S.acc_full[buf], (i >> 1) & 1);
        tc_fence_after();
        const int64_t m0 = mt * TE;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          float hh[32], xx[32];
          tmem_ld32x2(tlane + 256 + buf * 128 + half * 32, tlane + 256 + buf * 128 + 64 + half * 32, hh, xx);
          if (half == 1) {  // every column of this buffer has been read
            tc_fence_before();
            mbar_arrive(&S.acc_empty[buf]);
          }
          float* crow = C + (m0 + half * 32) * ldc + n;
          const int jmax = (int)min((int64_t)32, M - (m0 + half * 32));
          if (n_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < jmax) crow[(int64_t)j * ldc] = hh[j] + xx[j];
          }
        }
      }
    }
  } else if (warp == 12) {
    // =========================== MMA issue ================================================================
    const bool leader = elect_one();
    constexpr uint32_t SBO = (KMAX / 4) * 128, LBO = 128;
    const uint64_t dR0 = make_desc(smem_u32(S.araw[0]), LBO, SBO);
    const uint64_t dL0 = make_desc(smem_u32(S.alo[0]), LBO, SBO);
    constexpr uint32_t STAGE = (TILE_FLOATS * sizeof(float)) >> 4;
    const uint32_t idesc = make_idesc(TW, TE);
    uint32_t i = 0, s = 0, ph = 0, wq = 0;
    for (int q = sch.q; q < sch.T && sch.m_start < mtiles; q += sch.q_step, ++wq) {
      mbar_wait(&S.w_full, wq & 1);
      tc_fence_after();
      for (int64_t mt = sch.m_start; mt < mtiles; mt += sch.m_step, ++i) {
        const uint32_t buf = i & 1;
        if (i >= 2) { mbar_wait(&S.acc_empty[buf], ((i >> 1) - 1) & 1); tc_fence_after(); }
        const uint64_t b_hi = dR0 + (uint64_t)(s * STAGE), b_lo = dL0 + (uint64_t)((i % NLOT) * STAGE);
        const uint32_t d_hh = tmem + 256 + buf * 128, d_x = d_hh + 64;
        mbar_wait(&S.a_full[s], ph);
        if (leader) {
          // one k-step = 8 tf32 = 8 TMEM columns of the weights, 2 core matrices (16 descriptor units) of the rows
#pragma unroll 4
          for (int ks = 0; ks < ksteps; ++ks) umma_tf32_ts(d_hh, tmem + ks * 8, b_hi + ks * 16, idesc, ks > 0);
#pragma unroll 4
          for (int ks = 0; ks < ksteps; ++ks) umma_tf32_ts(d_x, tmem + 128 + ks * 8, b_hi + ks * 16, idesc, ks > 0);
#pragma unroll 4
          for (int ks = 0; ks < ksteps; ++ks) umma_tf32_ts(d_x, tmem + ks * 8, b_lo + ks * 16, idesc, 1);
          umma_commit(&S.a_done[s]);
          umma_commit(&S.acc_full[buf]);
        }
        __syncwarp();
        if (++s == RAWT) { s = 0; ph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem, 512);
}

// prepared layout: [N-tile][hi | lo][128 columns n][KMAX] floats; zero beyond N and K
__global__ void k_gemm_t_prepare(const float* __restrict__ B, int64_t ldb, int K, int N, int transposed, float scale,
                                 float* __restrict__ out, int ntiles) {
  const int64_t total = (int64_t)ntiles * TW * KMAX;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % KMAX);
    const int nn = (int)((idx / KMAX) % TW);
    const int j = (int)(idx / ((int64_t)KMAX * TW));
    const int n = j * TW + nn;
    float v = 0.f;
    if (k < K && n < N) v = (transposed ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n]) * scale;
    const float hi = tf32_rn(v), lo = v - hi;
    float* base = out + (int64_t)j * 2 * TW * KMAX;
    base[(int64_t)nn * KMAX + k] = hi;
    base[(int64_t)(TW + nn) * KMAX + k] = lo;
  }
}

}  // namespace

extern "C" int nqb_set_error(const char* msg);
extern "C" void nqb_count_launch(void);

extern "C" int64_t nqb_gemm_t_prepared_floats(int K, int N) {
  if (K <= 0 || N <= 0 || K > KMAX) return 0;
  return (int64_t)((N + TW - 1) / TW) * 2 * TW * KMAX;
}

extern "C" int nqb_gemm_t_prepare(const float* B, int64_t ldb, int K, int N, int transposed, float scale, float* prepared,
                                  nqb_stream_t st) {
  if (!B || !prepared) return nqb_set_error("nqb_gemm_t_prepare: null pointer");
  if (K <= 0 || N <= 0 || K > KMAX) return nqb_set_error("nqb_gemm_t_prepare: needs 0 < K <= 128");
  const int ntiles = (N + TW - 1) / TW;
  const int64_t total = (int64_t)ntiles * TW * KMAX;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  k_gemm_t_prepare<<<blocks, 256, 0, (cudaStream_t)st>>>(B, ldb, K, N, transposed, scale, prepared, ntiles);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}

extern "C" int nqb_gemm_t_run(const float* prepared, int K, int N, const float* A, int64_t lda, float* C, int64_t ldc,
                              int64_t M, nqb_stream_t st) {
  if (M < 0) return nqb_set_error("nqb_gemm_t_run: negative M");
  if (M == 0) return 0;
  if (!prepared || !A || !C) return nqb_set_error("nqb_gemm_t_run: null pointer");
  if (K <= 0 || K > KMAX || N <= 0 || (K % 4) || (lda % 4)) return nqb_set_error("nqb_gemm_t_run: needs 0 < K <= 128, K and lda multiples of 4");
  static bool attr_set[64] = {false};
  static int sms_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm3x_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT) + 1024);
    if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
    cudaDeviceGetAttribute(&sms_dev[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms_dev[dev] <= 0) sms_dev[dev] = 148;
    attr_set[dev] = true;
  }
  const int sms = sms_dev[dev];
  const int ntiles = (N + TW - 1) / TW;
  const int64_t nwork = ((M + TE - 1) / TE) * (int64_t)ntiles;
  const int grid = (int)(nwork < sms ? nwork : sms);
  k_gemm3x_t<<<grid, NT_THREADS, sizeof(SmemT) + 1024, (cudaStream_t)st>>>(A, lda, prepared, C, ldc, M, K, N);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}
