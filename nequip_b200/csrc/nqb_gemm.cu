// Grouped fp32-accurate GEMM on the 5th-gen tensor cores (tcgen05 kind::tf32, 3xTF32 split), sm_100a.
//
//   C_p[M, N_p] (ldc)  (+)=  rowscale_p[m] * ( A_p[M, K_p] (lda, fp32 row-major) @ B_p[K_p, N_p] )
//
// for a list of problems p that share M (number of edges or atoms) -- one launch per dense layer:
//   * radial MLP second layer and its backward         nequip/nn/mlp.py:262-268 (torch.mm)
//   * o3.Linear per-irrep channel mixing (linear_1/2)   nequip/nn/interaction_block.py:82-87,129-138
//   * self-connection FullyConnectedTensorProduct       nequip/nn/interaction_block.py:140-146
//     (per-type effective weights; rowscale = one-hot column of the atom type)
// in the channel-contiguous (ir_mul) node layout every (irrep component, chunk pair) is a plain
// strided GEMM, so no transposes/copies are needed around these calls.
//
// fp32 parity on a TF32 pipe: a = hi + lo (hi = cvt.rna.tf32), D = A_hi B_hi + (A_lo B_hi + A_hi B_lo).
// The tensor core's fp32 accumulate truncates (measured ~3e-8 relative bias per accumulation
// step), so the two cross terms go to their own TMEM accumulator and long reductions are cut into
// segments of SEG_CHUNKS*32 in K whose partial sums are added in registers (round-to-nearest).
//
// Roles per CTA (384 threads = 3 warpgroups with setmaxnreg-rebalanced registers, persistent over (M-tile, N-tile) work items, N-tile fastest so that
// the CTAs working on one A tile run together and share it in L2):
//   warps 0-3  producers: stream the A chunk [128 x 32] through registers (coalesced 16-byte loads, one
//              chunk ahead, across work-item boundaries), split hi/lo into the canonical K-major
//              core-matrix layout in shared memory
//   warps 4-7  epilogue: drain finished accumulator segments TMEM -> registers (fp32 adds), then write C
//              through a swizzled staging tile so that the global stores are full 128-byte rows;
//              overlaps with the production / MMA of the next work item (TMEM is double buffered)
//   warp  8    lane 0: cp.async.bulk of the pre-split weight chunk (+ mbarrier complete_tx)
//   warp  9    lane 0: tcgen05.mma M=128, N<=128, K=8: 4 k-steps x 3 terms per chunk; tcgen05.commit
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nqb.h"
#include "nqb_tc.cuh"

namespace {

constexpr int TM = 128;          // rows per tile
constexpr int TN = 128;          // max columns per tile
constexpr int KC = 32;           // K chunk
constexpr int ASTAGES = 2;       // A ring (the producers hold one more chunk in registers)
constexpr int BSLOTS = 4;        // B slots: a ring when K > 128, resident per N-tile when K <= 128
constexpr int SEG_CHUNKS = 10;   // chunks per accumulation segment (40 accumulate steps on the hi*hi accumulator)
constexpr int BLOCK_FLOATS = 2 * TN * KC;  // one prepared weight block: [hi | lo] x [128 x 32]

struct GemmDesc {  // mirrored by nequip_b200/ops.py (int64 fields)
  int64_t a_off, c_off, b_off, rs_off;  // element offsets from the base pointers; rs_off = ROW of the
                                        // [R, rs_ld] row-scale matrix (< 0: no row scale)
  int64_t lda, ldc, K, N;
  int64_t kchunks, ntiles, tile0, flags;  // flags: see the epilogue
};

struct Smem {
  float a[ASTAGES][2 * TM * KC];   // 2 x 32 KB
  float b[BSLOTS][BLOCK_FLOATS];   // 4 x 32 KB
  float stage[4][32 * 32];         // epilogue staging, one 32x32 tile per warp (swizzled)
  uint64_t a_full[ASTAGES], a_empty[ASTAGES];
  uint64_t b_full[BSLOTS], b_empty[BSLOTS];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ const GemmDesc* find_desc(const GemmDesc* d, int nd, int q) {
  int i = 0;
  while (i + 1 < nd && d[i + 1].tile0 <= q) ++i;
  return d + i;
}

// Work schedule (identical in every role).  T = N-tiles over all problems, G = CTAs.
//   T <= G: CTA b owns ONE N-tile q = b % T and the M-tiles r, r+R, r+2R, ... (r = b / T, R = G / T):
//           the T CTAs that share r sweep the same M-tiles in lockstep (A tiles are shared in L2)
//           and the weight tile of a K <= 128 problem stays resident in shared memory;
//   T >  G: CTA b owns N-tiles b, b+G, ... and sweeps all M-tiles for each.
struct Sched {
  int q, q_step, nq_total;
  int64_t m_start, m_step;
  __device__ Sched(int b, int G, int T) {
    if (T <= G) {
      const int R = G / T;
      q = (b < T * R) ? (b % T) : T;  // T = no work
      q_step = T;
      m_start = b / T;
      m_step = R;
    } else {
      q = b; q_step = G; m_start = 0; m_step = 1;
    }
    nq_total = T;
  }
};

struct WorkQ {  // one N-tile of one problem
  const GemmDesc* d;
  int nt, kchunks, K, ncols;
  bool resident;
};
__device__ __forceinline__ void decode_q(const GemmDesc* descs, int ndesc, int q, WorkQ& w) {
  w.d = find_desc(descs, ndesc, q);
  w.nt = q - (int)w.d->tile0;
  w.kchunks = (int)w.d->kchunks;
  w.K = (int)w.d->K;
  w.ncols = min(TN, (int)w.d->N - w.nt * TN);
  w.resident = w.kchunks <= BSLOTS;
}

__global__ void __launch_bounds__(384, 1)
k_gemm3x(const GemmDesc* __restrict__ descs, int ndesc, int ntiles_total, const float* __restrict__ a_base,
         const float* __restrict__ b_base, float* __restrict__ c_base, const float* __restrict__ rs_base, int64_t rs_ld,
         int64_t M) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t mtiles = (M + TM - 1) / TM;

  if (tid == 0) {
    for (int s = 0; s < ASTAGES; ++s) { mbar_init(&S.a_full[s], 128); mbar_init(&S.a_empty[s], 1); }
    for (int s = 0; s < BSLOTS; ++s) { mbar_init(&S.b_full[s], 1); mbar_init(&S.b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&S.acc_full[s], 1); mbar_init(&S.acc_empty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&S.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const Sched sch(blockIdx.x, gridDim.x, ntiles_total);

  // register budget per warpgroup (launch: 65536 / 384 = 168 each): the epilogue keeps a 128-value row of
  // partial sums in registers, the producers and the two single-lane roles need few
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // =========================== A producer: global -> registers (one chunk ahead) -> split -> smem ======
    const int r8 = lane & 7, kq = lane >> 3;
    uint32_t it = 0;
    float4 v[8], vn[8];
    for (int q = sch.q; q < sch.nq_total; q += sch.q_step) {
      WorkQ w;
      decode_q(descs, ndesc, q, w);
      const float* A = a_base + w.d->a_off;
      const int64_t lda = w.d->lda;
      auto load_chunk = [&](int64_t mt, int c, float4* dst) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int64_t m = mt * TM + warp * 32 + g * 8 + r8;
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            const int k = c * KC + (hp * 4 + kq) * 4;
            dst[g * 2 + hp] = (m < M && k < w.K) ? __ldg(reinterpret_cast<const float4*>(A + m * lda + k))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      int64_t mt = sch.m_start;
      int c = 0;
      if (mt < mtiles) load_chunk(mt, 0, v);
      while (mt < mtiles) {
        int cn = c + 1;
        int64_t mtn = mt;
        if (cn == w.kchunks) { cn = 0; mtn = mt + sch.m_step; }
        if (mtn < mtiles) load_chunk(mtn, cn, vn);
        const uint32_t s = it % ASTAGES, ph = (it / ASTAGES) & 1;
        if (it >= ASTAGES) mbar_wait(&S.a_empty[s], ph ^ 1);
        float* ahi = S.a[s];
        float* alo = S.a[s] + TM * KC;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            const int r = warp * 32 + g * 8 + r8, kg = hp * 4 + kq;
            const float4 a = v[g * 2 + hp];
            const float4 hi = make_float4(tf32_rn(a.x), tf32_rn(a.y), tf32_rn(a.z), tf32_rn(a.w));
            const float4 lo = make_float4(a.x - hi.x, a.y - hi.y, a.z - hi.z, a.w - hi.w);
            const int off = (r >> 3) * (KC / 4 * 32) + kg * 32 + (r & 7) * 4;
            *reinterpret_cast<float4*>(ahi + off) = hi;
            *reinterpret_cast<float4*>(alo + off) = lo;
          }
        }
        fence_proxy_async();
        mbar_arrive(&S.a_full[s]);
        ++it;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = vn[j];
        c = cn; mt = mtn;
      }
    }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // =========================== epilogue: TMEM segments -> registers -> C ============================
    const int ew = warp - 4;           // TMEM lane quadrant == warp % 4
    const int row = ew * 32 + lane;    // accumulator row owned by this thread
    uint32_t gseg = 0;
    float acc[TN];
    for (int q = sch.q; q < sch.nq_total; q += sch.q_step) {
      WorkQ w;
      decode_q(descs, ndesc, q, w);
      const GemmDesc* d = w.d;
      const int nseg = (w.kchunks + SEG_CHUNKS - 1) / SEG_CHUNKS;
      const int ncols = w.ncols;
      const int64_t ldc = d->ldc;
      float* C = c_base + d->c_off + (int64_t)w.nt * TN;
      // flags: bit0 read-modify-write accumulate (single writer per element within the launch),
      //        bit1 rows whose row scale is zero are not touched (disjoint row-masked writers),
      //        bit2 accumulate with red.global.add (several problems of this launch add into the same C)
      const bool accumulate = (d->flags & 1) != 0, skipz = (d->flags & 2) != 0, atomic = (d->flags & 4) != 0;
      for (int64_t mt = sch.m_start; mt < mtiles; mt += sch.m_step) {
        const int64_t m0 = mt * TM;
        float rs = 1.0f;
        if (d->rs_off >= 0) {
          const int64_t m = m0 + row;
          rs = (m < M) ? __ldg(rs_base + d->rs_off * rs_ld + m) : 0.f;
        }
        const uint32_t tlane = tmem + ((uint32_t)(ew * 32) << 16);
        float4* st = reinterpret_cast<float4*>(S.stage[ew]);
        // 32 summed + scaled values of my row -> swizzled staging tile -> full 128-byte row segments of C
        auto emit = [&](int cb, const float* v) {
          const int rl = lane;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            st[rl * 8 + (u ^ (rl & 7))] = make_float4(v[4 * u] * rs, v[4 * u + 1] * rs, v[4 * u + 2] * rs, v[4 * u + 3] * rs);
          __syncwarp();
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const int rr = p * 4 + (lane >> 3), u = lane & 7;
            const int64_t m = m0 + ew * 32 + rr;
            const int col = cb * 32 + u * 4;
            const float rsr = __shfl_sync(0xffffffffu, rs, rr);
            if (m < M && col < ncols && !(skipz && rsr == 0.f)) {
              float4 val = st[rr * 8 + (u ^ (rr & 7))];
              float4* dst = reinterpret_cast<float4*>(C + m * ldc + col);
              if (atomic) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(val.x), "f"(val.y),
                             "f"(val.z), "f"(val.w) : "memory");
              } else {
                if (accumulate) {
                  const float4 old = *dst;
                  val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                }
                *dst = val;
              }
            }
          }
          __syncwarp();
        };
        if (nseg == 1) {
          // K <= 320: TMEM -> staging -> C directly, 32 columns at a time
          const uint32_t buf = gseg & 1;
          mbar_wait(&S.acc_full[buf], (gseg >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int cb = 0; cb < TN / 32; ++cb) {
            float hh[32], xx[32];
            const bool need = cb * 32 < ncols;
            if (need) tmem_ld32x2(tlane + buf * 256 + cb * 32, tlane + buf * 256 + 128 + cb * 32, hh, xx);
            if (cb == (ncols + 31) / 32 - 1) {  // exactly once per tile: every needed column has been read
              tc_fence_before();
              mbar_arrive(&S.acc_empty[buf]);
            }
            if (need) {
#pragma unroll
              for (int j = 0; j < 32; ++j) hh[j] += xx[j];
              emit(cb, hh);
            }
          }
          ++gseg;
        } else {
          // long reductions: drain each segment into registers (round-to-nearest partial sums)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[j] = 0.f;
          for (int sidx = 0; sidx < nseg; ++sidx, ++gseg) {
            const uint32_t buf = gseg & 1;
            mbar_wait(&S.acc_full[buf], (gseg >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int cb = 0; cb < TN / 32; ++cb) {
              float hh[32], xx[32];
              tmem_ld32x2(tlane + buf * 256 + cb * 32, tlane + buf * 256 + 128 + cb * 32, hh, xx);
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[cb * 32 + j] += hh[j] + xx[j];
            }
            tc_fence_before();
            mbar_arrive(&S.acc_empty[buf]);
          }
#pragma unroll
          for (int cb = 0; cb < TN / 32; ++cb)
            if (cb * 32 < ncols) emit(cb, acc + cb * 32);
        }
      }
    }
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  }
  if (warp == 8) {
    // =========================== weight-chunk loader ===================================================
    if (lane == 0) {
      uint32_t bcount[BSLOTS] = {0, 0, 0, 0};  // loads issued per slot
      uint32_t bit = 0;                         // ring position (streaming mode)
      auto load_slot = [&](int slot, const float* src, uint32_t bytes) {
        if (bcount[slot] > 0) mbar_wait(&S.b_empty[slot], (bcount[slot] - 1) & 1);
        mbar_expect_tx(&S.b_full[slot], 2 * bytes);
        bulk_g2s(S.b[slot], src, bytes, &S.b_full[slot]);
        bulk_g2s(S.b[slot] + TN * KC, src + TN * KC, bytes, &S.b_full[slot]);
        ++bcount[slot];
      };
      for (int q = sch.q; q < sch.nq_total; q += sch.q_step) {
        WorkQ w;
        decode_q(descs, ndesc, q, w);
        const int nrows = (w.ncols + 15) & ~15;  // MMA N (multiple of 16); prepared blocks are zero padded
        const uint32_t bytes = (uint32_t)nrows * KC * sizeof(float);
        const float* B = b_base + w.d->b_off + (int64_t)w.nt * w.kchunks * BLOCK_FLOATS;
        if (sch.m_start >= mtiles) continue;
        if (w.resident) {
          for (int c = 0; c < w.kchunks; ++c) load_slot(c, B + (int64_t)c * BLOCK_FLOATS, bytes);
        } else {
          for (int64_t mt = sch.m_start; mt < mtiles; mt += sch.m_step)
            for (int c = 0; c < w.kchunks; ++c, ++bit) load_slot(bit % BSLOTS, B + (int64_t)c * BLOCK_FLOATS, bytes);
        }
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ============================================================
    if (lane == 0) {
      constexpr uint32_t SBO = (KC / 4) * 128, LBO = 128;
      uint32_t it = 0, gseg = 0, bit = 0;
      uint32_t bcount[BSLOTS] = {0, 0, 0, 0};  // loads consumed (waited for) per slot
      for (int q = sch.q; q < sch.nq_total; q += sch.q_step) {
        WorkQ w;
        decode_q(descs, ndesc, q, w);
        const int nmma = (w.ncols + 15) & ~15;
        const uint32_t idesc = make_idesc(TM, nmma);
        bool first_mt = true;
        for (int64_t mt = sch.m_start; mt < mtiles; mt += sch.m_step) {
          for (int c0 = 0; c0 < w.kchunks; c0 += SEG_CHUNKS, ++gseg) {
            const uint32_t buf = gseg & 1;
            if (gseg >= 2) mbar_wait(&S.acc_empty[buf], ((gseg >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t d_hh = tmem + buf * 256, d_x = d_hh + 128;
            uint32_t acc_hh = 0, acc_x = 0;
            const int c1 = min(w.kchunks, c0 + SEG_CHUNKS);
            for (int c = c0; c < c1; ++c, ++it) {
              const uint32_t s = it % ASTAGES, ph = (it / ASTAGES) & 1;
              int slot;
              if (w.resident) {
                slot = c;
                if (first_mt) { mbar_wait(&S.b_full[slot], bcount[slot] & 1); ++bcount[slot]; }
              } else {
                slot = bit % BSLOTS;
                ++bit;
                mbar_wait(&S.b_full[slot], bcount[slot] & 1);
                ++bcount[slot];
              }
              mbar_wait(&S.a_full[s], ph);
              tc_fence_after();
              const uint32_t a_hi = smem_u32(S.a[s]), a_lo = a_hi + TM * KC * sizeof(float);
              const uint32_t b_hi = smem_u32(S.b[slot]), b_lo = b_hi + TN * KC * sizeof(float);
#pragma unroll
              for (int ks = 0; ks < KC / 8; ++ks) {
                umma_tf32(d_hh, make_desc(a_hi + ks * 256, LBO, SBO), make_desc(b_hi + ks * 256, LBO, SBO), idesc, acc_hh);
                acc_hh = 1;
              }
#pragma unroll
              for (int ks = 0; ks < KC / 8; ++ks) {
                umma_tf32(d_x, make_desc(a_lo + ks * 256, LBO, SBO), make_desc(b_hi + ks * 256, LBO, SBO), idesc, acc_x);
                acc_x = 1;
                umma_tf32(d_x, make_desc(a_hi + ks * 256, LBO, SBO), make_desc(b_lo + ks * 256, LBO, SBO), idesc, 1);
              }
              umma_commit(&S.a_empty[s]);
              if (!w.resident) umma_commit(&S.b_empty[slot]);
            }
            umma_commit(&S.acc_full[buf]);
          }
          first_mt = false;
        }
        // resident weights: release the slots once every MMA of this N-tile has retired
        if (w.resident && sch.m_start < mtiles)
          for (int c = 0; c < w.kchunks; ++c) umma_commit(&S.b_empty[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

// prepared layout: for n-tile j, k-chunk c: block (j * kchunks + c) of BLOCK_FLOATS floats = [hi | lo],
// each [128 rows (n) x 32 (k)] K-major canonical; rows >= N and k >= K are zero.
__global__ void k_gemm_prepare(const float* __restrict__ B, int64_t ldb, int K, int N, int transposed, float scale,
                               float* __restrict__ out, int kchunks, int ntiles) {
  const int64_t total = (int64_t)ntiles * kchunks * TN * KC;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % KC);
    const int nn = (int)((idx / KC) % TN);
    const int64_t blk = idx / (KC * TN);
    const int c = (int)(blk % kchunks), j = (int)(blk / kchunks);
    const int k = c * KC + kk, n = j * TN + nn;
    float v = 0.f;
    if (k < K && n < N) v = (transposed ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n]) * scale;
    const float hi = tf32_rn(v), lo = v - hi;
    float* tile = out + blk * BLOCK_FLOATS;
    const int off = canon_off(nn, kk, KC / 4);
    tile[off] = hi;
    tile[TN * KC + off] = lo;
  }
}

}  // namespace

extern "C" int nqb_set_error(const char* msg);
extern "C" void nqb_count_launch(void);

static int gemm_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

extern "C" int64_t nqb_gemm_prepared_floats(int K, int N) {
  const int64_t kchunks = (K + KC - 1) / KC, ntiles = (N + TN - 1) / TN;
  return kchunks * ntiles * BLOCK_FLOATS;
}

extern "C" int nqb_gemm_prepare(const float* B, int64_t ldb, int K, int N, int transposed, float scale, float* prepared,
                                nqb_stream_t st) {
  if (!B || !prepared) return nqb_set_error("nqb_gemm_prepare: null pointer");
  if (K <= 0 || N <= 0) return nqb_set_error("nqb_gemm_prepare: bad shape");
  const int kchunks = (K + KC - 1) / KC, ntiles = (N + TN - 1) / TN;
  const int64_t total = (int64_t)ntiles * kchunks * TN * KC;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  k_gemm_prepare<<<blocks, 256, 0, (cudaStream_t)st>>>(B, ldb, K, N, transposed, scale, prepared, kchunks, ntiles);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}

extern "C" int nqb_gemm_grouped(const void* descs_dev, int ndesc, int ntiles_total, const float* a_base,
                                const float* prepared_base, float* c_base, const float* rowscale_base, int64_t rs_ld,
                                int64_t M, nqb_stream_t st) {
  if (ndesc <= 0 || ntiles_total <= 0) return nqb_set_error("nqb_gemm_grouped: empty problem list");
  if (M < 0) return nqb_set_error("nqb_gemm_grouped: negative M");
  if (M == 0) return 0;
  if (!descs_dev || !a_base || !prepared_base || !c_base) return nqb_set_error("nqb_gemm_grouped: null pointer");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm3x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
    if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t nwork = ((M + TM - 1) / TM) * (int64_t)ntiles_total;
  const int grid = (int)(nwork < gemm_sm_count() ? nwork : gemm_sm_count());
  k_gemm3x<<<grid, 384, sizeof(Smem) + 1024, (cudaStream_t)st>>>((const GemmDesc*)descs_dev, ndesc, ntiles_total, a_base,
                                                                 prepared_base, c_base, rowscale_base, rs_ld, M);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}
