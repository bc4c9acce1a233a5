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
// segments of SEG_H*16 = 320 in K whose partial sums are added in registers (round-to-nearest).
//
// Roles per CTA (384 threads = 3 warpgroups with setmaxnreg-rebalanced registers, persistent over
// (M-tile, N-tile) work items; the CTAs working on one A tile run together and share it in L2):
//   warps 0-3   producers: stream the A chunk [128 x 32] through registers (coalesced 16-byte loads,
//               THREE chunks ahead across work-item boundaries = 48 KB in flight per SM: with one chunk
//               ahead the role was L2-latency bound at ~2600 cycles per chunk, profiles/r01_gemm_roles.txt),
//               split hi/lo into the canonical K-major core-matrix layout in shared memory
//   warps 4-7   epilogue: drain finished accumulator segments TMEM -> registers (fp32 adds), then write C
//               through a swizzled staging tile so that the global stores are full 128-byte rows;
//               overlaps with the production / MMA of the next work item (TMEM is double buffered)
//   warp  8     lane 0: cp.async.bulk of the pre-split weight chunk (+ mbarrier complete_tx)
//   warp  9     tcgen05.mma M=128, N<=128, K=8: 4 k-steps x 3 terms per chunk; tcgen05.commit
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/nqb.h"
#include "nqb_tc.cuh"

namespace {

// Per-role stall accounting (tools/bench_gemm.py --prof, build with -DNQB_GEMM_PROF): cycles CTA 0's
// lane 0 of each role spends in each mbarrier wait, and the role's total.
#ifdef NQB_GEMM_PROF
__device__ unsigned long long g_prof[16];
#define PROF_DECL long long pw_[4] = {0, 0, 0, 0}; const long long pt0_ = clock64();
#define PROF_WAIT(i, stmt) { const long long t_ = clock64(); stmt; pw_[i] += clock64() - t_; }
#define PROF_END(base) { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { \
    for (int i_ = 0; i_ < 3; ++i_) g_prof[(base) + i_] = (unsigned long long)pw_[i_]; \
    g_prof[(base) + 3] = (unsigned long long)(clock64() - pt0_); } }
#else
#define PROF_DECL
#define PROF_WAIT(i, stmt) { stmt; }
#define PROF_END(base) {}
#endif

constexpr int TM = 128;          // rows per tile
constexpr int TN = 128;          // max columns per tile
constexpr int KC = 32;           // K chunk
#ifndef NQB_KH
#define NQB_KH 32
#define NQB_RAW 3
#define NQB_NLO 2
#define NQB_DEPTH 2
#endif
constexpr int KH = NQB_KH;       // A is staged in pieces of KH k (16 or 32 = half or whole weight chunk)
constexpr int RAW = NQB_RAW;     // ring of raw A pieces (TM x KH fp32), filled by cp.async
constexpr int DEPTH = NQB_DEPTH; // pieces in flight per CTA
constexpr int NLO = NQB_NLO;     // ring of A low-part pieces
constexpr int HPC = 32 / KH;     // pieces per weight chunk
static_assert(KH == 16 || KH == 32, "KH");
constexpr bool PRESPLIT_OK = true;
constexpr int BSLOTS = 4;        // B slots: a ring when K > 128, resident per N-tile when K <= 128
constexpr int SEG_H = 320 / KH;  // pieces per accumulation segment (40 accumulate steps on the hi*hi accumulator)
#ifndef NQB_NPG
#define NQB_NPG 1
#endif
constexpr int NPW = 4;            // producer warps
constexpr int NPG = NQB_NPG;      // independent producer groups (pieces are dealt round-robin)
static_assert(DEPTH * NPG <= RAW - 1, "a group refills the stage of an already consumed piece");
static_assert(NLO >= 2 && NPW % NPG == 0, "rings");
constexpr int NPROD = NPW * 32;   // producer threads
constexpr int NTHREADS = NPROD + 256;  // + epilogue warpgroup + {loader, MMA, 2 idle} warps
constexpr int BLOCK_FLOATS = 2 * TN * KC;  // one prepared weight block: [hi | lo] x [128 x 32]

struct GemmDesc {  // mirrored by nequip_b200/ops.py (int64 fields)
  int64_t a_off, c_off, b_off, rs_off;  // element offsets from the base pointers; rs_off = ROW of the
                                        // [R, rs_ld] row-scale matrix (< 0: no row scale)
  int64_t lda, ldc, K, N;
  int64_t kchunks, ntiles, tile0, flags;  // flags: see the epilogue
};

struct Smem {
  float araw[RAW][TM * KH];        // 3 x 16 KB: fp32 A pieces, canonical K-major core-matrix layout
  float alo[NLO][TM * KH];         // 2 x 16 KB: their tf32 low parts
  float b[BSLOTS][BLOCK_FLOATS];   // 4 x 32 KB
  float stage[4][32 * 32];         // epilogue staging, one 32x32 tile per warp (swizzled)
  uint64_t a_full[RAW], a_done[RAW];
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
  // tile_ctas (nullable): per N-tile {first CTA, number of CTAs} of a cost-weighted split computed by the
  // host for a grid of exactly sched_ctas CTAs (problems of one launch differ in K, N and store mode, so
  // an even split leaves most CTAs idle while the expensive tiles finish)
  __device__ Sched(int b, int G, int T, const int32_t* tile_ctas, int sched_ctas) {
    nq_total = T;
    if (tile_ctas != nullptr && G == sched_ctas) {
      q = T; q_step = T; m_start = 0; m_step = 1;
      for (int t = 0; t < T; ++t) {
        const int c0 = tile_ctas[2 * t], n = tile_ctas[2 * t + 1];
        if (b >= c0 && b < c0 + n) { q = t; m_start = b - c0; m_step = n; break; }
      }
    } else if (T <= G) {
      const int R = G / T;
      q = (b < T * R) ? (b % T) : T;  // T = no work
      q_step = T;
      m_start = b / T;
      m_step = R;
    } else {
      q = b; q_step = G; m_start = 0; m_step = 1;
    }
  }
};

struct WorkQ {  // one N-tile of one problem
  const GemmDesc* d;
  int nt, kchunks, nh, K, ncols;  // kchunks: weight chunks of 32 k; nh: A half-chunks of 16 k
  bool resident;
};
__device__ __forceinline__ void decode_q(const GemmDesc* descs, int ndesc, int q, WorkQ& w) {
  w.d = find_desc(descs, ndesc, q);
  w.nt = q - (int)w.d->tile0;
  w.kchunks = (int)w.d->kchunks;
  w.K = (int)w.d->K;
  w.nh = (w.K + KH - 1) / KH;
  w.ncols = min(TN, (int)w.d->N - w.nt * TN);
  w.resident = w.kchunks <= BSLOTS;
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm3x(const GemmDesc* __restrict__ descs, int ndesc, int ntiles_total, const int32_t* __restrict__ tile_ctas,
         int sched_ctas, const float* __restrict__ a_base,
         const float* __restrict__ a_lo_base, const float* __restrict__ b_base, float* __restrict__ c_base, const float* __restrict__ rs_base, int64_t rs_ld,
         int64_t M) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t mtiles = (M + TM - 1) / TM;

  if (tid == 0) {
    for (int s = 0; s < RAW; ++s) { mbar_init(&S.a_full[s], NPROD / NPG); mbar_init(&S.a_done[s], 1); }
    for (int s = 0; s < BSLOTS; ++s) { mbar_init(&S.b_full[s], 1); mbar_init(&S.b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&S.acc_full[s], 1); mbar_init(&S.acc_empty[s], 128); }
    fence_barrier_init();
  }
  if (warp == NPW + 4) tmem_alloc(&S.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const Sched sch(blockIdx.x, gridDim.x, ntiles_total, tile_ctas, sched_ctas);

  // register budget per warpgroup (launch: 65536 / 384 = 168 each): the epilogue keeps a 128-value row of
  // partial sums in registers (232), producers, loader and MMA warps need few
  if (warp < NPW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // =========================== A producers ========================================================
    // NPG independent groups of NPW / NPG warps; group g owns the pieces i = g (mod NPG) of the CTA's flat
    // piece sequence, so the latency chains of consecutive pieces (copy -> read back -> split -> fence ->
    // arrive -> wait for a free stage -> next copy) overlap instead of adding up.
    // Per owned piece: cp.async (16 bytes = one core-matrix row) global -> raw stage, DEPTH own pieces ahead
    // and across work-item boundaries.  The fp32 tile itself is the tf32 high operand (the tensor core
    // ignores the low 13 mantissa bits); each thread then reads back ITS OWN 16-byte pieces and writes
    // lo = rna_tf32(a - trunc_tf32(a)).  (Register prefetching was stuck at ~7 B/clk/SM no matter how many
    // loads were nominally in flight, profiles/r01_gemm_roles.txt.)
    // thread -> rows gw*(8 RG) + g*8 + r8 (g < RG), k-groups kh*4 + kq
    constexpr int WPG = NPW / NPG;          // warps per group
    constexpr int RG = TM / WPG / 8;        // 8-row groups per warp
    constexpr int KQ = KH / 16;             // 16-k halves per piece
    const int grp = warp / WPG, gw = warp % WPG;
    const int r8 = lane & 7, kq = lane >> 3;
    PROF_DECL
    // cursor over the CTA's flat sequence of (N-tile q, M-tile mt, piece h); n = index of the cursor's piece
    int pq = sch.q, ph_ = 0, pnh = 1, pK = 0;
    int64_t pmt = sch.m_start, plda = 0;
    const float* pA = nullptr;
    const float* pAlo = nullptr;
    const bool presplit = PRESPLIT_OK && a_lo_base != nullptr;  // caller-supplied low parts: producers only copy
    bool pvalid = pq < sch.nq_total && sch.m_start < mtiles;
    auto open_q = [&]() {
      WorkQ w;
      decode_q(descs, ndesc, pq, w);
      pA = a_base + w.d->a_off;
      if (presplit) pAlo = a_lo_base + w.d->a_off;
      plda = w.d->lda;
      pnh = w.nh;
      pK = w.K;
    };
    if (pvalid) open_q();
    auto advance = [&]() {
      if (++ph_ == pnh) {
        ph_ = 0;
        pmt += sch.m_step;
        if (pmt >= mtiles) {
          pmt = sch.m_start;
          pq += sch.q_step;
          pvalid = pq < sch.nq_total;
          if (pvalid) open_q();
        }
      }
    };
    uint32_t n = 0, own_issued = 0;
    for (int j = 0; j < grp && pvalid; ++j) { advance(); ++n; }  // first own piece
    const int my_off = (gw * RG) * (KH / 4 * 32) + kq * 32 + r8 * 4;  // float offset of piece (g = 0, kh = 0)
    auto issue = [&]() {  // cp.async the cursor's (own) piece into its stage, move to the next own piece, commit
      if (pvalid) {
        float* dst = S.araw[n % RAW] + my_off;
        float* dlo = S.alo[n % NLO] + my_off;
#pragma unroll
        for (int g = 0; g < RG; ++g) {
          const int64_t m = pmt * TM + gw * (8 * RG) + g * 8 + r8;
#pragma unroll
          for (int kh = 0; kh < KQ; ++kh) {
            const int k = ph_ * KH + (kh * 4 + kq) * 4;
            const bool in = m < M && k < pK;
            const int64_t off = in ? m * plda + k : 0;
#ifndef NQB_X_NOLOAD
            cp_async16(dst + g * (KH / 4 * 32) + kh * 128, pA + off, in ? 16u : 0u);
#endif
            if (presplit) cp_async16(dlo + g * (KH / 4 * 32) + kh * 128, pAlo + off, in ? 16u : 0u);
          }
        }
        ++own_issued;
#pragma unroll 1
        for (int j = 0; j < NPG && pvalid; ++j) { advance(); ++n; }
      }
      cp_async_commit();
    };
#pragma unroll 1
    for (int j = 0; j < DEPTH; ++j) issue();
#pragma unroll 1
    for (uint32_t k = 0; k < own_issued; ++k) {
      const uint32_t i = grp + k * NPG;  // global index of this own piece
      PROF_WAIT(1, cp_async_wait<DEPTH - 1>())  // my parts of piece i have landed
#ifndef NQB_X_NOLO
      if (!presplit) {
        // lo stage i % NLO was read by the MMAs of piece i - NLO
        if (i >= NLO) PROF_WAIT(0, mbar_wait(&S.a_done[(i - NLO) % RAW], ((i - NLO) / RAW) & 1))
        const float* raw = S.araw[i % RAW] + my_off;
        float* lo = S.alo[i % NLO] + my_off;
#pragma unroll
        for (int g0 = 0; g0 < RG; g0 += 4) {
          float4 a[4 * KQ];
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int kh = 0; kh < KQ; ++kh)
              a[g * KQ + kh] = *reinterpret_cast<const float4*>(raw + (g0 + g) * (KH / 4 * 32) + kh * 128);
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int kh = 0; kh < KQ; ++kh) {
              const float4 t = a[g * KQ + kh];
              *reinterpret_cast<float4*>(lo + (g0 + g) * (KH / 4 * 32) + kh * 128) =
                  make_float4(tf32_lo(t.x), tf32_lo(t.y), tf32_lo(t.z), tf32_lo(t.w));
            }
        }
      }
#endif
      PROF_WAIT(2, fence_proxy_async(); mbar_arrive(&S.a_full[i % RAW]))
      // the next own piece to load, j = i + DEPTH * NPG, reuses the stages of pieces j - RAW and j - NLO
      const uint32_t j = i + DEPTH * NPG;
      if (j >= RAW) PROF_WAIT(0, mbar_wait(&S.a_done[(j - RAW) % RAW], ((j - RAW) / RAW) & 1))
      if (presplit && NLO != RAW && j >= NLO) mbar_wait(&S.a_done[(j - NLO) % RAW], ((j - NLO) / RAW) & 1);
      issue();
    }
    cp_async_wait<0>();
    if (warp == 0) PROF_END(0)
  } else if (warp < NPW + 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // =========================== epilogue: TMEM segments -> registers -> C ============================
    const int ew = warp - NPW;         // TMEM lane quadrant == warp % 4
    const int row = ew * 32 + lane;    // accumulator row owned by this thread
    uint32_t gseg = 0;
    float acc[TN];
    PROF_DECL
    for (int q = sch.q; q < sch.nq_total; q += sch.q_step) {
      WorkQ w;
      decode_q(descs, ndesc, q, w);
      const GemmDesc* d = w.d;
      const int nseg = (w.nh + SEG_H - 1) / SEG_H;
      const int ncols = w.ncols;
      const int64_t ldc = d->ldc;
      float* C = c_base + d->c_off + (int64_t)w.nt * TN;
      // flags: bit0 read-modify-write accumulate (single writer per element within the launch),
      //        bit1 rows whose row scale is zero are not touched (disjoint row-masked writers),
      //        bit2 accumulate with red.global.add (several problems of this launch add into the same C)
      // both accumulate modes add with red.global (performed at L2)
      const bool reduce = (d->flags & 5) != 0, skipz = (d->flags & 2) != 0;
      for (int64_t mt = sch.m_start; mt < mtiles; mt += sch.m_step) {
        const int64_t m0 = mt * TM;
        float rs = 1.0f;
        if (d->rs_off >= 0) {
          const int64_t m = m0 + row;
          rs = (m < M) ? __ldg(rs_base + d->rs_off * rs_ld + m) : 0.f;
        }
        const uint32_t tlane = tmem + ((uint32_t)(ew * 32) << 16);
        // 32 summed + scaled values of my row -> swizzled staging tile -> full 128-byte row segments of C
        // (lane -> row p*4 + lane/8, 16-byte column group lane%8).  Everything that does not depend on p is
        // hoisted: a branchy first version of this block cost ~2000 cycles per call and bound the K = 128
        // problems (profiles/r01_gemm_roles.txt).
        float4* st = reinterpret_cast<float4*>(S.stage[ew]);
        const int sub = lane >> 3, u = lane & 7;
        const uint32_t rowmask = __ballot_sync(0xffffffffu, (m0 + row < M) && !(skipz && rs == 0.f));
        float* c_lane = C + (m0 + ew * 32 + sub) * ldc + u * 4;  // row `sub` of this warp's 32, column group u
        const int64_t pstep = 4 * ldc;
        // MODE 0: plain stores, 1: reduce-adds, 2: decided at run time (one short branch per store)
        auto emit = [&](auto mode_tag, int cb, const float* v) {
          constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            st[lane * 8 + (k ^ (lane & 7))] = make_float4(v[4 * k] * rs, v[4 * k + 1] * rs, v[4 * k + 2] * rs, v[4 * k + 3] * rs);
          __syncwarp();
          const bool colok = cb * 32 + u * 4 < ncols;
          float* dst = c_lane + cb * 32;
          float4 val[8];
#pragma unroll
          for (int p = 0; p < 8; ++p) val[p] = st[(p * 4 + sub) * 8 + (u ^ ((p * 4 + sub) & 7))];
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const bool ok = colok && ((rowmask >> (p * 4 + sub)) & 1u);
            float* d4 = dst + p * pstep;
            if (ok) {
              if (MODE == 1 || (MODE == 2 && reduce))
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d4), "f"(val[p].x), "f"(val[p].y),
                             "f"(val[p].z), "f"(val[p].w) : "memory");
              else
                *reinterpret_cast<float4*>(d4) = val[p];
            }
          }
          __syncwarp();
        };
        if (nseg == 1) {
          // K <= 320: TMEM -> staging -> C directly, 32 columns at a time
          const uint32_t buf = gseg & 1;
          PROF_WAIT(0, mbar_wait(&S.acc_full[buf], (gseg >> 1) & 1))
          tc_fence_after();
          const int ncb = (ncols + 31) / 32;
#pragma unroll 1
          for (int cb = 0; cb < ncb; ++cb) {
            float hh[32], xx[32];
            PROF_WAIT(1, tmem_ld32x2(tlane + buf * 256 + cb * 32, tlane + buf * 256 + 128 + cb * 32, hh, xx))
            if (cb == ncb - 1) {  // exactly once per tile: every needed column has been read
              tc_fence_before();
              mbar_arrive(&S.acc_empty[buf]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) hh[j] += xx[j];
#ifndef NQB_X_NOEMIT
            if (reduce) PROF_WAIT(2, emit(std::integral_constant<int, 1>{}, cb, hh))
            else PROF_WAIT(2, emit(std::integral_constant<int, 0>{}, cb, hh))
#else
            if (hh[0] == 123.456f) emit(std::integral_constant<int, 0>{}, cb, hh);
#endif
          }
          ++gseg;
        } else {
          // long reductions: drain each segment into registers (round-to-nearest partial sums)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[j] = 0.f;
          for (int sidx = 0; sidx < nseg; ++sidx, ++gseg) {
            const uint32_t buf = gseg & 1;
            PROF_WAIT(0, mbar_wait(&S.acc_full[buf], (gseg >> 1) & 1))
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
            if (cb * 32 < ncols) emit(std::integral_constant<int, 2>{}, cb, acc + cb * 32);
        }
      }
    }
    if (warp == NPW) PROF_END(4)
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  }
  if (warp == NPW + 4) {
    // =========================== weight-chunk loader ===================================================
    if (lane == 0) {
      uint32_t bused = 0, bpar = 0;             // per-slot bit: slot loaded before / parity of its load count
      uint32_t bit = 0;                         // ring position (streaming mode)
      auto load_slot = [&](int slot, const float* src, uint32_t bytes) {
        if ((bused >> slot) & 1) mbar_wait(&S.b_empty[slot], ((bpar >> slot) & 1) ^ 1);
        mbar_expect_tx(&S.b_full[slot], 2 * bytes);
        bulk_g2s(S.b[slot], src, bytes, &S.b_full[slot]);
        bulk_g2s(S.b[slot] + TN * KC, src + TN * KC, bytes, &S.b_full[slot]);
        bused |= 1u << slot;
        bpar ^= 1u << slot;
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
  } else if (warp == NPW + 5) {
    // =========================== MMA issuer ============================================================
    // The tensor pipe retires an M128 x N128 x K8 MMA every 64 cycles but one thread issues them, and every
    // instruction between two tcgen05.mma costs issue latency (tools/microbench/umma.cu: a branchy
    // 40-instruction body drops the rate to ~250 cycles/MMA).  So: the WHOLE warp runs the (warp-uniform)
    // loops, which lets the compiler keep descriptors in uniform registers instead of broadcasting them
    // from lane 0 around each MMA; one elected lane issues; descriptors are built once and advanced by adds.
    const bool leader = elect_one();
    constexpr uint32_t SBO_A = (KH / 4) * 128, SBO_B = (KC / 4) * 128, LBO = 128;
    const uint64_t dR0 = make_desc(smem_u32(S.araw[0]), LBO, SBO_A);
    const uint64_t dL0 = make_desc(smem_u32(S.alo[0]), LBO, SBO_A);
    const uint64_t dB0 = make_desc(smem_u32(S.b[0]), LBO, SBO_B);
    constexpr uint32_t A_STAGE = (TM * KH * sizeof(float)) >> 4;
    constexpr uint32_t B_SLOT = (BLOCK_FLOATS * sizeof(float)) >> 4, B_LO = (TN * KC * sizeof(float)) >> 4;
    constexpr uint32_t B_HALF = (KH / 4 * 128) >> 4;  // the second piece of a weight chunk (KH = 16)
    uint32_t gseg = 0, bit = 0;
    uint32_t s = 0, ph = 0, s_lo = 0;  // raw ring stage + phase, lo ring stage of the next piece
    uint32_t bpar = 0;  // per-slot bit: parity of the loads consumed (waited for)
    PROF_DECL
    for (int q = sch.q; q < sch.nq_total; q += sch.q_step) {
      WorkQ w;
      decode_q(descs, ndesc, q, w);
      const int nmma = (w.ncols + 15) & ~15;
      const uint32_t idesc = make_idesc(TM, nmma);
      bool first_mt = true;
      for (int64_t mt = sch.m_start; mt < mtiles; mt += sch.m_step) {
        uint32_t slot = 0;
        for (int h0 = 0; h0 < w.nh; h0 += SEG_H, ++gseg) {
          const uint32_t buf = gseg & 1;
          if (gseg >= 2) PROF_WAIT(0, mbar_wait(&S.acc_empty[buf], ((gseg >> 1) - 1) & 1))
          tc_fence_after();
          const uint32_t d_hh = tmem + buf * 256, d_x = d_hh + 128;
          uint32_t fresh = 1;  // the first half-chunk of a segment overwrites the accumulators
          const int h1 = min(w.nh, h0 + SEG_H);
          for (int h = h0; h < h1; ++h) {
            const uint32_t half = h % HPC;
            if (!half) {  // first piece of a weight chunk
              if (w.resident) {
                slot = h / HPC;
                if (first_mt) { PROF_WAIT(1, mbar_wait(&S.b_full[slot], (bpar >> slot) & 1)) bpar ^= 1u << slot; }
              } else {
                slot = bit % BSLOTS;
                ++bit;
                PROF_WAIT(1, mbar_wait(&S.b_full[slot], (bpar >> slot) & 1))
                bpar ^= 1u << slot;
              }
            }
            const uint64_t a_hi = dR0 + (uint64_t)(s * A_STAGE), a_lo = dL0 + (uint64_t)(s_lo * A_STAGE);
            const uint64_t b_hi = dB0 + (uint64_t)(slot * B_SLOT + half * B_HALF), b_lo = b_hi + B_LO;
            // (the stage was filled through the generic proxy and fenced by the producers: no tcgen05 fence)
            PROF_WAIT(2, mbar_wait(&S.a_full[s], ph))
            if (leader) {
#ifndef NQB_X_NOMMA
              // k-step advance = 2 core matrices = 256 bytes = 16 descriptor units.  MMAs on the same
              // accumulator are issued back to back: switching accumulators costs tensor-pipe time
#ifndef NQB_X_ORDER
#define NQB_X_ORDER 1
#endif
#if NQB_X_ORDER == 0
              umma_tf32(d_hh, a_hi, b_hi, idesc, fresh ^ 1);
              umma_tf32(d_x, a_lo, b_hi, idesc, fresh ^ 1);
              umma_tf32_acc(d_x, a_hi, b_lo, idesc);
#pragma unroll
              for (int ks = 1; ks < KH / 8; ++ks) {
                umma_tf32_acc(d_hh, a_hi + ks * 16, b_hi + ks * 16, idesc);
                umma_tf32_acc(d_x, a_lo + ks * 16, b_hi + ks * 16, idesc);
                umma_tf32_acc(d_x, a_hi + ks * 16, b_lo + ks * 16, idesc);
              }
#else
              const uint32_t d_x2 = (NQB_X_ORDER == 2) ? d_hh : d_x;  // ORDER 2: timing experiment only
              umma_tf32(d_hh, a_hi, b_hi, idesc, fresh ^ 1);
#pragma unroll
              for (int ks = 1; ks < KH / 8; ++ks) umma_tf32_acc(d_hh, a_hi + ks * 16, b_hi + ks * 16, idesc);
              umma_tf32(d_x2, a_lo, b_hi, idesc, fresh ^ 1);
#pragma unroll
              for (int ks = 1; ks < KH / 8; ++ks) umma_tf32_acc(d_x2, a_lo + ks * 16, b_hi + ks * 16, idesc);
#pragma unroll
              for (int ks = 0; ks < KH / 8; ++ks) umma_tf32_acc(d_x2, a_hi + ks * 16, b_lo + ks * 16, idesc);
#endif
#endif
              umma_commit(&S.a_done[s]);
              if (!w.resident && (half == HPC - 1 || h == w.nh - 1)) umma_commit(&S.b_empty[slot]);
            }
            fresh = 0;
            if (++s == RAW) { s = 0; ph ^= 1; }
            if (++s_lo == NLO) s_lo = 0;
          }
          if (leader) umma_commit(&S.acc_full[buf]);
          __syncwarp();
        }
        first_mt = false;
      }
      // resident weights: release the slots once every MMA of this N-tile has retired
      if (w.resident && sch.m_start < mtiles && leader)
        for (int c = 0; c < w.kchunks; ++c) umma_commit(&S.b_empty[c]);
      __syncwarp();
    }
    PROF_END(8)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPW + 4) tmem_dealloc(tmem, 512);
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
#ifdef NQB_GEMM_PROF
extern "C" int nqb_gemm_prof_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * 16);
}
#endif

// per-device state (one process may drive several GPUs: attributes and SM counts are per device)
static int gemm_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 63;
}
static int gemm_sm_count() {
  static int n[64] = {0};
  const int dev = gemm_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
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

extern "C" int nqb_gemm_grouped(const void* descs_dev, int ndesc, int ntiles_total, const int32_t* tile_ctas_dev,
                                int sched_ctas, const float* a_base,
                                const float* a_lo_base, const float* prepared_base, float* c_base,
                                const float* rowscale_base, int64_t rs_ld, int64_t M, nqb_stream_t st) {
  if (ndesc <= 0 || ntiles_total <= 0) return nqb_set_error("nqb_gemm_grouped: empty problem list");
  if (M < 0) return nqb_set_error("nqb_gemm_grouped: negative M");
  if (M == 0) return 0;
  if (!descs_dev || !a_base || !prepared_base || !c_base) return nqb_set_error("nqb_gemm_grouped: null pointer");
  static bool attr_set[64] = {false};
  const int dev = gemm_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm3x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
    if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  const int64_t nwork = ((M + TM - 1) / TM) * (int64_t)ntiles_total;
  int grid = (int)(nwork < gemm_sm_count() ? nwork : gemm_sm_count());
  if (tile_ctas_dev != nullptr && sched_ctas > 0 && sched_ctas <= gemm_sm_count()) grid = sched_ctas;
  else tile_ctas_dev = nullptr;
  k_gemm3x<<<grid, 384, sizeof(Smem) + 1024, (cudaStream_t)st>>>((const GemmDesc*)descs_dev, ndesc, ntiles_total, tile_ctas_dev, sched_ctas, a_base, a_lo_base,
                                                                 prepared_base, c_base, rowscale_base, rs_ld, M);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}
