/*
 * nqb.h -- C ABI of the B200-native NequIP hot path (libnqb.so).
 *
 * Plain C: raw device pointers, sizes, a CUDA stream.  No torch / C++ types cross
 * this boundary.  Every tensor is caller-allocated (torch caching allocator on the
 * Python side) and only borrowed for the stream-ordered duration of the call; the
 * library never allocates device memory per call and never synchronises the
 * device.  All functions return 0 on success; otherwise nqb_last_error() holds a
 * thread-local message (the Python wrapper raises RuntimeError, in the style of
 * the reference's modifiers, nequip/nn/_tp_scatter_base.py:57-58).
 *
 * Reference interfaces these entry points replace (paths under /root/reference):
 *   nqb_tp_scatter_fwd/bwd  TensorProductScatter.forward + its autograd
 *                           nequip/nn/_tp_scatter_base.py:35-38
 *                           (e3nn o3.TensorProduct 'uvu' + nequip/nn/utils.py:24-53 scatter;
 *                            same seat as OpenEquivariance's TensorProductConv,
 *                            nequip/nn/_tp_scatter_oeq.py:29-57, and cuEquivariance's
 *                            fused_tp, nequip/nn/_tp_scatter_cueq.py:90-122)
 *   nqb_plan_create         TensorProductScatter.__init__  nequip/nn/_tp_scatter_base.py:10-33
 *                           (path table built at nequip/nn/interaction_block.py:89-116)
 *   nqb_csr_*               the (dst,src)-sorted edge contract of
 *                           nequip/data/transforms/neighborlist.py:120-157
 *   nqb_edge_embed_fwd/bwd  with_edge_vectors_  nequip/nn/utils.py:68-118,
 *                           SphericalHarmonicEdgeAttrs.forward  nequip/nn/embedding/_edge.py:193-198,
 *                           EdgeLengthNormalizer :65-80, BesselEdgeLengthEncoding :136-150,
 *                           PolynomialCutoff  nequip/nn/embedding/cutoffs.py:17-27,
 *                           ApplyFactor  nequip/nn/misc.py:46-48
 *   nqb_sh_fwd/bwd          e3nn o3.SphericalHarmonics(normalize=True, "component") as
 *                           constructed at nequip/nn/embedding/_edge.py:187-189
 */
#ifndef NQB_H
#define NQB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* nqb_stream_t; /* == cudaStream_t */
typedef struct nqb_plan nqb_plan;

typedef struct nqb_irrep {
  int32_t mul;
  int32_t l;
  int32_t p; /* +1 even, -1 odd */
} nqb_irrep;

typedef struct nqb_instruction {
  int32_t i_in1;
  int32_t i_in2;
  int32_t i_out; /* 'uvu', has_weight=True */
} nqb_instruction;

enum { NQB_F32 = 0, NQB_F64 = 1 };

/* library / error handling */
int nqb_abi_version(void);
const char* nqb_last_error(void);

/* Plan: immutable description of one TensorProductScatter signature bound to the
 * specialised kernel library generated for it (spec_lib_path, built by
 * nequip_b200.build; its embedded signature string is validated against the
 * descriptors).  Thread-safe to share between the forward and autograd threads. */
int nqb_plan_create(const nqb_irrep* in1, int n_in1, const nqb_irrep* in2, int n_in2,
                    const nqb_irrep* out, int n_out, const nqb_instruction* ins, int n_ins,
                    const char* spec_lib_path, nqb_plan** plan);
void nqb_plan_destroy(nqb_plan* plan);
int nqb_plan_dims(const nqb_plan* plan, int* d_in, int* s_dim, int* weight_numel, int* d_out);
/* writes the canonical signature (NUL terminated) into buf; returns needed length */
int nqb_plan_signature(const nqb_plan* plan, char* buf, int buflen);

/* Destination-CSR helpers.  edge_dst must be non-decreasing for
 * nqb_csr_from_sorted (the reference's neighbour lists are grouped by centre atom);
 * nqb_csr_check_sorted writes 1/0 into *flag_dev.  For unsorted edges the caller
 * supplies perm (stable argsort of edge_dst) and the sorted keys. */
int nqb_csr_check_sorted(const int64_t* keys, int64_t E, int32_t* flag_dev, nqb_stream_t st);
int nqb_csr_from_sorted(const int64_t* sorted_keys, int64_t E, int64_t N, int64_t* row_ptr /* [N+1] */,
                        nqb_stream_t st);

/* out[N, D_mid] = scatter_dst( TP_uvu( x[src], y, w ) ).  Every element of out is
 * written exactly once (no pre-zeroing needed, deterministic).
 *   x [N, D_in], y [E, S], w [E, W], out [N, D_mid]: dtype NQB_F32/NQB_F64, row-major, mul_ir layout
 *   row_ptr [N+1]: CSR over edges ordered by destination; perm [E] or NULL (identity):
 *   slot s of the CSR refers to edge perm[s]; src [E]: source node of each edge (original order). */
int nqb_tp_scatter_fwd(const nqb_plan* plan, int dtype, const void* x, const void* y, const void* w,
                       const int64_t* row_ptr, const int64_t* perm, const int64_t* src, int64_t N,
                       int64_t E, void* out, nqb_stream_t st);

/* Gradients of the above.  grad_w [E, W] is fully written.  grad_y [E, S] and
 * grad_x [N, D_in] are ACCUMULATED INTO (caller zero-fills); grad_x may be NULL
 * (skips the source-row reduction, e.g. first layer at inference). */
int nqb_tp_scatter_bwd(const nqb_plan* plan, int dtype, const void* x, const void* y, const void* w,
                       const int64_t* row_ptr, const int64_t* perm, const int64_t* src,
                       const void* grad_out, int64_t N, int64_t E, void* grad_x, void* grad_y,
                       void* grad_w, int deterministic,
                       nqb_stream_t st);
/* deterministic != 0 (bitwise-repeatable backward; the default accumulates grad_x / grad_Y with red.global.add in
 * whatever order the edges retire, as the reference's OpenEquivariance back-end does, nequip/nn/_tp_scatter_oeq.py:46):
 *   grad_x is then an [E, D_in] buffer -- every edge stores its contribution to its SOURCE atom in its own row -- to be
 *   reduced over the source-sorted edges with nqb_segment_sum (perm = stable argsort of edge_src, the
 *   edge_transpose_perm of nequip/data/transforms/neighborlist.py:150-155; seg_ptr = CSR over the sorted sources);
 *   grad_y is [nqb_tp_scatter_gy_slices(plan, dtype), E, S], zero-initialised by the caller: each writer owns a slice,
 *   the caller sums the slices in index order. */
int nqb_tp_scatter_gy_slices(const nqb_plan* plan, int dtype);
int nqb_segment_sum(int dtype, const void* rows /* [R, D] */, int D, const int64_t* perm, const int64_t* seg_ptr /* [N+1] */,
                    int64_t N, void* out /* [N, D] */, nqb_stream_t st);

/* Fused "last radial-MLP layer -> tensor product -> scatter" forward:
 *   out[n] = sum_{e: dst[e] = n} TP_uvu(x[src[e]], y[e], w[e]),   w[e] = h[e, :K] @ (W2 * alpha2)
 * i.e. nequip/nn/mlp.py:262-268 (the last ScalarLinearLayer built at nequip/nn/interaction_block.py:119-127)
 * composed with TensorProductScatter.forward (nequip/nn/_tp_scatter_base.py:35-38) so that the [E, W] weight
 * tensor is never written (w_out == NULL) -- or is written once on the side for an unfused backward.
 * float32, ir_mul node layout, edges grouped by destination (row_ptr; no permutation), K <= 128, K % 8 == 0.
 * nqb_tp_fused_slices(plan): number of 128-row weight slices of the signature, 0 = no fused kernel built.
 * w2_prepared: nqb_gemm_t_prepare() of the [K, 128 * slices] weight matrix whose COLUMNS are in slice order
 *   (nequip_b200/codegen.py TPGenerator.fused_layout()["cols"], -1 = zero column), scaled by alpha2.
 * slice_cta0 [slices + 1] (device, int32): CTAs [cta0[s], cta0[s+1]) process slice s; nctas = cta0[slices]
 *   (one CTA per SM; the host splits the grid in proportion to the slices' costs). */
int nqb_tp_fused_slices(const nqb_plan* plan);
int nqb_tp_fused_fwd(const nqb_plan* plan, const float* x, const float* y, const float* h, int64_t ldh, int K,
                     const float* w2_prepared, const int64_t* row_ptr, const int64_t* src, int64_t N, int64_t E,
                     float* out, float* w_out, const int32_t* slice_cta0, int nctas, nqb_stream_t st);

/* Real spherical harmonics, "component" normalisation, input normalised (lmax <= 3).
 *   vec [E,3] f64 -> y [E,(lmax+1)^2] of out_dtype (computed in f64, then cast). */
int nqb_sh_fwd(int lmax, const double* vec, int64_t E, int out_dtype, void* y, nqb_stream_t st);
/*   grad_vec [E,3] f64 = J^T grad_y (includes the normalisation Jacobian); overwritten */
int nqb_sh_bwd(int lmax, const double* vec, int64_t E, int out_dtype, const void* grad_y,
               double* grad_vec, nqb_stream_t st);

/* Fused edge geometry + embeddings:
 *   r_ij = pos[idx1] - pos[idx0] + shift @ cell ; Y = SH(r_ij) ;
 *   emb[:, n] = sinc(n x) n * f_cut(x) * prefactor , x = |r|/r_max, n = 1..num_bessel
 * edge_index [2,E] i64; shift [E,3] f64 or NULL; cell [3,3] f64 (rows = lattice vectors) or NULL.
 * Outputs: vec [E,3] f64 (kept for backward), y [E,S], emb [E,num_bessel] (out_dtype). */
int nqb_edge_embed_fwd(int lmax, int num_bessel, double r_max, double poly_p, double prefactor,
                       const double* pos, const int64_t* edge_index, const double* shift,
                       const double* cell, int64_t N, int64_t E, int out_dtype, double* vec, void* y,
                       void* emb, nqb_stream_t st);
/* grad_pos [N,3] f64 is ACCUMULATED INTO (caller zero-fills):
 *   g = J_Y^T grad_y + J_emb^T grad_emb ;  grad_pos[idx1] += g ; grad_pos[idx0] -= g.
 * grad_vec [E,3] f64 (may be NULL) receives g itself (per-edge forces / virial assembly). */
int nqb_edge_embed_bwd(int lmax, int num_bessel, double r_max, double poly_p, double prefactor,
                       const double* vec, const int64_t* edge_index, int64_t N, int64_t E,
                       int out_dtype, const void* grad_y, const void* grad_emb, double* grad_pos,
                       double* grad_vec, nqb_stream_t st);

/* Neighbour list on the device (cell list; full list, both directions, periodic images, no self edge in the home
 * image) -- replaces the host construction of nequip/data/_nl.py:60-152,292-361 and emits what
 * SortedNeighborListTransform (nequip/data/transforms/neighborlist.py:120-157) produces: edges sorted by
 * (centre, neighbour) = the destination CSR of the convolution.  Three calls around two host-side scans
 * (nequip_b200/ops.py neighbor_list): bins -> [sort atoms by bin] -> counts -> [exclusive scan] -> fill.
 * cell/inv: 3x3 row-major HOST arrays (rows = lattice vectors, inv = cell^-1); pbc/nbins/search: 3 ints;
 * lo/width: bounding box of the fractional coordinates in non-periodic directions (NULL if all periodic).
 * edge vector = pos[j] - pos[i] + shifts @ cell, i = edge_index[0][e] (centre), j = edge_index[1][e]. */
int nqb_nl_bin(const double* pos, int64_t N, const double* cell_host, const double* inv_host, const int* pbc,
               const int* nbins, const int* search, const double* lo, const double* width, double r_max,
               double* wpos /* [N,3] */, int32_t* base /* [N,3] */, int64_t* bin /* [N] */, int32_t* cidx /* [N,3] */,
               nqb_stream_t st);
int nqb_nl_count(int64_t N, const double* cell_host, const double* inv_host, const int* pbc, const int* nbins,
                 const int* search, double r_max, const double* wpos, const int32_t* cidx, const int64_t* order,
                 const int64_t* bin_start, int64_t* counts /* [N] */, nqb_stream_t st);
int nqb_nl_fill(int64_t N, int64_t E, const double* cell_host, const double* inv_host, const int* pbc, const int* nbins,
                const int* search, double r_max, const double* wpos, const int32_t* cidx, const int32_t* base,
                const int64_t* order, const int64_t* bin_start, const int64_t* row_ptr /* [N+1] */,
                int64_t* edge_index /* [2,E] */, double* shifts /* [E,3] */, nqb_stream_t st);

/* First radial layer (K = 8, CUDA cores):  h[E,128] = silu(emb[E,8] @ W1s[8,128])  and
 * grad_emb[E,8] = (grad_h * silu'(emb @ W1s)) @ W1s^T  (pre-activation recomputed, nothing saved).
 * Together with nqb_gemm_grouped for the second layer this is ScalarMLPFunction (nequip/nn/mlp.py:80-195). */
/* h_lo (nullable): the part of h the tensor core does not see, h_lo = rna_tf32(h - trunc_tf32(h)).  It can be
 * handed to nqb_gemm_grouped as a_lo_base, but that form doubles the A reads and limits the producer ring to one
 * piece in flight (measured 1.7x slower, profiles/r01_gemm_roles.txt): the product path passes NULL for both. */
int nqb_mlp_hidden_fwd(const float* emb, const float* W1s, int64_t E, int num_bessel, int hidden, float* h,
                       float* h_lo, nqb_stream_t st);
int nqb_mlp_hidden_bwd(const float* emb, const float* W1s, const float* grad_h, int64_t E, int num_bessel,
                       int hidden, float* grad_emb, nqb_stream_t st);
/* Kernel generation of the two calls above: 2 = batched kernels (32 edges per warp, prefetched basis values, packed
 * FFMA2, one 128-byte grad_emb row per four edges), 1 = the round-1 kernels (one edge per warp iteration).  Returns the
 * previous value; 0 only queries.  Default: the library's build-time choice, overridden by the environment variable
 * NQB_HIDDEN_VARIANT=1|2.  Not thread safe -- call between launches (A/B timing, parity tests). */
int nqb_mlp_hidden_set_variant(int variant);

/* Grouped fp32-accurate GEMM on the tensor cores (tcgen05 kind::tf32, 3xTF32, segmented fp32
 * accumulation):  C_p[M, N_p] (+)= rowscale_p[m] * A_p[M, K_p] @ B_p[K_p, N_p]  for a list of problems
 * sharing M.  Replaces the dense algebra around the convolution: ScalarMLPFunction's torch.mm
 * (nequip/nn/mlp.py:262-268), e3nn o3.Linear linear_1/linear_2 and the self-connection
 * FullyConnectedTensorProduct (nequip/nn/interaction_block.py:82-87,129-146) in the ir_mul layout.
 * descs_dev: device array of ndesc records of 12 int64:
 *   {a_off, c_off, b_off, rs_off (row of the [R, rs_ld] row-scale matrix, -1 = none), lda, ldc, K, N, kchunks=ceil(K/32), ntiles=ceil(N/128),
 *    tile0 (prefix sum of ntiles), flags};  offsets in floats from the bases.
 *   flags: bit0 C += (one writer per element within the launch), bit1 rows whose row scale is 0 are left
 *   untouched (disjoint row-masked writers), bit2 C += with red.global.add (several problems add into the same C).
 * Requirements: K, N, lda, ldc, a_off, c_off multiples of 4; bases 16-byte aligned.
 * B_p is prepared once (split hi/lo, tiled) with nqb_gemm_prepare into nqb_gemm_prepared_floats(K,N) floats.
 * tile_ctas_dev (nullable): int32 {first CTA, CTAs} per N-tile -- a cost-weighted split of sched_ctas CTAs over
 * the N-tiles computed by the host (problems of one launch differ in K, N and store mode); used when
 * sched_ctas <= #SMs, otherwise the even split is used.
 * a_lo_base (nullable): pre-split low parts of A, same offsets/strides as a_base (see nqb_mlp_hidden_fwd);
 * when null the kernel derives them itself. */
int64_t nqb_gemm_prepared_floats(int K, int N);
int nqb_gemm_prepare(const float* B, int64_t ldb, int K, int N, int transposed, float scale, float* prepared,
                     nqb_stream_t st);
int nqb_gemm_grouped(const void* descs_dev, int ndesc, int ntiles_total, const int32_t* tile_ctas_dev,
                     int sched_ctas, const float* a_base,
                     const float* a_lo_base, const float* prepared_base, float* c_base,
                     const float* rowscale_base, int64_t rs_ld, int64_t M, nqb_stream_t st);

/* EXPERIMENTAL (not used by the model yet): the same product for ONE problem with K <= 128, computed transposed
 * with the weights resident in tensor memory as the MMA's A operand (nequip_b200/csrc/nqb_gemm_t.cu).
 * prepared: nqb_gemm_t_prepared_floats(K, N) floats written by nqb_gemm_t_prepare. */
int64_t nqb_gemm_t_prepared_floats(int K, int N);
int nqb_gemm_t_prepare(const float* B, int64_t ldb, int K, int N, int transposed, float scale, float* prepared,
                       nqb_stream_t st);
int nqb_gemm_t_run(const float* prepared, int K, int N, const float* A, int64_t lda, float* C, int64_t ldc, int64_t M,
                   nqb_stream_t st);

/* Gate nonlinearity (e3nn nn.Gate with normalize2mom'd SiLU for even / tanh for odd scalars and gates,
 * nequip/nn/convnetlayer.py:42-56,104-112), one kernel per direction.  Column tables (device, int32) are
 * built by the host from the irreps, for either layout:
 *   forward, per OUTPUT column j: src[j], gate[j] (-1: scalar), kind[j] (0 silu, 1 tanh):
 *     out[n,j] = gate[j] < 0 ? act(x[n,src[j]]) : x[n,src[j]] * act(x[n,gate[j]])
 *   backward, per INPUT column i: tab[6*i..] = {role, a, b, c, d, kind}
 *     role 0 scalar (a = output column); role 1 gated value (a = output column, b = gate input column);
 *     role 2 gate (a = first output column, b = first gated input column, c = component stride, d = 2l+1). */
int nqb_gate_fwd(int dtype, const void* x, int64_t N, int d_in, int d_out, const int32_t* src,
                 const int32_t* gate, const int32_t* kind, void* out, nqb_stream_t st);
int nqb_gate_bwd(int dtype, const void* x, const void* grad_out, int64_t N, int d_in, int d_out,
                 const int32_t* tab, void* grad_x, nqb_stream_t st);

/* number of kernels the library has launched in this process (bench accounting) */
int64_t nqb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NQB_H */
