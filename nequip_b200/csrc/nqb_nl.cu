// Neighbour list on the GPU (cell list, full list, periodic images), sm_100a.        SURVEY.md section 8(f)-2
//
// Reference contract (paths under /root/reference):
//   compute_neighborlist_ / backends            nequip/data/_nl.py:60-152, 292-361 -- full list (both directions), no
//     self interaction in the home image, edge vector = pos[j] - pos[i] + shift @ cell, shifts integer-valued;
//     the reference builds it on the host from pos.detach().cpu().numpy() (a serial bottleneck at >= 10k atoms)
//   SortedNeighborListTransform                 nequip/data/transforms/neighborlist.py:120-157 -- edges sorted by
//     (centre, neighbour) and the permutation to the (neighbour, centre) order
// Output here: edges grouped by centre i (= edge_index[0], the scatter destination of the convolution) and sorted
// by neighbour j inside a row -- i.e. the destination CSR the TP kernels want, written directly in device memory.
//
// Two passes over the 27 (or more, for small cells) neighbouring bins of every atom: count, exclusive scan (host
// side: torch.cumsum), fill + in-row sort.  All arithmetic that decides membership (wrapped coordinates, image
// offsets, squared distance) is done in the same order as the host reference list (nequip_b200/data.py) with
// explicitly rounded operations (no FMA contraction), so the edge set is bit-identical.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nqb.h"

extern "C" int nqb_set_error(const char* msg);
extern "C" void nqb_count_launch(void);

namespace {

struct NlParams {
  double cell[9];     // rows = lattice vectors
  double inv[9];      // inverse (columns give fractional coordinates: frac = pos @ inv)
  double diag[3];     // diagonal of an orthorhombic cell (frac = pos / diag, exactly as the host list does)
  int orthorhombic;
  int pbc[3];
  int nb[3];          // bins per direction
  int sr[3];          // bin search range per direction
  double lo[3], width[3];  // non-periodic directions: bounding box origin / extent in fractional units
  double r2;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// wrapped cartesian position, integer base shift (pos + base @ cell is the wrapped position) and bin of one atom
__global__ void k_nl_bin(NlParams p, const double* __restrict__ pos, int64_t N, double* __restrict__ wpos,
                         int32_t* __restrict__ base, int64_t* __restrict__ bin, int32_t* __restrict__ cidx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  double f[3];
  if (p.orthorhombic) {
    f[0] = __ddiv_rn(x, p.diag[0]); f[1] = __ddiv_rn(y, p.diag[1]); f[2] = __ddiv_rn(z, p.diag[2]);
  } else {
    for (int d = 0; d < 3; ++d) f[d] = dadd(dadd(dmul(x, p.inv[d]), dmul(y, p.inv[3 + d])), dmul(z, p.inv[6 + d]));
  }
  double w[3];
  int c[3];
  for (int d = 0; d < 3; ++d) {
    if (p.pbc[d]) {
      const double fl = floor(f[d]);
      w[d] = dadd(f[d], -fl);
      base[3 * i + d] = (int32_t)(-fl);
      int q = (int)dmul(w[d], (double)p.nb[d]);
      c[d] = q < p.nb[d] - 1 ? q : p.nb[d] - 1;
    } else {
      w[d] = f[d];
      base[3 * i + d] = 0;
      int q = (int)floor(dmul(__ddiv_rn(dadd(f[d], -p.lo[d]), p.width[d]), (double)p.nb[d]));
      c[d] = q < 0 ? 0 : (q < p.nb[d] - 1 ? q : p.nb[d] - 1);
    }
    cidx[3 * i + d] = c[d];
  }
  if (p.orthorhombic) {
    wpos[3 * i] = dmul(w[0], p.diag[0]); wpos[3 * i + 1] = dmul(w[1], p.diag[1]); wpos[3 * i + 2] = dmul(w[2], p.diag[2]);
  } else {
    for (int d = 0; d < 3; ++d)
      wpos[3 * i + d] = dadd(dadd(dmul(w[0], p.cell[d]), dmul(w[1], p.cell[3 + d])), dmul(w[2], p.cell[6 + d]));
  }
  bin[i] = ((int64_t)c[2] * p.nb[1] + c[1]) * p.nb[0] + c[0];
}

// visit every (neighbour atom j, image) candidate of atom i; F(j, img[3], within cutoff)
template <class F>
__device__ __forceinline__ void nl_visit(const NlParams& p, int64_t i, const double* __restrict__ wpos,
                                         const int32_t* __restrict__ cidx, const int64_t* __restrict__ order,
                                         const int64_t* __restrict__ bin_start, F&& f) {
  const double xi = wpos[3 * i], yi = wpos[3 * i + 1], zi = wpos[3 * i + 2];
  const int c0 = cidx[3 * i], c1 = cidx[3 * i + 1], c2 = cidx[3 * i + 2];
  for (int oz = -p.sr[2]; oz <= p.sr[2]; ++oz) {
    int bz = c2 + oz, iz = 0;
    if (p.pbc[2]) { iz = (bz >= 0) ? bz / p.nb[2] : -((-bz + p.nb[2] - 1) / p.nb[2]); bz -= iz * p.nb[2]; }
    else if (bz < 0 || bz >= p.nb[2]) continue;
    for (int oy = -p.sr[1]; oy <= p.sr[1]; ++oy) {
      int by = c1 + oy, iy = 0;
      if (p.pbc[1]) { iy = (by >= 0) ? by / p.nb[1] : -((-by + p.nb[1] - 1) / p.nb[1]); by -= iy * p.nb[1]; }
      else if (by < 0 || by >= p.nb[1]) continue;
      for (int ox = -p.sr[0]; ox <= p.sr[0]; ++ox) {
        int bx = c0 + ox, ix = 0;
        if (p.pbc[0]) { ix = (bx >= 0) ? bx / p.nb[0] : -((-bx + p.nb[0] - 1) / p.nb[0]); bx -= ix * p.nb[0]; }
        else if (bx < 0 || bx >= p.nb[0]) continue;
        // image offset in cartesian coordinates: img @ cell (exact for an orthorhombic cell)
        double sx, sy, sz;
        if (p.orthorhombic) {
          sx = dmul((double)ix, p.diag[0]); sy = dmul((double)iy, p.diag[1]); sz = dmul((double)iz, p.diag[2]);
        } else {
          sx = dadd(dadd(dmul((double)ix, p.cell[0]), dmul((double)iy, p.cell[3])), dmul((double)iz, p.cell[6]));
          sy = dadd(dadd(dmul((double)ix, p.cell[1]), dmul((double)iy, p.cell[4])), dmul((double)iz, p.cell[7]));
          sz = dadd(dadd(dmul((double)ix, p.cell[2]), dmul((double)iy, p.cell[5])), dmul((double)iz, p.cell[8]));
        }
        const int64_t b = ((int64_t)bz * p.nb[1] + by) * p.nb[0] + bx;
        const bool home = (ix == 0) && (iy == 0) && (iz == 0);
        for (int64_t q = bin_start[b]; q < bin_start[b + 1]; ++q) {
          const int64_t j = order[q];
          if (home && j == i) continue;
          const double dx = dadd(dadd(wpos[3 * j], sx), -xi), dy = dadd(dadd(wpos[3 * j + 1], sy), -yi),
                       dz = dadd(dadd(wpos[3 * j + 2], sz), -zi);
          const double d2 = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
          if (d2 < p.r2) f(j, ix, iy, iz);
        }
      }
    }
  }
}

__global__ void k_nl_count(NlParams p, int64_t N, const double* __restrict__ wpos, const int32_t* __restrict__ cidx,
                           const int64_t* __restrict__ order, const int64_t* __restrict__ bin_start,
                           int64_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int64_t n = 0;
  nl_visit(p, i, wpos, cidx, order, bin_start, [&](int64_t, int, int, int) { ++n; });
  counts[i] = n;
}

__device__ __forceinline__ bool nl_less(int64_t ja, const double* sa, int64_t jb, const double* sb) {
  if (ja != jb) return ja < jb;
  if (sa[0] != sb[0]) return sa[0] < sb[0];
  if (sa[1] != sb[1]) return sa[1] < sb[1];
  return sa[2] < sb[2];
}

__global__ void k_nl_fill(NlParams p, int64_t N, const double* __restrict__ wpos, const int32_t* __restrict__ cidx,
                          const int32_t* __restrict__ base, const int64_t* __restrict__ order,
                          const int64_t* __restrict__ bin_start, const int64_t* __restrict__ row_ptr, int64_t E,
                          int64_t* __restrict__ edge_index, double* __restrict__ shifts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t beg = row_ptr[i];
  int64_t n = 0;
  int64_t* ej = edge_index + E;  // neighbours (row 1)
  nl_visit(p, i, wpos, cidx, order, bin_start, [&](int64_t j, int ix, int iy, int iz) {
    // insertion into the sorted prefix of the row (rows hold a few dozen neighbours)
    double s[3] = {(double)(ix + base[3 * j] - base[3 * i]), (double)(iy + base[3 * j + 1] - base[3 * i + 1]),
                   (double)(iz + base[3 * j + 2] - base[3 * i + 2])};
    int64_t q = beg + n;
    while (q > beg && nl_less(j, s, ej[q - 1], shifts + 3 * (q - 1))) {
      ej[q] = ej[q - 1];
      shifts[3 * q] = shifts[3 * (q - 1)]; shifts[3 * q + 1] = shifts[3 * (q - 1) + 1]; shifts[3 * q + 2] = shifts[3 * (q - 1) + 2];
      --q;
    }
    ej[q] = j;
    shifts[3 * q] = s[0]; shifts[3 * q + 1] = s[1]; shifts[3 * q + 2] = s[2];
    edge_index[beg + n] = i;
    ++n;
  });
}

}  // namespace

// Step 1: bins.  cell/inv: row-major 3x3 on the HOST (9 doubles each); nbins/search: per direction.
extern "C" int nqb_nl_bin(const double* pos, int64_t N, const double* cell_host, const double* inv_host, const int* pbc,
                          const int* nbins, const int* search, const double* lo, const double* width, double r_max,
                          double* wpos, int32_t* base, int64_t* bin, int32_t* cidx, nqb_stream_t st) {
  if (N < 0) return nqb_set_error("nqb_nl_bin: negative size");
  if (N == 0) return 0;
  if (!pos || !cell_host || !inv_host || !pbc || !nbins || !search || !wpos || !base || !bin || !cidx)
    return nqb_set_error("nqb_nl_bin: null pointer");
  NlParams p;
  bool ortho = true;
  for (int k = 0; k < 9; ++k) { p.cell[k] = cell_host[k]; p.inv[k] = inv_host[k]; if ((k % 4) != 0 && cell_host[k] != 0.0) ortho = false; }
  p.orthorhombic = ortho ? 1 : 0;
  for (int d = 0; d < 3; ++d) {
    p.diag[d] = cell_host[4 * d]; p.pbc[d] = pbc[d]; p.nb[d] = nbins[d]; p.sr[d] = search[d];
    p.lo[d] = lo ? lo[d] : 0.0; p.width[d] = width ? width[d] : 1.0;
    if (nbins[d] < 1 || search[d] < 0) return nqb_set_error("nqb_nl_bin: bad bin grid");
  }
  p.r2 = r_max * r_max;
  k_nl_bin<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)st>>>(p, pos, N, wpos, base, bin, cidx);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}

static int nl_params(const double* cell_host, const double* inv_host, const int* pbc, const int* nbins, const int* search,
                     double r_max, NlParams& p) {
  bool ortho = true;
  for (int k = 0; k < 9; ++k) { p.cell[k] = cell_host[k]; p.inv[k] = inv_host[k]; if ((k % 4) != 0 && cell_host[k] != 0.0) ortho = false; }
  p.orthorhombic = ortho ? 1 : 0;
  for (int d = 0; d < 3; ++d) {
    p.diag[d] = cell_host[4 * d]; p.pbc[d] = pbc[d]; p.nb[d] = nbins[d]; p.sr[d] = search[d]; p.lo[d] = 0.0; p.width[d] = 1.0;
  }
  p.r2 = r_max * r_max;
  return 0;
}

// Step 2: neighbours per atom.  order = atom ids sorted by bin, bin_start [nbins + 1].
extern "C" int nqb_nl_count(int64_t N, const double* cell_host, const double* inv_host, const int* pbc, const int* nbins,
                            const int* search, double r_max, const double* wpos, const int32_t* cidx,
                            const int64_t* order, const int64_t* bin_start, int64_t* counts, nqb_stream_t st) {
  if (N <= 0) return 0;
  if (!wpos || !cidx || !order || !bin_start || !counts) return nqb_set_error("nqb_nl_count: null pointer");
  NlParams p;
  nl_params(cell_host, inv_host, pbc, nbins, search, r_max, p);
  k_nl_count<<<(unsigned)((N + 63) / 64), 64, 0, (cudaStream_t)st>>>(p, N, wpos, cidx, order, bin_start, counts);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}

// Step 3: fill.  row_ptr [N + 1] = exclusive scan of counts; edge_index [2, E] (row 0 = centre i, row 1 = neighbour j,
// sorted by (i, j, shift)); shifts [E, 3] (integer-valued doubles): pos[j] - pos[i] + shifts @ cell is the edge vector.
extern "C" int nqb_nl_fill(int64_t N, int64_t E, const double* cell_host, const double* inv_host, const int* pbc,
                           const int* nbins, const int* search, double r_max, const double* wpos, const int32_t* cidx,
                           const int32_t* base, const int64_t* order, const int64_t* bin_start, const int64_t* row_ptr,
                           int64_t* edge_index, double* shifts, nqb_stream_t st) {
  if (N <= 0 || E <= 0) return 0;
  if (!wpos || !cidx || !base || !order || !bin_start || !row_ptr || !edge_index || !shifts)
    return nqb_set_error("nqb_nl_fill: null pointer");
  NlParams p;
  nl_params(cell_host, inv_host, pbc, nbins, search, r_max, p);
  k_nl_fill<<<(unsigned)((N + 63) / 64), 64, 0, (cudaStream_t)st>>>(p, N, wpos, cidx, base, order, bin_start, row_ptr, E,
                                                                    edge_index, shifts);
  nqb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nqb_set_error(cudaGetErrorString(e));
  return 0;
}
