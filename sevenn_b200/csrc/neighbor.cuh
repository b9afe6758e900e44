// Device neighbour list / graph build: positions + cell -> CSR over centre atoms.
//
// Replaces, for the calculator front-end, the CPU graph build the reference runs every step
//   sevenn/train/dataload.py:32-129  (matscipy / ASE neighbour_list 'ijDS' -> edge_index, edge_vec)
//   sevenn/pair_e3gnn/pair_e3gnn.cpp:118-170 (LAMMPS full neighbour list -> edge arrays)
// Semantics restated: a directed edge i <- j for every pair, every periodic image included (also
// self-images at non-zero shift), with |r_j - r_i + S.cell| < cutoff; edge_vec = r_j - r_i + S.cell
// evaluated in double and stored as float (the reference builds it in numpy float64, then casts).
//
// Method: atoms are binned on a grid of the fractional cell (bin width >= cutoff unless the cell is
// thinner than the cutoff), sorted by bin with a stable radix sort (deterministic neighbour order),
// then one thread per atom visits the (2R+1)^3 surrounding bins, R = ceil(cutoff / bin width), with
// the periodic image shift of every visited bin -- correct for any cell size, including cells much
// smaller than the cutoff.  Two passes (count, exclusive scan, fill) emit the CSR directly, so the
// engine needs no sort of the edge list.  Non-periodic directions use the bounding box and no images.
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace s7b {

struct NLGrid {
  double cell[9];      // rows = lattice vectors a, b, c (for non-periodic systems: a bounding frame)
  double inv[9];       // inverse: frac = pos * inv  (row-vector convention)
  double fmin[3];      // offset of the binned fractional range (0 for periodic directions)
  double fspan[3];     // length of the binned fractional range (1 for periodic directions)
  int nb[3];           // bins per direction
  int R[3];            // search radius in bins
  int pbc[3];
  double cutoff2;
};

__device__ __forceinline__ void nl_frac(const NLGrid& g, const double* p, double* f) {
#pragma unroll
  for (int a = 0; a < 3; ++a) f[a] = p[0] * g.inv[0 * 3 + a] + p[1] * g.inv[1 * 3 + a] + p[2] * g.inv[2 * 3 + a];
}

// bin key per atom + wrapped cartesian position
static __global__ void nl_bin_kernel(const NLGrid g, const double* __restrict__ pos, int n, int* __restrict__ key,
                              int* __restrict__ idx, double* __restrict__ wrapped) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
  double f[3];
  nl_frac(g, p, f);
  int b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (g.pbc[a]) f[a] -= floor(f[a]);
    int q = (int)floor((f[a] - g.fmin[a]) / g.fspan[a] * g.nb[a]);
    b[a] = q < 0 ? 0 : (q >= g.nb[a] ? g.nb[a] - 1 : q);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) wrapped[3 * i + c] = f[0] * g.cell[0 * 3 + c] + f[1] * g.cell[1 * 3 + c] + f[2] * g.cell[2 * 3 + c];
  key[i] = (b[0] * g.nb[1] + b[1]) * g.nb[2] + b[2];
  idx[i] = i;
}

// first sorted position of every bin (bin_start[nbins] = n)
static __global__ void nl_bin_start_kernel(const int* __restrict__ key_sorted, int n, int nbins, int* __restrict__ bin_start) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n) return;
  const int prev = (s == 0) ? -1 : key_sorted[s - 1];
  const int cur = (s == n) ? nbins : key_sorted[s];
  for (int b = prev + 1; b <= cur; ++b) bin_start[b] = s;
}

// One thread per centre atom (centre c = centres[tid], or tid itself when centres == nullptr; original
// numbering).  FILL = false: count[tid] = neighbours; FILL = true: write src / edge_vec at rowptr[tid] +
// running offset.  A centre subset is what a rank of the multi-GPU runner asks for: the rows of its own atoms.
template <bool FILL>
__global__ void nl_pairs_kernel(const NLGrid g, const double* __restrict__ wrapped, const int* __restrict__ key,
                                const int* __restrict__ idx_sorted, const int* __restrict__ bin_start, int n,
                                int* __restrict__ count, const int* __restrict__ rowptr,
                                int* __restrict__ src, float* __restrict__ edge_vec,
                                const int* __restrict__ centres = nullptr) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  const int i = centres != nullptr ? centres[tid] : tid;
  const double xi = wrapped[3 * i], yi = wrapped[3 * i + 1], zi = wrapped[3 * i + 2];
  const int k = key[i];
  const int b2 = k % g.nb[2], b1 = (k / g.nb[2]) % g.nb[1], b0 = k / (g.nb[2] * g.nb[1]);
  int out = FILL ? rowptr[tid] : 0;
  for (int d0 = -g.R[0]; d0 <= g.R[0]; ++d0) {
    int q0 = b0 + d0, s0 = 0;
    if (g.pbc[0]) { s0 = (q0 >= 0) ? q0 / g.nb[0] : -((-q0 + g.nb[0] - 1) / g.nb[0]); q0 -= s0 * g.nb[0]; }
    else if (q0 < 0 || q0 >= g.nb[0]) continue;
    for (int d1 = -g.R[1]; d1 <= g.R[1]; ++d1) {
      int q1 = b1 + d1, s1 = 0;
      if (g.pbc[1]) { s1 = (q1 >= 0) ? q1 / g.nb[1] : -((-q1 + g.nb[1] - 1) / g.nb[1]); q1 -= s1 * g.nb[1]; }
      else if (q1 < 0 || q1 >= g.nb[1]) continue;
      for (int d2 = -g.R[2]; d2 <= g.R[2]; ++d2) {
        int q2 = b2 + d2, s2 = 0;
        if (g.pbc[2]) { s2 = (q2 >= 0) ? q2 / g.nb[2] : -((-q2 + g.nb[2] - 1) / g.nb[2]); q2 -= s2 * g.nb[2]; }
        else if (q2 < 0 || q2 >= g.nb[2]) continue;
        const double sx = s0 * g.cell[0] + s1 * g.cell[3] + s2 * g.cell[6];
        const double sy = s0 * g.cell[1] + s1 * g.cell[4] + s2 * g.cell[7];
        const double sz = s0 * g.cell[2] + s1 * g.cell[5] + s2 * g.cell[8];
        const int nbin = (q0 * g.nb[1] + q1) * g.nb[2] + q2;
        const bool same_image = (s0 == 0 && s1 == 0 && s2 == 0);
        for (int s = bin_start[nbin]; s < bin_start[nbin + 1]; ++s) {
          const int j = idx_sorted[s];
          if (same_image && j == i) continue;
          const double dx = wrapped[3 * j] + sx - xi, dy = wrapped[3 * j + 1] + sy - yi, dz = wrapped[3 * j + 2] + sz - zi;
          if (dx * dx + dy * dy + dz * dz < g.cutoff2) {
            if (FILL) {
              src[out] = j;
              edge_vec[3 * (size_t)out] = (float)dx;
              edge_vec[3 * (size_t)out + 1] = (float)dy;
              edge_vec[3 * (size_t)out + 2] = (float)dz;
            }
            ++out;
          }
        }
      }
    }
  }
  if (!FILL) count[tid] = out;
}

}  // namespace s7b
