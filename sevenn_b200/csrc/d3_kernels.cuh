// DFT-D3 dispersion correction (Grimme 2010; zero and Becke-Johnson damping) on a cell list.
//
// Replaces the reference's CUDA D3 (sevenn/pair_e3gnn/pair_d3_for_ase.cu = pair_d3.cu without LAMMPS):
//   kernel_get_coordination_number        :1004-1058
//   kernel_get_dC6_dCNij                  :765-845
//   kernel_get_forces_without_dC6_{zero,bj}  :1263-1496, 1534-1745
//   kernel_get_forces_with_dC6            :1797-1962
// The reference enumerates all N(N+1)/2 atom pairs x all lattice translations (O(N^2 tau) work, an
// N(N+1)/2-sized C6 table, int indexing that overflows at 46 340 atoms, one GPU).  Here:
//   * atoms are binned on the fractional cell (neighbor.cuh) and one WARP per atom sweeps the bin images
//     inside the cutoff, one lane per bin image -> O(N * neighbours) work, any cell size (also cells
//     much smaller than the 50 A cutoff), 64-bit-free indexing (no pair table at all);
//   * C6_ij(CN_i, CN_j) is not tabulated per pair: the Gaussian weight of reference (a, b) factorises,
//     L_ij(a,b) = w_i(a) w_j(b), because the reference coordination numbers depend only on (element,
//     index) (asserted by tools/convert_d3_params.py), so with per-atom normalised weights W_i
//         C6_ij = sum_b V_i[t_j][b] W_j[b],   V_i[t][b] = sum_a c6ref[t_i][t][a][b] W_i[a]
//     costs 5 FMAs per pair (and 5 more for dC6/dCN_i) instead of 25 exponentials;
//   * every ordered pair (i <- j) is evaluated by i's warp only: forces, dE/dCN and the virial are plain
//     per-warp sums, no atomics except one per warp for the scalars; pair math in fp32 as the reference,
//     accumulation in fp64 (the reference's float image sums lose ~4.5e-5 of the NaCl golden energy);
//   * an atom range [i_begin, i_end) per launch: multi-GPU = atom decomposition with replicated positions
//     (the 50 A range is of the order of the box), three small all-gathers per step (sevenn_b200/d3.py).
// Units inside: bohr and hartree, as in the reference.
#pragma once
#include "common.cuh"
#include "neighbor.cuh"

namespace s7b {

constexpr int kD3MaxTypes = 16;
constexpr int kD3WarpsPerBlock = 4;
constexpr float kD3K1 = 16.0f;
constexpr double kD3K3 = -4.0;

struct D3Atoms {              // arrays over atoms in bin-sorted order
  const double* x;            // [n,3] wrapped cartesian positions (bohr)
  const int* type;            // [n]
  const float* W;             // [n,5] normalised C6 reference weights
  const float* logD;          // [n]   log of the weight sum (den <= 1e-99 fallback, pair_d3_for_ase.cu:824-844)
  const int* near;            // [n]   nearest reference index
  const double* dc6i;         // [n]   -dE/dCN (after the pair pass)
  const int* bin_start;       // [nbins + 1]
  const int* bin_of;          // [n]   bin of each (sorted) atom
};

struct D3Params {
  int ntypes, damping;        // damping: 0 = zero, 1 = Becke-Johnson
  float s6, s8, a1, a2, alp6, alp8;
  double rthr, cnthr;         // squared cutoffs (bohr^2)
  float rcov[kD3MaxTypes], r2r4[kD3MaxTypes];
  const float* r0ab;          // [ntypes, ntypes] (bohr)
  const float* c6ref;         // [ntypes, ntypes, 5, 5]
};

struct D3Out {
  double* cn;                 // [n]   (sorted order)
  double* dc6i;               // [n]
  double* force;              // [n,3]
  double* energy;             // [1]
  double* sigma;              // [9]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Sweep of all (bin image, atom) candidates around atom i with one lane per bin image.
// f(j, dx, dy, dz, r2, self) is called for every candidate with r2 <= cut2 (self: j == i, other image).
template <class F>
__device__ __forceinline__ void d3_sweep(const NLGrid& g, const D3Atoms& A, int i, const int (&R)[3], double cut2,
                                         int lane, F&& f) {
  const double xi = A.x[3 * i], yi = A.x[3 * i + 1], zi = A.x[3 * i + 2];
  const int k = A.bin_of[i];
  const int b2 = k % g.nb[2], b1 = (k / g.nb[2]) % g.nb[1], b0 = k / (g.nb[2] * g.nb[1]);
  const int n1 = 2 * R[1] + 1, n2 = 2 * R[2] + 1;
  const int total = (2 * R[0] + 1) * n1 * n2;
  for (int m = lane; m < total; m += 32) {
    const int d2 = m % n2 - R[2], d1 = (m / n2) % n1 - R[1], d0 = m / (n2 * n1) - R[0];
    int q0 = b0 + d0, q1 = b1 + d1, q2 = b2 + d2, s0 = 0, s1 = 0, s2 = 0;
    if (g.pbc[0]) { s0 = (q0 >= 0) ? q0 / g.nb[0] : -((-q0 + g.nb[0] - 1) / g.nb[0]); q0 -= s0 * g.nb[0]; }
    else if (q0 < 0 || q0 >= g.nb[0]) continue;
    if (g.pbc[1]) { s1 = (q1 >= 0) ? q1 / g.nb[1] : -((-q1 + g.nb[1] - 1) / g.nb[1]); q1 -= s1 * g.nb[1]; }
    else if (q1 < 0 || q1 >= g.nb[1]) continue;
    if (g.pbc[2]) { s2 = (q2 >= 0) ? q2 / g.nb[2] : -((-q2 + g.nb[2] - 1) / g.nb[2]); q2 -= s2 * g.nb[2]; }
    else if (q2 < 0 || q2 >= g.nb[2]) continue;
    const double sx = s0 * g.cell[0] + s1 * g.cell[3] + s2 * g.cell[6] - xi;
    const double sy = s0 * g.cell[1] + s1 * g.cell[4] + s2 * g.cell[7] - yi;
    const double sz = s0 * g.cell[2] + s1 * g.cell[5] + s2 * g.cell[8] - zi;
    const int nbin = (q0 * g.nb[1] + q1) * g.nb[2] + q2;
    const bool same_image = (s0 == 0 && s1 == 0 && s2 == 0);
    const int e = A.bin_start[nbin + 1];
    for (int j = A.bin_start[nbin]; j < e; ++j) {
      if (same_image && j == i) continue;
      const double dx = A.x[3 * j] + sx, dy = A.x[3 * j + 1] + sy, dz = A.x[3 * j + 2] + sz;
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 <= cut2) f(j, (float)dx, (float)dy, (float)dz, (float)r2, j == i);
    }
  }
}

// ---- pass 1: coordination numbers (:1004-1058) -----------------------------------------------------
__global__ void __launch_bounds__(32 * kD3WarpsPerBlock)
d3_cn_kernel(const NLGrid g, const D3Atoms A, const D3Params P, int3 R, int i_begin, int i_end, D3Out out) {
  const int i = i_begin + blockIdx.x * kD3WarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= i_end) return;
  const int Rr[3] = {R.x, R.y, R.z};
  const float rci = P.rcov[A.type[i]];
  double cn = 0.0;
  d3_sweep(g, A, i, Rr, P.cnthr, lane, [&](int j, float, float, float, float r2, bool) {
    const float rc = rci + P.rcov[A.type[j]];
    const float rr = rsqrtf(r2);
    cn += (double)(1.0f / (1.0f + __expf(-kD3K1 * (rc * rr - 1.0f))));
  });
  cn = warp_sum(cn);
  if (lane == 0) out.cn[i] = cn;
}

// ---- per atom: normalised Gaussian weights of the C6 references and their CN derivative (:765-845) --
// W[a] = w_a / D, dW[a] = d W[a] / d CN, w_a = exp(K3 (CN_ref[a] - CN)^2), D = sum_a w_a (double: the
// exponents reach -400 for highly coordinated atoms)
__global__ void d3_weights_kernel(int n, const int* __restrict__ type, const double* __restrict__ cn,
                                  const float* __restrict__ cnref /*[ntypes,5]*/, const int* __restrict__ mxc,
                                  float* __restrict__ W, float* __restrict__ dW, float* __restrict__ logD,
                                  int* __restrict__ near) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = type[i], m = mxc[t];
  const float cni = (float)cn[i];                           // the reference holds CN as a float here (:780)
  double w[5], dw[5], D = 0.0, dD = 0.0;
  float best = 3.0e38f;
  int nb = 0;
  for (int a = 0; a < 5; ++a) {
    w[a] = dw[a] = 0.0;
    if (a >= m) continue;
    const float cr = cnref[t * 5 + a];
    const float d2 = (cr - cni) * (cr - cni);
    if (d2 < best) { best = d2; nb = a; }
    w[a] = exp(kD3K3 * (double)d2);
    dw[a] = w[a] * 2.0 * kD3K3 * (double)(cni - cr);
    D += w[a];
    dD += dw[a];
  }
  // exponent bookkeeping for the reference's "denominator <= 1e-99" branch: log D without underflow
  double lD;
  if (D > 1e-300) lD = log(D);
  else lD = kD3K3 * (double)best;                            // dominated by the nearest reference
  for (int a = 0; a < 5; ++a) {
    const double Wn = D > 1e-300 ? w[a] / D : (a == nb ? 1.0 : 0.0);
    const double dWn = D > 1e-300 ? (dw[a] - Wn * dD) / D : 0.0;
    W[i * 5 + a] = (float)Wn;
    dW[i * 5 + a] = (float)dWn;
  }
  logD[i] = (float)lD;
  near[i] = nb;
}

// ---- pass 2: pair energy, explicit-r forces, dE/dCN (:1263-1745) -----------------------------------
__global__ void __launch_bounds__(32 * kD3WarpsPerBlock)
d3_pair_kernel(const NLGrid g, const D3Atoms A, const D3Params P, const float* __restrict__ dW, int3 R,
               int i_begin, int i_end, D3Out out) {
  __shared__ float sV[kD3WarpsPerBlock][kD3MaxTypes][10];      // V_i[t][b], dV_i[t][b]
  __shared__ double sred[kD3WarpsPerBlock][7];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = i_begin + blockIdx.x * kD3WarpsPerBlock + wib;
  const bool active = i < i_end;
  const int Rr[3] = {R.x, R.y, R.z};
  double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0, dc = 0.0;
  double sg[6] = {0, 0, 0, 0, 0, 0};                              // xx, yy, zz, xy, xz, yz
  if (active) {
    const int ti = A.type[i];
    // V_i[t][b] = sum_a c6ref[ti][t][a][b] W_i[a]  (lanes over (t, b))
    for (int q = lane; q < P.ntypes * 5; q += 32) {
      const int t = q / 5, b = q % 5;
      float v = 0.0f, dv = 0.0f;
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        const float c = __ldg(P.c6ref + ((ti * P.ntypes + t) * 5 + a) * 5 + b);
        v = fmaf(c, A.W[i * 5 + a], v);
        dv = fmaf(c, dW[i * 5 + a], dv);
      }
      sV[wib][t][b] = v;
      sV[wib][t][5 + b] = dv;
    }
    __syncwarp();
    const float logDi = A.logD[i];
    const int near_i = A.near[i];
    const float r2r4i = P.r2r4[ti];
    d3_sweep(g, A, i, Rr, P.rthr, lane, [&](int j, float dx, float dy, float dz, float r2, bool self) {
      const int tj = A.type[j];
      float c6 = 0.0f, dc6 = 0.0f;
      if (logDi + A.logD[j] > -227.95593f) {                      // den > 1e-99
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          const float wj = A.W[j * 5 + b];
          c6 = fmaf(sV[wib][tj][b], wj, c6);
          dc6 = fmaf(sV[wib][tj][5 + b], wj, dc6);
        }
      } else {
        c6 = __ldg(P.c6ref + ((ti * P.ntypes + tj) * 5 + near_i) * 5 + A.near[j]);
      }
      float gfun, dgdr;                                           // E_pair = -C6 g(r)
      const float r = sqrtf(r2);
      if (P.damping == 1) {
        const float r42x3 = r2r4i * P.r2r4[tj] * 3.0f;
        const float R0 = fmaf(P.a1, sqrtf(r42x3), P.a2);
        const float R0_2 = R0 * R0, R0_6 = R0_2 * R0_2 * R0_2, R0_8 = R0_6 * R0_2;
        const float r5 = r2 * r2 * r, r7 = r5 * r2;
        const float t6 = 1.0f / fmaf(r5, r, R0_6), t8 = 1.0f / fmaf(r7, r, R0_8);
        const float s8r = P.s8 * r42x3;
        gfun = fmaf(s8r, t8, P.s6 * t6);
        dgdr = -fmaf(8.0f * s8r * r7, t8 * t8, 6.0f * P.s6 * r5 * t6 * t6);
      } else {
        const float r0 = __ldg(P.r0ab + ti * P.ntypes + tj);
        const float r42 = r2r4i * P.r2r4[tj];
        const float rr = 1.0f / r;
        const float t6 = __powf(P.a1 * r0 * rr, P.alp6), t8 = __powf(P.a2 * r0 * rr, P.alp8);
        const float d6 = 1.0f / fmaf(6.0f, t6, 1.0f), d8 = 1.0f / fmaf(6.0f, t8, 1.0f);
        const float r2_rc = rr * rr, r6_rc = r2_rc * r2_rc * r2_rc, r8_rc = r6_rc * r2_rc;
        const float s8r = P.s8 * r42;
        gfun = r6_rc * fmaf(3.0f * r2_rc, s8r * d8, P.s6 * d6);
        // d/dr [ s6 d6 r^-6 + 3 s8 r42 d8 r^-8 ]
        dgdr = 6.0f * r8_rc * r * (P.s6 * d6 * fmaf(P.alp6 * t6, d6, -1.0f) + r2_rc * s8r * d8 * fmaf(3.0f * P.alp8 * t8, d8, -4.0f));
      }
      e -= 0.5 * (double)(c6 * gfun);
      dc += (double)(gfun * dc6);                                 // dc6i_i = sum g dC6/dCN_i
      const float dEdr = -c6 * dgdr;                              // of the full pair
      const float s = dEdr / r;
      const float vx = s * dx, vy = s * dy, vz = s * dz;          // dE/d(r_ij) direction (r_ij = x_j - x_i + tau)
      if (!self) { fx += (double)vx; fy += (double)vy; fz += (double)vz; }
      sg[0] -= 0.5 * (double)(vx * dx); sg[1] -= 0.5 * (double)(vy * dy); sg[2] -= 0.5 * (double)(vz * dz);
      sg[3] -= 0.5 * (double)(vx * dy); sg[4] -= 0.5 * (double)(vx * dz); sg[5] -= 0.5 * (double)(vy * dz);
    });
    e = warp_sum(e); fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz); dc = warp_sum(dc);
#pragma unroll
    for (int q = 0; q < 6; ++q) sg[q] = warp_sum(sg[q]);
    if (lane == 0) {
      out.force[3 * i] = fx; out.force[3 * i + 1] = fy; out.force[3 * i + 2] = fz;
      out.dc6i[i] = dc;
    }
  }
  if (lane == 0) {
    sred[wib][0] = active ? e : 0.0;
    for (int q = 0; q < 6; ++q) sred[wib][1 + q] = active ? sg[q] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double t = 0.0;
    for (int w = 0; w < kD3WarpsPerBlock; ++w) t += sred[w][threadIdx.x];
    if (threadIdx.x == 0) atomicAdd(out.energy, t);
    else atomicAdd(out.sigma + (threadIdx.x - 1), t);
  }
}

// ---- pass 3: chain rule through the coordination numbers (:1797-1962) ------------------------------
__global__ void __launch_bounds__(32 * kD3WarpsPerBlock)
d3_chain_kernel(const NLGrid g, const D3Atoms A, const D3Params P, int3 R, int i_begin, int i_end, D3Out out) {
  __shared__ double sred[kD3WarpsPerBlock][6];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = i_begin + blockIdx.x * kD3WarpsPerBlock + wib;
  const bool active = i < i_end;
  const int Rr[3] = {R.x, R.y, R.z};
  double fx = 0.0, fy = 0.0, fz = 0.0;
  double sg[6] = {0, 0, 0, 0, 0, 0};
  if (active) {
    const float rci = P.rcov[A.type[i]];
    const double di = A.dc6i[i];
    const float cn2 = (float)P.cnthr;
    d3_sweep(g, A, i, Rr, P.cnthr, lane, [&](int j, float dx, float dy, float dz, float r2, bool self) {
      if (r2 >= cn2) return;                                      // the reference uses a strict bound here (:1843)
      const float rc = rci + P.rcov[A.type[j]];
      const float rr = rsqrtf(r2);
      const float ex = __expf(-kD3K1 * (rc * rr - 1.0f));
      const float dcnn = -kD3K1 * rc * ex / (r2 * (ex + 1.0f) * (ex + 1.0f));     // d cnf / dr
      const float x1 = dcnn * (float)(di + A.dc6i[j]);            // -dE/dr of the pair through CN_i and CN_j
      const float s = x1 * rr;
      const float vx = s * dx, vy = s * dy, vz = s * dz;
      if (!self) { fx -= (double)vx; fy -= (double)vy; fz -= (double)vz; }
      sg[0] += 0.5 * (double)(vx * dx); sg[1] += 0.5 * (double)(vy * dy); sg[2] += 0.5 * (double)(vz * dz);
      sg[3] += 0.5 * (double)(vx * dy); sg[4] += 0.5 * (double)(vx * dz); sg[5] += 0.5 * (double)(vy * dz);
    });
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
#pragma unroll
    for (int q = 0; q < 6; ++q) sg[q] = warp_sum(sg[q]);
    if (lane == 0) { out.force[3 * i] += fx; out.force[3 * i + 1] += fy; out.force[3 * i + 2] += fz; }
  }
  if (lane == 0)
    for (int q = 0; q < 6; ++q) sred[wib][q] = active ? sg[q] : 0.0;
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int w = 0; w < kD3WarpsPerBlock; ++w) t += sred[w][threadIdx.x];
    atomicAdd(out.sigma + threadIdx.x, t);
  }
}

// gather / scatter between the caller's atom order and the bin-sorted order
__global__ void d3_sort_gather_kernel(int n, const int* __restrict__ idx_sorted, const int* __restrict__ key_sorted,
                                      const double* __restrict__ wrapped, const int* __restrict__ type,
                                      double* __restrict__ xs, int* __restrict__ ts, int* __restrict__ bin_of) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int i = idx_sorted[s];
  xs[3 * s] = wrapped[3 * i]; xs[3 * s + 1] = wrapped[3 * i + 1]; xs[3 * s + 2] = wrapped[3 * i + 2];
  ts[s] = type[i];
  bin_of[s] = key_sorted[s];
}
__global__ void d3_unsort_kernel(int n, int width, const int* __restrict__ idx_sorted, const double* __restrict__ in,
                                 double scale, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const int s = t / width, c = t - s * width;
  out[(size_t)idx_sorted[s] * width + c] = in[t] * scale;
}

}  // namespace s7b
