// Fused neighbour-gather -> Clebsch-Gordan tensor product -> scatter-to-centre kernels.
//
// Replaces, for one interaction layer, the reference's
//   x[edge_src] gather                     sevenn/nn/convolution.py:131
//   e3nn TensorProduct ('uvu', per-edge w) sevenn/nn/convolution.py:84-100,131
//   message_gather scatter_reduce_         sevenn/nn/convolution.py:17-26,133
// and their autograd backward (sevenn/nn/force_output.py:177-182) with one forward and one
// backward kernel per l1 "kind" (see csrc/gen_kernels.py).
//
// Mapping: one warp owns one destination atom n and one l1 block; lane <-> channel u (NCH
// channels per lane, 32 apart), so every global access is a 128-byte contiguous segment in the
// component-major ("cm") layout.  The edge loop runs over the CSR row of n; the accumulators
// for all paths of the kind stay in registers and are written once -- no atomics in the forward.
// The radial weights w_p,u(r) come either from a cubic-Hermite table indexed by the edge
// length (TABLE = true; L2-resident, [knots][W] float4) or from a stored [E, W] array
// (TABLE = false; the reference's plug-in boundary, where the radial MLP stays outside).
#pragma once
#include "common.cuh"
#include "generated/tp_kinds.cuh"

namespace s7b {

constexpr int kConvWarpsPerBlock = 4;

// Sum M values (M = 8 or 16) over the 32 lanes with M-1+log2(32/M) shuffles instead of 5*M.
// On return v[0] of lane L holds the total of value index (L >> (5 - log2 M)) & (M-1).
template <int M>
__device__ __forceinline__ void warp_reduce_multi(float (&v)[M], int lane) {
  static_assert(M == 8 || M == 16, "M must be 8 or 16");
  int off = 16;
#pragma unroll
  for (int m = M / 2; m >= 1; m >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < m; ++j) {
      const float send = up ? v[j] : v[j + m];
      const float keep = up ? v[j + m] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

template <class Kind>
__device__ __forceinline__ void load_Y(const float* __restrict__ Yrow, float (&Y)[Kind::NY]) {
  Y[0] = 1.0f;
  constexpr int NV = (Kind::NY - 1 + 3) / 4;
  const float4* p = reinterpret_cast<const float4*>(Yrow);
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const float4 v = __ldg(p + q);
    if (4 * q + 1 < Kind::NY) Y[4 * q + 1] = v.x;
    if (4 * q + 2 < Kind::NY) Y[4 * q + 2] = v.y;
    if (4 * q + 3 < Kind::NY) Y[4 * q + 3] = v.z;
    if (4 * q + 4 < Kind::NY) Y[4 * q + 4] = v.w;
  }
}

// ------------------------------------------------------------------------------------------
// forward:  out[n, path block] = sum_{e in row n} w_e * CG(x[src_e], Y_e)
// grid = (ceil(n_dst / kConvWarpsPerBlock), mul / (32 * NCH)), block = 32 * kConvWarpsPerBlock
// ------------------------------------------------------------------------------------------
template <class Kind, int NCH, bool TABLE>
__global__ void __launch_bounds__(32 * kConvWarpsPerBlock)
conv_fwd_kernel(const ConvArgs a, const ConvRole role, float* __restrict__ out) {
  const int n = blockIdx.x * kConvWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= a.n_dst) return;
  const int lane = threadIdx.x & 31;
  const int u0 = blockIdx.y * (32 * NCH) + lane;
  const int e0 = __ldg(a.rowptr + n), e1 = __ldg(a.rowptr + n + 1);

  float acc[NCH][Kind::NACC];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int q = 0; q < Kind::NACC; ++q) acc[c][q] = 0.0f;

  for (int e = e0; e < e1; ++e) {
    const int4 rec = __ldg(a.rec + e);
    float Y[Kind::NY];
    load_Y<Kind>(a.Y + (size_t)e * a.ny_stride, Y);
    const float* __restrict__ xrow = a.x + (size_t)rec.x * a.dim_x + role.x_off;
    const float tt = __int_as_float(rec.z);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int u = u0 + 32 * c;
      float x[Kind::D1], w[Kind::NPATH];
#pragma unroll
      for (int i = 0; i < Kind::D1; ++i) x[i] = __ldg(xrow + i * role.mul + u);
#pragma unroll
      for (int p = 0; p < Kind::NPATH; ++p) {
        if (TABLE) {
          const float4 cf = __ldg(a.table + (size_t)rec.y * a.w_numel + role.w_off[p] + u);
          w[p] = fmaf(tt, fmaf(tt, fmaf(tt, cf.w, cf.z), cf.y), cf.x);
        } else {
          w[p] = __ldg(a.w + (size_t)e * a.w_numel + role.w_off[p] + u);
        }
      }
      Kind::fwd(x, Y, w, acc[c]);
    }
  }

  float* __restrict__ orow = out + (size_t)n * a.dim_mid;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int u = u0 + 32 * c;
#pragma unroll
    for (int p = 0; p < Kind::NPATH; ++p) {
#pragma unroll
      for (int k = 0; k < 2 * Kind::path_l3(p) + 1; ++k)
        orow[role.out_off[p] + k * role.out_stride[p] + u] = acc[c][Kind::acc_off(p) + k];
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward (centre-major): given ga = dE/d out[n, :], per edge of row n
//   TABLE : dEdr_acc[e] += sum_{p,u} (dE/dw_{p,u}) * w'_{p,u}(r_e)         (radial chain rule)
//   !TABLE: dw[e, :]     = dE/dw                                          (plug-in boundary)
//   dY_acc[e, 1..]      += sum_u dE/dY                                    (warp reduction)
//   dx[src_e, :]        += dE/dx                                          (RED.ADD, NEED_DX)
// dY_acc / dEdr_acc / dw rows are owned by exactly one warp of one launch: plain read-modify-write.
// ------------------------------------------------------------------------------------------
template <class Kind, int NCH, bool TABLE, bool NEED_DX>
__global__ void __launch_bounds__(32 * kConvWarpsPerBlock)
conv_bwd_kernel(const ConvArgs a, const ConvRole role, const float* __restrict__ gout,
                float* __restrict__ dx, float* __restrict__ dY_acc, float* __restrict__ dEdr_acc,
                float* __restrict__ dw) {
  const int n = blockIdx.x * kConvWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= a.n_dst) return;
  const int lane = threadIdx.x & 31;
  const int u0 = blockIdx.y * (32 * NCH) + lane;
  const int e0 = __ldg(a.rowptr + n), e1 = __ldg(a.rowptr + n + 1);

  float ga[NCH][Kind::NACC];
  {
    const float* __restrict__ grow = gout + (size_t)n * a.dim_mid;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int u = u0 + 32 * c;
#pragma unroll
      for (int p = 0; p < Kind::NPATH; ++p)
#pragma unroll
        for (int k = 0; k < 2 * Kind::path_l3(p) + 1; ++k)
          ga[c][Kind::acc_off(p) + k] = __ldg(grow + role.out_off[p] + k * role.out_stride[p] + u);
    }
  }

  constexpr int NR = (Kind::NY <= 9) ? 8 : 16;   // values reduced with the transposing butterfly
  for (int e = e0; e < e1; ++e) {
    const int4 rec = __ldg(a.rec + e);
    float Y[Kind::NY];
    load_Y<Kind>(a.Y + (size_t)e * a.ny_stride, Y);
    const float* __restrict__ xrow = a.x + (size_t)rec.x * a.dim_x + role.x_off;
    const float tt = __int_as_float(rec.z);
    float dY[Kind::NY];
#pragma unroll
    for (int j = 0; j < Kind::NY; ++j) dY[j] = 0.0f;
    float dEdr = 0.0f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int u = u0 + 32 * c;
      float x[Kind::D1], w[Kind::NPATH], wd[Kind::NPATH], dwv[Kind::NPATH], dxv[Kind::D1];
#pragma unroll
      for (int i = 0; i < Kind::D1; ++i) x[i] = __ldg(xrow + i * role.mul + u);
#pragma unroll
      for (int p = 0; p < Kind::NPATH; ++p) {
        if (TABLE) {
          const float4 cf = __ldg(a.table + (size_t)rec.y * a.w_numel + role.w_off[p] + u);
          w[p] = fmaf(tt, fmaf(tt, fmaf(tt, cf.w, cf.z), cf.y), cf.x);
          wd[p] = fmaf(tt, fmaf(tt, 3.0f * cf.w, 2.0f * cf.z), cf.y) * a.inv_h;
        } else {
          w[p] = __ldg(a.w + (size_t)e * a.w_numel + role.w_off[p] + u);
          wd[p] = 0.0f;
        }
      }
      Kind::bwd(x, Y, w, ga[c], dwv, dxv, dY);
#pragma unroll
      for (int p = 0; p < Kind::NPATH; ++p) {
        if (TABLE) dEdr = fmaf(dwv[p], wd[p], dEdr);
        else dw[(size_t)e * a.w_numel + role.w_off[p] + u] = dwv[p];
      }
      if (NEED_DX) {
        float* __restrict__ dxrow = dx + (size_t)rec.x * a.dim_x + role.x_off;
#pragma unroll
        for (int i = 0; i < Kind::D1; ++i) atomicAdd(dxrow + i * role.mul + u, dxv[i]);
      }
    }
    // cross-channel reduction of dE/dY (NY-1 values) and dE/dr (1 value)
    if (Kind::NY > 1) {
      float red[NR];
#pragma unroll
      for (int j = 0; j < NR; ++j) red[j] = (j + 1 < Kind::NY) ? dY[j + 1] : 0.0f;
      if (TABLE && Kind::NY - 1 < NR) red[NR - 1] = dEdr;     // free slot: ride along
      warp_reduce_multi<NR>(red, lane);
      constexpr int SH = (NR == 8) ? 2 : 1;
      const int idx = (lane >> SH) & (NR - 1);
      if ((lane & ((1 << SH) - 1)) == 0) {
        if (idx + 1 < Kind::NY) {
          float* p = dY_acc + (size_t)e * a.ny_stride + idx;
          *p += red[0];
        } else if (TABLE && Kind::NY - 1 < NR && idx == NR - 1) {
          dEdr_acc[e] += red[0];
        }
      }
      if (TABLE && !(Kind::NY - 1 < NR)) {
        const float s = warp_sum(dEdr);
        if (lane == 0) dEdr_acc[e] += s;
      }
    } else if (TABLE) {
      const float s = warp_sum(dEdr);
      if (lane == 0) dEdr_acc[e] += s;
    }
  }
}

}  // namespace s7b
