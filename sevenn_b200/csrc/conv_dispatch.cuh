// Launch dispatch for one (lmax_filter, lmax_out) group of tensor-product kinds.  Each group is
// its own translation unit (conv_group_*.cu) so that the kinds compile in parallel.
#pragma once
#include "conv_kernels.cuh"

namespace s7b {

template <class Kind, int NCH, bool TABLE>
static int launch_fwd_one(const ConvArgs& a, const ConvRole& role, float* out, cudaStream_t st) {
  dim3 grid((a.n_dst + kConvWarpsPerBlock - 1) / kConvWarpsPerBlock, role.mul / (32 * NCH));
  conv_fwd_kernel<Kind, NCH, TABLE><<<grid, 32 * kConvWarpsPerBlock, 0, st>>>(a, role, out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <class Kind, int NCH, bool TABLE, bool NEED_DX>
static int launch_bwd_one(const ConvArgs& a, const ConvRole& role, const float* gout, float* dx,
                          float* dY_acc, float* dEdr_acc, float* dw, cudaStream_t st) {
  dim3 grid((a.n_dst + kConvWarpsPerBlock - 1) / kConvWarpsPerBlock, role.mul / (32 * NCH));
  conv_bwd_kernel<Kind, NCH, TABLE, NEED_DX><<<grid, 32 * kConvWarpsPerBlock, 0, st>>>(a, role, gout, dx, dY_acc, dEdr_acc, dw);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// channels per lane: as many as divide mul/32 among the instantiated {MAXCH, .., 1}
template <class Kind, int MAXCH>
static int fwd_kind(bool table, const ConvArgs& a, const ConvRole& role, float* out, cudaStream_t st) {
  const int chunks = role.mul / 32;
  if (MAXCH >= 4 && chunks % 4 == 0)
    return table ? launch_fwd_one<Kind, (MAXCH >= 4 ? 4 : 1), true>(a, role, out, st)
                 : launch_fwd_one<Kind, (MAXCH >= 4 ? 4 : 1), false>(a, role, out, st);
  if (MAXCH >= 2 && chunks % 2 == 0)
    return table ? launch_fwd_one<Kind, (MAXCH >= 2 ? 2 : 1), true>(a, role, out, st)
                 : launch_fwd_one<Kind, (MAXCH >= 2 ? 2 : 1), false>(a, role, out, st);
  return table ? launch_fwd_one<Kind, 1, true>(a, role, out, st)
               : launch_fwd_one<Kind, 1, false>(a, role, out, st);
}

template <class Kind, int NCH, bool ALLOW_NODX>
static int bwd_kind_nch(bool table, bool need_dx, const ConvArgs& a, const ConvRole& role,
                        const float* gout, float* dx, float* dY, float* dEdr, float* dw,
                        cudaStream_t st) {
  if (!need_dx && ALLOW_NODX) {
    return table ? launch_bwd_one<Kind, NCH, true, !ALLOW_NODX>(a, role, gout, dx, dY, dEdr, dw, st)
                 : launch_bwd_one<Kind, NCH, false, !ALLOW_NODX>(a, role, gout, dx, dY, dEdr, dw, st);
  }
  return table ? launch_bwd_one<Kind, NCH, true, true>(a, role, gout, dx, dY, dEdr, dw, st)
               : launch_bwd_one<Kind, NCH, false, true>(a, role, gout, dx, dY, dEdr, dw, st);
}

// ALLOW_NODX: only the l1 = 0 kinds are ever run without dx (first layer: x depends on species only)
template <class Kind, int MAXCH, bool ALLOW_NODX>
static int bwd_kind(bool table, bool need_dx, const ConvArgs& a, const ConvRole& role,
                    const float* gout, float* dx, float* dY, float* dEdr, float* dw, cudaStream_t st) {
  const int chunks = role.mul / 32;
  if (MAXCH >= 4 && chunks % 4 == 0)
    return bwd_kind_nch<Kind, (MAXCH >= 4 ? 4 : 1), ALLOW_NODX>(table, need_dx, a, role, gout, dx, dY, dEdr, dw, st);
  if (MAXCH >= 2 && chunks % 2 == 0)
    return bwd_kind_nch<Kind, (MAXCH >= 2 ? 2 : 1), ALLOW_NODX>(table, need_dx, a, role, gout, dx, dY, dEdr, dw, st);
  return bwd_kind_nch<Kind, 1, ALLOW_NODX>(table, need_dx, a, role, gout, dx, dY, dEdr, dw, st);
}

}  // namespace s7b

// Defines  launch_conv_fwd_LF_LO / launch_conv_bwd_LF_LO  for l1 = 0..LF.
#define S7B_DEFINE_CONV_GROUP(LF, LO)                                                              \
  namespace s7b {                                                                                  \
  int launch_conv_fwd_##LF##_##LO(int l1, bool table, const ConvArgs& a, const ConvRole& role,     \
                                  float* out, cudaStream_t st) {                                   \
    switch (l1) {                                                                                  \
      case 0: return fwd_kind<TPKind<0, LF, LO>, 4>(table, a, role, out, st);                      \
      case 1: return fwd_kind<TPKind<1, LF, LO>, 2>(table, a, role, out, st);                      \
      case 2: return fwd_kind<TPKind<2, LF, LO>, 1>(table, a, role, out, st);                      \
      case 3: return fwd_kind<TPKind<(LF >= 3 ? 3 : 2), LF, LO>, 1>(table, a, role, out, st);      \
    }                                                                                              \
    return 1;                                                                                      \
  }                                                                                                \
  int launch_conv_bwd_##LF##_##LO(int l1, bool table, bool need_dx, const ConvArgs& a,             \
                                  const ConvRole& role, const float* gout, float* dx, float* dY,   \
                                  float* dEdr, float* dw, cudaStream_t st) {                       \
    switch (l1) {                                                                                  \
      case 0: return bwd_kind<TPKind<0, LF, LO>, 4, true>(table, need_dx, a, role, gout, dx, dY, dEdr, dw, st);  \
      case 1: return bwd_kind<TPKind<1, LF, LO>, 2, false>(table, need_dx, a, role, gout, dx, dY, dEdr, dw, st); \
      case 2: return bwd_kind<TPKind<2, LF, LO>, 1, false>(table, need_dx, a, role, gout, dx, dY, dEdr, dw, st); \
      case 3: return bwd_kind<TPKind<(LF >= 3 ? 3 : 2), LF, LO>, 1, false>(table, need_dx, a, role, gout, dx, dY, dEdr, dw, st); \
    }                                                                                              \
    return 1;                                                                                      \
  }                                                                                                \
  }
