// Host-side harness: exposes the GENERATED per-lane arithmetic (tp_kinds.cuh, sh.cuh) through a
// C interface so that tests/test_generated_math.py can check it against numpy einsum on the CPU.
// (Test infrastructure only; compiled with g++ by the test.)
#include "../../sevenn_b200/csrc/generated/tp_kinds.cuh"
#include "../../sevenn_b200/csrc/generated/sh.cuh"

using namespace s7b;

#define FOR_KINDS(X) \
  X(0,2,2) X(1,2,2) X(2,2,2) X(0,2,0) X(1,2,0) X(2,2,0) \
  X(0,3,3) X(1,3,3) X(2,3,3) X(3,3,3) X(0,3,0) X(1,3,0) X(2,3,0) X(3,3,0)

extern "C" {

int tp_info(int l1, int lf, int lo, int* npath, int* nacc, int* l2s, int* l3s, int* offs) {
#define X(a,b,c) if (l1==a && lf==b && lo==c) { using K = TPKind<a,b,c>; *npath = K::NPATH; *nacc = K::NACC; \
    for (int p = 0; p < K::NPATH; ++p) { l2s[p] = K::path_l2(p); l3s[p] = K::path_l3(p); offs[p] = K::acc_off(p); } return 0; }
  FOR_KINDS(X)
#undef X
  return 1;
}

int tp_fwd(int l1, int lf, int lo, const float* x, const float* Y, const float* w, float* acc) {
#define X(a,b,c) if (l1==a && lf==b && lo==c) { TPKind<a,b,c>::fwd(x, Y, w, acc); return 0; }
  FOR_KINDS(X)
#undef X
  return 1;
}

int tp_bwd(int l1, int lf, int lo, const float* x, const float* Y, const float* w, const float* ga,
           float* dw, float* dx, float* dY) {
#define X(a,b,c) if (l1==a && lf==b && lo==c) { TPKind<a,b,c>::bwd(x, Y, w, ga, dw, dx, dY); return 0; }
  FOR_KINDS(X)
#undef X
  return 1;
}

// packed-pair instantiation (two channels per "lane"): arrays of V2 = {x, y}
int tp_fwd2(int l1, int lf, int lo, const float* x, const float* Y, const float* w, float* acc) {
#define X(a,b,c) if (l1==a && lf==b && lo==c) { TPKind<a,b,c>::fwd(reinterpret_cast<const V2*>(x), Y, reinterpret_cast<const V2*>(w), reinterpret_cast<V2*>(acc)); return 0; }
  FOR_KINDS(X)
#undef X
  return 1;
}

int tp_bwd2(int l1, int lf, int lo, const float* x, const float* Y, const float* w, const float* ga,
            float* dw, float* dx, float* dY) {
#define X(a,b,c) if (l1==a && lf==b && lo==c) { TPKind<a,b,c>::bwd(reinterpret_cast<const V2*>(x), Y, reinterpret_cast<const V2*>(w), reinterpret_cast<const V2*>(ga), reinterpret_cast<V2*>(dw), reinterpret_cast<V2*>(dx), reinterpret_cast<V2*>(dY)); return 0; }
  FOR_KINDS(X)
#undef X
  return 1;
}

int sh_eval(int lmax, float x, float y, float z, float* Y) {
  if (lmax == 1) { SH<1>::eval(x, y, z, Y); return 0; }
  if (lmax == 2) { SH<2>::eval(x, y, z, Y); return 0; }
  if (lmax == 3) { SH<3>::eval(x, y, z, Y); return 0; }
  return 1;
}

int sh_vjp(int lmax, float x, float y, float z, const float* gY, float* g) {
  if (lmax == 1) { SH<1>::vjp(x, y, z, gY, g[0], g[1], g[2]); return 0; }
  if (lmax == 2) { SH<2>::vjp(x, y, z, gY, g[0], g[1], g[2]); return 0; }
  if (lmax == 3) { SH<3>::vjp(x, y, z, gY, g[0], g[1], g[2]); return 0; }
  return 1;
}
}
