// Host side of the D3 dispersion C ABI (include/sevenn_b200.h, "D3" section): parameters, cell list,
// stage launches and the reference-named entry points (pair_init ... pair_fin) that
// sevenn/calculator.py:430-483 binds with ctypes.  Kernels: d3_kernels.cuh.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sevenn_b200.h"
#include "common.cuh"
#include "d3_kernels.cuh"

namespace s7b {
static int d3_fail(const std::string& m) {
  set_error(__FILE__, 0, m.c_str());
  return 1;
}
constexpr double kAuToAng = 0.52917726, kAuToEv = 27.21138505;   // pair_d3_for_ase.h:200-201

struct D3Buf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    if (cudaMalloc(&p, need + need / 8 + 256) != cudaSuccess) { cudaGetLastError(); return 1; }
    bytes = need + need / 8 + 256;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
}  // namespace s7b

using namespace s7b;

struct S7bD3 {
  D3Params P;
  bool have_params = false, have_damping = false, have_system = false;
  int n = 0;
  NLGrid grid;
  int3 R_vdw, R_cn;
  D3Buf r0ab, c6ref, cnref, mxc;
  D3Buf pos, wrapped, type, key, key_sorted, idx, idx_sorted, bin_start, tmp;
  D3Buf xs, ts, bin_of, W, dW, logD, near_, cn, dc6i, force, energy, sigma, out_force;
  std::vector<double> host_force;      // reference ABI: pair_get_force returns a pointer
  double host_energy = 0.0, host_sigma[6] = {0, 0, 0, 0, 0, 0};
  // reference-ABI staging (pair_set_atom / pair_set_domain / pair_run_settings / pair_run_coeff)
  std::vector<int> ref_types;
  std::vector<double> ref_pos;
  double ref_cell[9] = {0}, ref_rthr = 9000.0, ref_cnthr = 1600.0;
  int ref_pbc[3] = {1, 1, 1}, ref_ntypes = 0;
  std::string ref_damp = "damp_bj", ref_func = "pbe";
};

static D3Atoms d3_atoms(const S7bD3* d) {
  D3Atoms A;
  A.x = d->xs.as<double>();
  A.type = d->ts.as<int>();
  A.W = d->W.as<float>();
  A.logD = d->logD.as<float>();
  A.near = d->near_.as<int>();
  A.dc6i = d->dc6i.as<double>();
  A.bin_start = d->bin_start.as<int>();
  A.bin_of = d->bin_of.as<int>();
  return A;
}
static D3Out d3_out(const S7bD3* d) {
  D3Out o;
  o.cn = d->cn.as<double>();
  o.dc6i = d->dc6i.as<double>();
  o.force = d->force.as<double>();
  o.energy = d->energy.as<double>();
  o.sigma = d->sigma.as<double>();
  return o;
}

extern "C" {

int s7b_d3_create(S7bD3** out) {
  if (!out) return d3_fail("null argument");
  *out = new S7bD3();
  memset(&(*out)->P, 0, sizeof(D3Params));
  return 0;
}

void s7b_d3_destroy(S7bD3* d) {
  if (!d) return;
  D3Buf* bufs[] = {&d->r0ab, &d->c6ref, &d->cnref, &d->mxc, &d->pos, &d->wrapped, &d->type, &d->key, &d->key_sorted, &d->idx,
                   &d->idx_sorted, &d->bin_start, &d->tmp, &d->xs, &d->ts, &d->bin_of, &d->W, &d->dW, &d->logD, &d->near_,
                   &d->cn, &d->dc6i, &d->force, &d->energy, &d->sigma, &d->out_force};
  for (D3Buf* b : bufs) b->release();
  delete d;
}

int s7b_d3_set_params(S7bD3* d, int32_t ntypes, const double* rcov, const double* r2r4, const double* r0ab,
                      const double* c6ref, const double* cnref, const int32_t* mxc) {
  if (!d || !rcov || !r2r4 || !r0ab || !c6ref || !cnref || !mxc) return d3_fail("null argument");
  if (ntypes < 1 || ntypes > kD3MaxTypes) return d3_fail("D3: 1.." + std::to_string(kD3MaxTypes) + " atom types are supported");
  d->P.ntypes = ntypes;
  std::vector<float> f0((size_t)ntypes * ntypes), fc((size_t)ntypes * ntypes * 25), fr((size_t)ntypes * 5);
  for (int t = 0; t < ntypes; ++t) { d->P.rcov[t] = (float)rcov[t]; d->P.r2r4[t] = (float)r2r4[t]; }
  for (size_t k = 0; k < f0.size(); ++k) f0[k] = (float)(r0ab[k] / kAuToAng);     // table is in Angstrom (:326-337)
  for (size_t k = 0; k < fc.size(); ++k) fc[k] = (float)c6ref[k];
  for (size_t k = 0; k < fr.size(); ++k) fr[k] = (float)cnref[k];
  if (d->r0ab.ensure(f0.size() * 4) || d->c6ref.ensure(fc.size() * 4) || d->cnref.ensure(fr.size() * 4) || d->mxc.ensure(ntypes * 4))
    return d3_fail("cudaMalloc failed for the D3 tables");
  S7B_CUDA_CHECK(cudaMemcpy(d->r0ab.p, f0.data(), f0.size() * 4, cudaMemcpyHostToDevice));
  S7B_CUDA_CHECK(cudaMemcpy(d->c6ref.p, fc.data(), fc.size() * 4, cudaMemcpyHostToDevice));
  S7B_CUDA_CHECK(cudaMemcpy(d->cnref.p, fr.data(), fr.size() * 4, cudaMemcpyHostToDevice));
  S7B_CUDA_CHECK(cudaMemcpy(d->mxc.p, mxc, ntypes * 4, cudaMemcpyHostToDevice));
  d->P.r0ab = d->r0ab.as<float>();
  d->P.c6ref = d->c6ref.as<float>();
  d->have_params = true;
  return 0;
}

int s7b_d3_set_damping(S7bD3* d, int32_t damping, double s6, double s8, double a1, double a2, double alp6, double alp8,
                       double vdw_cutoff_au2, double cn_cutoff_au2) {
  if (!d) return d3_fail("null argument");
  if (damping != 0 && damping != 1) return d3_fail("D3 damping must be 0 (zero) or 1 (Becke-Johnson)");
  if (!(vdw_cutoff_au2 > 0) || !(cn_cutoff_au2 > 0)) return d3_fail("D3 cutoffs must be positive");
  d->P.damping = damping;
  d->P.s6 = (float)s6; d->P.s8 = (float)s8; d->P.a1 = (float)a1; d->P.a2 = (float)a2;
  d->P.alp6 = (float)alp6; d->P.alp8 = (float)alp8;
  d->P.rthr = (double)(float)vdw_cutoff_au2;      // the reference compares float r^2 with float thresholds
  d->P.cnthr = (double)(float)cn_cutoff_au2;
  d->have_damping = true;
  return 0;
}

static int d3_invert3(const double* m, double* inv) {
  const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (fabs(det) < 1e-12) return 1;
  const double id = 1.0 / det;
  inv[0] = (m[4] * m[8] - m[5] * m[7]) * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  inv[3] = (m[5] * m[6] - m[3] * m[8]) * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  inv[6] = (m[3] * m[7] - m[4] * m[6]) * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return 0;
}

// positions [n,3] and cell rows in Angstrom; types 0-based indices into the tables of set_params
int s7b_d3_set_system(S7bD3* d, int32_t n, const int32_t* types, const double* positions, const double* cell9,
                      const int32_t* pbc3, void* stream) {
  if (!d || !types || !positions || !cell9 || !pbc3) return d3_fail("null argument");
  if (!d->have_params || !d->have_damping) return d3_fail("D3: set_params and set_damping come first");
  if (n < 1) return d3_fail("D3 needs at least one atom");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  NLGrid& g = d->grid;
  memset(&g, 0, sizeof(g));
  for (int k = 0; k < 9; ++k) g.cell[k] = cell9[k] / kAuToAng;
  if (d3_invert3(g.cell, g.inv)) return d3_fail("D3 requires a cell (non-singular lattice vectors)");
  for (int a = 0; a < 3; ++a) { g.pbc[a] = pbc3[a] ? 1 : 0; g.fmin[a] = 0.0; g.fspan[a] = 1.0; }
  g.cutoff2 = d->P.rthr;
  double height[3];
  for (int a = 0; a < 3; ++a) {
    const double nx = g.inv[0 * 3 + a], ny = g.inv[1 * 3 + a], nz = g.inv[2 * 3 + a];
    height[a] = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
  }
  const double w_target = 6.0 / kAuToAng;           // ~10 atoms per bin in a dense solid
  long long nbins = 1;
  int Rv[3], Rc[3];
  const double rc_v = sqrt(d->P.rthr), rc_c = sqrt(d->P.cnthr);
  for (int a = 0; a < 3; ++a) {
    int nb = (int)floor(height[a] / w_target);
    nb = std::max(1, std::min(nb, 128));
    g.nb[a] = nb;
    const double w = height[a] / nb;
    Rv[a] = (int)ceil(rc_v / w - 1e-12);
    Rc[a] = (int)ceil(rc_c / w - 1e-12);
    if (!g.pbc[a]) { Rv[a] = std::min(Rv[a], nb - 1); Rc[a] = std::min(Rc[a], nb - 1); }
    g.R[a] = Rv[a];
    nbins *= nb;
  }
  d->R_vdw = make_int3(Rv[0], Rv[1], Rv[2]);
  d->R_cn = make_int3(Rc[0], Rc[1], Rc[2]);
  // bohr, wrapped into the cell in ALL directions as the reference does (pair_d3_for_ase.cu:1198-1212)
  std::vector<double> x((size_t)n * 3);
  for (int i = 0; i < n; ++i) {
    double f[3];
    for (int a = 0; a < 3; ++a) {
      f[a] = (positions[3 * i] * g.inv[0 * 3 + a] + positions[3 * i + 1] * g.inv[1 * 3 + a] + positions[3 * i + 2] * g.inv[2 * 3 + a]) / kAuToAng;
      f[a] -= floor(f[a]);
    }
    for (int c = 0; c < 3; ++c) x[3 * (size_t)i + c] = f[0] * g.cell[0 * 3 + c] + f[1] * g.cell[1 * 3 + c] + f[2] * g.cell[2 * 3 + c];
  }
  for (int i = 0; i < n; ++i)
    if (types[i] < 0 || types[i] >= d->P.ntypes) return d3_fail("D3: atom type out of range");
  const size_t N = (size_t)n;
  int rc = 0;
  rc |= d->pos.ensure(N * 24); rc |= d->wrapped.ensure(N * 24); rc |= d->type.ensure(N * 4);
  rc |= d->key.ensure(N * 4); rc |= d->key_sorted.ensure(N * 4); rc |= d->idx.ensure(N * 4); rc |= d->idx_sorted.ensure(N * 4);
  rc |= d->bin_start.ensure(((size_t)nbins + 1) * 4);
  rc |= d->xs.ensure(N * 24); rc |= d->ts.ensure(N * 4); rc |= d->bin_of.ensure(N * 4);
  rc |= d->W.ensure(N * 20); rc |= d->dW.ensure(N * 20); rc |= d->logD.ensure(N * 4); rc |= d->near_.ensure(N * 4);
  rc |= d->cn.ensure(N * 8); rc |= d->dc6i.ensure(N * 8); rc |= d->force.ensure(N * 24); rc |= d->out_force.ensure(N * 24);
  rc |= d->energy.ensure(8); rc |= d->sigma.ensure(72);
  if (rc) return d3_fail("cudaMalloc failed for the D3 system");
  S7B_CUDA_CHECK(cudaMemcpyAsync(d->pos.p, x.data(), N * 24, cudaMemcpyHostToDevice, st));
  S7B_CUDA_CHECK(cudaMemcpyAsync(d->type.p, types, N * 4, cudaMemcpyHostToDevice, st));
  const int blk = 128, grd = (n + blk - 1) / blk;
  nl_bin_kernel<<<grd, blk, 0, st>>>(g, d->pos.as<double>(), n, d->key.as<int>(), d->idx.as<int>(), d->wrapped.as<double>());
  S7B_CUDA_CHECK(cudaGetLastError());
  size_t tmp_sort = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, d->key.as<int>(), d->key_sorted.as<int>(), d->idx.as<int>(), d->idx_sorted.as<int>(), n, 0, 32, st);
  if (d->tmp.ensure(tmp_sort + 256)) return d3_fail("cudaMalloc failed for cub workspace");
  size_t tmp = d->tmp.bytes;
  S7B_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d->tmp.p, tmp, d->key.as<int>(), d->key_sorted.as<int>(), d->idx.as<int>(), d->idx_sorted.as<int>(), n, 0, 32, st));
  nl_bin_start_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(d->key_sorted.as<int>(), n, (int)nbins, d->bin_start.as<int>());
  S7B_CUDA_CHECK(cudaGetLastError());
  d3_sort_gather_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d->idx_sorted.as<int>(), d->key_sorted.as<int>(), d->wrapped.as<double>(),
                                                        d->type.as<int>(), d->xs.as<double>(), d->ts.as<int>(), d->bin_of.as<int>());
  S7B_CUDA_CHECK(cudaGetLastError());
  S7B_CUDA_CHECK(cudaStreamSynchronize(st));     // `x` is a host temporary
  d->n = n;
  d->have_system = true;
  return 0;
}

// stage 1: coordination numbers of atoms [i_begin, i_end) (bin-sorted order);
// stage 2: reference weights of ALL atoms from cn[], then pair energy / forces / dE/dCN of the range
//          (zeroes energy and sigma first);  stage 3: CN chain-rule forces of the range (needs dc6i[] of all atoms)
int s7b_d3_run_stage(S7bD3* d, int32_t stage, int32_t i_begin, int32_t i_end, void* stream) {
  if (!d || !d->have_system) return d3_fail("D3: no system set");
  if (i_begin < 0 || i_end > d->n || i_begin > i_end) return d3_fail("D3: bad atom range");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n = d->n, cnt = i_end - i_begin;
  const int grd = (cnt + kD3WarpsPerBlock - 1) / kD3WarpsPerBlock, blk = 32 * kD3WarpsPerBlock;
  const D3Atoms A = d3_atoms(d);
  const D3Out O = d3_out(d);
  if (stage == 1) {
    if (cnt > 0) d3_cn_kernel<<<grd, blk, 0, st>>>(d->grid, A, d->P, d->R_cn, i_begin, i_end, O);
  } else if (stage == 2) {
    d3_weights_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, d->ts.as<int>(), d->cn.as<double>(), d->cnref.as<float>(), d->mxc.as<int>(),
                                                      d->W.as<float>(), d->dW.as<float>(), d->logD.as<float>(), d->near_.as<int>());
    S7B_CUDA_CHECK(cudaMemsetAsync(d->energy.p, 0, 8, st));
    S7B_CUDA_CHECK(cudaMemsetAsync(d->sigma.p, 0, 72, st));
    S7B_CUDA_CHECK(cudaMemsetAsync(d->force.p, 0, (size_t)n * 24, st));
    if (cnt > 0) d3_pair_kernel<<<grd, blk, 0, st>>>(d->grid, A, d->P, d->dW.as<float>(), d->R_vdw, i_begin, i_end, O);
  } else if (stage == 3) {
    if (cnt > 0) d3_chain_kernel<<<grd, blk, 0, st>>>(d->grid, A, d->P, d->R_cn, i_begin, i_end, O);
  } else {
    return d3_fail("D3: unknown stage");
  }
  S7B_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// device buffers, bin-sorted atom order: "cn" double[n], "dc6i" double[n], "force" double[n,3] (hartree/bohr),
// "energy" double[1] (hartree), "sigma" double[6] (hartree; xx,yy,zz,xy,xz,yz), "order" int[n] (sorted -> caller index)
void* s7b_d3_buffer(S7bD3* d, const char* name, size_t* numel) {
  if (!d || !name) return nullptr;
  const std::string nm(name);
  void* p = nullptr;
  size_t n = 0;
  if (nm == "cn") { p = d->cn.p; n = d->n; }
  else if (nm == "dc6i") { p = d->dc6i.p; n = d->n; }
  else if (nm == "force") { p = d->force.p; n = (size_t)d->n * 3; }
  else if (nm == "energy") { p = d->energy.p; n = 1; }
  else if (nm == "sigma") { p = d->sigma.p; n = 6; }
  else if (nm == "order") { p = d->idx_sorted.p; n = d->n; }
  if (numel) *numel = n;
  return p;
}

// energy (eV), forces [n,3] (eV/A, caller's atom order), sigma6 (eV; xx,yy,zz,xy,xz,yz of sum f (x) r) -> host
int s7b_d3_results_host(S7bD3* d, double* energy, double* forces, double* sigma6, void* stream) {
  if (!d || !d->have_system) return d3_fail("D3: no system set");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n = d->n;
  d3_unsort_kernel<<<(n * 3 + 255) / 256, 256, 0, st>>>(n, 3, d->idx_sorted.as<int>(), d->force.as<double>(), kAuToEv / kAuToAng, d->out_force.as<double>());
  S7B_CUDA_CHECK(cudaGetLastError());
  double e = 0.0, s[6];
  S7B_CUDA_CHECK(cudaMemcpyAsync(&e, d->energy.p, 8, cudaMemcpyDeviceToHost, st));
  S7B_CUDA_CHECK(cudaMemcpyAsync(s, d->sigma.p, 48, cudaMemcpyDeviceToHost, st));
  if (forces) S7B_CUDA_CHECK(cudaMemcpyAsync(forces, d->out_force.p, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
  S7B_CUDA_CHECK(cudaStreamSynchronize(st));
  if (energy) *energy = e * kAuToEv;
  if (sigma6) for (int q = 0; q < 6; ++q) sigma6[q] = s[q] * kAuToEv;
  return 0;
}

int s7b_d3_compute_host(S7bD3* d, double* energy, double* forces, double* sigma6, void* stream) {
  if (!d || !d->have_system) return d3_fail("D3: no system set");
  for (int stage = 1; stage <= 3; ++stage)
    if (s7b_d3_run_stage(d, stage, 0, d->n, stream)) return 1;
  return s7b_d3_results_host(d, energy, forces, sigma6, stream);
}

}  // extern "C"

// ---- the reference's own entry points (sevenn/pair_e3gnn/pair_d3_for_ase.cu:2034-2082) ------------------
// Same names, argument meaning and call order as the ctypes binding in sevenn/calculator.py:430-483, so the
// reference's D3Calculator can load this library in place of its pair_d3.so.  The parameter tables come from
// weights/d3_params.bin (tools/convert_d3_params.py), found through $S7B_D3_PARAMS or relative to this library.
namespace s7b {
struct D3Tables {
  std::vector<double> r0ab, c6ref, cnref, r2r4, rcov;
  std::vector<int> mxc;
  std::map<std::string, std::vector<double>> func;     // "damp_bj/pbe" -> {s6, rs6, s18, rs18, alp}
  bool ok = false;
  std::string error;
};
static D3Tables& d3_tables() {
  static D3Tables T;
  if (T.ok || !T.error.empty()) return T;
  std::string path;
  if (const char* env = getenv("S7B_D3_PARAMS")) path = env;
  else {
    Dl_info info;
    if (dladdr((void*)&d3_tables, &info) && info.dli_fname) {
      std::string lib(info.dli_fname);
      const size_t k = lib.rfind('/');
      path = (k == std::string::npos ? std::string(".") : lib.substr(0, k)) + "/../../weights/d3_params.bin";
    }
  }
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { T.error = "cannot open " + path; return T; }
  auto rd = [&](std::vector<double>& v, size_t n) { v.resize(n); return fread(v.data(), 8, n, f) == n; };
  std::vector<double> m;
  bool good = rd(T.r0ab, 94 * 94) && rd(T.c6ref, 94 * 94 * 25) && rd(T.cnref, 94 * 5) && rd(m, 94) && rd(T.r2r4, 94) && rd(T.rcov, 94);
  if (good) {
    T.mxc.assign(94, 0);
    for (int i = 0; i < 94; ++i) T.mxc[i] = (int)m[i];
    char line[256];
    while (fgets(line, sizeof(line), f)) {
      char damp[64], name[64];
      double v[5];
      if (sscanf(line, "%63s %63s %lf %lf %lf %lf %lf", damp, name, &v[0], &v[1], &v[2], &v[3], &v[4]) == 7)
        T.func[std::string(damp) + "/" + name] = std::vector<double>(v, v + 5);
    }
  }
  fclose(f);
  if (!good || T.func.empty()) T.error = "malformed " + path;
  else T.ok = true;
  return T;
}
}  // namespace s7b

extern "C" {

S7bD3* pair_init(void) {
  S7bD3* d = nullptr;
  return s7b_d3_create(&d) ? nullptr : d;
}

void pair_set_atom(S7bD3* d, int natoms, int ntypes, int* type, double* x_flat) {
  if (!d) return;
  d->ref_types.assign(type, type + natoms);                  // 1-based, as LAMMPS / the reference
  d->ref_pos.assign(x_flat, x_flat + (size_t)natoms * 3);
  d->ref_ntypes = ntypes;
}

void pair_set_domain(S7bD3* d, int xperiodic, int yperiodic, int zperiodic, double* boxlo, double* boxhi, double xy,
                     double xz, double yz) {
  if (!d) return;
  const double c[9] = {boxhi[0] - boxlo[0], 0, 0, xy, boxhi[1] - boxlo[1], 0, xz, yz, boxhi[2] - boxlo[2]};   // :889-897
  memcpy(d->ref_cell, c, sizeof(c));
  d->ref_pbc[0] = xperiodic; d->ref_pbc[1] = yperiodic; d->ref_pbc[2] = zperiodic;
}

void pair_run_settings(S7bD3* d, double rthr, double cnthr, const char* damp_name, const char* func_name) {
  if (!d) return;
  d->ref_rthr = rthr;
  d->ref_cnthr = cnthr;
  d->ref_damp = damp_name ? damp_name : "";
  d->ref_func = func_name ? func_name : "";
}

void pair_run_coeff(S7bD3* d, int* atomic_numbers) {
  if (!d) return;
  D3Tables& T = d3_tables();
  if (!T.ok) { d3_fail("D3 tables: " + T.error); fprintf(stderr, "Error: %s\n", s7b_last_error()); return; }
  const int nt = d->ref_ntypes;
  std::vector<double> rcov(nt), r2r4(nt), r0((size_t)nt * nt), c6((size_t)nt * nt * 25), cr((size_t)nt * 5);
  std::vector<int> mxc(nt);
  for (int a = 0; a < nt; ++a) {
    const int za = atomic_numbers[a] - 1;
    if (za < 0 || za >= 94) { d3_fail("D3: atomic number out of range"); return; }
    rcov[a] = T.rcov[za]; r2r4[a] = T.r2r4[za]; mxc[a] = T.mxc[za];
    for (int q = 0; q < 5; ++q) cr[a * 5 + q] = T.cnref[za * 5 + q];
    for (int b = 0; b < nt; ++b) {
      const int zb = atomic_numbers[b] - 1;
      r0[a * nt + b] = T.r0ab[za * 94 + zb];
      for (int q = 0; q < 25; ++q) c6[((size_t)a * nt + b) * 25 + q] = T.c6ref[((size_t)za * 94 + zb) * 25 + q];
    }
  }
  if (s7b_d3_set_params(d, nt, rcov.data(), r2r4.data(), r0.data(), c6.data(), cr.data(), mxc.data())) { fprintf(stderr, "Error: %s\n", s7b_last_error()); return; }
  auto it = T.func.find(d->ref_damp + "/" + d->ref_func);
  if (it == T.func.end()) { d3_fail("Functional name unknown"); fprintf(stderr, "Error: Functional name unknown\n"); return; }
  const std::vector<double>& p = it->second;    // s6, rs6, s18, rs18, alp  ->  setfuncpar (:608-631)
  if (s7b_d3_set_damping(d, d->ref_damp == "damp_bj" ? 1 : 0, p[0], p[2], p[1], p[3], p[4], p[4] + 2.0, d->ref_rthr, d->ref_cnthr))
    fprintf(stderr, "Error: %s\n", s7b_last_error());
}

void pair_run_compute(S7bD3* d) {
  if (!d) return;
  const int n = (int)d->ref_types.size();
  std::vector<int> t0(n);
  for (int i = 0; i < n; ++i) t0[i] = d->ref_types[i] - 1;
  d->host_force.assign((size_t)n * 3, 0.0);
  if (s7b_d3_set_system(d, n, t0.data(), d->ref_pos.data(), d->ref_cell, d->ref_pbc, nullptr) ||
      s7b_d3_compute_host(d, &d->host_energy, d->host_force.data(), d->host_sigma, nullptr))
    fprintf(stderr, "Error: %s\n", s7b_last_error());
}

double pair_get_energy(S7bD3* d) { return d ? d->host_energy : 0.0; }
double* pair_get_force(S7bD3* d) { return d ? d->host_force.data() : nullptr; }
double* pair_get_stress(S7bD3* d) { return d ? d->host_sigma : nullptr; }
void pair_fin(S7bD3* d) { s7b_d3_destroy(d); }

}  // extern "C"
