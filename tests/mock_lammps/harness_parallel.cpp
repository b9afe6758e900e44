// Mock-LAMMPS harness for examples/lammps/pair_e3gnn_b200_parallel.cpp and pair_e3gnn_b200.cpp (CPU, test infrastructure): one rank, a
// periodic box whose ghost atoms are the periodic images of its own atoms, a full neighbour list with a skin, and a
// Comm whose forward / reverse communication copies owner rows to the image ghosts and sums them back through the
// pair style's own pack / unpack hooks.  The engine is the toy stage-protocol double of stub_s7b.cpp.  The pair
// style's result (graph with ghost rows + exchanges between the stages) must equal the evaluation of the same toy
// model on the ghost-free graph (every neighbour mapped to its owner, as the serial pair style builds it).
#include <array>
#include <cmath>
#include <cstdio>
#include <random>
#include <stdexcept>
#include <vector>

#include "pair_e3gnn_b200.h"
#include "pair_e3gnn_b200_parallel.h"

#include "../../include/sevenn_b200.h"
#ifdef REAL_ENGINE
#include "../../examples/s7b_model_file.h"
#endif

namespace LAMMPS_NS {
static std::vector<int> g_owner;          // owner (local index) of every ghost
static Atom *g_atom = nullptr;
void Error::all(const char *f, int l, const char *m) { throw std::runtime_error(std::string(f) + ":" + std::to_string(l) + " " + m); }
void Error::one(const char *f, int l, const char *m) { throw std::runtime_error(std::string(f) + ":" + std::to_string(l) + " " + m); }
int Atom::map(tagint t) { return t - 1; }
int Atom::tag_consecutive() { return 1; }
void *Neighbor::add_request(Pair *, int) { return nullptr; }
void Pair::ev_init(int eflag, int vflag, int) {
  eflag_global = eflag & 1; eflag_atom = eflag & 2; vflag_global = vflag & 1; vflag_atom = vflag & 2;
  eng_vdwl = 0.0;
  for (double &v : virial) v = 0.0;
}
void Comm::forward_comm(Pair *p, int) {
  std::vector<double> buf(p->comm_forward);
  for (size_t g = 0; g < g_owner.size(); ++g) {
    int o = g_owner[g];
    p->pack_forward_comm(1, &o, buf.data(), 0, nullptr);
    p->unpack_forward_comm(1, g_atom->nlocal + (int)g, buf.data());
  }
}
void Comm::reverse_comm(Pair *p, int) {
  std::vector<double> buf(p->comm_reverse);
  for (size_t g = 0; g < g_owner.size(); ++g) {
    int o = g_owner[g];
    p->pack_reverse_comm(1, g_atom->nlocal + (int)g, buf.data());
    p->unpack_reverse_comm(1, &o, buf.data());
  }
}
}  // namespace LAMMPS_NS

using namespace LAMMPS_NS;

static const int kLayers = 3, kWidth = 4;

#ifdef REAL_ENGINE
static const char *g_model_path = nullptr;
template <class P>
static void wire_real(P &p) {           // the user-facing path: pair_style ... / pair_coeff * * model.s7b Si
  p.settings(0, nullptr);
  char a0[] = "*", a1[] = "*", sym[] = "Si";
  std::string mp(g_model_path);
  char *args[4] = {a0, a1, mp.data(), sym};
  p.coeff(4, args);
  p.init_style();
}
#endif

struct TestPair : PairE3GNNB200Parallel {
  Error err; Memory mem; Force frc; Neighbor nb; Comm cm;
  explicit TestPair(Atom *a) : PairE3GNNB200Parallel(nullptr) {
    error = &err; memory = &mem; atom = a; force = &frc; neighbor = &nb; comm = &cm;
    frc.newton_pair = 1;
#ifdef REAL_ENGINE
    wire_real(*this);
#else
    S7bModelDesc d{};
    d.n_layers = kLayers;
    d.cutoff = 3.0f;
    for (int t = 0; t <= kLayers; ++t) { d.n_l[t] = 1; d.muls[t][0] = kWidth; }
    if (s7b_engine_create(&d, &engine)) throw std::runtime_error("create");
    allocate();
    cutoff = d.cutoff;
    n_layers = kLayers;
    dim_x.assign(kLayers, kWidth);
    comm_width = kWidth;
    comm_forward = comm_reverse = kWidth;
    for (int t = 1; t <= a->ntypes; ++t) species_of_type[t] = t - 1;
#endif
  }
  void set_list(NeighList *l) { list = l; }
};

struct TestSerialPair : PairE3GNNB200 {
  Error err; Memory mem; Force frc; Neighbor nb; Comm cm;
  explicit TestSerialPair(Atom *a) : PairE3GNNB200(nullptr) {
    error = &err; memory = &mem; atom = a; force = &frc; neighbor = &nb; comm = &cm;
    frc.newton_pair = 1;
#ifdef REAL_ENGINE
    wire_real(*this);
#else
    S7bModelDesc d{};
    d.n_layers = kLayers;
    d.cutoff = 3.0f;
    for (int t = 0; t <= kLayers; ++t) { d.n_l[t] = 1; d.muls[t][0] = kWidth; }
    if (s7b_engine_create(&d, &engine)) throw std::runtime_error("create");
    allocate();
    cutoff = d.cutoff;
    for (int t = 1; t <= a->ntypes; ++t) species_of_type[t] = t - 1;
#endif
  }
  void set_list(NeighList *l) { list = l; }
};

int main(int argc, char **argv) {
  (void)argc; (void)argv;
#ifdef REAL_ENGINE
  // the real library on the GPU: diamond Si 2x2x2 (64 atoms, a = 5.431), positions jittered by up to +-0.05 A
  if (argc < 2) { std::printf("usage: harness <model.s7b>\n"); return 2; }
  g_model_path = argv[1];
  const double a0 = 5.431, L = 2 * a0, rc = 5.0, skin = 0.5;
  const int n = 64, ntypes = 1;
  const double basis[8][3] = {{0, 0, 0}, {0, .5, .5}, {.5, 0, .5}, {.5, .5, 0}, {.25, .25, .25}, {.25, .75, .75}, {.75, .25, .75}, {.75, .75, .25}};
  std::mt19937 rng(11);
  std::uniform_real_distribution<double> J(-0.05, 0.05);
  std::vector<std::vector<double>> pos;
  std::vector<int> type;
  for (int ix = 0; ix < 2; ++ix) for (int iy = 0; iy < 2; ++iy) for (int iz = 0; iz < 2; ++iz)
    for (const auto &b : basis) {
      double p[3] = {(ix + b[0]) * a0 + J(rng), (iy + b[1]) * a0 + J(rng), (iz + b[2]) * a0 + J(rng)};
      for (double &c : p) c -= L * std::floor(c / L);            // LAMMPS keeps owned atoms inside the box
      pos.push_back({p[0], p[1], p[2]});
      type.push_back(1);
    }
#else
  const double L = 6.0, rc = 3.0, skin = 0.5;
  const int n = 11, ntypes = 2;
  std::mt19937 rng(7);
  std::uniform_real_distribution<double> U(0.0, L);
  std::vector<std::vector<double>> pos;
  std::vector<int> type;
  for (int i = 0; i < n; ++i) { pos.push_back({U(rng), U(rng), U(rng)}); type.push_back(1 + i % 2); }
#endif
  // ghosts: periodic images within rc + skin of the box
  std::vector<int> owner;
  for (int sx = -1; sx <= 1; ++sx) for (int sy = -1; sy <= 1; ++sy) for (int sz = -1; sz <= 1; ++sz) {
    if (!sx && !sy && !sz) continue;
    for (int i = 0; i < n; ++i) {
      const double p[3] = {pos[i][0] + sx * L, pos[i][1] + sy * L, pos[i][2] + sz * L};
      bool in = true;
      for (int a = 0; a < 3; ++a) in = in && p[a] > -(rc + skin) && p[a] < L + rc + skin;
      if (in) { pos.push_back({p[0], p[1], p[2]}); type.push_back(type[i]); owner.push_back(i); }
    }
  }
  const int nall = (int)pos.size();
  Atom atom{};
  atom.ntypes = ntypes; atom.nlocal = n; atom.nghost = nall - n; atom.map_style = Atom::MAP_ARRAY;
  std::vector<double *> xp(nall), fp(nall);
  std::vector<double> fflat((size_t)nall * 3, 0.0);
  std::vector<tagint> tag(nall);
  for (int i = 0; i < nall; ++i) { xp[i] = pos[i].data(); fp[i] = &fflat[3 * (size_t)i]; tag[i] = (i < n ? i : owner[i - n]) + 1; }
  atom.x = xp.data(); atom.f = fp.data(); atom.type = type.data(); atom.tag = tag.data();
  g_owner = owner;
  g_atom = &atom;
  // full neighbour list with the skin
  std::vector<std::vector<int>> neigh(n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < nall; ++j) {
      if (j == i) continue;
      double d2 = 0;
      for (int a = 0; a < 3; ++a) d2 += (pos[j][a] - pos[i][a]) * (pos[j][a] - pos[i][a]);
      if (d2 < (rc + skin) * (rc + skin)) neigh[i].push_back(j);
    }
  std::vector<int> ilist(n), numneigh(n);
  std::vector<int *> firstneigh(n);
  for (int i = 0; i < n; ++i) { ilist[i] = n - 1 - i; numneigh[i] = (int)neigh[i].size(); firstneigh[i] = neigh[i].data(); }   // ilist in reverse order on purpose
  NeighList list{n, ilist.data(), numneigh.data(), firstneigh.data()};

  std::vector<double> eatom(nall, 0.0);
  std::vector<double> vflat((size_t)nall * 6, 0.0), vflat_s((size_t)nall * 6, 0.0);
  std::vector<double *> vrow(nall), vrow_s(nall);
  for (int i = 0; i < nall; ++i) { vrow[i] = &vflat[6 * (size_t)i]; vrow_s[i] = &vflat_s[6 * (size_t)i]; }
  double e_pair = 0, v_pair[6];
  try {
    TestPair pair(&atom);
    pair.set_list(&list);
    pair.eatom = eatom.data();
    pair.vatom = vrow.data();
    pair.compute(3, 3);
    e_pair = pair.eng_vdwl;
    for (int q = 0; q < 6; ++q) v_pair[q] = pair.virial[q];
  } catch (const std::exception &ex) {
    std::printf("FAIL pair style raised: %s\n", ex.what());
    return 1;
  }
  // the serial pair style on the same mock system (it maps image neighbours to their owners through the atom map)
  std::vector<double> f_serial((size_t)n * 3, 0.0), eatom_s(nall, 0.0);
  double e_serial = 0, v_serial[6];
  {
    std::vector<double> keep = fflat;
    std::fill(fflat.begin(), fflat.end(), 0.0);
    try {
      TestSerialPair sp(&atom);
      sp.set_list(&list);
      sp.eatom = eatom_s.data();
      sp.vatom = vrow_s.data();
      sp.compute(3, 3);
      e_serial = sp.eng_vdwl;
      for (int q = 0; q < 6; ++q) v_serial[q] = sp.virial[q];
    } catch (const std::exception &ex) {
      std::printf("FAIL serial pair style raised: %s\n", ex.what());
      return 1;
    }
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) f_serial[3 * (size_t)i + a] = fflat[3 * (size_t)i + a];
    fflat = keep;
  }
  // LAMMPS' newton reverse communication of the forces: ghosts -> owners
  std::vector<double> f_pair((size_t)n * 3, 0.0);
  for (int i = 0; i < nall; ++i)
    for (int a = 0; a < 3; ++a) f_pair[3 * (size_t)(i < n ? i : owner[i - n]) + a] += fflat[3 * (size_t)i + a];

#ifdef REAL_ENGINE
  // reference: the library's own positions-in entry (device neighbour list) on the periodic cell
  std::vector<float> f_ref((size_t)n * 3), e_ref_atom(n);
  double e_ref = 0, v_ref[6];
  std::vector<std::array<int, 2>> centre;      // only its size is reported
  {
    s7b_file::Model m;
    const std::string err = s7b_file::load(g_model_path, m);
    if (!err.empty()) { std::printf("FAIL %s\n", err.c_str()); return 1; }
    std::vector<int> species(n, m.species_of_z.at(14));
    std::vector<double> flat((size_t)n * 3);
    for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) flat[3 * (size_t)i + a] = pos[i][a];
    const double cell9[9] = {L, 0, 0, 0, L, 0, 0, 0, L};
    const int pbc3[3] = {1, 1, 1};
    int64_t ne = 0;
    if (s7b_engine_compute_positions_host(m.engine, n, species.data(), flat.data(), cell9, pbc3, &e_ref, e_ref_atom.data(), f_ref.data(), v_ref, &ne, nullptr)) {
      std::printf("FAIL reference: %s\n", s7b_last_error());
      return 1;
    }
    centre.resize((size_t)ne);
    s7b_engine_destroy(m.engine);
  }
  const double tolE = 2e-4, tolF = 1e-4, tolV = 5e-3, tolEa = 2e-5;
  std::vector<float> av_ref;              // the positions-in entry returns no per-atom virial: only the sum rule is checked
#else
  // the same toy model on the ghost-free graph
  std::vector<int> species(n), centre, nbr;
  std::vector<float> vec;
  for (int i = 0; i < n; ++i) species[i] = type[i] - 1;
  for (int i = 0; i < n; ++i)
    for (int j : neigh[i]) {
      double d2 = 0, d[3];
      for (int a = 0; a < 3; ++a) { d[a] = pos[j][a] - pos[i][a]; d2 += d[a] * d[a]; }
      if (d2 >= rc * rc) continue;
      centre.push_back(i); nbr.push_back(j < n ? j : owner[j - n]);
      for (int a = 0; a < 3; ++a) vec.push_back((float)d[a]);
    }
  S7bModelDesc d{};
  d.n_layers = kLayers;
  for (int t = 0; t <= kLayers; ++t) { d.n_l[t] = 1; d.muls[t][0] = kWidth; }
  S7bEngine *ref = nullptr;
  s7b_engine_create(&d, &ref);
  s7b_engine_set_atomic_virial(ref, 1);
  if (s7b_engine_set_graph_host(ref, n, n, (int64_t)centre.size(), species.data(), centre.data(), nbr.data(), vec.data(), nullptr)) {
    std::printf("FAIL reference graph: %s\n", s7b_last_error());
    return 1;
  }
  s7b_engine_run_stage(ref, S7B_STAGE_FWD_BEGIN, 0, nullptr);
  for (int t = 0; t < kLayers; ++t) s7b_engine_run_stage(ref, S7B_STAGE_FWD_LAYER, t, nullptr);
  s7b_engine_run_stage(ref, S7B_STAGE_FWD_END, 0, nullptr);
  for (int t = kLayers - 1; t >= 0; --t) {
    s7b_engine_run_stage(ref, S7B_STAGE_BWD_LAYER_A, t, nullptr);
    if (t > 0) s7b_engine_run_stage(ref, S7B_STAGE_BWD_LAYER_B, t, nullptr);
  }
  s7b_engine_run_stage(ref, S7B_STAGE_BWD_END, 0, nullptr);
  std::vector<float> f_ref((size_t)n * 3), e_ref_atom(n), av_ref((size_t)n * 6);
  double e_ref = 0, v_ref[6];
  s7b_engine_read_rows_host(ref, "atomic_virial", 0, 0, n, 6, av_ref.data(), nullptr);
  s7b_engine_read_rows_host(ref, "forces", 0, 0, n, 3, f_ref.data(), nullptr);
  s7b_engine_read_rows_host(ref, "atomic_energy", 0, 0, n, 1, e_ref_atom.data(), nullptr);
  s7b_engine_read_scalars_host(ref, &e_ref, v_ref, nullptr);

  s7b_engine_destroy(ref);
  const double tolE = 1e-5, tolF = 1e-5, tolV = 1e-4, tolEa = 1e-6;
#endif
  double de = std::fabs(e_pair - e_ref), df = 0, dv = 0, dea = 0, fmax = 0;
  for (size_t q = 0; q < f_pair.size(); ++q) { df = std::fmax(df, std::fabs(f_pair[q] - f_ref[q])); fmax = std::fmax(fmax, std::fabs(f_ref[q])); }
  const int lm[6] = {0, 1, 2, 3, 5, 4};      // LAMMPS (xx,yy,zz,xy,xz,yz) <- library (xx,yy,zz,xy,yz,zx)
  for (int q = 0; q < 6; ++q) dv = std::fmax(dv, std::fabs(v_pair[q] - v_ref[lm[q]]));
  for (int i = 0; i < n; ++i) dea = std::fmax(dea, std::fabs(eatom[i] - e_ref_atom[i]));
  double des = std::fabs(e_serial - e_ref), dfs = 0, dvs = 0, deas = 0;
  for (size_t q = 0; q < f_serial.size(); ++q) dfs = std::fmax(dfs, std::fabs(f_serial[q] - f_ref[q]));
  for (int q = 0; q < 6; ++q) dvs = std::fmax(dvs, std::fabs(v_serial[q] - v_ref[lm[q]]));
  for (int i = 0; i < n; ++i) deas = std::fmax(deas, std::fabs(eatom_s[i] - e_ref_atom[i]));
  // per-atom virial: ghosts folded into their owners (what the reverse communication of compute stress/atom does);
  // must sum to the global virial and, in stub mode, equal the ghost-free per-atom values
  double dva = 0, dvas = 0, dsum = 0, dsum_s = 0;
  {
    std::vector<double> va((size_t)n * 6, 0.0), vas((size_t)n * 6, 0.0);
    for (int i = 0; i < nall; ++i)
      for (int q = 0; q < 6; ++q) {
        va[6 * (size_t)(i < n ? i : owner[i - n]) + q] += vflat[6 * (size_t)i + q];
        vas[6 * (size_t)(i < n ? i : owner[i - n]) + q] += vflat_s[6 * (size_t)i + q];
      }
    for (int q = 0; q < 6; ++q) {
      double t = 0, ts = 0;
      for (int i = 0; i < n; ++i) { t += va[6 * (size_t)i + q]; ts += vas[6 * (size_t)i + q]; }
      dsum = std::fmax(dsum, std::fabs(t - v_pair[q]));
      dsum_s = std::fmax(dsum_s, std::fabs(ts - v_serial[q]));
    }
    if (!av_ref.empty())
      for (int i = 0; i < n; ++i)
        for (int q = 0; q < 6; ++q) {
          dva = std::fmax(dva, std::fabs(va[6 * (size_t)i + q] - av_ref[6 * (size_t)i + lm[q]]));
          dvas = std::fmax(dvas, std::fabs(vas[6 * (size_t)i + q] - av_ref[6 * (size_t)i + lm[q]]));
        }
  }
  std::printf("per-atom virial: |sum - virial| %.2e (parallel) %.2e (serial); vs ghost-free per atom %.2e / %.2e\n", dsum, dsum_s, dva, dvas);
  std::printf("serial style: |dE| %.2e  max|dF| %.2e  max|dV| %.2e  max|dEatom| %.2e\n", des, dfs, dvs, deas);
  std::printf("atoms %d ghosts %d edges %zu  E %.6f  |dE| %.2e  max|dF| %.2e (max|F| %.2e)  max|dV| %.2e  max|dEatom| %.2e\n",
              n, nall - n, centre.size(), e_ref, de, df, fmax, dv, dea);
  const bool ok = de < tolE && df < tolF && dv < tolV && dea < tolEa && fmax > 1e-3 && nall > n && centre.size() > (size_t)n &&
                  des < tolE && dfs < tolF && dvs < tolV && deas < tolEa &&
                  dsum < tolV && dsum_s < tolV && dva < tolV && dvas < tolV;
  std::printf(ok ? "OK\n" : "FAIL\n");
  return ok ? 0 : 1;
}
