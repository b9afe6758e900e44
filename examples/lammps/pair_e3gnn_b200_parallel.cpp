/* See pair_e3gnn_b200_parallel.h.  Per step, on every rank:
 *   full neighbour list -> graph over (owned atoms + the ghost atoms that are neighbours of an owned atom)
 *   -> s7b_engine_set_graph_host -> the stage loop of PairE3GNNParallel::compute (pair_e3gnn_parallel.cpp:345-441):
 *        FWD_BEGIN | for t: FWD_LAYER(t), forward_comm of x(t+1) | FWD_END |
 *        for t = T-1..0: BWD_LAYER_A(t), reverse_comm of dx(t), BWD_LAYER_B(t) | BWD_END
 *   -> forces of owned AND ghost rows into f (LAMMPS' own newton reverse communication finishes them),
 *      energy / virial of the owned atoms (LAMMPS sums over ranks).
 * Layer 0 needs no exchange: ghost species are known locally (the reference's segment-0 trick,
 * sevenn/model_build.py:383-421). */
#include "pair_e3gnn_b200_parallel.h"

#include <algorithm>
#include <string>

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neigh_list.h"
#include "neighbor.h"

#include "../s7b_model_file.h"

using namespace LAMMPS_NS;

PairE3GNNB200Parallel::PairE3GNNB200Parallel(LAMMPS *lmp) : Pair(lmp) {
  single_enable = 0;
  restartinfo = 0;
  one_coeff = 1;
  manybody_flag = 1;
  no_virial_fdotr_compute = 1;     // the virial comes from the edge forces, not from f . r
}

PairE3GNNB200Parallel::~PairE3GNNB200Parallel() {
  if (engine) s7b_engine_destroy(engine);
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
    memory->destroy(species_of_type);
  }
}

void PairE3GNNB200Parallel::allocate() {
  allocated = 1;
  const int n = atom->ntypes;
  memory->create(setflag, n + 1, n + 1, "pair:setflag");
  memory->create(cutsq, n + 1, n + 1, "pair:cutsq");
  memory->create(species_of_type, n + 1, "pair:species_of_type");
}

void PairE3GNNB200Parallel::settings(int narg, char ** /*arg*/) {
  if (narg != 0) error->all(FLERR, "Illegal pair_style command: e3gnn/b200/parallel takes no arguments");
}

// pair_coeff * * <model.s7b> <symbol of type 1> <symbol of type 2> ...
void PairE3GNNB200Parallel::coeff(int narg, char **arg) {
  if (!allocated) allocate();
  const int ntypes = atom->ntypes;
  if (narg != 3 + ntypes) error->all(FLERR, "Incorrect args for pair coefficients: * * model.s7b + one element per atom type");
  if (std::string(arg[0]) != "*" || std::string(arg[1]) != "*") error->all(FLERR, "e3gnn/b200/parallel: pair_coeff must start with * *");

  s7b_file::Model model;
  const std::string err = s7b_file::load(arg[2], model);
  if (!err.empty()) error->all(FLERR, ("e3gnn/b200/parallel: " + err).c_str());
  if (engine) s7b_engine_destroy(engine);
  engine = model.engine;
  cutoff = model.desc.cutoff;
  n_layers = model.desc.n_layers;
  dim_x.assign(n_layers, 0);
  comm_width = 0;
  for (int t = 0; t < n_layers; ++t) {
    for (int l = 0; l < model.desc.n_l[t]; ++l) dim_x[t] += (2 * l + 1) * model.desc.muls[t][l];
    if (t > 0) comm_width = std::max(comm_width, dim_x[t]);      // x(0) / dx(0) never travel
  }
  // LAMMPS sizes its swap buffers from these at init: one feature row per exchanged atom
  // (the reference: comm_forward = comm_reverse = comm_size from the .pt metadata, pair_e3gnn_parallel.cpp:611-612)
  comm_forward = comm_width;
  comm_reverse = comm_width;

  for (int t = 1; t <= ntypes; ++t) {
    const int z = s7b_file::atomic_number(arg[2 + t]);
    const auto it = model.species_of_z.find(z);
    if (z == 0 || it == model.species_of_z.end())
      error->all(FLERR, (std::string("e3gnn/b200/parallel: element ") + arg[2 + t] + " is not known to the model").c_str());
    species_of_type[t] = it->second;
  }
  for (int i = 1; i <= ntypes; ++i)
    for (int j = 1; j <= ntypes; ++j) {
      setflag[i][j] = 1;
      cutsq[i][j] = cutoff * cutoff;
    }
}

void PairE3GNNB200Parallel::init_style() {
  if (force->newton_pair == 0) error->all(FLERR, "Pair style e3gnn/b200/parallel requires newton pair on");
  neighbor->add_request(this, NeighConst::REQ_FULL);
}

double PairE3GNNB200Parallel::init_one(int /*i*/, int /*j*/) { return cutoff; }

// ---- the four hooks of Comm::forward_comm(Pair*) / reverse_comm(Pair*) ------------------------------------------
// atom_rows holds, per LAMMPS atom index, the row that is travelling (width floats); LAMMPS hands us doubles.
// A swap may forward atoms that are themselves ghosts (multi-hop brick communication), so every atom index has a
// slot, also the ghosts that are not in the graph.

int PairE3GNNB200Parallel::pack_forward_comm(int n, int *list, double *buf, int /*pbc_flag*/, int * /*pbc*/) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const float *row = atom_rows.data() + (size_t)list[i] * comm_width;
    for (int c = 0; c < comm_width; ++c) buf[m++] = row[c];
  }
  return m;
}

void PairE3GNNB200Parallel::unpack_forward_comm(int n, int first, double *buf) {
  int m = 0;
  for (int i = first; i < first + n; ++i) {
    float *row = atom_rows.data() + (size_t)i * comm_width;
    for (int c = 0; c < comm_width; ++c) row[c] = (float)buf[m++];
  }
}

int PairE3GNNB200Parallel::pack_reverse_comm(int n, int first, double *buf) {
  int m = 0;
  for (int i = first; i < first + n; ++i) {
    const float *row = atom_rows.data() + (size_t)i * comm_width;
    for (int c = 0; c < comm_width; ++c) buf[m++] = row[c];
  }
  return m;
}

void PairE3GNNB200Parallel::unpack_reverse_comm(int n, int *list, double *buf) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    float *row = atom_rows.data() + (size_t)list[i] * comm_width;
    for (int c = 0; c < comm_width; ++c) row[c] += (float)buf[m++];
  }
}

// ghost rows of x(layer) <- their owners' rows
void PairE3GNNB200Parallel::exchange_forward(int layer, int width) {
  const int nlocal = atom->nlocal, nall = atom->nlocal + atom->nghost;
  std::fill(atom_rows.begin(), atom_rows.begin() + (size_t)nall * comm_width, 0.0f);
  row_stage.resize((size_t)std::max(n_rows, 1) * width);
  if (s7b_engine_read_rows_host(engine, "x", layer, 0, n_owned, width, row_stage.data(), nullptr)) error->one(FLERR, s7b_last_error());
  for (int i = 0; i < nlocal; ++i) std::copy_n(row_stage.data() + (size_t)i * width, width, atom_rows.data() + (size_t)i * comm_width);
  comm->forward_comm(this);        // collective: called on every rank and layer, with or without ghosts
  const int n_ghost_rows = n_rows - n_owned;
  for (int g = 0; g < n_ghost_rows; ++g)
    std::copy_n(atom_rows.data() + (size_t)atom_of_row[n_owned + g] * comm_width, width, row_stage.data() + (size_t)g * width);
  if (s7b_engine_write_rows_host(engine, "x", layer, n_owned, n_ghost_rows, width, row_stage.data(), nullptr)) error->one(FLERR, s7b_last_error());
}

// owned rows of dx(layer) += the ghost rows that stand for them on this and other ranks
void PairE3GNNB200Parallel::exchange_reverse(int layer, int width) {
  const int nall = atom->nlocal + atom->nghost;
  std::fill(atom_rows.begin(), atom_rows.begin() + (size_t)nall * comm_width, 0.0f);
  row_stage.resize((size_t)std::max(n_rows, 1) * width);
  if (s7b_engine_read_rows_host(engine, "dx", layer, 0, n_rows, width, row_stage.data(), nullptr)) error->one(FLERR, s7b_last_error());
  for (int r = 0; r < n_rows; ++r) std::copy_n(row_stage.data() + (size_t)r * width, width, atom_rows.data() + (size_t)atom_of_row[r] * comm_width);
  comm->reverse_comm(this);
  for (int i = 0; i < n_owned; ++i) std::copy_n(atom_rows.data() + (size_t)i * comm_width, width, row_stage.data() + (size_t)i * width);
  if (s7b_engine_write_rows_host(engine, "dx", layer, 0, n_owned, width, row_stage.data(), nullptr)) error->one(FLERR, s7b_last_error());
}

void PairE3GNNB200Parallel::compute(int eflag, int vflag) {
  ev_init(eflag, vflag);

  double **x = atom->x;
  double **f = atom->f;
  const int *type = atom->type;
  const int nlocal = atom->nlocal, nall = atom->nlocal + atom->nghost;
  if (list->inum != nlocal) error->one(FLERR, "e3gnn/b200/parallel: the neighbour list must cover every owned atom");
  const int *numneigh = list->numneigh;
  int **firstneigh = list->firstneigh;
  const double cut2 = cutoff * cutoff;

  // rows: owned atoms by LAMMPS index, then the ghosts that are within the cutoff of an owned atom, in the order
  // of first appearance (the reference prunes its ghost set the same way, pair_e3gnn_parallel.cpp:282-290).
  // Every LAMMPS ghost atom -- also a periodic image of an owned atom -- is its own row: features are exchanged
  // per ghost, LAMMPS' communication pattern knows the owner.
  n_owned = nlocal;
  row_of_ghost.assign(atom->nghost, -1);
  atom_of_row.resize(nlocal);
  for (int i = 0; i < nlocal; ++i) atom_of_row[i] = i;
  edge_centre.clear();
  edge_neighbour.clear();
  edge_vec.clear();
  for (int i = 0; i < nlocal; ++i) {
    const int *jlist = firstneigh[i];                            // per-atom arrays; centres in index order: edges sorted by centre
    const int jnum = numneigh[i];
    for (int jj = 0; jj < jnum; ++jj) {
      const int j = jlist[jj] & NEIGHMASK;
      const double dx = x[j][0] - x[i][0], dy = x[j][1] - x[i][1], dz = x[j][2] - x[i][2];
      if (dx * dx + dy * dy + dz * dz >= cut2) continue;          // the list carries the skin
      int row = j;
      if (j >= nlocal) {
        int &slot = row_of_ghost[j - nlocal];
        if (slot < 0) {
          slot = (int)atom_of_row.size();
          atom_of_row.push_back(j);
        }
        row = slot;
      }
      edge_centre.push_back(i);
      edge_neighbour.push_back(row);
      edge_vec.push_back((float)dx);
      edge_vec.push_back((float)dy);
      edge_vec.push_back((float)dz);
    }
  }
  n_rows = (int)atom_of_row.size();
  species.resize(n_rows);
  for (int r = 0; r < n_rows; ++r) species[r] = species_of_type[type[atom_of_row[r]]];
  atom_rows.resize((size_t)std::max(nall, 1) * comm_width);

  if (vflag_atom && !atomic_virial_on) {       // per-atom virial on demand (pair_e3gnn.cpp:263-275)
    if (s7b_engine_set_atomic_virial(engine, 1)) error->one(FLERR, s7b_last_error());
    atomic_virial_on = true;
  }
  if (s7b_engine_set_graph_host(engine, n_rows, n_owned, (int64_t)edge_centre.size(), species.data(), edge_centre.data(),
                                edge_neighbour.data(), edge_vec.data(), nullptr))
    error->one(FLERR, s7b_last_error());

  auto stage = [&](int st, int t) {
    if (s7b_engine_run_stage(engine, st, t, nullptr)) error->one(FLERR, s7b_last_error());
  };
  stage(S7B_STAGE_FWD_BEGIN, 0);
  for (int t = 0; t < n_layers; ++t) {
    stage(S7B_STAGE_FWD_LAYER, t);
    if (t + 1 < n_layers) exchange_forward(t + 1, dim_x[t + 1]);
  }
  stage(S7B_STAGE_FWD_END, 0);
  for (int t = n_layers - 1; t >= 0; --t) {
    stage(S7B_STAGE_BWD_LAYER_A, t);
    if (t > 0) {
      exchange_reverse(t, dim_x[t]);
      stage(S7B_STAGE_BWD_LAYER_B, t);
    }
  }
  stage(S7B_STAGE_BWD_END, 0);

  // forces of every row (owned and ghost); LAMMPS' reverse communication (newton on) sums the ghost parts
  row_stage.resize((size_t)std::max(n_rows, 1) * 3);
  if (s7b_engine_read_rows_host(engine, "forces", 0, 0, n_rows, 3, row_stage.data(), nullptr)) error->one(FLERR, s7b_last_error());
  for (int r = 0; r < n_rows; ++r) {
    const int a = atom_of_row[r];
    f[a][0] += row_stage[3 * (size_t)r];
    f[a][1] += row_stage[3 * (size_t)r + 1];
    f[a][2] += row_stage[3 * (size_t)r + 2];
  }
  double energy = 0.0, v6[6] = {0, 0, 0, 0, 0, 0};
  if (s7b_engine_read_scalars_host(engine, &energy, v6, nullptr)) error->one(FLERR, s7b_last_error());
  if (eflag_global) eng_vdwl += energy;          // this rank's owned atoms; LAMMPS sums over ranks
  if (eflag_atom) {
    row_stage.resize((size_t)std::max(n_owned, 1));
    if (s7b_engine_read_rows_host(engine, "atomic_energy", 0, 0, n_owned, 1, row_stage.data(), nullptr)) error->one(FLERR, s7b_last_error());
    for (int i = 0; i < n_owned; ++i) eatom[i] += row_stage[i];
  }
  if (vflag_global) {
    // library order (xx, yy, zz, xy, yz, zx), value -sum r (x) dE/dr over the edges centred on owned atoms
    // ->  LAMMPS (xx, yy, zz, xy, xz, yz)
    virial[0] += v6[0];
    virial[1] += v6[1];
    virial[2] += v6[2];
    virial[3] += v6[3];
    virial[4] += v6[5];
    virial[5] += v6[4];
  }
  if (vflag_atom) {
    // per-atom virial: -(v (x) f_e) of every edge on its NEIGHBOUR row (owned or ghost), order xx yy zz xy yz zx; the
    // ghost parts are summed into their owners by the reverse communication of the compute that asked for vatom
    row_stage.resize((size_t)std::max(n_rows, 1) * 6);
    if (s7b_engine_read_rows_host(engine, "atomic_virial", 0, 0, n_rows, 6, row_stage.data(), nullptr)) error->one(FLERR, s7b_last_error());
    const int lm[6] = {0, 1, 2, 3, 5, 4};      // LAMMPS (xx, yy, zz, xy, xz, yz)
    for (int r = 0; r < n_rows; ++r)
      for (int q = 0; q < 6; ++q) vatom[atom_of_row[r]][q] += row_stage[(size_t)r * 6 + lm[q]];
  }
}
