/* Multi-rank LAMMPS pair style backed by libsevenn_b200.so.
 *
 *   pair_style e3gnn/b200/parallel
 *   pair_coeff * * model.s7b Hf O        # one chemical symbol per LAMMPS atom type
 *
 * The user contract of the reference's `pair_style e3gnn/parallel` (sevenn/pair_e3gnn/
 * pair_e3gnn_parallel.cpp:194-528, 547-678): metal units, `newton_pair on`, a full neighbour list, one GPU per
 * MPI rank, LAMMPS owns the spatial decomposition.  Differences: ONE model file instead of one TorchScript
 * segment per layer (the stage API of the library plays the role of the segments), and the ghost-feature
 * exchanges go through LAMMPS' stock `Comm::forward_comm(Pair*)` / `reverse_comm(Pair*)` with the standard
 * pack/unpack hooks -- no patched comm_brick.cpp (the reference adds two methods to it,
 * comm_brick.cpp:1057-1123).  Rows travel through host buffers like the reference's without CUDA-aware MPI
 * (pair_e3gnn_parallel.cpp:698-799): s7b_engine_read_rows_host / s7b_engine_write_rows_host.
 * Written against LAMMPS stable_2Aug2023.  LAMMPS is not in the image: compiled against tests/mock_lammps/ and run there
 * on the CPU (one rank, periodic image ghosts, stock Comm hooks) against a toy double of the stage protocol
 * (tests/mock_lammps/harness_parallel.cpp, tests/test_host_logic.py). */
#ifdef PAIR_CLASS
// clang-format off
PairStyle(e3gnn/b200/parallel, PairE3GNNB200Parallel)
// clang-format on
#else
#ifndef LMP_PAIR_E3GNN_B200_PARALLEL_H
#define LMP_PAIR_E3GNN_B200_PARALLEL_H

#include <vector>

#include "pair.h"

struct S7bEngine;

namespace LAMMPS_NS {

class PairE3GNNB200Parallel : public Pair {
 public:
  PairE3GNNB200Parallel(class LAMMPS *);
  ~PairE3GNNB200Parallel() override;
  void compute(int, int) override;
  void settings(int, char **) override;
  void coeff(int, char **) override;
  void init_style() override;
  double init_one(int, int) override;
  int pack_forward_comm(int, int *, double *, int, int *) override;
  void unpack_forward_comm(int, int, double *) override;
  int pack_reverse_comm(int, int, double *) override;
  void unpack_reverse_comm(int, int *, double *) override;

 protected:
  void allocate();
  void exchange_forward(int layer, int width);
  void exchange_reverse(int layer, int width);

  S7bEngine *engine = nullptr;
  double cutoff = 0.0;
  int n_layers = 0;
  std::vector<int> dim_x;             // row width of x(t) / dx(t)
  int *species_of_type = nullptr;     // LAMMPS type -> species index of the model

  // graph rows: 0..nlocal-1 = owned atoms (LAMMPS index), then one row per ghost atom that is a neighbour
  int n_rows = 0, n_owned = 0;
  std::vector<int> row_of_ghost;      // LAMMPS ghost index - nlocal -> graph row, -1: not a neighbour of an owned atom
  std::vector<int> atom_of_row;       // graph row -> LAMMPS atom index
  std::vector<int> species, edge_centre, edge_neighbour;
  std::vector<float> edge_vec;

  // per-atom staging of the row being exchanged: [nlocal + nghost, comm_width] (LAMMPS atom index)
  int comm_width = 0;
  std::vector<float> atom_rows, row_stage;
  bool atomic_virial_on = false;
};

}  // namespace LAMMPS_NS
#endif
#endif
