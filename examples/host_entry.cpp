// A Python-free host of the C ABI: what a LAMMPS pair style (pair_e3gnn.cpp:74-289 in the reference)
// would do with libsevenn_b200.so -- load a model, hand over centre-major edge arrays, read back
// energy / forces / virial.  Built and run by tests/test_cpp_host_gpu.py on the GPU box.
//
//   host_entry model.s7b graph.bin            graph-in  (s7b_engine_compute_host)
//   host_entry model.s7b positions.bin pos    positions-in (s7b_engine_compute_positions_host)
//
// graph.bin:     int32 n_atoms, int64 n_edges, int32 Z[n], int32 centre[E], int32 neighbour[E], float vec[E][3]
// positions.bin: int32 n_atoms, int32 pbc[3], double cell[9], int32 Z[n], double pos[n][3]
// Output (stdout): energy, then one line per atom: fx fy fz, then the 6 virial components.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "s7b_model_file.h"

#define CHECK(call)                                                     \
  do {                                                                  \
    if ((call) != 0) {                                                  \
      std::fprintf(stderr, "sevenn_b200 error: %s\n", s7b_last_error()); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

using s7b_file::read_exact;

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s model.s7b input.bin [pos]\n", argv[0]);
    return 2;
  }
  // ---- model -------------------------------------------------------------------------------------
  s7b_file::Model model;
  const std::string err = s7b_file::load(argv[1], model);
  if (!err.empty()) {
    std::fprintf(stderr, "%s\n", err.c_str());
    return 2;
  }
  S7bEngine* eng = model.engine;
  const std::map<int, int>& type_map = model.species_of_z;

  // ---- input -------------------------------------------------------------------------------------
  FILE* f = std::fopen(argv[2], "rb");
  if (!f) { std::perror(argv[2]); return 2; }
  int32_t n = 0;
  double energy = 0.0, virial[6];
  std::vector<float> forces;
  std::vector<int32_t> species;
  auto to_species = [&](const std::vector<int32_t>& z) {
    species.resize(z.size());
    for (size_t i = 0; i < z.size(); ++i) {
      auto it = type_map.find(z[i]);
      if (it == type_map.end()) { std::fprintf(stderr, "unknown element Z=%d\n", z[i]); return false; }
      species[i] = it->second;
    }
    return true;
  };
  if (argc > 3 && std::string(argv[3]) == "pos") {
    int32_t pbc[3];
    double cell[9];
    if (!read_exact(f, &n, 4) || !read_exact(f, pbc, 12) || !read_exact(f, cell, 72)) return 2;
    std::vector<int32_t> z(n);
    std::vector<double> pos((size_t)n * 3);
    if (!read_exact(f, z.data(), (size_t)n * 4) || !read_exact(f, pos.data(), (size_t)n * 24)) return 2;
    if (!to_species(z)) return 2;
    forces.resize((size_t)n * 3);
    int64_t n_edges = 0;
    CHECK(s7b_engine_compute_positions_host(eng, n, species.data(), pos.data(), cell, pbc, &energy, nullptr,
                                            forces.data(), virial, &n_edges, nullptr));
    std::fprintf(stderr, "edges: %lld\n", (long long)n_edges);
  } else {
    int64_t E = 0;
    if (!read_exact(f, &n, 4) || !read_exact(f, &E, 8)) return 2;
    std::vector<int32_t> z(n), centre((size_t)E), neigh((size_t)E);
    std::vector<float> vec((size_t)E * 3);
    if (!read_exact(f, z.data(), (size_t)n * 4) || !read_exact(f, centre.data(), (size_t)E * 4) ||
        !read_exact(f, neigh.data(), (size_t)E * 4) || !read_exact(f, vec.data(), (size_t)E * 12)) return 2;
    if (!to_species(z)) return 2;
    forces.resize((size_t)n * 3);
    CHECK(s7b_engine_compute_host(eng, n, E, species.data(), centre.data(), neigh.data(), vec.data(), &energy,
                                  nullptr, forces.data(), virial, nullptr));
  }
  std::fclose(f);
  std::printf("%.9f\n", energy);
  for (int i = 0; i < n; ++i) std::printf("%.7e %.7e %.7e\n", forces[3 * i], forces[3 * i + 1], forces[3 * i + 2]);
  std::printf("%.9e %.9e %.9e %.9e %.9e %.9e\n", virial[0], virial[1], virial[2], virial[3], virial[4], virial[5]);
  s7b_engine_destroy(eng);
  return 0;
}
