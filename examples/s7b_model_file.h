// Loader for the flat model file written by sevenn_b200/export.py:export_flat -- shared by the
// stand-alone host (host_entry.cpp) and the LAMMPS pair style (lammps/pair_e3gnn_b200.cpp).
// Header-only, C++17, no dependency besides the C ABI in include/sevenn_b200.h.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../include/sevenn_b200.h"

namespace s7b_file {

struct Model {
  S7bEngine* engine = nullptr;
  S7bModelDesc desc{};
  std::map<int, int> species_of_z;   // atomic number -> species index of the model
};

inline bool read_exact(FILE* f, void* p, size_t n) { return std::fread(p, 1, n, f) == n; }

// Returns an empty string on success, otherwise what went wrong.
inline std::string load(const char* path, Model& m) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return std::string("cannot open ") + path;
  auto bail = [&](const std::string& why) {
    std::fclose(f);
    if (m.engine) { s7b_engine_destroy(m.engine); m.engine = nullptr; }
    return why;
  };
  char magic[8];
  int32_t version = 0, n_arrays = 0, n_types = 0;
  if (!read_exact(f, magic, 8) || std::memcmp(magic, "S7BMODEL", 8) != 0) return bail("not a sevenn_b200 model file");
  if (!read_exact(f, &version, 4) || version != 1) return bail("unsupported model file version");
  if (!read_exact(f, &m.desc, sizeof(m.desc)) || !read_exact(f, &n_arrays, 4) || !read_exact(f, &n_types, 4))
    return bail("truncated model file");
  for (int i = 0; i < n_types; ++i) {
    int32_t zi[2];
    if (!read_exact(f, zi, 8)) return bail("truncated model file");
    m.species_of_z[zi[0]] = zi[1];
  }
  if (s7b_engine_create(&m.desc, &m.engine) != 0) return bail(s7b_last_error());
  std::vector<float> buf;
  for (int a = 0; a < n_arrays; ++a) {
    char name[33] = {0};
    int32_t layer = 0;
    int64_t numel = 0;
    if (!read_exact(f, name, 32) || !read_exact(f, &layer, 4) || !read_exact(f, &numel, 8) || numel < 0)
      return bail("truncated model file");
    buf.resize((size_t)numel);
    if (!read_exact(f, buf.data(), (size_t)numel * sizeof(float))) return bail("truncated model file");
    if (s7b_engine_set_param(m.engine, name, layer, buf.data(), (size_t)numel) != 0) return bail(s7b_last_error());
  }
  std::fclose(f);
  return "";
}

// Atomic number of a chemical symbol (0 if unknown).
inline int atomic_number(const std::string& sym) {
  static const char* kSym[] = {
      "X",  "H",  "He", "Li", "Be", "B",  "C",  "N",  "O",  "F",  "Ne", "Na", "Mg", "Al", "Si", "P",  "S",  "Cl", "Ar", "K",
      "Ca", "Sc", "Ti", "V",  "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr", "Rb", "Sr", "Y",
      "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I",  "Xe", "Cs", "Ba", "La", "Ce", "Pr",
      "Nd", "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W",  "Re", "Os", "Ir", "Pt", "Au",
      "Hg", "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U",  "Np", "Pu"};
  for (int z = 1; z < (int)(sizeof(kSym) / sizeof(kSym[0])); ++z)
    if (sym == kSym[z]) return z;
  return 0;
}

}  // namespace s7b_file
