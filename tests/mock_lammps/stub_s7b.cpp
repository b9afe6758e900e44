// Test double for the handful of libsevenn_b200 entry points the LAMMPS pair styles call -- CPU only, TEST
// INFRASTRUCTURE (tests/test_host_logic.py builds it into the mock-LAMMPS harness; nothing under sevenn_b200/ uses
// it).  It implements the STAGE PROTOCOL of include/sevenn_b200.h with a small message-passing toy model that has the
// same data dependencies as the real network (a layer needs the ghost rows of its input; the backward leaves
// contributions in ghost rows of dx and of the forces), so that a pair style's graph construction, row maps and
// exchange hooks can be checked end to end:
//   x_0[r][c]     = 0.1 (species[r] + 1)(c + 1)                                   every row, ghosts included
//   x_{t+1}[i][c] = tanh(a x_t[i][c] + b sum_{e: centre i} f(|v_e|) x_t[src_e][(c+1) % W])   owned rows; f(r) = exp(-r/2)
//   E             = sum_{owned i} sum_c x_T[i][c],     forces / virial from dE/dv_e as the engine defines them.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sevenn_b200.h"

struct S7bEngine {
  S7bModelDesc desc{};
  int T = 0, W = 0;
  int n_nodes = 0, n_local = 0;
  int64_t n_edges = 0;
  std::vector<int> species, centre, src;
  std::vector<float> vec;
  std::vector<std::vector<float>> x;      // x[t] : [n_nodes, W], t = 0..T
  std::vector<float> dx;                  // [n_nodes, W]: the rows travelling in the backward
  std::vector<float> g;                   // dE/dx_{t+1} of the owned rows
  std::vector<float> forces, edge_grad, atomic_energy, atomic_virial;
  bool want_atomic_virial = false;
  double energy = 0.0, virial[6] = {0, 0, 0, 0, 0, 0};
};

static std::string g_err;
static const float kA = 0.7f, kB = 0.2f;
static int fail(const std::string& m) { g_err = m; return 1; }

extern "C" {

const char* s7b_last_error(void) { return g_err.c_str(); }

int s7b_engine_create(const S7bModelDesc* desc, S7bEngine** out) {
  auto* e = new S7bEngine;
  e->desc = *desc;
  e->T = desc->n_layers;
  e->W = desc->muls[0][0];
  *out = e;
  return 0;
}
void s7b_engine_destroy(S7bEngine* e) { delete e; }
int s7b_engine_set_param(S7bEngine*, const char*, int, const float*, size_t) { return 0; }
int s7b_engine_set_atomic_virial(S7bEngine* e, int enable) { e->want_atomic_virial = enable != 0; return 0; }

int s7b_engine_set_graph_host(S7bEngine* e, int32_t n_nodes, int32_t n_local, int64_t n_edges, const int32_t* species,
                              const int32_t* edge_centre, const int32_t* edge_neighbour, const float* edge_vec, void*) {
  if (n_local > n_nodes) return fail("bad sizes");
  for (int64_t k = 0; k < n_edges; ++k) {
    if (edge_centre[k] < 0 || edge_centre[k] >= n_local) return fail("edge centres must be owned atoms (< n_local)");
    if (k > 0 && edge_centre[k] < edge_centre[k - 1]) return fail("edges must be sorted by centre");
    if (edge_neighbour[k] < 0 || edge_neighbour[k] >= n_nodes) return fail("edge neighbour index out of range");
  }
  e->n_nodes = n_nodes; e->n_local = n_local; e->n_edges = n_edges;
  e->species.assign(species, species + n_nodes);
  e->centre.assign(edge_centre, edge_centre + n_edges);
  e->src.assign(edge_neighbour, edge_neighbour + n_edges);
  e->vec.assign(edge_vec, edge_vec + 3 * n_edges);
  e->x.assign(e->T + 1, std::vector<float>((size_t)n_nodes * e->W, 0.0f));
  e->dx.assign((size_t)n_nodes * e->W, 0.0f);
  e->g.assign((size_t)n_local * e->W, 0.0f);
  e->forces.assign((size_t)n_nodes * 3, 0.0f);
  e->edge_grad.assign((size_t)n_edges * 3, 0.0f);
  e->atomic_energy.assign(n_local, 0.0f);
  e->atomic_virial.assign(e->want_atomic_virial ? (size_t)n_nodes * 6 : 0, 0.0f);
  return 0;
}

static float fcut(float r) { return std::exp(-0.5f * r); }

int s7b_engine_run_stage(S7bEngine* e, int stage, int t, void*) {
  const int W = e->W, T = e->T;
  switch (stage) {
    case S7B_STAGE_FWD_BEGIN:
      for (int r = 0; r < e->n_nodes; ++r)
        for (int c = 0; c < W; ++c) e->x[0][(size_t)r * W + c] = 0.1f * (e->species[r] + 1) * (c + 1);
      std::fill(e->edge_grad.begin(), e->edge_grad.end(), 0.0f);
      return 0;
    case S7B_STAGE_FWD_LAYER: {
      std::vector<float> acc((size_t)e->n_local * W, 0.0f);
      for (int64_t k = 0; k < e->n_edges; ++k) {
        const float* v = &e->vec[3 * k];
        const float f = fcut(std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
        for (int c = 0; c < W; ++c) acc[(size_t)e->centre[k] * W + c] += f * e->x[t][(size_t)e->src[k] * W + (c + 1) % W];
      }
      for (int i = 0; i < e->n_local; ++i)
        for (int c = 0; c < W; ++c)
          e->x[t + 1][(size_t)i * W + c] = std::tanh(kA * e->x[t][(size_t)i * W + c] + kB * acc[(size_t)i * W + c]);
      return 0;   // ghost rows of x[t+1] are the caller's job (exchange), exactly as with the real engine
    }
    case S7B_STAGE_FWD_END:
      e->energy = 0.0;
      for (int i = 0; i < e->n_local; ++i) {
        float s = 0.0f;
        for (int c = 0; c < W; ++c) s += e->x[T][(size_t)i * W + c];
        e->atomic_energy[i] = s;
        e->energy += s;
      }
      std::fill(e->g.begin(), e->g.end(), 1.0f);
      return 0;
    case S7B_STAGE_BWD_LAYER_A: {
      std::fill(e->dx.begin(), e->dx.end(), 0.0f);
      for (int i = 0; i < e->n_local; ++i)
        for (int c = 0; c < W; ++c) {
          const float y = e->x[t + 1][(size_t)i * W + c];
          e->dx[(size_t)i * W + c] += kA * e->g[(size_t)i * W + c] * (1.0f - y * y);
        }
      for (int64_t k = 0; k < e->n_edges; ++k) {
        const float* v = &e->vec[3 * k];
        const float r = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), f = fcut(r);
        const int i = e->centre[k], j = e->src[k];
        float dEdf = 0.0f;
        for (int c = 0; c < W; ++c) {
          const float y = e->x[t + 1][(size_t)i * W + c];
          const float d = kB * e->g[(size_t)i * W + c] * (1.0f - y * y);
          e->dx[(size_t)j * W + (c + 1) % W] += d * f;                  // j may be a ghost row
          dEdf += d * e->x[t][(size_t)j * W + (c + 1) % W];
        }
        const float dEdr = dEdf * (-0.5f * f);
        for (int a = 0; a < 3; ++a) e->edge_grad[3 * k + a] += dEdr * v[a] / r;
      }
      return 0;
    }
    case S7B_STAGE_BWD_LAYER_B:
      for (size_t q = 0; q < e->g.size(); ++q) e->g[q] = e->dx[q];      // owned rows, after the caller's reverse exchange
      return 0;
    case S7B_STAGE_BWD_END: {
      std::fill(e->forces.begin(), e->forces.end(), 0.0f);
      for (double& v : e->virial) v = 0.0;
      std::fill(e->atomic_virial.begin(), e->atomic_virial.end(), 0.0f);
      for (int64_t k = 0; k < e->n_edges; ++k) {
        const float* gk = &e->edge_grad[3 * k];          // dE/dv_e, v_e = r_src - r_centre
        const float* v = &e->vec[3 * k];
        for (int a = 0; a < 3; ++a) {
          e->forces[(size_t)e->centre[k] * 3 + a] += gk[a];
          e->forces[(size_t)e->src[k] * 3 + a] -= gk[a];
        }
        // the engine's conventions (edge_kernels.cuh): f_e = gk is what the centre atom receives; virial6 = -sum v (x) f_e in
        // the order xx yy zz xy yz zx; the per-atom virial is the same 6-vector, negated, on the NEIGHBOUR row
        const int ia[6] = {0, 1, 2, 0, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 0};
        for (int q = 0; q < 6; ++q) {
          const float w = v[ia[q]] * gk[ib[q]];
          e->virial[q] -= (double)w;
          if (e->want_atomic_virial) e->atomic_virial[(size_t)e->src[k] * 6 + q] -= w;
        }
      }
      return 0;
    }
    default:
      return fail("stub: stage not implemented");
  }
}

static float* buffer(S7bEngine* e, const std::string& nm, int layer, size_t* numel, int* width) {
  if (nm == "x" && layer >= 0 && layer <= e->T) { *numel = e->x[layer].size(); *width = e->W; return e->x[layer].data(); }
  if (nm == "dx") { *numel = e->dx.size(); *width = e->W; return e->dx.data(); }
  if (nm == "forces") { *numel = e->forces.size(); *width = 3; return e->forces.data(); }
  if (nm == "atomic_energy") { *numel = e->atomic_energy.size(); *width = 1; return e->atomic_energy.data(); }
  if (nm == "atomic_virial" && e->want_atomic_virial) { *numel = e->atomic_virial.size(); *width = 6; return e->atomic_virial.data(); }
  return nullptr;
}

static int rows_copy(S7bEngine* e, const char* name, int layer, int32_t row_begin, int32_t n_rows, int32_t width, float* host, bool to_host) {
  if (n_rows <= 0) return 0;
  size_t numel = 0;
  int w = 0;
  float* base = buffer(e, name, layer, &numel, &w);
  if (!base) return fail(std::string("no such buffer: ") + name);
  if (w != width) return fail("row width mismatch");
  if ((size_t)(row_begin + n_rows) * width > numel) return fail("row range exceeds the buffer");
  float* dev = base + (size_t)row_begin * width;
  if (to_host) std::memcpy(host, dev, sizeof(float) * n_rows * width);
  else std::memcpy(dev, host, sizeof(float) * n_rows * width);
  return 0;
}

int s7b_engine_read_rows_host(S7bEngine* e, const char* name, int layer, int32_t row_begin, int32_t n_rows, int32_t width,
                              float* host_out, void*) {
  return rows_copy(e, name, layer, row_begin, n_rows, width, host_out, true);
}
int s7b_engine_write_rows_host(S7bEngine* e, const char* name, int layer, int32_t row_begin, int32_t n_rows, int32_t width,
                               const float* host_in, void*) {
  return rows_copy(e, name, layer, row_begin, n_rows, width, const_cast<float*>(host_in), false);
}
int s7b_engine_read_scalars_host(S7bEngine* e, double* energy, double* virial6, void*) {
  if (energy) *energy = e->energy;
  if (virial6) std::memcpy(virial6, e->virial, sizeof(e->virial));
  return 0;
}

// what the serial pair style calls: graph without ghosts, all stages, results to the host
int s7b_engine_compute_host(S7bEngine* e, int32_t n_nodes, int64_t n_edges, const int32_t* species, const int32_t* edge_centre,
                            const int32_t* edge_neighbour, const float* edge_vec, double* energy, float* atomic_energy,
                            float* forces, double* virial, void* stream) {
  if (s7b_engine_set_graph_host(e, n_nodes, n_nodes, n_edges, species, edge_centre, edge_neighbour, edge_vec, stream)) return 1;
  s7b_engine_run_stage(e, S7B_STAGE_FWD_BEGIN, 0, stream);
  for (int t = 0; t < e->T; ++t) s7b_engine_run_stage(e, S7B_STAGE_FWD_LAYER, t, stream);
  s7b_engine_run_stage(e, S7B_STAGE_FWD_END, 0, stream);
  for (int t = e->T - 1; t >= 0; --t) {
    s7b_engine_run_stage(e, S7B_STAGE_BWD_LAYER_A, t, stream);
    if (t > 0) s7b_engine_run_stage(e, S7B_STAGE_BWD_LAYER_B, t, stream);
  }
  s7b_engine_run_stage(e, S7B_STAGE_BWD_END, 0, stream);
  if (energy) *energy = e->energy;
  if (virial) std::memcpy(virial, e->virial, sizeof(e->virial));
  if (atomic_energy) std::memcpy(atomic_energy, e->atomic_energy.data(), sizeof(float) * n_nodes);
  if (forces) std::memcpy(forces, e->forces.data(), sizeof(float) * 3 * n_nodes);
  return 0;
}

}  // extern "C"
