// pion_b200/host/pion_ugs_gpu.cpp -- uniform-grid driver on the B200 library, the
// counterpart of the reference's `pion-ugs` main (source/main.cpp:120-260 +
// ics/get_sim_info.cpp:90-600 for the parameter file): reads a PION parameter
// file (the `key value` text format of test_problems/*/params_*.txt), reads the initial
// primitive state as raw FP64 SoA [nvar][NZ+2g][NY+2g][NX+2g] (what dataio->ReadData
// delivers; file-format readers are out of scope), runs
// sim_control_gpu::Init / Time_Int / Finalise and writes the final state the same way.
//
//   pion_ugs_gpu <paramfile> --in P0.bin [--out P.bin] [--steps N] [--device D]
//               [--tables tables.bin] [--verbose]
// Without --in a built-in uniform state with a central over-pressured sphere is used
// (ambient = refvec, p x 200 inside r < L/4), enough to time the step from C++.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "sim_control_gpu.h"

using namespace pion_b200;

static std::map<std::string, std::string> read_params(const char* path) {
  std::map<std::string, std::string> kv;
  std::ifstream f(path);
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string k, v;
    if (ss >> k >> v) kv[k] = v;
  }
  return kv;
}
static double getd(const std::map<std::string, std::string>& kv, const char* k, double def) {
  auto it = kv.find(k);
  return it == kv.end() ? def : atof(it->second.c_str());
}
static int geti(const std::map<std::string, std::string>& kv, const char* k, int def) {
  auto it = kv.find(k);
  return it == kv.end() ? def : atoi(it->second.c_str());
}
static std::string gets(const std::map<std::string, std::string>& kv, const char* k, const char* def) {
  auto it = kv.find(k);
  return it == kv.end() ? std::string(def) : it->second;
}
// boundary names as in boundaries/assign_update_bcs.cpp / setup_fixed_grid.cpp:700-860
static int bc_code(const std::string& s) {
  if (s == "periodic") return PION_BC_PERIODIC;
  if (s == "outflow" || s == "zero-gradient") return PION_BC_OUTFLOW;
  if (s == "inflow") return PION_BC_INFLOW;
  if (s == "reflecting") return PION_BC_REFLECTING;
  if (s == "fixed") return PION_BC_FIXED;
  if (s == "DMR") return PION_BC_DMACH;
  if (s == "DMR2") return PION_BC_DMACH2;
  if (s == "one-way-outflow") return PION_BC_ONEWAY_OUT;
  if (s == "stellar-wind") return PION_BC_STWIND;
  return -1;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <paramfile> [--in P0.bin] [--out P.bin] [--steps N] [--device D] [--tables t.bin] [--verbose]\n", argv[0]);
    return 2;
  }
  const char *in = nullptr, *out = nullptr, *tables = nullptr;
  long steps = -1;
  int device = 0;
  bool verbose = false;
  for (int i = 2; i < argc; i++) {
    if (!strcmp(argv[i], "--in") && i + 1 < argc) in = argv[++i];
    else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
    else if (!strcmp(argv[i], "--tables") && i + 1 < argc) tables = argv[++i];
    else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atol(argv[++i]);
    else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--verbose")) verbose = true;
  }
  auto kv = read_params(argv[1]);
  sim_control_gpu sim;
  SimParamsGPU& p = sim.SimPM;
  p.ndim = geti(kv, "ndim", 3);
  const std::string eqn = gets(kv, "eqn", "glm-mhd");
  p.eqntype = (eqn == "euler") ? PION_EQEUL : (eqn == "i-mhd") ? PION_EQMHD : PION_EQGLM;
  p.ntracer = geti(kv, "ntracer", 0);
  p.nvar = ((p.eqntype == PION_EQEUL) ? 5 : (p.eqntype == PION_EQMHD) ? 8 : 9) + p.ntracer;
  p.solverType = geti(kv, "solver", 7);
  p.spOOA = geti(kv, "OrderOfAccSpace", 2);
  p.tmOOA = geti(kv, "OrderOfAccTime", 2);
  p.artviscosity = geti(kv, "ArtificialViscosity", 1);
  p.etav = getd(kv, "EtaViscosity", 0.15);
  p.gamma = getd(kv, "GAMMA", 5.0 / 3.0);
  p.CFL = getd(kv, "CFL", 0.3);
  const char* ax = "XYZ";
  const char* face[6] = {"BC_XN", "BC_XP", "BC_YN", "BC_YP", "BC_ZN", "BC_ZP"};
  for (int a = 0; a < 3; a++) {
    p.NG[a] = (a < p.ndim) ? geti(kv, (std::string("NGrid") + ax[a]).c_str(), 1) : 1;
    p.Xmin[a] = getd(kv, (std::string(1, ax[a]) + "min").c_str(), 0.0);
    p.Xmax[a] = getd(kv, (std::string(1, ax[a]) + "max").c_str(), 1.0);
  }
  for (int f = 0; f < 2 * p.ndim; f++) {
    p.BC[f] = bc_code(gets(kv, face[f], "outflow"));
    if (p.BC[f] < 0) { fprintf(stderr, "unsupported boundary %s\n", gets(kv, face[f], "?").c_str()); return 1; }
  }
  for (int i = 0; i < geti(kv, "BC_Ninternal", 0); i++) {
    char key[32];
    snprintf(key, sizeof key, "BC_INTERNAL_%03d", i);
    p.BC_internal.push_back(bc_code(gets(kv, key, "")));
  }
  for (int v = 0; v < PION_GPU_MAXVAR; v++) {
    char key[16];
    snprintf(key, sizeof key, "refvec%d", v);
    p.RefVec[v] = getd(kv, key, 1.0);
  }
  // stellar wind sources (ics/get_sim_info.cpp:702-875), constant type only
  for (int i = 0; i < geti(kv, "WIND_NSRC", 0); i++) {
    auto key = [&](const char* suffix) { return "WIND_" + std::to_string(i) + "_" + suffix; };
    if (geti(kv, key("type").c_str(), 0) != 0) { fprintf(stderr, "only constant winds (WIND_%d_type 0)\n", i); return 1; }
    pion_gpu_wind_source w;
    memset(&w, 0, sizeof w);
    for (int a = 0; a < 3; a++) w.dpos[a] = getd(kv, key(("pos" + std::to_string(a)).c_str()).c_str(), 0.0);
    w.radius = getd(kv, key("radius").c_str(), 0.0);
    w.mdot = getd(kv, key("mdot").c_str(), 0.0);
    w.vinf = getd(kv, key("vinf").c_str(), 0.0);
    w.vrot = getd(kv, key("vrot").c_str(), 0.0);
    w.temp = getd(kv, key("temp").c_str(), 0.0);
    w.rstar = getd(kv, key("Rstr").c_str(), 0.0);
    w.bsrf = getd(kv, key("Bsrf").c_str(), 0.0);
    for (int t = 0; t < p.ntracer && t < PION_GPU_MAXTR; t++) w.tr[t] = getd(kv, key(("TR" + std::to_string(t)).c_str()).c_str(), 0.0);
    p.SWP.push_back(w);
  }
  p.starttime = p.simtime = getd(kv, "StartTime", 0.0);
  p.finishtime = getd(kv, "FinishTime", 1e30);
  p.op_criterion = geti(kv, "OutputCriterion", 0);
  p.opfreq_time = getd(kv, "OPfreqTime", 0.0);
  p.opfreq = geti(kv, "OutputFrequency", 0);
  p.min_timestep = getd(kv, "min_timestep", 0.0);
  p.EP.cooling = geti(kv, "EP_cooling", 0);
  p.EP.MP_timestep_limit = geti(kv, "EP_MP_timestep_limit", 0);
  p.EP.MinTemperature = getd(kv, "EP_Min_Temperature", 0.0);
  p.EP.MaxTemperature = getd(kv, "EP_Max_Temperature", 1e99);
  if (p.EP.cooling) {
    if (!tables) { fprintf(stderr, "EP_cooling needs --tables (6 x n FP64 columns: T rrhp C_rrh C_ffhe C_fbdn C_cie)\n"); return 1; }
    std::ifstream tf(tables, std::ios::binary | std::ios::ate);
    const size_t n = (size_t)tf.tellg() / (6 * sizeof(double));
    tf.seekg(0);
    std::vector<double>* cols[6] = {&p.table_T, &p.table_rrhp, &p.table_C_rrh, &p.table_C_ffhe, &p.table_C_fbdn, &p.table_C_cie};
    for (auto* c : cols) {
      c->resize(n);
      tf.read(reinterpret_cast<char*>(c->data()), n * sizeof(double));
    }
  }

  int ext[3];
  sim.padded_extents(ext);
  std::vector<double> P(sim.state_size());
  const size_t plane = (size_t)ext[0] * ext[1] * ext[2];
  if (in) {
    std::ifstream f(in, std::ios::binary);
    f.read(reinterpret_cast<char*>(P.data()), P.size() * sizeof(double));
    if ((size_t)f.gcount() != P.size() * sizeof(double)) { fprintf(stderr, "%s: expected %zu doubles\n", in, P.size()); return 1; }
  } else {
    const int g = (p.spOOA == 2) ? 2 : 1;
    const double dx = (p.Xmax[0] - p.Xmin[0]) / p.NG[0], R = 0.25 * (p.Xmax[0] - p.Xmin[0]);
    for (int k = 0; k < ext[2]; k++)
      for (int j = 0; j < ext[1]; j++)
        for (int i = 0; i < ext[0]; i++) {
          const int ijk[3] = {i, j, k};
          double r2 = 0;
          for (int a = 0; a < p.ndim; a++) {
            const double x = p.Xmin[a] + (ijk[a] - g + 0.5) * dx - 0.5 * (p.Xmin[a] + p.Xmax[a]);
            r2 += x * x;
          }
          const size_t c = ((size_t)k * ext[1] + j) * ext[0] + i;
          for (int v = 0; v < p.nvar; v++) P[v * plane + c] = (v >= 2 && v <= 4) ? 0.0 : p.RefVec[v];
          if (p.eqntype == PION_EQGLM) P[8 * plane + c] = 0.0;
          if (r2 < R * R) P[plane + c] *= 200.0;
        }
  }
  if (sim.Init(device, P.data())) { fprintf(stderr, "Init: %s\n", sim.error().c_str()); return 1; }
  if (sim.Time_Int(steps, verbose)) { fprintf(stderr, "Time_Int: %s\n", sim.error().c_str()); return 1; }
  if (out) {
    if (sim.download_state(P.data())) { fprintf(stderr, "output_data: %s\n", sim.error().c_str()); return 1; }
    std::ofstream f(out, std::ios::binary);
    f.write(reinterpret_cast<const char*>(P.data()), P.size() * sizeof(double));
  }
  printf("final time %.12e after %d steps\n", sim.SimPM.simtime, sim.SimPM.timestep);
  sim.Finalise();
  return 0;
}
