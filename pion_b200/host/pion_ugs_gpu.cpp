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
//               [--in-text ics.txt] [--out-text base] [--convert]
// --in-text reads the state from the reference's ASCII format (dataio_text::output_ascii_data,
// dataIO/dataio_text.cpp:477-555: what icgen writes with "OutputFileType text"), --out-text writes
// <base>.<timestep, 8 digits>.txt in that format (dataio_text::OutputData / set_filename, :130-175);
// --convert only translates between the formats (no device needed; ghost cells stay zero).
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

// ---- the reference's ASCII format (dataio_text::output_ascii_data) --------------------------------------------
// "# format" / "# time" header lines, a blank line before every row of cells, then per cell (x fastest):
//   x [y [z]]  P[0..nvar-1]  T|eint  [ptot  divB]      scientific, 14 digits; B is written in Gauss-type
// units, B sqrt(4 pi) (NEW_B_NORM, defines/functionality_flags.h:42); T = p Mu_tot/(kB rho) with microphysics
// (mp_only_cooling::Temperature), else eint = p/((gamma-1) rho); the MHD columns are p + B^2/2 and the central-
// difference div B (VectorOps Divergence) of the unscaled field.
static const double SQRT4PI = std::sqrt(4.0 * M_PI);
static const double MU_TOT_OVER_KB = (0.609 * 1.672621898e-24) / 1.38064852e-16;  // mp_only_cooling.cpp:79-81

static int read_text_state(const char* path, const SimParamsGPU& p, const int ext[3], int g, std::vector<double>& P,
                           double* simtime, int* timestep) {
  std::ifstream f(path);
  if (!f.is_open()) { fprintf(stderr, "%s: cannot open\n", path); return 1; }
  const size_t plane = (size_t)ext[0] * ext[1] * ext[2];
  const bool mhd = p.eqntype != PION_EQEUL;
  const long ncell = (long)p.NG[0] * p.NG[1] * p.NG[2];
  long n = 0;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    if (line[0] == '#') {
      double t; int ts;
      if (sscanf(line.c_str(), "# time = %lf timestep = %d", &t, &ts) == 2) { if (simtime) *simtime = t; if (timestep) *timestep = ts; }
      continue;
    }
    std::istringstream ss(line);
    double x;
    for (int a = 0; a < p.ndim; a++) ss >> x;  // cell-centre position: the order in the file IS the grid order
    if (n >= ncell) { fprintf(stderr, "%s: more than %ld cells\n", path, ncell); return 1; }
    const int i = (int)(n % p.NG[0]), j = (int)((n / p.NG[0]) % p.NG[1]), k = (int)(n / ((long)p.NG[0] * p.NG[1]));
    const size_t c = ((size_t)(k + (p.ndim > 2 ? g : 0)) * ext[1] + (j + (p.ndim > 1 ? g : 0))) * ext[0] + i + g;
    for (int v = 0; v < p.nvar; v++) {
      double val;
      if (!(ss >> val)) { fprintf(stderr, "%s: cell %ld has fewer than %d variables\n", path, n, p.nvar); return 1; }
      if (mhd && v >= 5 && v <= 7) val /= SQRT4PI;
      P[v * plane + c] = val;
    }
    n++;
  }
  if (n != ncell) { fprintf(stderr, "%s: %ld cells, expected %ld\n", path, n, ncell); return 1; }
  return 0;
}

static int write_text_state(const std::string& path, const SimParamsGPU& p, const int ext[3], int g, const std::vector<double>& P) {
  std::ofstream outf(path.c_str());
  if (!outf.is_open()) { fprintf(stderr, "Error opening file %s for writing.\n", path.c_str()); return 1; }
  const size_t plane = (size_t)ext[0] * ext[1] * ext[2];
  const bool mhd = p.eqntype != PION_EQEUL;
  const double dx = (p.Xmax[0] - p.Xmin[0]) / p.NG[0];
  const long st[3] = {1, ext[0], (long)ext[0] * ext[1]};
  outf << "# format: x,[y,z,],rho,pg,vx,vy,vz,[Bx,By,Bz],[Tr0,Tr1,Tr2,..],T,[Tau0,Tau1,...]\n";
  outf << "# time = " << p.simtime << "  timestep = " << p.timestep << "\n";
  outf.setf(std::ios_base::scientific);
  outf.precision(14);
  for (int k = 0; k < p.NG[2]; k++)
    for (int j = 0; j < p.NG[1]; j++)
      for (int i = 0; i < p.NG[0]; i++) {
        const size_t c = ((size_t)(k + (p.ndim > 2 ? g : 0)) * ext[1] + (j + (p.ndim > 1 ? g : 0))) * ext[0] + i + g;
        if (i == 0) outf << "\n";  // "put in a blank line for gnuplot"
        outf << p.Xmin[0] + (i + 0.5) * dx << "  ";
        if (p.ndim > 1) outf << p.Xmin[1] + (j + 0.5) * dx << "  ";
        if (p.ndim > 2) outf << p.Xmin[2] + (k + 0.5) * dx << "\t";
        for (int v = 0; v < p.nvar; v++) outf << ((mhd && v >= 5 && v <= 7) ? P[v * plane + c] * SQRT4PI : P[v * plane + c]) << "  ";
        const double ro = P[c], pg = P[plane + c];
        if (p.EP.cooling) outf << pg * MU_TOT_OVER_KB / ro;
        else outf << pg / (p.gamma - 1.) / ro;
        if (mhd) {
          const double b2 = P[5 * plane + c] * P[5 * plane + c] + P[6 * plane + c] * P[6 * plane + c] + P[7 * plane + c] * P[7 * plane + c];
          outf << "  " << pg + b2 / 2.;
          double div = 0.0;
          for (int a = 0; a < p.ndim; a++) div += (P[(5 + a) * plane + c + st[a]] - P[(5 + a) * plane + c - st[a]]) / (2.0 * dx);
          outf << "  " << div;
        }
        outf << "\n";
      }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <paramfile> [--in P0.bin] [--out P.bin] [--steps N] [--device D] [--tables t.bin] [--verbose]\n", argv[0]);
    return 2;
  }
  const char *in = nullptr, *out = nullptr, *tables = nullptr, *in_text = nullptr, *out_text = nullptr;
  long steps = -1;
  int device = 0;
  bool verbose = false, convert = false;
  for (int i = 2; i < argc; i++) {
    if (!strcmp(argv[i], "--in") && i + 1 < argc) in = argv[++i];
    else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
    else if (!strcmp(argv[i], "--in-text") && i + 1 < argc) in_text = argv[++i];
    else if (!strcmp(argv[i], "--out-text") && i + 1 < argc) out_text = argv[++i];
    else if (!strcmp(argv[i], "--convert")) convert = true;
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
  if (p.EP.cooling && !convert) {
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
  const int gtxt = (p.spOOA == 2) ? 2 : 1;
  // dataio_text::set_filename (dataio_text.cpp:158-175): <base>.<counter, 8 digits>.txt, no counter for ICs
  auto text_name = [&](long counter) {
    char buf[32] = "";
    if (counter >= 0) snprintf(buf, sizeof buf, "%08ld.", counter);
    return std::string(out_text) + "." + buf + "txt";
  };
  if (in_text) {
    if (read_text_state(in_text, p, ext, gtxt, P, &p.simtime, &p.timestep)) return 1;
  } else if (in) {
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
  if (convert) {  // format translation only: no device, ghost cells as read (zero from a text file)
    if (out) {
      std::ofstream f(out, std::ios::binary);
      f.write(reinterpret_cast<const char*>(P.data()), P.size() * sizeof(double));
    }
    if (out_text && write_text_state(text_name(in_text ? -1 : p.timestep), p, ext, gtxt, P)) return 1;
    return 0;
  }
  if (sim.Init(device, P.data())) { fprintf(stderr, "Init: %s\n", sim.error().c_str()); return 1; }
  if (sim.Time_Int(steps, verbose)) { fprintf(stderr, "Time_Int: %s\n", sim.error().c_str()); return 1; }
  if (out || out_text) {
    if (sim.download_state(P.data())) { fprintf(stderr, "output_data: %s\n", sim.error().c_str()); return 1; }
    if (out) {
      std::ofstream f(out, std::ios::binary);
      f.write(reinterpret_cast<const char*>(P.data()), P.size() * sizeof(double));
    }
    if (out_text && write_text_state(text_name(sim.SimPM.timestep), sim.SimPM, ext, gtxt, P)) return 1;
  }
  printf("final time %.12e after %d steps\n", sim.SimPM.simtime, sim.SimPM.timestep);
  sim.Finalise();
  return 0;
}
