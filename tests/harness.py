"""Test harness (TEST INFRASTRUCTURE): one problem description, three backends.

* ``RefSim``    -- the UNMODIFIED reference translation units compiled into
                   oracle/_ref/libpion_ref.so (only present when /root/reference
                   was available at build time; it travels to the GPU box).
* ``OracleSim`` -- oracle/libpion_oracle.so, our plain-C restatement.
* ``GpuSim``    -- the product: libpion_b200.so through its C ABI (tests marked gpu).

All three expose the same methods and exchange state as float64 arrays of shape
[nvar, NZ+2g, NY+2g, NX+2g] (ghost cells included, unused dimensions have
extent 1), so a parity test is "run the same calls on two backends, compare".
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_LIB = ROOT / "oracle" / "_ref" / "libpion_ref.so"
ORACLE_LIB = ROOT / "oracle" / "libpion_oracle.so"

EQN_NAMES = {"euler": 1, "i-mhd": 2, "glm-mhd": 3}
EQN_NVAR = {"euler": 5, "i-mhd": 8, "glm-mhd": 9}
BC_CODES = {
    "periodic": 1, "outflow": 2, "inflow": 3, "reflecting": 4, "fixed": 5,
    "DMR": 8, "DMR2": 9, "one-way-outflow": 13, "stellar-wind": 14,
}
PO_MAXVAR = 16
COORD_CODES = {"cartesian": 1, "cylindrical": 2, "spherical": 3}


@dataclass
class Problem:
    """A PION problem description (mirrors the reference's parameter file keys)."""
    ndim: int = 2
    NG: tuple = (32, 32, 1)
    eqn: str = "glm-mhd"
    solver: int = 7
    artviscosity: int = 1
    etav: float = 0.15
    gamma: float = 5.0 / 3.0
    cfl: float = 0.3
    xmin: tuple = (0.0, 0.0, 0.0)
    xmax: tuple = (1.0, 1.0, 1.0)
    bcs: tuple = ("periodic",) * 6
    internal_bcs: tuple = ()
    ntracer: int = 0
    ooa: int = 2
    starttime: float = 0.0
    finishtime: float = 1.0e30
    op_criterion: int = 0
    opfreq_time: float = 0.0
    refvec: tuple = (1.0,) * PO_MAXVAR
    ics: str = "Uniform"
    extra: dict = field(default_factory=dict)  # extra parameter-file keys (IC parameters etc.)
    # microphysics (mp_only_cooling)
    cooling: int = 0
    mp_timestep_limit: int = 0
    min_temperature: float = 0.0
    max_temperature: float = 1.0e99
    # constant stellar-wind sources (internal boundary "stellar-wind"): dicts with keys
    # pos (3), radius, mdot [Msun/yr], vinf, vrot [km/s], temp [K], rstar [cm], bsrf [G], tr (tuple)
    winds: tuple = ()
    coords: str = "cartesian"  # "cartesian" | "cylindrical" (2-D z,R) | "spherical" (1-D r)

    @property
    def nvar(self):
        return EQN_NVAR[self.eqn] + self.ntracer

    def __post_init__(self):
        # "Force Nbc=1 if using Lax-Friedrichs flux" (setup_fixed_grid.cpp:188-190): first order in space and time
        if self.solver == 0:
            self.ooa = 1

    @property
    def nbc(self):
        return 2 if self.ooa == 2 else 1

    @property
    def dx(self):
        return (self.xmax[0] - self.xmin[0]) / self.NG[0]

    def padded_shape(self):
        g = self.nbc
        ext = [self.NG[a] + 2 * g if a < self.ndim else 1 for a in range(3)]
        return (self.nvar, ext[2], ext[1], ext[0])

    def interior(self):
        g = self.nbc
        sl = [slice(g, -g) if a < self.ndim else slice(None) for a in range(3)]
        return (slice(None), sl[2], sl[1], sl[0])

    def paramfile_text(self) -> str:
        f = repr  # repr() round-trips through the reference's atof()
        ax = "XYZ"
        L = [f"ndim {self.ndim}", f"eqn {self.eqn}", f"coordinates {self.coords}", f"solver {self.solver}",
             f"OrderOfAccSpace {self.ooa}", f"OrderOfAccTime {self.ooa}", f"ics {self.ics}",
             "OutputFile none", "OutputPath ./", "OutputFileType text",
             f"StartTime {f(self.starttime)}", f"FinishTime {f(self.finishtime)}",
             "OutputFrequency 1000000", f"OutputCriterion {self.op_criterion}", f"OPfreqTime {f(self.opfreq_time)}",
             "EP_dynamics 1", "EP_raytracing 0", "EP_phot_ionisation 0", f"EP_cooling {self.cooling}",
             "EP_chemistry 0", "EP_coll_ionisation 0", "EP_rad_recombination 0",
             f"EP_update_erg {1 if self.cooling else 0}", f"EP_MP_timestep_limit {self.mp_timestep_limit}",
             f"EP_Min_Temperature {f(self.min_temperature)}", f"EP_Max_Temperature {f(self.max_temperature)}",
             "EP_Hydrogen_MassFrac 1.0", "EP_Helium_MassFrac 0.0", "EP_Metal_MassFrac 0.0",
             f"ntracer {self.ntracer}", "chem_code none", "smooth -1", "noise -1.0"]
        for t in range(self.ntracer):
            L.append(f"Tracer{t:03d} colour{t}")
        for a in range(3):
            L.append(f"NGrid{ax[a]} {self.NG[a] if a < self.ndim else 1}")
            L.append(f"{ax[a]}min {f(float(self.xmin[a]))}")
            L.append(f"{ax[a]}max {f(float(self.xmax[a]))}")
        L += ["grid_nlevels 1"]
        for a in range(3):
            L += [f"grid_aspect_ratio_{ax[a]*2} 1", f"NG_centre_{ax[a]*2} 0.0", f"NG_refine_{ax[a]*2} 0"]
        names = ["XN", "XP", "YN", "YP", "ZN", "ZP"]
        for d in range(2 * self.ndim):
            L.append(f"BC_{names[d]} {self.bcs[d]}")
        L.append(f"BC_Ninternal {len(self.internal_bcs)}")
        for i, b in enumerate(self.internal_bcs):
            L.append(f"BC_INTERNAL_{i:03d} {b}")
        L += [f"GAMMA {f(self.gamma)}", f"CFL {f(self.cfl)}", f"ArtificialViscosity {self.artviscosity}",
              f"EtaViscosity {f(self.etav)}", "units SI", "rhoval 1.0", "lenval 1.0", "velval 1.0", "magval 1.0",
              "RT_Nsources 0", f"WIND_NSRC {len(self.winds)}", "N_JET 0"]
        for i, w in enumerate(self.winds):
            L += [f"WIND_{i}_pos{a} {f(float(w['pos'][a]))}" for a in range(3)]
            L += [f"WIND_{i}_radius {f(float(w['radius']))}", f"WIND_{i}_type 0", f"WIND_{i}_mdot {f(float(w['mdot']))}",
                  f"WIND_{i}_vinf {f(float(w['vinf']))}", f"WIND_{i}_vrot {f(float(w.get('vrot', 0.0)))}",
                  f"WIND_{i}_temp {f(float(w['temp']))}", f"WIND_{i}_Rstr {f(float(w['rstar']))}",
                  f"WIND_{i}_Bsrf {f(float(w.get('bsrf', 0.0)))}", f"WIND_{i}_evofile NOFILE", f"WIND_{i}_t_offset 0.0",
                  f"WIND_{i}_t_scalefac 1.0", f"WIND_{i}_updatefreq 1.0", f"WIND_{i}_enhance_mdot 0", f"WIND_{i}_xi 0.0",
                  f"WIND_{i}_ecentricity_fac 0.0", f"WIND_{i}_orbital_period 0.0", f"WIND_{i}_periastron_vec_x 0.0",
                  f"WIND_{i}_periastron_vec_y 0.0"]
            L += [f"WIND_{i}_TR{t} {f(float(w.get('tr', (1.0,) * 4)[t]))}" for t in range(self.ntracer)]
        for v in range(PO_MAXVAR):
            L.append(f"refvec{v} {f(float(self.refvec[v]))}")
        for k, v in self.extra.items():
            L.append(f"{k} {v if isinstance(v, str) else f(v) if isinstance(v, float) else v}")
        return "\n".join(L) + "\n"


# --------------------------------------------------------------------------
class _CSim:
    """Shared ctypes plumbing: both oracle libraries export the same calls with
    prefix `pref_` / `po_`."""
    prefix = ""
    lib = None

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _bind(self):
        vp, d, i = C.c_void_p, C.c_double, C.c_int
        sig = {
            "info": (i, [vp, C.POINTER(i), C.POINTER(d)]),
            "get_state": (i, [vp, i, vp]), "set_state": (i, [vp, i, vp]),
            "get_flags": (i, [vp, vp]), "get_extra": (i, [vp, i, i, vp]),
            "init_after_state": (i, [vp]), "calc_timestep": (d, [vp]), "advance": (d, [vp]),
            "dynamics_dt": (d, [vp]), "microphysics_dt": (d, [vp]),
            "run": (i, [vp, i, vp]), "update_bcs": (i, [vp, i, i]),
            "dynamics_dU": (i, [vp, d, i]), "microphysics_dU": (i, [vp, d]),
            "update_state": (i, [vp, d, i, i]), "set_dt": (None, [vp, d]),
            "set_glm_speeds": (None, [vp, d, d, d]), "set_time": (None, [vp, d, d, i]),
            "destroy": (None, [vp]),
        }
        for n, (r, a) in sig.items():
            fn = self._f(n)
            fn.restype, fn.argtypes = r, a

    def info(self):
        ii = (C.c_int * 24)()
        dd = (C.c_double * 24)()
        self._f("info")(self.h, ii, dd)
        return list(ii), list(dd)

    def shape(self):
        ii, _ = self.info()
        return (ii[6], ii[2], ii[1], ii[0])

    def get_state(self, which=0):
        out = np.zeros(self.shape())
        self._f("get_state")(self.h, which, out.ctypes.data)
        return out

    def set_state(self, arr, which=0):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == self.shape(), (a.shape, self.shape())
        self._f("set_state")(self.h, which, a.ctypes.data)

    def get_flags(self):
        s = self.shape()
        out = np.zeros(s[1:], dtype=np.int32)
        self._f("get_flags")(self.h, out.ctypes.data)
        return out

    def get_extra(self, what, axis=0):
        s = self.shape()
        out = np.zeros(s[1:])
        self._f("get_extra")(self.h, what, axis, out.ctypes.data)
        return out

    def init_after_state(self):
        return self._f("init_after_state")(self.h)

    def calc_timestep(self):
        return self._f("calc_timestep")(self.h)

    def advance(self):
        return self._f("advance")(self.h)

    def dynamics_dt(self):
        return self._f("dynamics_dt")(self.h)

    def microphysics_dt(self):
        return self._f("microphysics_dt")(self.h)

    def run(self, n):
        dts = np.zeros(n)
        got = self._f("run")(self.h, n, dts.ctypes.data)
        assert got == n, f"only {got} of {n} steps taken"
        return dts

    def update_bcs(self, cstep, maxstep):
        return self._f("update_bcs")(self.h, cstep, maxstep)

    def dynamics_dU(self, dt, step):
        return self._f("dynamics_dU")(self.h, dt, step)

    def microphysics_dU(self, dt):
        return self._f("microphysics_dU")(self.h, dt)

    def update_state(self, dt, step, ooa):
        return self._f("update_state")(self.h, dt, step, ooa)

    def set_dt(self, dt):
        self._f("set_dt")(self.h, dt)

    def set_glm_speeds(self, tdyn, dx, cr):
        self._f("set_glm_speeds")(self.h, tdyn, dx, cr)

    def set_time(self, simtime, last_dt, timestep):
        self._f("set_time")(self.h, simtime, last_dt, timestep)

    def close(self):
        if getattr(self, "h", None):
            self._f("destroy")(self.h)
            self.h = None


_ref_lib = None
TABLE_KEYS = ["T", "rrhp", "C_rrh", "C_ffhe", "C_fbdn", "C_cie"]
COOLING_TABLES = ROOT / "tests" / "golden" / "cooling_tables_wss09_T5e3_1e8.npz"


COOLING_SPLINES = ROOT / "tests" / "golden" / "cooling_splines_sd93_wss09.npz"


def load_cooling_spline(flag):
    """Committed fixture: knots and end slopes of the reference's cooling-curve spline for EP_cooling 4, 5
    (setup_SD93_cie) or 6, 7 (setup_WSS09_CIE), read out of the reference's MP object by
    tests/golden/make_golden.py (RefSim.cooling_spline)."""
    z = np.load(COOLING_SPLINES)
    pre = "sd93" if flag in (4, 5) else "wss09"
    return {"spline_logT": z[pre + "_logT"], "spline_logL": z[pre + "_logL"], "spline_slopes": z[pre + "_slopes"]}


def tables_for(prob):
    """The committed cooling fixture a problem's EP_cooling flag needs (None without tabulated data)."""
    if prob.cooling == 8:
        return load_cooling_tables()
    if prob.cooling in (4, 5, 6, 7):
        return load_cooling_spline(prob.cooling)
    return None


def load_cooling_tables():
    """Committed fixture: the reference's EP_cooling=8 lookup tables for T in [5e3, 1e8] K
    (generated by tests/golden/make_golden.py from oracle/_ref)."""
    z = np.load(COOLING_TABLES)
    return {k: z[k] for k in TABLE_KEYS}


def have_ref():
    return REF_LIB.exists()


class RefSim(_CSim):
    """The compiled, unmodified reference (oracle/_ref)."""
    prefix = "pref_"

    def __init__(self, prob: Problem, run_ics: bool = False):
        global _ref_lib
        if _ref_lib is None:
            _ref_lib = C.CDLL(str(REF_LIB))
        self.lib = _ref_lib
        self._bind()
        self.lib.pref_create.restype = C.c_void_p
        self.lib.pref_create.argtypes = [C.c_char_p, C.c_int]
        self.lib.pref_intercell_flux.restype = C.c_int
        self.lib.pref_intercell_flux.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_double] * 4 + [C.c_void_p]
        self.prob = prob
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(prob.paramfile_text())
            path = f.name
        try:
            self.h = self.lib.pref_create(path.encode(), 1 if run_ics else 0)
        finally:
            os.unlink(path)
        assert self.h, "reference set-up failed"

    def cooling_tables(self, n=200):
        """mp_only_cooling's 200-point lookup tables, re-evaluated through the reference's own
        rate functions (pref_cooling_tables in oracle/ref_driver.cpp)."""
        arrs = {k: np.zeros(n) for k in TABLE_KEYS}
        self.lib.pref_cooling_tables.restype = C.c_int
        self.lib.pref_cooling_tables.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        err = self.lib.pref_cooling_tables(self.h, n, *[arrs[k].ctypes.data for k in TABLE_KEYS])
        assert err == 0, "reference has no mp_only_cooling object"
        return arrs

    def output_text(self, base, counter=-1):
        """The reference's ASCII writer (dataio_text::OutputData) on the current state: <base>.txt or
        <base>.<counter:08d>.txt."""
        self.lib.pref_output_text.restype = C.c_int
        self.lib.pref_output_text.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        err = self.lib.pref_output_text(self.h, str(base).encode(), counter)
        assert err == 0, "dataio_text::OutputData failed"

    def cooling_spline(self):
        """Knots and out-of-table slopes of the reference MP object's cooling-curve spline
        (pref_cooling_spline in oracle/ref_driver.cpp)."""
        self.lib.pref_cooling_spline.restype = C.c_int
        self.lib.pref_cooling_spline.argtypes = [C.c_void_p] * 4
        n = self.lib.pref_cooling_spline(self.h, None, None, None)
        assert n > 2, "reference has no cooling_function_SD93CIE object"
        x, y, sl = np.zeros(n), np.zeros(n), np.zeros(2)
        self.lib.pref_cooling_spline(self.h, x.ctypes.data, y.ctypes.data, sl.ctypes.data)
        return {"spline_logT": x, "spline_logL": y, "spline_slopes": sl}

    def intercell_flux(self, axis, Pl, Pr, divv_l=0.0, gradp_l=0.0, divv_r=0.0, gradp_r=0.0):
        Pl = np.ascontiguousarray(Pl, dtype=np.float64)
        Pr = np.ascontiguousarray(Pr, dtype=np.float64)
        out = np.zeros(self.prob.nvar)
        self.lib.pref_intercell_flux(self.h, axis, Pl.ctypes.data, Pr.ctypes.data, divv_l, gradp_l, divv_r, gradp_r,
                                     out.ctypes.data)
        return out


class WindSource(C.Structure):
    """struct po_wind_source == struct pion_gpu_wind_source."""
    _fields_ = [("dpos", C.c_double * 3), ("radius", C.c_double), ("mdot", C.c_double), ("vinf", C.c_double),
                ("vrot", C.c_double), ("temp", C.c_double), ("rstar", C.c_double), ("bsrf", C.c_double),
                ("tr", C.c_double * 4)]


class _OracleConfig(C.Structure):
    _fields_ = [
        ("ndim", C.c_int), ("NG", C.c_int * 3), ("nvar", C.c_int), ("ntracer", C.c_int), ("eqntype", C.c_int),
        ("coord_sys", C.c_int), ("solver", C.c_int), ("artviscosity", C.c_int), ("spOOA", C.c_int), ("tmOOA", C.c_int),
        ("gamma", C.c_double), ("cfl", C.c_double), ("etav", C.c_double),
        ("xmin", C.c_double * 3), ("xmax", C.c_double * 3),
        ("bc", C.c_int * 6), ("n_internal_bc", C.c_int), ("internal_bc", C.c_int * 4),
        ("refvec", C.c_double * PO_MAXVAR),
        ("starttime", C.c_double), ("finishtime", C.c_double),
        ("op_criterion", C.c_int), ("opfreq_time", C.c_double),
        ("cooling", C.c_int), ("mp_timestep_limit", C.c_int),
        ("min_temperature", C.c_double), ("max_temperature", C.c_double),
        ("n_table", C.c_int),
        ("table_T", C.c_void_p), ("table_rrhp", C.c_void_p), ("table_C_rrh", C.c_void_p),
        ("table_C_ffhe", C.c_void_p), ("table_C_fbdn", C.c_void_p), ("table_C_cie", C.c_void_p),
        ("n_wind", C.c_int), ("wind", WindSource * 2),
        ("n_spline", C.c_int), ("spline_logT", C.c_void_p), ("spline_logL", C.c_void_p),
        ("spline_min_slope", C.c_double), ("spline_max_slope", C.c_double),
    ]


def fill_spline(c, tables, keep):
    """tables["spline_logT" / "spline_logL" / "spline_slopes"] -> the n_spline / spline_* members shared by the
    oracle and the GPU config (EP_cooling 4..7)."""
    if tables is None or "spline_logT" not in tables:
        return
    x = np.ascontiguousarray(tables["spline_logT"], dtype=np.float64)
    y = np.ascontiguousarray(tables["spline_logL"], dtype=np.float64)
    keep += [x, y]
    c.n_spline = len(x)
    c.spline_logT, c.spline_logL = x.ctypes.data, y.ctypes.data
    c.spline_min_slope, c.spline_max_slope = float(tables["spline_slopes"][0]), float(tables["spline_slopes"][1])


def fill_winds(c, prob):
    """Problem.winds -> the n_wind / wind[] members shared by the oracle and the GPU config."""
    c.n_wind = len(prob.winds)
    for i, w in enumerate(prob.winds):
        ws = c.wind[i]
        for a in range(3):
            ws.dpos[a] = w["pos"][a]
        ws.radius, ws.mdot, ws.vinf, ws.vrot = w["radius"], w["mdot"], w["vinf"], w.get("vrot", 0.0)
        ws.temp, ws.rstar, ws.bsrf = w["temp"], w["rstar"], w.get("bsrf", 0.0)
        for t in range(4):
            ws.tr[t] = w.get("tr", (1.0,) * 4)[t]


def effective_etav(prob: Problem):
    # ics/get_sim_info.cpp:452-468: etav is 0 for AV=0 and forced to 0.1 for AV=3
    if prob.artviscosity == 0:
        return 0.0
    if prob.artviscosity == 3:
        return 0.1
    return prob.etav


def oracle_config(prob: Problem, tables=None):
    c = _OracleConfig()
    c.ndim = prob.ndim
    for a in range(3):
        c.NG[a] = prob.NG[a] if a < prob.ndim else 1
        c.xmin[a] = prob.xmin[a]
        c.xmax[a] = prob.xmax[a]
    c.nvar, c.ntracer, c.eqntype = prob.nvar, prob.ntracer, EQN_NAMES[prob.eqn]
    c.coord_sys, c.solver, c.artviscosity = COORD_CODES[prob.coords], prob.solver, prob.artviscosity
    c.spOOA = c.tmOOA = prob.ooa
    c.gamma, c.cfl, c.etav = prob.gamma, prob.cfl, effective_etav(prob)
    for d in range(6):
        c.bc[d] = BC_CODES[prob.bcs[d]] if d < 2 * prob.ndim else 0
    c.n_internal_bc = len(prob.internal_bcs)
    for i, b in enumerate(prob.internal_bcs):
        c.internal_bc[i] = BC_CODES[b]
    for v in range(PO_MAXVAR):
        c.refvec[v] = prob.refvec[v]
    c.starttime, c.finishtime = prob.starttime, prob.finishtime
    c.op_criterion, c.opfreq_time = prob.op_criterion, prob.opfreq_time
    c.cooling, c.mp_timestep_limit = prob.cooling, prob.mp_timestep_limit
    c.min_temperature, c.max_temperature = prob.min_temperature, prob.max_temperature
    fill_winds(c, prob)
    keep = []
    if tables is not None and "T" in tables:
        c.n_table = len(tables["T"])
        for name, key in [("table_T", "T"), ("table_rrhp", "rrhp"), ("table_C_rrh", "C_rrh"),
                          ("table_C_ffhe", "C_ffhe"), ("table_C_fbdn", "C_fbdn"), ("table_C_cie", "C_cie")]:
            arr = np.ascontiguousarray(tables[key], dtype=np.float64)
            keep.append(arr)
            setattr(c, name, arr.ctypes.data)
    fill_spline(c, tables, keep)
    return c, keep


_oracle_lib = None


class OracleSim(_CSim):
    """oracle/libpion_oracle.so -- the plain-C restatement."""
    prefix = "po_"

    def __init__(self, prob: Problem, tables=None):
        global _oracle_lib
        if _oracle_lib is None:
            _oracle_lib = C.CDLL(str(ORACLE_LIB))
        self.lib = _oracle_lib
        self._bind()
        self.lib.po_create.restype = C.c_void_p
        self.lib.po_create.argtypes = [C.c_void_p]
        self.lib.po_intercell_flux.restype = C.c_int
        self.lib.po_intercell_flux.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_double] * 5 + [C.c_void_p]
        self.lib.po_error_counts.argtypes = [C.c_void_p, C.c_void_p]
        self.prob = prob
        cfg, self._keep = oracle_config(prob, tables)
        self.h = self.lib.po_create(C.addressof(cfg))
        assert self.h

    def intercell_flux(self, axis, Pl, Pr, divv_l=0.0, gradp_l=0.0, divv_r=0.0, gradp_r=0.0, hc_etamax=0.0):
        Pl = np.ascontiguousarray(Pl, dtype=np.float64)
        Pr = np.ascontiguousarray(Pr, dtype=np.float64)
        out = np.zeros(self.prob.nvar)
        self.lib.po_intercell_flux(self.h, axis, Pl.ctypes.data, Pr.ctypes.data, divv_l, gradp_l, divv_r, gradp_r,
                                   hc_etamax, out.ctypes.data)
        return out

    def error_counts(self):
        out = (C.c_long * 2)()
        self.lib.po_error_counts(self.h, out)
        return list(out)


# --------------------------------------------------------------------------
def splitmix64(seed, n):
    """Deterministic uniform [0,1) stream (seed 12345 branch-coverage state)."""
    x = np.uint64(seed)
    out = np.empty(n)
    with np.errstate(over="ignore"):
        for i in range(n):
            x = x + np.uint64(0x9E3779B97F4A7C15)
            z = x
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[i] = float(z >> np.uint64(11)) * (1.0 / 9007199254740992.0)
    return out


def random_state(prob: Problem, seed=12345, smooth=True, amp=0.5):
    """Seeded synthetic primitive state on the padded grid: rho,p in [0.5,1.5],
    v,B in [-amp,amp], psi=0, tracers in [0,1].  `smooth` low-passes it so the
    state has resolved gradients (hits every HLLD fan region without NaNs)."""
    shp = prob.padded_shape()
    rng = np.random.Generator(np.random.PCG64(seed))
    P = rng.random(shp)
    if smooth:
        for ax in range(1, 4):
            if shp[ax] > 1:
                for _ in range(2):
                    P = (np.roll(P, 1, ax) + 2 * P + np.roll(P, -1, ax)) / 4
        P = (P - P.min()) / (P.max() - P.min())
    out = np.zeros(shp)
    nv_phys = EQN_NVAR[prob.eqn]
    out[0] = 0.5 + P[0]
    out[1] = 0.5 + P[1]
    for v in range(2, min(nv_phys, 8)):
        out[v] = amp * (2 * P[v] - 1)
    for v in range(nv_phys, prob.nvar):
        out[v] = P[v]
    return out


def hot_sphere_state(prob: Problem, seed=7, amp=0.5, fac=100.0):
    """random_state with the pressure multiplied by `fac` inside a central ellipsoid (30 % of each extent):
    |grad p|/p >> 5 and converging/diverging flow at its surface, so that a few thousand cells trip the
    HLLD -> HLL switch (solver_eqn_mhd_adi.cpp:167-177) -- the smooth random state never does."""
    P = random_state(prob, seed, amp=amp)
    shp = P.shape[1:]
    z, y, x = np.meshgrid(*[np.arange(n) - (n - 1) / 2 for n in shp], indexing="ij")
    r2 = sum((c / (0.3 * n)) ** 2 for c, n in zip((z, y, x), shp) if n > 1)
    P[1] = np.where(r2 < 1.0, P[1] * fac, P[1])
    return P


MU_TOT_OVER_KB = (0.609 * 1.672621898e-24) / 1.38064852e-16  # mp_only_cooling.cpp:79-81 with constants.h:53,64


def cooling_state(prob: Problem, seed=4242, rho0=2.0e-24, Tlo=6.0e3, Thi=5.0e7):
    """Seeded cgs state for the cooling tests: rho ~ 2e-24 g/cm3 x [0.3,3], T log-uniform in
    [6e3, 5e7] K (both sides of the cooling-curve peak), |v| <~ 3e6 cm/s, weak B, tracers in
    [0,1.1] (exercises the sCMA clamp).  rho0 = 2e-22 makes the cooling time comparable to the
    CFL step, so the adaptive integrator bisects and sub-steps."""
    P = random_state(prob, seed)
    shp = prob.padded_shape()
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    nv_phys = EQN_NVAR[prob.eqn]
    rho = rho0 * 10.0 ** (P[0] - 1.0)
    T = 10.0 ** (np.log10(Tlo) + (np.log10(Thi) - np.log10(Tlo)) * np.clip((P[1] - 0.5) * 1.6 - 0.3 + 0.4 * rng.random(shp[1:]), 0, 1))
    out = np.zeros(shp)
    out[0] = rho
    out[1] = rho * T / MU_TOT_OVER_KB
    out[2:5] = 6.0e6 * P[2:5]
    if nv_phys >= 8:
        out[5:8] = 2.0e-6 * P[5:8]
    for v in range(nv_phys, prob.nvar):
        out[v] = 1.1 * P[v]
    return out


def state_scales(b, gamma=5.0 / 3.0, nphys=None):
    """Per-variable normalisation for primitive-state comparisons: density and pressure by
    their maxima; every velocity component by max(|v|, sound speed); every B component and
    psi by max |B| (vector components that are identically zero in the problem, e.g. v_z in
    a 2-D run, only carry rounding noise and must not be normalised by themselves)."""
    b = np.asarray(b)
    nv = b.shape[0]
    if nphys is None:  # without the equation set, assume no tracers beyond the 9 GLM variables
        nphys = 9 if nv >= 9 else (8 if nv >= 8 else 5)
    sc = np.ones(nv)
    sc[0] = np.max(np.abs(b[0]))
    sc[1] = np.max(np.abs(b[1]))
    cs = float(np.sqrt(gamma * np.max(np.abs(b[1])) / max(np.min(np.abs(b[0])), 1e-300)))
    sc[2:5] = max(float(np.max(np.abs(b[2:5]))), cs)
    if nphys >= 8:
        bmax = float(np.max(np.abs(b[5:8])))
        sc[5:8] = bmax if bmax > 0 else 1.0
        if nphys >= 9:
            sc[8] = sc[5]
    for v in range(nphys, nv):
        m = float(np.max(np.abs(b[v])))
        sc[v] = m if m > 0 else 1.0
    return sc


def rel_err(a, b, scale=None, primitive=True, nphys=None):
    """max |a-b| / scale per variable (scale: state_scales(b) for primitive states)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if scale is None:
        scale = state_scales(b, nphys=nphys) if primitive and b.shape[0] >= 5 else [np.max(np.abs(b[v])) for v in range(b.shape[0])]
    errs = []
    for v in range(a.shape[0]):
        s = scale[v] if scale[v] > 0 else 1.0
        errs.append(float(np.max(np.abs(a[v] - b[v])) / s))
    return np.array(errs)


def ulp_response(prob: Problem, P0, nsteps, seeds=(1, 2, 3), sim=None):
    """How far the REFERENCE algorithm's own answer moves when its input moves by one unit in the last place:
    max over `seeds` of rel_err(run(P0 with rho, p, B_x each multiplied by 1 + {-1,0,1} x 1.1e-16), run(P0)) per
    variable, with the CPU checker (`sim`, default OracleSim -- bit-exact with the compiled reference).  Discrete
    decisions in the scheme (HLLD fan region, HLLD -> HLL switch, minmod sign tests) make some states --
    exact symmetries, exact zeros, strong discontinuities -- respond to rounding noise at 1e-11..1e-9 within a few
    steps; a GPU/CPU difference below that level is not a defect of either."""
    sim = sim or OracleSim

    def run(P):
        o = sim(prob)
        o.set_state(P)
        o.init_after_state()
        o.run(nsteps)
        R = o.get_state(0)
        o.close()
        return R

    base = run(P0)
    nphys = EQN_NVAR[prob.eqn]
    resp = np.zeros(prob.nvar)
    for seed in seeds:
        rng = np.random.default_rng(seed)
        P1 = np.array(P0, copy=True)
        for v in (0, 1, 5) if nphys >= 8 else (0, 1):
            P1[v] *= 1.0 + rng.integers(-1, 2, size=P1[v].shape) * 1.1e-16
        resp = np.maximum(resp, rel_err(run(P1), base, nphys=nphys))
    return resp, base


# --------------------------------------------------------------------------
def gpu_config(prob: Problem, device=0, tables=None):
    """Problem -> struct pion_gpu_config of the product library."""
    import sys
    sys.path.insert(0, str(ROOT))
    from pion_b200.capi import GpuConfig
    c = GpuConfig()
    c.device, c.ndim = device, prob.ndim
    for a in range(3):
        c.NG[a] = prob.NG[a] if a < prob.ndim else 1
        c.xmin[a] = prob.xmin[a]
        c.xmax[a] = prob.xmax[a]
        c.sim_xmin[a] = prob.xmin[a]
    c.nvar, c.ntracer, c.eqntype = prob.nvar, prob.ntracer, EQN_NAMES[prob.eqn]
    c.coord_sys, c.solver, c.artviscosity = COORD_CODES[prob.coords], prob.solver, prob.artviscosity
    c.spOOA = c.tmOOA = prob.ooa
    c.gamma, c.cfl, c.etav = prob.gamma, prob.cfl, prob.etav
    for d in range(6):
        c.bc[d] = BC_CODES[prob.bcs[d]] if d < 2 * prob.ndim else 0
        c.ngbprocs[d] = -1
    c.n_internal_bc = len(prob.internal_bcs)
    for i, b in enumerate(prob.internal_bcs):
        c.internal_bc[i] = BC_CODES[b]
    for v in range(PO_MAXVAR):
        c.refvec[v] = prob.refvec[v]
    c.starttime, c.finishtime = prob.starttime, prob.finishtime
    c.op_criterion, c.opfreq_time = prob.op_criterion, prob.opfreq_time
    c.cooling, c.mp_timestep_limit = prob.cooling, prob.mp_timestep_limit
    c.min_temperature, c.max_temperature = prob.min_temperature, prob.max_temperature
    c.rank, c.nproc = 0, 1
    fill_winds(c, prob)
    keep = []
    if tables is not None and "T" in tables:
        c.n_table = len(tables["T"])
        for name, key in [("table_T", "T"), ("table_rrhp", "rrhp"), ("table_C_rrh", "C_rrh"),
                          ("table_C_ffhe", "C_ffhe"), ("table_C_fbdn", "C_fbdn"), ("table_C_cie", "C_cie")]:
            arr = np.ascontiguousarray(tables[key], dtype=np.float64)
            keep.append(arr)
            setattr(c, name, arr.ctypes.data)
    fill_spline(c, tables, keep)
    return c, keep


class GpuSim:
    """The product (libpion_b200.so through its C ABI) behind the same method
    names as RefSim / OracleSim."""

    def __init__(self, prob: Problem, device=0, tables=None):
        import sys
        sys.path.insert(0, str(ROOT))
        from pion_b200.capi import Context
        self.prob = prob
        cfg, keep = gpu_config(prob, device, tables)
        self.ctx = Context(cfg, keep)

    def shape(self):
        return self.ctx.shape

    def get_state(self, which=0):
        return self.ctx.download(which)

    def set_state(self, arr, which=0):
        self.ctx.upload(arr, which)

    def init_after_state(self):
        self.ctx.init_after_upload()
        return 0

    def calc_timestep(self):
        return self.ctx.calculate_timestep()

    def advance(self):
        return self.ctx.advance_time()

    def dynamics_dt(self):
        return self.ctx.calc_dt()[0]

    def microphysics_dt(self):
        return self.ctx.calc_dt()[1]

    def run(self, n):
        return self.ctx.run(n)

    def update_bcs(self, cstep, maxstep):
        self.ctx.time_update_bcs(self.ctx.get_time()[0], cstep, maxstep)
        return 0

    def dynamics_dU(self, dt, step):
        self.ctx.calc_dynamics_dU(dt, step)
        return 0

    def microphysics_dU(self, dt):
        self.ctx.calc_microphysics_dU(dt)
        return 0

    def update_state(self, dt, step, ooa):
        self.ctx.grid_update_state_vector(dt, step, ooa)
        return 0

    def set_dt(self, dt):
        self.ctx.set_dt(dt)

    def set_glm_speeds(self, tdyn, dx, cr):
        self.ctx.set_glm_speeds(tdyn, dx, cr)

    def set_time(self, simtime, last_dt, timestep):
        self.ctx.set_time(simtime, last_dt, timestep)

    def error_counts(self):
        return self.ctx.counters()[:2]

    def close(self):
        self.ctx.close()
