"""ctypes binding of include/pion_b200.h (no numerics here)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libpion_b200.so"
MAXVAR = 16

_lib = None


class WindSource(C.Structure):
    """struct pion_gpu_wind_source (include/pion_b200.h)."""
    _fields_ = [("dpos", C.c_double * 3), ("radius", C.c_double), ("mdot", C.c_double), ("vinf", C.c_double),
                ("vrot", C.c_double), ("temp", C.c_double), ("rstar", C.c_double), ("bsrf", C.c_double),
                ("tr", C.c_double * 4)]


class GpuConfig(C.Structure):
    """struct pion_gpu_config (include/pion_b200.h)."""
    _fields_ = [
        ("device", C.c_int), ("ndim", C.c_int), ("NG", C.c_int * 3), ("nvar", C.c_int), ("ntracer", C.c_int),
        ("eqntype", C.c_int), ("coord_sys", C.c_int), ("solver", C.c_int), ("artviscosity", C.c_int),
        ("spOOA", C.c_int), ("tmOOA", C.c_int),
        ("gamma", C.c_double), ("cfl", C.c_double), ("etav", C.c_double),
        ("xmin", C.c_double * 3), ("xmax", C.c_double * 3), ("sim_xmin", C.c_double * 3),
        ("bc", C.c_int * 6), ("n_internal_bc", C.c_int), ("internal_bc", C.c_int * 4),
        ("refvec", C.c_double * MAXVAR),
        ("starttime", C.c_double), ("finishtime", C.c_double),
        ("op_criterion", C.c_int), ("opfreq_time", C.c_double),
        ("cooling", C.c_int), ("mp_timestep_limit", C.c_int),
        ("min_temperature", C.c_double), ("max_temperature", C.c_double),
        ("n_table", C.c_int),
        ("table_T", C.c_void_p), ("table_rrhp", C.c_void_p), ("table_C_rrh", C.c_void_p),
        ("table_C_ffhe", C.c_void_p), ("table_C_fbdn", C.c_void_p), ("table_C_cie", C.c_void_p),
        ("rank", C.c_int), ("nproc", C.c_int), ("ngbprocs", C.c_int * 6),
        ("n_wind", C.c_int), ("wind", WindSource * 2),
        ("min_timestep", C.c_double),
        ("n_spline", C.c_int), ("spline_logT", C.c_void_p), ("spline_logL", C.c_void_p),
        ("spline_min_slope", C.c_double), ("spline_max_slope", C.c_double),
    ]


def load_library(path=None):
    """Load libpion_b200.so; there is no fallback if it is missing.  `path` (first call only) selects another
    build of the SAME library -- kernel tuning experiments pass it explicitly (bench.py --lib); there is no
    environment switch."""
    global _lib, LIB_PATH
    if _lib is not None:
        if path is not None and Path(path).resolve() != LIB_PATH.resolve():
            raise RuntimeError(f"library already loaded from {LIB_PATH}")
        return _lib
    if path is not None:
        LIB_PATH = Path(path)
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(pion_b200 has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "pion_gpu_create": (vp, [C.POINTER(GpuConfig)]),
        "pion_gpu_destroy": (None, [vp]),
        "pion_gpu_last_error": (C.c_char_p, []),
        "pion_gpu_upload": (i, [vp, i, vp]),
        "pion_gpu_download": (i, [vp, i, vp]),
        "pion_gpu_init_after_upload": (i, [vp]),
        "pion_gpu_calc_dt": (i, [vp, pd, pd]),
        "pion_gpu_calculate_timestep": (i, [vp, pd]),
        "pion_gpu_set_dt": (i, [vp, d]),
        "pion_gpu_set_glm_speeds": (i, [vp, d, d, d]),
        "pion_gpu_set_time": (i, [vp, d, d, i]),
        "pion_gpu_get_time": (i, [vp, pd, pd, pd, pi]),
        "pion_gpu_calc_microphysics_dU": (i, [vp, d]),
        "pion_gpu_calc_dynamics_dU": (i, [vp, d, i]),
        "pion_gpu_grid_update_state_vector": (i, [vp, d, i, i]),
        "pion_gpu_time_update_bcs": (i, [vp, d, i, i]),
        "pion_gpu_time_update_internal_bcs": (i, [vp, d, i, i]),
        "pion_gpu_time_update_external_bcs": (i, [vp, d, i, i]),
        "pion_gpu_advance_time": (i, [vp, pd]),
        "pion_gpu_run": (i, [vp, i, vp]),
        "pion_gpu_output_due": (i, [vp, i, pi]),
        "pion_gpu_counters": (i, [vp, vp]),
        "pion_gpu_mp_failures": (i, [vp, C.POINTER(C.c_longlong)]),
        "pion_gpu_riemann_failures": (i, [vp, C.POINTER(C.c_longlong)]),
        "pion_gpu_sync": (i, [vp]),
        "pion_gpu_stream": (vp, [vp]),
        "pion_gpu_stage_timing": (i, [vp, i, pd, C.POINTER(C.c_longlong)]),
        "pion_gpu_describe": (i, [vp, vp, i]),
        "pion_gpu_nccl_unique_id": (i, [vp]),
        "pion_gpu_nccl_init": (i, [vp, vp]),
        "pion_gpu_decompose_domain": (i, [C.POINTER(GpuConfig), i, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "pion_gpu_create", "pion_gpu_destroy", "pion_gpu_last_error", "pion_gpu_upload", "pion_gpu_download",
    "pion_gpu_init_after_upload", "pion_gpu_calc_dt", "pion_gpu_calculate_timestep", "pion_gpu_set_dt",
    "pion_gpu_set_glm_speeds", "pion_gpu_set_time", "pion_gpu_get_time", "pion_gpu_calc_microphysics_dU",
    "pion_gpu_calc_dynamics_dU", "pion_gpu_grid_update_state_vector", "pion_gpu_time_update_bcs",
    "pion_gpu_time_update_internal_bcs", "pion_gpu_time_update_external_bcs",
    "pion_gpu_advance_time", "pion_gpu_run", "pion_gpu_output_due", "pion_gpu_counters", "pion_gpu_mp_failures", "pion_gpu_riemann_failures", "pion_gpu_sync", "pion_gpu_stream",
    "pion_gpu_nccl_unique_id", "pion_gpu_nccl_init", "pion_gpu_decompose_domain", "pion_gpu_stage_timing",
    "pion_gpu_describe",
]


class Context:
    """RAII wrapper of pion_gpu_ctx*.  Every method is one C-ABI call."""

    def __init__(self, cfg: GpuConfig, keepalive=()):
        self.lib = load_library()
        self.cfg = cfg
        self._keep = keepalive
        self.h = self.lib.pion_gpu_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("pion_gpu_create: " + self.lib.pion_gpu_last_error().decode())
        g = 2 if cfg.spOOA == 2 else 1
        ext = [cfg.NG[a] + 2 * g if a < cfg.ndim else 1 for a in range(3)]
        self.shape = (cfg.nvar, ext[2], ext[1], ext[0])

    def _ck(self, err, what):
        if err:
            raise RuntimeError(f"{what}: " + self.lib.pion_gpu_last_error().decode())

    def upload(self, arr, which=0):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == self.shape, (a.shape, self.shape)
        self._ck(self.lib.pion_gpu_upload(self.h, which, a.ctypes.data), "upload")

    def download(self, which=0):
        out = np.empty(self.shape)
        self._ck(self.lib.pion_gpu_download(self.h, which, out.ctypes.data), "download")
        return out

    def upload_ptr(self, ptr, which=0):
        self._ck(self.lib.pion_gpu_upload(self.h, which, ptr), "upload")

    def download_ptr(self, ptr, which=0):
        self._ck(self.lib.pion_gpu_download(self.h, which, ptr), "download")

    def init_after_upload(self):
        self._ck(self.lib.pion_gpu_init_after_upload(self.h), "init_after_upload")

    def calc_dt(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.pion_gpu_calc_dt(self.h, C.byref(a), C.byref(b)), "calc_dt")
        return a.value, b.value

    def calculate_timestep(self):
        a = C.c_double()
        self._ck(self.lib.pion_gpu_calculate_timestep(self.h, C.byref(a)), "calculate_timestep")
        return a.value

    def set_dt(self, dt):
        self._ck(self.lib.pion_gpu_set_dt(self.h, dt), "set_dt")

    def set_glm_speeds(self, tdyn, dx, cr):
        self._ck(self.lib.pion_gpu_set_glm_speeds(self.h, tdyn, dx, cr), "set_glm_speeds")

    def set_time(self, simtime, last_dt, timestep):
        self._ck(self.lib.pion_gpu_set_time(self.h, simtime, last_dt, timestep), "set_time")

    def get_time(self):
        a, b, c_, t = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        self.lib.pion_gpu_get_time(self.h, C.byref(a), C.byref(b), C.byref(c_), C.byref(t))
        return a.value, b.value, c_.value, t.value

    def calc_microphysics_dU(self, dt):
        self._ck(self.lib.pion_gpu_calc_microphysics_dU(self.h, dt), "calc_microphysics_dU")

    def calc_dynamics_dU(self, dt, step):
        self._ck(self.lib.pion_gpu_calc_dynamics_dU(self.h, dt, step), "calc_dynamics_dU")

    def grid_update_state_vector(self, dt, step, ooa):
        self._ck(self.lib.pion_gpu_grid_update_state_vector(self.h, dt, step, ooa), "grid_update_state_vector")

    def time_update_bcs(self, simtime, cstep, maxstep):
        self._ck(self.lib.pion_gpu_time_update_bcs(self.h, simtime, cstep, maxstep), "time_update_bcs")

    def time_update_internal_bcs(self, simtime, cstep, maxstep):
        self._ck(self.lib.pion_gpu_time_update_internal_bcs(self.h, simtime, cstep, maxstep), "time_update_internal_bcs")

    def time_update_external_bcs(self, simtime, cstep, maxstep):
        self._ck(self.lib.pion_gpu_time_update_external_bcs(self.h, simtime, cstep, maxstep), "time_update_external_bcs")

    def advance_time(self):
        a = C.c_double()
        self._ck(self.lib.pion_gpu_advance_time(self.h, C.byref(a)), "advance_time")
        return a.value

    def run(self, nsteps):
        dts = np.zeros(nsteps)
        self._ck(self.lib.pion_gpu_run(self.h, nsteps, dts.ctypes.data), "run")
        return dts

    def output_due(self, opfreq=0):
        d = C.c_int()
        self._ck(self.lib.pion_gpu_output_due(self.h, opfreq, C.byref(d)), "output_due")
        return bool(d.value)

    def counters(self):
        out = (C.c_longlong * 3)()
        self._ck(self.lib.pion_gpu_counters(self.h, out), "counters")
        return list(out)

    def riemann_failures(self):
        out = C.c_longlong(0)
        self._ck(self.lib.pion_gpu_riemann_failures(self.h, C.byref(out)), "riemann_failures")
        return int(out.value)

    def mp_failures(self):
        out = C.c_longlong(0)
        self._ck(self.lib.pion_gpu_mp_failures(self.h, C.byref(out)), "mp_failures")
        return int(out.value)

    def sync(self):
        self._ck(self.lib.pion_gpu_sync(self.h), "sync")

    def stream(self):
        return self.lib.pion_gpu_stream(self.h)

    def stage_timing(self, enable: bool):
        ms, n = C.c_double(), C.c_longlong()
        self._ck(self.lib.pion_gpu_stage_timing(self.h, 1 if enable else 0, C.byref(ms), C.byref(n)), "stage_timing")
        return ms.value, n.value

    def describe(self) -> str:
        buf = C.create_string_buffer(512)
        self._ck(self.lib.pion_gpu_describe(self.h, buf, 512), "describe")
        return buf.value.decode()

    def nccl_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.pion_gpu_nccl_init(self.h, buf), "nccl_init")

    def close(self):
        if getattr(self, "h", None):
            self.lib.pion_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
