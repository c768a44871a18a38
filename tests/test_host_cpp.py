"""The C++ host side (pion_b200/host: class sim_control_gpu + the pion_ugs_gpu driver).

CPU: the host library and the driver are built, load, and export the mirrored seam methods.
GPU: the driver, fed a reference-format parameter file and the golden initial state, must
reproduce the compiled reference's golden final state (same tolerance as the ctypes path) --
this is the end-to-end C++ path: parameter file -> sim_control_gpu::Init -> Time_Int ->
output_data."""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

from harness import load_cooling_tables, rel_err, TABLE_KEYS

ROOT = Path(__file__).resolve().parent.parent
HOSTLIB = ROOT / "pion_b200" / "libpion_b200_host.so"
DRIVER = ROOT / "pion_b200" / "pion_ugs_gpu"
GOLD = Path(__file__).resolve().parent / "golden"


def test_host_library_exports_the_seam_methods():
    assert HOSTLIB.exists() and DRIVER.exists(), "run __graft_entry__.build()"
    syms = subprocess.run(["nm", "-DC", str(HOSTLIB)], capture_output=True, text=True, check=True).stdout
    for m in ["Init", "Time_Int", "Finalise", "calculate_timestep", "calc_dynamics_dt", "calc_microphysics_dt",
              "advance_time", "calc_microphysics_dU", "calc_dynamics_dU", "grid_update_state_vector",
              "TimeUpdateInternalBCs", "TimeUpdateExternalBCs", "output_data", "check_eosim"]:
        assert f"pion_b200::sim_control_gpu::{m}(" in syms, m


def test_driver_usage_without_arguments():
    r = subprocess.run([str(DRIVER)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def _load_case(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    prob, nsteps = mod.CASES[name]
    return prob, nsteps, np.load(GOLD / f"{name}.npz")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tp_BWcrt3D_octant_n016", "tp_DMR_n065_roe", "tp_FieldLoop_64x32_hlld",
                                  "glm_hlld_fkj_3d_outflow", "cool_euler_hll_3d_dense", "wind3d_euler_hll_cool_n016",
                                  "wind3d_glm_hlld_nocool_n012"])
def test_cpp_driver_reproduces_reference_golden(name):
    prob, nsteps, z = _load_case(name)
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        (d / "params.txt").write_text(prob.paramfile_text())
        np.ascontiguousarray(z["P0"], dtype=np.float64).tofile(d / "P0.bin")
        cmd = [str(DRIVER), str(d / "params.txt"), "--in", str(d / "P0.bin"), "--out", str(d / "P.bin"), "--steps", str(nsteps)]
        if prob.cooling:
            tab = load_cooling_tables()
            np.concatenate([np.ascontiguousarray(tab[k], dtype=np.float64) for k in TABLE_KEYS]).tofile(d / "tables.bin")
            cmd += ["--tables", str(d / "tables.bin")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        assert "STEPS:" in r.stdout  # the reference's cell-updates/s line (sim_control.cpp:270-277)
        P = np.fromfile(d / "P.bin", dtype=np.float64).reshape(z["P"].shape)
    err = rel_err(P, z["P"], nphys=prob.nvar - prob.ntracer)
    assert err.max() < 5e-12, err
