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


# ------------------------------------------------------------------------------------------------------------
# the reference's ASCII format (dataio_text::output_ascii_data, what icgen writes with "OutputFileType text"):
# pion_ugs_gpu reads it (--in-text) and writes it (--out-text)
def _parse_text(path, ndim, ncol=None):
    rows = [ln.split() for ln in Path(path).read_text().splitlines() if ln and not ln.startswith("#")]
    return np.array(rows, dtype=np.float64)


TEXT_CASES = [("glm-mhd", 7, 3, (6, 5, 4), 0), ("euler", 8, 2, (12, 8, 1), 1), ("i-mhd", 8, 3, (5, 4, 3), 2)]


@pytest.mark.parametrize("eqn,solver,ndim,NG,ntr", TEXT_CASES)
def test_text_format_round_trip_against_the_reference_writer(eqn, solver, ndim, NG, ntr):
    """CPU: the reference writes its ASCII dump of a seeded state (oracle/_ref, dataio_text::OutputData); the driver's
    --convert mode turns it into the raw SoA state and back into the ASCII format.  State: 14 printed digits;
    text: every column of every cell, div B included."""
    from cases import case_2d, case_3d
    from harness import RefSim, have_ref, random_state
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    prob = (case_3d if ndim == 3 else case_2d)(eqn, solver, 1, ntracer=ntr, NG=NG)
    r = RefSim(prob)
    r.set_state(random_state(prob, 5))
    r.init_after_state()
    P = r.get_state(0)
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        r.output_text(d / "ref")
        r.close()
        (d / "params.txt").write_text(prob.paramfile_text())
        # text -> raw
        run = subprocess.run([str(DRIVER), str(d / "params.txt"), "--convert", "--in-text", str(d / "ref.txt"), "--out", str(d / "P.bin")],
                             capture_output=True, text=True)
        assert run.returncode == 0, run.stderr
        Q = np.fromfile(d / "P.bin", dtype=np.float64).reshape(P.shape)
        inner = prob.interior()
        assert np.allclose(Q[inner], P[inner], rtol=2e-14, atol=1e-300)
        # raw (with the reference's ghost cells) -> text
        np.ascontiguousarray(P).tofile(d / "P0.bin")
        run = subprocess.run([str(DRIVER), str(d / "params.txt"), "--convert", "--in", str(d / "P0.bin"), "--out-text", str(d / "ours")],
                             capture_output=True, text=True)
        assert run.returncode == 0, run.stderr
        ours, ref = _parse_text(d / "ours.00000000.txt", ndim), _parse_text(d / "ref.txt", ndim)
        assert ours.shape == ref.shape and ours.shape[1] == ndim + prob.nvar + (1 if eqn == "euler" else 3)
        ncol = ours.shape[1] - (0 if eqn == "euler" else 1)
        assert np.allclose(ours[:, :ncol], ref[:, :ncol], rtol=1e-13, atol=1e-300)
        if eqn != "euler":  # div B: a difference of O(1) numbers divided by dx
            assert np.max(np.abs(ours[:, -1] - ref[:, -1])) <= 1e-12 * np.max(np.abs(P[5:8])) / prob.dx
        # the two header lines and the blank-line structure are the reference's
        a, b = (d / "ours.00000000.txt").read_text().splitlines(), (d / "ref.txt").read_text().splitlines()
        assert a[:2] == b[:2] and [i for i, ln in enumerate(a) if not ln] == [i for i, ln in enumerate(b) if not ln]


@pytest.mark.gpu
def test_cpp_driver_runs_from_the_reference_text_file_and_writes_it_back():
    """GPU: initial conditions in the reference's ASCII format -> pion_ugs_gpu --in-text -> N steps on the device ->
    --out-text, against the reference's own text dump after the same N steps."""
    from cases import case_3d
    from harness import RefSim, have_ref, random_state
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    prob = case_3d("glm-mhd", 7, 1, bcs="outflow", NG=(12, 10, 8))
    nsteps = 4
    r = RefSim(prob)
    r.set_state(random_state(prob, 9))
    r.init_after_state()
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        r.output_text(d / "ics")
        r.run(nsteps)
        r.output_text(d / "ref", nsteps)
        r.close()
        (d / "params.txt").write_text(prob.paramfile_text())
        run = subprocess.run([str(DRIVER), str(d / "params.txt"), "--in-text", str(d / "ics.txt"), "--out-text", str(d / "gpu"),
                              "--steps", str(nsteps)], capture_output=True, text=True)
        assert run.returncode == 0, run.stderr + run.stdout
        ours, ref = _parse_text(d / f"gpu.{nsteps:08d}.txt", 3), _parse_text(d / f"ref.{nsteps:08d}.txt", 3)
        assert ours.shape == ref.shape
        scale = np.max(np.abs(ref), axis=0)
        assert np.max(np.abs(ours[:, :-1] - ref[:, :-1]) / scale[:-1]) < 1e-11  # 14 printed digits in, 14 out
        assert np.max(np.abs(ours[:, -1] - ref[:, -1])) <= 1e-11 * np.max(np.abs(ref[:, 8:11])) / prob.dx
