"""CPU: the C-ABI library loads and exports every symbol include/pion_b200.h declares;
without a GPU it refuses to create a context (there is no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    txt = (ROOT / "include" / "pion_b200.h").read_text()
    return sorted(set(re.findall(r"\b(pion_gpu_[A-Za-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pion_b200.capi import EXPORTED_SYMBOLS, LIB_PATH, load_library
    assert LIB_PATH.exists(), "libpion_b200.so not built: run __graft_entry__.build()"
    lib = load_library()
    decl = declared_symbols()
    assert len(decl) >= 24
    for name in decl:
        assert hasattr(lib, name), name
    assert set(decl) == set(EXPORTED_SYMBOLS)


def test_product_does_not_link_the_oracle():
    import subprocess
    from pion_b200.capi import LIB_PATH
    out = subprocess.run(["ldd", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in out and "pion_ref" not in out
    syms = subprocess.run(["nm", "-D", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert " po_" not in syms and " pref_" not in syms
    for src in (ROOT / "pion_b200").rglob("*"):
        if src.suffix in (".cu", ".cuh", ".cpp", ".h", ".py"):
            assert "oracle" not in src.read_text().replace("no CPU fallback", ""), src


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from harness import GpuSim, Problem
    with pytest.raises(RuntimeError, match="no CUDA device"):
        GpuSim(Problem())


def test_config_validation_matches_reference_errors():
    from pion_b200.capi import GpuConfig, load_library
    lib = load_library()
    c = GpuConfig()
    c.ndim = 2
    c.eqntype, c.nvar, c.solver, c.coord_sys, c.spOOA, c.tmOOA = 3, 9, 7, 1, 2, 1
    assert not lib.pion_gpu_create(ctypes.byref(c))
    assert b"Bad OOA requests" in lib.pion_gpu_last_error()  # time_integrator.cpp:130


def test_every_supported_equation_solver_pair_passes_the_configuration_check():
    """check_config runs before the device is touched, so it can be exercised here: every (equation set, solverType,
    artviscosity) the parity suite uses must get PAST it (on this GPU-less host that means the next error, "no CUDA device"),
    and what the reference rejects must be rejected with a message (solver_eqn_mhd_adi.cpp:132-198, riemannMHD.cpp:176-183,
    mp_only_cooling.cpp:418)."""
    import dataclasses
    import torch
    from cases import AVS, EQ_SOLVERS, case_2d, case_cooling
    from harness import gpu_config, load_cooling_tables
    from pion_b200.capi import load_library
    lib = load_library()

    def create_error(prob, tables=None):
        cfg, keep = gpu_config(prob, tables=tables)
        h = lib.pion_gpu_create(ctypes.byref(cfg))
        if h:
            lib.pion_gpu_destroy(h)
            return b""
        return lib.pion_gpu_last_error()

    have_gpu = torch.cuda.is_available()
    for eqn, solver in EQ_SOLVERS:
        for av in AVS:
            err = create_error(case_2d(eqn, solver, av))
            assert err == (b"" if have_gpu else err) and (have_gpu or b"no CUDA device" in err), (eqn, solver, av, err)
    for flag in (2, 4, 5, 6, 7, 8):
        prob = dataclasses.replace(case_cooling("euler", 8), cooling=flag)
        from harness import tables_for
        err = create_error(prob, tables=tables_for(prob))
        assert have_gpu or b"no CUDA device" in err, (flag, err)
    # rejected like the reference
    for eqn, solver, word in (("i-mhd", 2, b"Euler equations only"), ("glm-mhd", 3, b"Euler equations only"), ("i-mhd", 5, b"Euler equations only"),
                              ("glm-mhd", 6, b"Euler equations only"), ("euler", 7, b"HLLD needs MHD"), ("euler", 9, b"solver must be")):
        err = create_error(case_2d(eqn, solver, 0))
        assert word in err, (eqn, solver, err)
    err = create_error(dataclasses.replace(case_cooling("euler", 8), cooling=3), tables=load_cooling_tables())
    assert b"cooling flag" in err, err
    err = create_error(case_2d("euler", 8, 2))
    assert b"artviscosity" in err, err


def test_decompose_domain_matches_mcmd():
    """MCMDcontrol::decomposeDomain: 8 ranks on a cube -> 2x2x2, rank = nx*ny*iz + nx*iy + ix."""
    from harness import Problem, gpu_config
    from pion_b200.capi import load_library
    lib = load_library()
    prob = Problem(ndim=3, NG=(64, 64, 64), xmin=(0, 0, 0), xmax=(1, 1, 1), bcs=("periodic", "periodic", "outflow", "outflow", "reflecting", "outflow"))
    for rank in range(8):
        cfg, _ = gpu_config(prob)
        assert lib.pion_gpu_decompose_domain(cfg, rank, 8) == 0
        ix, iy, iz = rank % 2, (rank // 2) % 2, rank // 4
        assert list(cfg.NG) == [32, 32, 32]
        assert cfg.xmin[0] == 0.5 * ix and cfg.xmin[1] == 0.5 * iy and cfg.xmin[2] == 0.5 * iz
        # x is periodic: both faces talk to the other x-rank
        assert cfg.bc[0] == 10 and cfg.bc[1] == 10 and cfg.ngbprocs[0] == cfg.ngbprocs[1] == rank ^ 1
        assert cfg.bc[2] == (10 if iy else 2) and cfg.bc[3] == (2 if iy else 10)
        assert cfg.bc[4] == (10 if iz else 4) and cfg.bc[5] == (2 if iz else 10)
    cfg, _ = gpu_config(Problem(ndim=2, NG=(64, 32, 1), xmax=(2.0, 1.0, 1.0)))
    assert lib.pion_gpu_decompose_domain(cfg, 1, 2) == 0
    assert list(cfg.NG)[:2] == [32, 32] and cfg.xmin[0] == 1.0  # longest axis split first


@pytest.mark.skipif(not Path("/root/reference/source/constants.h").exists(), reason="reference tree not mounted")
def test_header_codes_are_the_reference_integers():
    """The integer codes of include/pion_b200.h are the reference's own (constants.h, boundaries/boundaries.h), so a
    maintainer can pass SimPM.eqntype / solverType / artviscosity / BC codes straight through."""
    hdr = (ROOT / "include" / "pion_b200.h").read_text()
    mine = {k: int(v) for k, v in re.findall(r"\b(PION_[A-Z0-9_]+)\s*=\s*(\d+)", hdr)}
    consts = Path("/root/reference/source/constants.h").read_text()
    ref = {k: int(v) for k, v in re.findall(r"#define\s+([A-Za-z0-9_]+)\s+(\d+)\b", consts)}
    bcs = Path("/root/reference/source/boundaries/boundaries.h").read_text()
    ref.update({k: int(v) for k, v in re.findall(r"\b([A-Z0-9_]+)\s*=\s*(\d+)\s*,", bcs)})
    pairs = {"PION_EQEUL": "EQEUL", "PION_EQMHD": "EQMHD", "PION_EQGLM": "EQGLM", "PION_COORD_CRT": "COORD_CRT",
             "PION_COORD_CYL": "COORD_CYL", "PION_COORD_SPH": "COORD_SPH", "PION_FLUX_LF": "FLUX_LF", "PION_FLUX_RSLINEAR": "FLUX_RSlinear", "PION_FLUX_RSEXACT": "FLUX_RSexact", "PION_FLUX_RSHYBRID": "FLUX_RShybrid", "PION_FLUX_ROE": "FLUX_RSroe",
             "PION_FLUX_ROE_PV": "FLUX_RSroe_pv", "PION_FLUX_FVS": "FLUX_FVS", "PION_FLUX_HLLD": "FLUX_RS_HLLD",
             "PION_FLUX_HLL": "FLUX_RS_HLL", "PION_AV_NONE": "AV_NONE", "PION_AV_FKJ98": "AV_FKJ98_1D",
             "PION_AV_HCORR": "AV_HCORRECTION", "PION_AV_HCORR_FKJ98": "AV_HCORR_FKJ98", "PION_BC_PERIODIC": "PERIODIC",
             "PION_BC_OUTFLOW": "OUTFLOW", "PION_BC_INFLOW": "INFLOW", "PION_BC_REFLECTING": "REFLECTING",
             "PION_BC_FIXED": "FIXED", "PION_BC_DMACH": "DMACH", "PION_BC_DMACH2": "DMACH2", "PION_BC_MPI": "BCMPI",
             "PION_BC_ONEWAY_OUT": "ONEWAY_OUT", "PION_BC_STWIND": "STWIND"}
    for m, r in pairs.items():
        assert m in mine and r in ref, (m, r)
        assert mine[m] == ref[r], (m, mine[m], r, ref[r])
