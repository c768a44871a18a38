"""N>1 host logic on CPU (world_size 2, gloo): the product's domain decomposition
(pion_gpu_decompose_domain = MCMDcontrol::decomposeDomain + pointToNeighbours) drives two
ranks that each own one block; halos travel with torch.distributed send/recv in the order the
library uses on the GPU (pack the 2-deep interior slab next to the face, exchange, unpack into the
ghost layers; MPI axis first, then the physical faces), dt is an all-reduce(min).  The per-block
numerics are the plain-C oracle (no GPU here), so the test pins the DECOMPOSITION + EXCHANGE +
REDUCTION logic: the stitched two-rank result must equal the single-domain run bit for bit (the
reference's own serial == parallel identity, solver_eqn_base.cpp:46-48)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _decompose(prob, rank, nproc):
    from harness import gpu_config
    from pion_b200.capi import load_library
    lib = load_library()
    cfg, _ = gpu_config(prob)
    assert lib.pion_gpu_decompose_domain(cfg, rank, nproc) == 0, lib.pion_gpu_last_error()
    return cfg


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_decomposition_tiles_the_domain_and_neighbours_agree(nproc):
    from cases import case_3d
    from harness import BC_CODES
    prob = case_3d("glm-mhd", 7, 1, bcs="reflect-outflow", NG=(16, 16, 8))
    cfgs = [_decompose(prob, r, nproc) for r in range(nproc)]
    cells = sum(c.NG[0] * c.NG[1] * c.NG[2] for c in cfgs)
    assert cells == 16 * 16 * 8
    seen = np.zeros((8, 16, 16), dtype=int)
    for r, c in enumerate(cfgs):
        off = [int(round((c.xmin[a] - prob.xmin[a]) / prob.dx)) for a in range(3)]
        seen[off[2]:off[2] + c.NG[2], off[1]:off[1] + c.NG[1], off[0]:off[0] + c.NG[0]] += 1
        for f in range(6):
            n = c.ngbprocs[f]
            if c.bc[f] == 10:  # PION_BC_MPI: the neighbour points back through the opposite face
                assert 0 <= n < nproc and cfgs[n].ngbprocs[f ^ 1] == r and cfgs[n].bc[f ^ 1] == 10
            else:  # physical face keeps the global boundary type
                assert n == -1 and c.bc[f] == BC_CODES[prob.bcs[f]]
    assert np.all(seen == 1)
    # MCMD_control.cpp:62-221: the longest local axis is halved, ties -> lowest axis.  Box 1 x 1 x 0.5:
    # 2 ranks -> 2x1x1, 4 -> 2x2x1, 8 -> 4x2x1 (after two cuts all three ranges are 0.5: x is cut again)
    expect = {2: (8, 16, 8), 4: (8, 8, 8), 8: (4, 8, 8)}[nproc]
    assert tuple(cfgs[0].NG) == expect


def _worker(rank, world, port, bcs, eqn, solver, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dataclasses
        from cases import case_3d
        from harness import BC_CODES, OracleSim, random_state
        gprob = case_3d(eqn, solver, 1, bcs=bcs, NG=(16, 8, 6))
        cfg = _decompose(gprob, rank, world)
        g = gprob.nbc
        names = {v: k for k, v in BC_CODES.items()}
        names[10] = "MPI"
        BC_CODES["MPI"] = 10
        lbcs = tuple(names[cfg.bc[f]] for f in range(6))
        lprob = dataclasses.replace(gprob, NG=tuple(cfg.NG), xmin=tuple(cfg.xmin), xmax=tuple(cfg.xmax), bcs=lbcs)
        off = [int(round((cfg.xmin[a] - gprob.xmin[a]) / gprob.dx)) for a in range(3)]
        Pg = random_state(gprob, seed=77)
        sl = [slice(off[a], off[a] + cfg.NG[a] + 2 * g) for a in range(3)]
        sim = OracleSim(lprob)
        sim.set_state(Pg[:, sl[2], sl[1], sl[0]].copy())

        def exchange(which):
            """BC_update_BCMPI for the x faces: interior slab next to the face -> neighbour's ghost layers."""
            A = sim.get_state(which)
            reqs, recv = [], {}
            for f in (0, 1):
                if cfg.bc[f] != 10:
                    continue
                peer = cfg.ngbprocs[f]
                send = A[..., g:2 * g] if f == 0 else A[..., -2 * g:-g]
                t = torch.from_numpy(np.ascontiguousarray(send))
                recv[f] = torch.empty_like(t)
                reqs.append(dist.isend(t, peer, tag=f))
                reqs.append(dist.irecv(recv[f], peer, tag=f ^ 1))
            for r_ in reqs:
                r_.wait()
            for f, t in recv.items():
                if f == 0:
                    A[..., :g] = t.numpy()
                else:
                    A[..., -g:] = t.numpy()
            sim.set_state(A, which)

        def bcs_update(cstep, maxstep):
            exchange(1)  # Ph
            if cstep == maxstep:
                exchange(0)
            sim.update_bcs(cstep, maxstep)  # physical faces (y, z incl. the x-ghost corners)

        # sim_init::Init: Ph=P, psi=0, assign + first update
        sim.init_after_state()
        exchange(0)
        exchange(1)
        sim.update_bcs(2, 2)
        dts = []
        for _ in range(3):
            td = torch.tensor([sim.dynamics_dt(), sim.microphysics_dt()], dtype=torch.float64)
            dist.all_reduce(td, op=dist.ReduceOp.MIN)  # sim_control_MPI.cpp:503-504
            t_dyn = float(td[0])
            simtime, last_dt = sim.info()[1][0], sim.info()[1][2]
            dt = min(t_dyn, float(td[1]), 1.3 * last_dt)
            if gprob.eqn == "glm-mhd":
                sim.set_glm_speeds(t_dyn, gprob.dx, 0.25 / gprob.dx)
            # advance_time, second order (time_integrator.cpp:72-142) through the seam calls
            sim.set_dt(0.5 * dt)
            sim.dynamics_dU(0.5 * dt, 1)
            sim.update_state(0.5 * dt, 1, 2)
            bcs_update(1, 2)
            sim.set_dt(dt)
            sim.dynamics_dU(dt, 2)
            sim.update_state(dt, 2, 2)
            bcs_update(2, 2)
            sim.set_time(simtime + dt, dt, len(dts) + 1)
            dts.append(dt)
        inner = sim.get_state(0)[lprob.interior()]
        q.put((rank, off, tuple(cfg.NG), inner, dts))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bcs,eqn,solver", [("periodic", "glm-mhd", 7), ("reflect-outflow", "euler", 4)])
def test_two_ranks_equal_single_domain(bcs, eqn, solver):
    import torch.multiprocessing as mp
    from cases import case_3d
    from harness import OracleSim, random_state
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bcs, eqn, solver, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    gprob = case_3d(eqn, solver, 1, bcs=bcs, NG=(16, 8, 6))
    o = OracleSim(gprob)
    o.set_state(random_state(gprob, seed=77))
    o.init_after_state()
    do = o.run(3)
    Po = o.get_state(0)[gprob.interior()]
    o.close()
    full = np.zeros_like(Po)
    for rank, off, NG, inner, dts in res:
        assert np.array_equal(np.array(dts), do), (dts, do)
        full[:, off[2]:off[2] + NG[2], off[1]:off[1] + NG[1], off[0]:off[0] + NG[0]] = inner
    assert np.array_equal(full, Po), float(np.max(np.abs(full - Po)))
