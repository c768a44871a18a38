#!/usr/bin/env python
"""BASELINE.json config 5 on N GPUs: Wind3D-style stellar-wind bubble (Euler, HLL, FKJ98, 1 tracer, per-cell radiative
cooling EP_cooling 8 with the microphysics timestep limit, constant wind source in the corner of the octant), the
GLOBAL size^3 grid block-decomposed over the ranks exactly like MCMDcontrol::decomposeDomain (MCMD_control.cpp:62-221),
halos by NCCL send/recv, dt (dynamics and cooling time) by a device-side NCCL all-reduce.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/bench_wind_mgpu.py [--size 384] [--steps 10] [--warmup 3]

Prints ONE JSON line on rank 0: throughput of the decomposed run, and `parity_mgpu`: a 64^3 global grid advanced
through the same N-rank path against the same grid on one rank.  Developer / profiling script (profiles/)."""
import argparse
import dataclasses
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))


def wind_problem(n):
    from cases import case_cooling
    base = case_cooling("euler", 8, NG=(n, n, n), ntracer=1, mp_limit=1)
    Lw = 3.160064e18
    wind = dict(pos=(0.0, 0.0, 0.0), radius=1.543e17, mdot=1.0e-7, vinf=1500.0, vrot=0.0, temp=3.0e4, rstar=6.96e11,
                bsrf=10.0, tr=(1.0, 0.0, 0.0, 0.0))
    return dataclasses.replace(base, xmax=(Lw,) * 3, internal_bcs=("stellar-wind",), winds=(wind,),
                               bcs=("reflecting", "one-way-outflow") * 3)


def ambient(shape):
    P = np.zeros(shape)
    P[0] = 2.124e-24
    P[1] = 2.209e-12
    return P


def local_context(gprob, lib, rank, world, local_rank, dist, torch, tables):
    from bench import nccl_attach
    from harness import gpu_config
    from pion_b200.capi import Context
    cfg, keep = gpu_config(gprob, device=local_rank, tables=tables)
    if world > 1:
        assert lib.pion_gpu_decompose_domain(cfg, rank, world) == 0, lib.pion_gpu_last_error()
    ctx = Context(cfg, keep)
    if world > 1:
        nccl_attach(ctx, lib, rank, dist, torch)
    g = gprob.nbc
    shp = (gprob.nvar, cfg.NG[2] + 2 * g, cfg.NG[1] + 2 * g, cfg.NG[0] + 2 * g)
    return ctx, cfg, shp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lib", default=None, help="another build of libpion_b200.so (kernel A/B experiments)")
    args = ap.parse_args()
    import torch
    from bench import timed_steps
    from harness import load_cooling_tables, rel_err
    from pion_b200.capi import load_library
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = load_library(args.lib)
    tables = load_cooling_tables()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity: 64^3 global grid, N ranks vs one rank, 6 steps (the wind switches on in the corner block)
    parity = None
    if world > 1:
        pprob = wind_problem(64)
        psteps = 6
        ctx, cfg, shp = local_context(pprob, lib, rank, world, local_rank, dist, torch, tables)
        ctx.upload(ambient(shp))
        ctx.init_after_upload()
        dts = ctx.run(psteps)
        desc = ctx.describe()
        g = pprob.nbc
        mine = torch.from_numpy(np.ascontiguousarray(ctx.download(0)[:, g:-g, g:-g, g:-g])).cuda()
        dx = pprob.dx
        meta = torch.tensor([int(round(cfg.xmin[a] / dx)) for a in range(3)] + [cfg.NG[a] for a in range(3)], dtype=torch.int64).cuda()
        metas = [torch.zeros_like(meta) for _ in range(world)]
        blocks = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(metas, meta)
        dist.all_gather(blocks, mine)
        fails = ctx.mp_failures()
        ctx.close()
        if rank == 0:
            one, _, shp1 = local_context(pprob, lib, 0, 1, local_rank, dist, torch, tables)
            one.upload(ambient(shp1))
            one.init_after_upload()
            d1 = one.run(psteps)
            P1 = one.download(0)[:, g:-g, g:-g, g:-g]
            one.close()
            full = np.zeros_like(P1)
            for m, b in zip(metas, blocks):
                m = m.cpu().numpy()
                full[:, m[2]:m[2] + m[5], m[1]:m[1] + m[4], m[0]:m[0] + m[3]] = b.cpu().numpy()
            err = rel_err(full, P1, nphys=5)
            parity = {"global_grid": [64] * 3, "ranks": world, "steps": psteps, "max_rel_err": float(err.max()),
                      "dt_max_rel_err": float(np.max(np.abs(dts - d1) / d1)), "tolerance": 1e-12, "ok": bool(err.max() <= 1e-12),
                      "checker": "the same global grid advanced by one rank through the same library",
                      "cooling_integration_failures": int(fails), "multi_rank_path": desc}

    # ---- the timed run
    S = args.size
    gprob = wind_problem(S)
    ctx, cfg, shp = local_context(gprob, lib, rank, world, local_rank, dist, torch, tables)
    ctx.upload(ambient(shp))
    ctx.init_after_upload()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    for _ in range(max(3, args.warmup)):
        ctx.calculate_timestep()
        ctx.advance_time()
    ctx.sync()
    ctx.stage_timing(True)
    l0 = ctx.counters()[2]
    ms = timed_steps(ctx, stream, torch, dist, world, args.steps, barrier)
    st_ms, st_n = ctx.stage_timing(False)
    if rank == 0:
        nvar = gprob.nvar
        value = S ** 3 * args.steps / (ms * 1e-3)
        line = {"metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "scaling": "strong", "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"Wind3D-style 3-D Euler {S}^3 GLOBAL, HLL + FKJ98(0.1), 2nd order, 1 tracer, EP_cooling 8 + "
                                       "MP_timestep_limit 1, constant stellar wind in the octant corner, reflecting / one-way-outflow",
                           "global_grid": [S] * 3, "local_grid": [cfg.NG[q] for q in range(3)]},
                "hbm_roofline_frac_per_gpu": value / world * 5 * nvar * 8 / 6556.5e9,
                "stage_share_of_step": st_ms / ms, "gpu_launches": int(ctx.counters()[2] - l0),
                "negative_density": int(ctx.counters()[0]), "cooling_integration_failures": int(ctx.mp_failures()),
                "multi_rank_path": ctx.describe()}
        if parity is not None:
            line["parity_mgpu"] = parity
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
