"""Multi-GPU parity check (launch with torchrun, one rank per GPU):
the block-decomposed run (NCCL halo exchange + dt all-reduce) against the single-domain
oracle on the same global state.  Rank 0 prints the worst relative error."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from harness import OracleSim, Problem, gpu_config, random_state, rel_err  # noqa: E402
from pion_b200.capi import Context, load_library  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    lib = load_library()
    import ctypes as C
    worst = 0.0
    cases = [
        # big enough for the split stage (boundary shell first, halo exchange overlapped with the interior)
        Problem(ndim=3, NG=(192, 96, 96), eqn="glm-mhd", solver=7, artviscosity=1, xmax=(2.0, 1.0, 1.0), bcs=("periodic",) * 6),
        Problem(ndim=3, NG=(128, 64, 64), eqn="glm-mhd", solver=7, artviscosity=1, xmax=(2.0, 1.0, 1.0), bcs=("periodic",) * 6),
        Problem(ndim=3, NG=(128, 64, 64), eqn="euler", solver=8, artviscosity=1, xmax=(2.0, 1.0, 1.0), bcs=("reflecting", "outflow") * 3, ntracer=1),
        Problem(ndim=3, NG=(32, 16, 16), eqn="glm-mhd", solver=7, artviscosity=1, xmax=(2.0, 1.0, 1.0), bcs=("periodic",) * 6),
        Problem(ndim=3, NG=(16, 16, 16), eqn="glm-mhd", solver=7, artviscosity=1, xmax=(1.0, 1.0, 1.0), bcs=("outflow",) * 6),
        Problem(ndim=3, NG=(16, 16, 16), eqn="euler", solver=4, artviscosity=4, xmax=(1.0, 1.0, 1.0), bcs=("reflecting", "outflow") * 3, ntracer=1),
        Problem(ndim=2, NG=(32, 32, 1), eqn="i-mhd", solver=8, artviscosity=1, xmax=(1.0, 1.0, 1.0), bcs=("periodic", "periodic", "reflecting", "outflow", "periodic", "periodic")),
    ]
    for prob in cases:
        Pg = random_state(prob, seed=99)
        cfg, keep = gpu_config(prob, device=lr)
        assert lib.pion_gpu_decompose_domain(cfg, rank, world) == 0, lib.pion_gpu_last_error()
        ctx = Context(cfg, keep)
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            assert lib.pion_gpu_nccl_unique_id(idbuf) == 0
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        ctx.nccl_init(bytes(t.cpu().numpy().tobytes()))
        # local block (with its ghost frame) cut out of the global padded state
        g = prob.nbc
        off = [int(round((cfg.xmin[a] - prob.xmin[a]) / prob.dx)) if a < prob.ndim else 0 for a in range(3)]
        sl = [slice(off[a], off[a] + cfg.NG[a] + 2 * g) if a < prob.ndim else slice(None) for a in range(3)]
        ctx.upload(Pg[:, sl[2], sl[1], sl[0]])
        ctx.init_after_upload()
        nsteps = 4
        dts = ctx.run(nsteps)
        Pl = ctx.download(0)
        # gather interiors on rank 0
        inner = [slice(g, g + cfg.NG[a]) if a < prob.ndim else slice(None) for a in range(3)]
        mine = torch.from_numpy(np.ascontiguousarray(Pl[:, inner[2], inner[1], inner[0]])).cuda()
        meta = torch.tensor(off + list(cfg.NG), dtype=torch.int64).cuda()
        metas = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta)
        blocks = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(blocks, mine)
        if rank == 0:
            o = OracleSim(prob)
            o.set_state(Pg)
            o.init_after_state()
            do = o.run(nsteps)
            Po = o.get_state(0)[prob.interior()]
            full = np.zeros_like(Po)
            for m, b in zip(metas, blocks):
                m = m.cpu().numpy()
                b = b.cpu().numpy()
                full[:, m[2]:m[2] + (m[5] if prob.ndim > 2 else 1), m[1]:m[1] + (m[4] if prob.ndim > 1 else 1), m[0]:m[0] + m[3]] = b
            e = rel_err(full, Po).max()
            dte = float(np.max(np.abs(dts - do) / do))
            worst = max(worst, e, dte)
            print(f"world={world} {prob.eqn} nd={prob.ndim} bcs={prob.bcs[:2*prob.ndim]} err={e:.2e} dt_err={dte:.1e} negs={ctx.counters()[:2]}", flush=True)
            o.close()
        ctx.close()
    if rank == 0:
        print("MGPU WORST", worst, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
