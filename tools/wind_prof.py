import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tools")
import bench_configs as bc
from harness import GpuSim, load_cooling_tables
name, prob, icfn, nsteps, _ = bc.configs(True)[-1]
g = GpuSim(prob, tables=load_cooling_tables()); g.set_state(icfn(prob)); g.init_after_state(); g.run(6); g.ctx.sync(); g.close()
