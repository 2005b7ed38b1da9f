"""torchrun worker: the reference's examples/parallel_tempering/runner.jl flow through the host mirror —
one process per GPU, each holding a block of temperature slots, `parallel_tempering` with output files."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    try:
        _main()
    except Exception:
        import traceback
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"pt_driver_rank{os.environ.get('RANK', '0')}.err"), "w") as f:
            traceback.print_exc(file=f)
        traceback.print_exc()
        raise


def _main():
    import torch
    import torch.distributed as dist

    import classicalspinmc.jl_b200 as csm
    from classicalspinmc.jl_b200 import hdf5 as h5
    from tests import models

    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per_rank = 3
    temp = np.geomspace(0.09 / 11.6, 14 / 11.6, per_rank * world)          # runner.jl:14
    T = temp[rank * per_rank:(rank + 1) * per_rank]
    P = models.pyrochlore_local()
    lat = csm.Lattice((4, 4, 4), P, 0.5, rng=np.random.default_rng(100 + rank))
    params = {"t_thermalization": 400, "t_measurement": 1200, "probe_rate": 20, "swap_rate": 10,
              "overrelaxation_rate": 5, "report_interval": 800, "checkpoint_rate": 600}
    mc = csm.MonteCarlo(T, lat, params, outpath=out, seed=4242)
    csm.parallel_tempering(mc, [0])
    st = mc.statistics
    ok = st["energy_series"].shape == (60, per_rank * world)
    ok &= sorted(st["slot_of_replica"].tolist()) == list(range(per_rank * world))
    ok &= st["exchanges"].sum() > 0
    means = [o.energy.mean(1) for o in mc.observables_all]
    ok &= all(np.isfinite(means))
    dist.barrier()
    if rank == 0:
        files = sorted(f for f in os.listdir(out) if f.endswith(".h5"))
        ok &= files == [f"configuration_{s}.h5" for s in range(per_rank * world)]
        E = []
        for s in range(per_rank * world):
            obs = h5.read_observables(os.path.join(out, f"configuration_{s}.h5"))
            E.append(obs["energy"])
        ok &= bool(E[0] < E[-1])
        ok &= os.path.isfile(os.path.join(out, "IC_0", "IC_0.h5"))
        print(json.dumps({"ok": bool(ok), "world": world, "E_cold": E[0], "E_hot": E[-1]}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
