"""Worker of tests/test_exchange_gloo.py: one rank of a domain-decomposed run under torch.distributed (gloo),
tracking done by the CPU oracle through the same driver code (quicksilver_b200.driver.exchange_rounds /
Simulation) that drives the GPUs under NCCL.  Rank 0 writes the global cycle rows and every rank writes its census
(cells as global ids) to the output directory."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    import helpers as H
    from quicksilver_b200 import driver

    out_dir, cycles = sys.argv[1], int(sys.argv[2])
    argv = sys.argv[3:]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if os.environ.get("QSB_TEST_BACKEND") == "device":
        # the product path: validation kernels on one GPU per rank, NCCL for the exchange
        import torch
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
        sim = driver.Simulation(argv, rank, world, device=local, validation=True, dist=dist, particle_capacity=1 << 20,
                                resident=os.environ.get("QSB_TEST_RESIDENT") == "1")
    else:
        dist.init_process_group("gloo")
        sim = driver.Simulation(argv, rank, world, dist=dist,
                                make_backend=lambda mc: H.OracleBackend(mc.image, mc.get_double("dt"), rank, world, strict=False))
    rows, info = [], []
    gid = sim.mc.image.array("cell_gid")
    domain_offset = sim.mc.image.array("domain_cell_offset")      # several domains per rank: records carry (rank-local domain, cell)
    for c in range(cycles):
        row, flux, meta = sim.cycle()
        rows.append([int(v) for v in row] + [flux])
        info.append({"rounds": meta["rounds"], "sent": meta["sent"]})
        census, _, _ = sim.backend.results()
        census = census.copy()
        census["cell"] = gid[domain_offset[census["domain"]] + census["cell"]]
        census["domain"] = 0
        np.save(os.path.join(out_dir, "census_c%d_r%d.npy" % (c, rank)), census)
    report, passed = sim.report()            # collective: timer table (min/avg/max over ranks) + FOM + CORAL self checks
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump({"rows": rows, "info": info, "exchange": getattr(sim, "exchange", "rounds"), "report": report,
                   "tracking_us": sim.mc.timer("cycleTracking")[0]}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
