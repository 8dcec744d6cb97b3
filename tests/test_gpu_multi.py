"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise): one process per GPU under torchrun, NCCL
send/recv of the boundary-particle slabs, validation kernels.  The N-GPU cycle rows and the union of the census
vaults must equal the single-rank CPU oracle run of the same global problem bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import decks, host
from test_exchange_gloo import _free_port

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _single_rank_strict(argv, cycles, strict_source=False):
    mc = host.MonteCarlo(argv)
    mc.set_strict_math(strict_source)
    dt = mc.get_double("dt")
    gid = mc.image.array("cell_gid")
    rows, censuses = [], []
    for _ in range(cycles):
        mc.cycle_init()
        r = H.oracle_track(mc.image, dt, mc.processing(), strict=True, threads=os.cpu_count() or 1)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        row, flux = mc.cycle_finalize()
        rows.append([int(v) for v in row] + [flux])
        census = r.census.copy()
        census["cell"] = gid[census["cell"]]
        census["domain"] = 0
        censuses.append(census)
    return rows, censuses


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("deck_name,grid,n,per_cell,n_ranks", [("CTS2", (2, 1, 1), 8, 10, 2), ("Coral2_P1", (2, 2, 1), 6, 40, 4),
                                                                 ("Coral2_P2", (2, 2, 2), 4, 40, 8), ("Coral2_P1", (2, 2, 1), 6, 40, 2)])
def test_n_gpu_run_equals_single_rank_oracle(tmp_path, deck_name, grid, n, per_cell, n_ranks, exchange):
    """exchange = "peer": the kernels deposit boundary particles in each other's rings over NVLink and decide termination
    on the devices (one launch per cycle); "nccl": per-round slabs moved with NCCL send/recv.  Last case: a 2 x 2 domain grid
    on 2 GPUs -- two domains per GPU (north_star: "one or more domains per GPU"; the reference stops there, src/initMC.cc:288-289)."""
    gx, gy, gz = grid
    world = n_ranks
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cycles = 3
    cell = 1.0 / 11.0 if deck_name == "Coral2_P2" else 1.0
    deck = decks.write_deck(decks.derive(deck_name, nSteps=cycles), str(tmp_path / "deck.inp"))
    sizes = ["-X", n * gx * cell, "-Y", n * gy * cell, "-Z", n * gz * cell, "-x", n * gx, "-y", n * gy, "-z", n * gz,
             "-n", per_cell * n ** 3 * gx * gy * gz]
    argv1 = [str(a) for a in ["-i", deck] + sizes + ["-I", 1, "-J", 1, "-K", 1]]
    argvN = [str(a) for a in ["-i", deck] + sizes + ["-I", gx, "-J", gy, "-K", gz]]
    want_rows, want_census = _single_rank_strict(argv1, cycles)

    out = tmp_path / "out"
    out.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(H.ROOT, "tests", "_exchange_worker.py"), str(out), str(cycles)] + argvN
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900,
                         env=dict(os.environ, QSB_TEST_BACKEND="device", QSB_EXCHANGE=exchange, QSB_PEER_WATCHDOG_S="20"))
    assert res.returncode == 0, res.stdout[-3000:]
    ranks = [json.load(open(out / ("rank%d.json" % r))) for r in range(world)]
    assert sum(i["sent"] for r in ranks for i in r["info"]) > 0
    assert all(r["exchange"] == exchange for r in ranks), [r["exchange"] for r in ranks]
    if exchange == "peer":
        assert all(i["rounds"] == 1 for r in ranks for i in r["info"])
    for c in range(cycles):
        got, want = ranks[0]["rows"][c], want_rows[c]
        assert got[:13] == want[:13], "cycle %d: %s != %s" % (c, got[:13], want[:13])
        assert abs(got[13] - want[13]) <= 1e-11 * abs(want[13])
        union = np.concatenate([np.load(out / ("census_c%d_r%d.npy" % (c, r))) for r in range(world)])
        assert H.sort_particles(union).tobytes() == H.sort_particles(want_census[c]).tobytes(), "cycle %d census" % c


@pytest.mark.parametrize("exchange,grid,boundary_first", [("peer", (2, 1, 1), "0"), ("nccl", (2, 1, 1), "0"), ("peer", (2, 2, 1), "0"),
                                                         ("peer", (2, 1, 1), "1"), ("peer", (2, 2, 1), "1")])
def test_n_gpu_resident_cycles_equal_single_rank_cpu_chain(tmp_path, exchange, grid, boundary_first):
    """the population stays on the GPUs from cycle to cycle (cycleInit on every device, split factor from the allreduced
    global count); rows and census must equal the single-rank CPU chain -- host cycleInit in strict-math mode + oracle.
    Grid (2, 2, 1) on 2 GPUs: two domains per GPU.  boundary_first = "1": the opt-in boundary-first list (QSB_BOUNDARY_FIRST,
    DESIGN.md section 6) -- the order in which histories are tracked must not change a bit of the result."""
    world, (gx, gy, gz), n, per_cell, cycles = 2, grid, 8, 10, 3
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    deck = decks.write_deck(decks.derive("CTS2", nSteps=cycles), str(tmp_path / "deck.inp"))
    sizes = ["-X", n * gx, "-Y", n * gy, "-Z", n * gz, "-x", n * gx, "-y", n * gy, "-z", n * gz, "-n", per_cell * n ** 3 * gx * gy * gz]
    argv1 = [str(a) for a in ["-i", deck] + sizes + ["-I", 1, "-J", 1, "-K", 1]]
    argvN = [str(a) for a in ["-i", deck] + sizes + ["-I", gx, "-J", gy, "-K", gz]]
    want_rows, want_census = _single_rank_strict(argv1, cycles, strict_source=True)
    out = tmp_path / "out"
    out.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(H.ROOT, "tests", "_exchange_worker.py"), str(out), str(cycles)] + argvN
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900,
                         env=dict(os.environ, QSB_TEST_BACKEND="device", QSB_EXCHANGE=exchange, QSB_PEER_WATCHDOG_S="20", QSB_TEST_RESIDENT="1",
                                  QSB_BOUNDARY_FIRST=boundary_first))
    assert res.returncode == 0, res.stdout[-3000:]
    ranks = [json.load(open(out / ("rank%d.json" % r))) for r in range(world)]
    assert sum(i["sent"] for r in ranks for i in r["info"]) > 0
    for c in range(cycles):
        got, want = ranks[0]["rows"][c], want_rows[c]
        assert got[:13] == want[:13], "cycle %d: %s != %s" % (c, got[:13], want[:13])
        assert abs(got[13] - want[13]) <= 1e-11 * abs(want[13])
        union = np.concatenate([np.load(out / ("census_c%d_r%d.npy" % (c, r))) for r in range(world)])
        assert H.sort_particles(union).tobytes() == H.sort_particles(want_census[c]).tobytes(), "cycle %d census" % c
