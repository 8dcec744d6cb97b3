"""The multi-rank path on CPU: world_size 2 and 4 under gloo.  The spatial domain decomposition gives each rank
one domain; boundary-crossing particles are exchanged as packed 160-byte records (MC_Base_Particle + direction
cosine) and the loop ends when no rank sent anything -- the same driver code that runs over NCCL on the GPUs, with
the CPU oracle standing in for the tracking kernels.  Because the benchmark decks use loadBalance 0, the N-rank
cycle rows AND the union of the ranks' census records must equal the single-rank run of the same global problem
bit for bit."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import decks, host


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _single_rank(argv, cycles):
    mc = host.MonteCarlo(argv)
    dt = mc.get_double("dt")
    gid = mc.image.array("cell_gid")
    rows, censuses = [], []
    for _ in range(cycles):
        mc.cycle_init()
        r = H.oracle_track(mc.image, dt, mc.processing(), strict=False, threads=1)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        row, flux = mc.cycle_finalize()
        rows.append([int(v) for v in row] + [flux])
        census = r.census.copy()
        census["cell"] = gid[census["cell"]]
        census["domain"] = 0
        censuses.append(census)
    return rows, censuses


@pytest.mark.parametrize("deck_name,grid,n,per_cell,ranks", [("CTS2", (2, 1, 1), 4, 10, 2), ("Coral2_P1", (2, 2, 1), 3, 40, 4),
                                                               ("Coral2_P1", (2, 2, 1), 3, 40, 2)])
def test_domain_decomposed_run_equals_single_rank_run(tmp_path, deck_name, grid, n, per_cell, ranks):
    """the last case: 4 domains on 2 ranks -- two domains per rank, which the reference stops at (src/initMC.cc:288-289) and
    north_star asks for ("one or more domains per GPU")"""
    gx, gy, gz = grid
    world = ranks
    cycles = 3
    deck = decks.write_deck(decks.derive(deck_name, nSteps=cycles), str(tmp_path / "deck.inp"))
    sizes = ["-X", n * gx, "-Y", n * gy, "-Z", n * gz, "-x", n * gx, "-y", n * gy, "-z", n * gz, "-n", per_cell * n ** 3 * gx * gy * gz]
    argv1 = [str(a) for a in ["-i", deck] + sizes + ["-I", 1, "-J", 1, "-K", 1]]
    argvN = [str(a) for a in ["-i", deck] + sizes + ["-I", gx, "-J", gy, "-K", gz]]
    want_rows, want_census = _single_rank(argv1, cycles)

    out = tmp_path / "out"
    out.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(H.ROOT, "tests", "_exchange_worker.py"), str(out), str(cycles)] + argvN
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:]

    ranks = [json.load(open(out / ("rank%d.json" % r))) for r in range(world)]
    for r in range(world):
        assert ranks[r]["rows"] == ranks[0]["rows"]                 # every rank holds the reduced row
    sent = sum(i["sent"] for r in ranks for i in r["info"])
    assert sent > 0 and max(i["rounds"] for i in ranks[0]["info"]) > 1
    for c in range(cycles):
        got, want = ranks[0]["rows"][c], want_rows[c]
        assert got[:13] == want[:13], "cycle %d: %s != %s" % (c, got[:13], want[:13])
        assert abs(got[13] - want[13]) <= 1e-11 * abs(want[13])
        union = np.concatenate([np.load(out / ("census_c%d_r%d.npy" % (c, r))) for r in range(world)])
        assert H.sort_particles(union).tobytes() == H.sort_particles(want_census[c]).tobytes(), "cycle %d census" % c
    # the closing report is rank 0's; its timer table is reduced over the ranks (src/MC_Fast_Timer.cc:58-105)
    assert all(ranks[r]["report"] == "" for r in range(1, world))
    report = ranks[0]["report"]
    line = [l for l in report.splitlines() if l.startswith("cycleTracking ")][0].split()
    times = [ranks[r]["tracking_us"] for r in range(world)]
    assert int(line[1]) == cycles
    assert float(line[2]) == pytest.approx(min(times), rel=2e-3) and float(line[4]) == pytest.approx(max(times), rel=2e-3)
    assert float(line[3]) == pytest.approx(sum(times) / world, rel=2e-3)
    segs = sum(row[12] for row in ranks[0]["rows"])
    fom = [l for l in report.splitlines() if l.startswith("Figure Of Merit")][0].split()
    assert float(fom[3]) == pytest.approx(segs / (max(times) * 1e-6), rel=2e-3)
    if deck_name == "Coral2_P1":
        assert "Test for lost / unaccounted for particles in this simulation\nPASS:: No Particles Lost During Run" in report
