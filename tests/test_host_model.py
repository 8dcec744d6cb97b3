"""Host-side logic around the hot path: the deck / command-line surface (src/Parameters.cc), the mesh tables
the kernels hard-code, the spatial decomposition over ranks (src/initMC.cc:240-320, src/MeshPartition.cc)."""
import os
import re

import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import decks, host
from test_oracle_golden import FACET_POINTS, OPPOSING_FACET


def _mc(tmp_path, deck, *cli, rank=0, n_ranks=1, name="d.inp"):
    path = decks.write_deck(deck, str(tmp_path / name))
    return host.MonteCarlo(["-i", path] + [str(c) for c in cli], rank, n_ranks)


def test_deck_overrides_command_line(tmp_path):
    """CLI is parsed first, the deck second (src/Parameters.cc:80-95): a key present in both keeps the deck's value."""
    deck = decks.derive("CTS2_1", nx=4, ny=4, nz=4, lx=4, ly=4, lz=4, nParticles=640, nSteps=2)
    mc = _mc(tmp_path, deck, "-x", 9, "-n", 12345, "-N", 7)
    assert mc.get_int("nx") == 4 and mc.get_int("nParticles") == 640 and mc.get_int("nSteps") == 2
    # a deck that omits the sizes takes them from the command line (Examples/CTS2_Benchmark/CTS2.inp)
    mc = _mc(tmp_path, decks.derive("CTS2", nSteps=1), "-X", 6, "-Y", 6, "-Z", 6, "-x", 6, "-y", 6, "-z", 6, "-n", 2160, name="e.inp")
    assert (mc.get_int("nx"), mc.get_int("ny"), mc.get_int("nz")) == (6, 6, 6)
    assert mc.get_int("nParticles") == 2160 and mc.image.n_cells == 216


def test_echoed_parameters_are_a_valid_deck(tmp_path):
    """'output is a valid input' (src/Parameters.cc:97-116): feeding the echo back builds the same problem."""
    mc = _mc(tmp_path, decks.derive("NonFlatXC", nx=4, ny=4, nz=4, nParticles=1000, nSteps=1, dt=5e-10))
    echo = mc.print_parameters()
    assert "Simulation:" in echo and "CrossSection:" in echo and "Material:" in echo
    path = tmp_path / "echo.inp"
    path.write_text(echo)
    mc2 = host.MonteCarlo(["-i", str(path)])
    for name in ("planes", "nodes", "xs_react", "xs_total", "energies", "cell_material", "face_event", "face_adj_cell"):
        assert np.array_equal(mc.image.array(name), mc2.image.array(name)), name
    mc.cycle_init(), mc2.cycle_init()
    assert mc.processing().tobytes() == mc2.processing().tobytes()


def test_unknown_keys_are_ignored_and_bad_decks_are_reported(tmp_path):
    good = decks.deck_text(decks.derive("CTS2_1", nx=4, ny=4, nz=4, lx=4, ly=4, lz=4, nParticles=640, nSteps=1))
    p = tmp_path / "u.inp"
    p.write_text(good.replace("Simulation:", "Simulation:\n   notAKey: 17"))
    assert host.MonteCarlo(["-i", str(p)]).get_int("nx") == 4
    with pytest.raises(host.QsbError):
        host.MonteCarlo(["-i", str(tmp_path / "missing.inp")])


def test_kernel_facet_tables_follow_from_the_reference_tables():
    """c_facet_points / c_facet_of_edge in track_physics.cuh: the first is the reference's nodeIndirect table
    (src/MC_Domain.cc:41-50); the second (face, edge) -> facet map is derived here from it and the 14-point
    layout of a cell (8 corners in 000,100,010,110,001,... order, then the +x,-x,+y,-y,+z,-z face centres)."""
    src = open(os.path.join(H.ROOT, "quicksilver_b200", "csrc", "device", "track_physics.cuh")).read()
    m = re.search(r"c_facet_points\[24\]\[4\]\s*=\s*\{(.*?)\};", src, re.S)
    pts = np.array([int(v) for v in re.findall(r"-?\d+", m.group(1))]).reshape(24, 4)[:, :3]
    assert np.array_equal(pts, np.array(FACET_POINTS))
    m = re.search(r"c_facet_of_edge\[6\]\[4\]\s*=\s*\{(.*?)\};", src, re.S)
    table = np.array([int(v) for v in re.findall(r"-?\d+", m.group(1))]).reshape(6, 4)
    corner = lambda k: np.array([k & 1, (k >> 1) & 1, (k >> 2) & 1])
    for f in range(24):
        face = f // 4
        axis = face // 2
        a, b = corner(FACET_POINTS[f][0]), corner(FACET_POINTS[f][1])
        assert FACET_POINTS[f][2] == 8 + face                    # third point = that face's centre
        assert a[axis] == b[axis] == (0 if face & 1 else 1)      # both corners lie on the face
        u, v = [ax for ax in range(3) if ax != axis]
        if a[u] == b[u]:
            edge = 1 if a[u] == 1 else 0                         # base edge at low / high u
        else:
            assert a[v] == b[v]
            edge = 3 if a[v] == 1 else 2
        assert table[face][edge] == f
    # opposing facet: same triangle seen from the neighbour (its face index flips, the corner set mirrors)
    for f in range(24):
        g = OPPOSING_FACET[f]
        assert OPPOSING_FACET[g] == f and (g // 4) == ((f // 4) ^ 1)


@pytest.mark.parametrize("grid", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_grid_decomposition_is_a_partition_with_symmetric_faces(tmp_path, grid):
    """One domain per rank on an xDom x yDom x zDom grid of centres (src/initMC.cc:288-289,371-384): ranks'
    cells partition the global grid, and every off-processor face points at the geometric neighbour's cell
    index on the owning rank, whose matching face points back."""
    gx, gy, gz = grid
    n_ranks = gx * gy * gz
    n = 4
    deck = decks.derive("CTS2", nSteps=1)
    cli = ["-X", n * gx, "-Y", n * gy, "-Z", n * gz, "-x", n * gx, "-y", n * gy, "-z", n * gz, "-I", gx, "-J", gy, "-K", gz,
           "-n", 10 * n ** 3 * n_ranks]
    models = [_mc(tmp_path, deck, *cli, rank=r, n_ranks=n_ranks, name="g%d.inp" % r) for r in range(n_ranks)]
    gids = [m.image.array("cell_gid").copy() for m in models]
    everything = np.concatenate(gids)
    assert len(everything) == (n ** 3) * n_ranks and len(np.unique(everything)) == len(everything)
    nx, ny = n * gx, n * gy
    step = [1, -1, nx, -nx, nx * ny, -nx * ny]
    for r, m in enumerate(models):
        im = m.image
        assert im.n_ranks == n_ranks and im.my_rank == r and im.n_domains == 1
        assert np.all(np.diff(gids[r]) > 0)                      # ascending global id (src/MeshPartition.cc:144-155)
        ev, adj, nbr = im.array("face_event"), im.array("face_adj_cell"), im.array("face_nbr_rank")
        off = ev == 4
        assert off.any() and np.all(nbr[off] >= 0) and np.all(nbr[~off] == -1)
        for cell, face in zip(*np.nonzero(off)):
            peer = int(nbr[cell, face])
            assert peer != r
            want_gid = int(gids[r][cell]) + step[face]
            assert int(gids[peer][adj[cell, face]]) == want_gid
            back = models[peer].image
            assert int(back.array("face_event")[adj[cell, face], face ^ 1]) == 4
            assert int(back.array("face_nbr_rank")[adj[cell, face], face ^ 1]) == r
            assert int(back.array("face_adj_cell")[adj[cell, face], face ^ 1]) == cell
        on = ev == 3
        for cell, face in zip(*np.nonzero(on)):
            assert int(gids[r][adj[cell, face]]) == int(gids[r][cell]) + step[face]


def test_source_is_decomposition_independent(tmp_path):
    """loadBalance 0 decks: seeds are keyed by global cell id and the split factor is global, so the union of the
    ranks' processing vaults equals the single-rank vault particle for particle."""
    n, gx = 4, 2
    deck = decks.derive("CTS2", nSteps=1)
    cli = ["-X", n * gx, "-Y", n, "-Z", n, "-x", n * gx, "-y", n, "-z", n, "-n", 10 * n ** 3 * gx]
    one = _mc(tmp_path, deck, *cli, "-I", 1, "-J", 1, "-K", 1, name="one.inp")
    one.cycle_init()
    want = one.processing()

    # two ranks in one process: the allreduce callback needs both contributions, so run the two cycle_inits
    # with a pre-computed sum (source weight and particle count are the only reductions)
    parts = []
    totals = {}

    def make_allreduce(rank):
        def fn(arr):
            key = (arr.dtype.str, len(arr), len(totals.get(rank, [])))
            totals.setdefault(rank, []).append(arr.copy())
            if rank == 1:
                arr += totals[0][len(totals[1]) - 1]             # rank 0 ran first and recorded its local value
        return fn

    r0 = _mc(tmp_path, deck, *cli, "-I", gx, "-J", 1, "-K", 1, rank=0, n_ranks=2, name="r0.inp")
    r1 = _mc(tmp_path, deck, *cli, "-I", gx, "-J", 1, "-K", 1, rank=1, n_ranks=2, name="r1.inp")
    r0.set_allreduce(make_allreduce(0))
    r1.set_allreduce(make_allreduce(1))
    r0.cycle_init()          # its reductions see only the local values: discarded, used to feed rank 1
    r1.cycle_init()          # sees local + rank 0
    got1 = r1.processing()
    gid1 = r1.image.array("cell_gid")[got1["cell"]]
    gid_want = one.image.array("cell_gid")[want["cell"]]
    mine = want[np.isin(gid_want, r1.image.array("cell_gid"))]
    a = H.sort_particles(got1).copy()
    b = H.sort_particles(mine).copy()
    a["cell"] = gid1[np.argsort(got1["identifier"], kind="stable")]
    b["cell"] = gid_want[np.isin(gid_want, r1.image.array("cell_gid"))][np.argsort(mine["identifier"], kind="stable")]
    a["domain"] = b["domain"] = 0
    assert a.tobytes() == b.tobytes()


def test_strict_math_source_is_the_libm_source_up_to_rounding(tmp_path):
    """qsb_mc_set_strict_math: MC_SourceNow with the portable log/sin/cos the device cycle-init kernel uses.  Same streams,
    same counts, same identifiers and seeds as the libm mode the golden fixtures pin to the reference; only the two
    direction components that go through sin/cos and the number of mean free paths (log) may move, by a few ulp."""
    deck = decks.derive("CTS2_1", nx=6, ny=6, nz=6, lx=6, ly=6, lz=6, nParticles=4320, nSteps=1)
    a, b = _mc(tmp_path, deck, name="libm.inp"), _mc(tmp_path, deck, name="strict.inp")
    b.set_strict_math(True)
    a.cycle_init(), b.cycle_init()
    pa, pb = a.processing(), b.processing()
    assert len(pa) == len(pb) > 4000
    for field in ("identifier", "random_number_seed", "coordinate", "kinetic_energy", "weight", "time_to_census", "cell", "domain"):
        assert np.array_equal(pa[field], pb[field]), field
    assert np.array_equal(pa["velocity"][:, 2], pb["velocity"][:, 2])            # gamma does not pass through sin/cos
    speed = np.linalg.norm(pa["velocity"], axis=1)
    assert np.abs(pa["velocity"] - pb["velocity"]).max() <= 4e-16 * speed.max()
    assert np.allclose(pa["num_mean_free_paths"], pb["num_mean_free_paths"], rtol=4e-16, atol=1e-18)
    # negative azimuths (half of all particles) go through the symmetry branch of the strict sin/cos
    assert (pa["velocity"][:, 1] < 0).sum() > 1000


def test_host_cycle_init_parts_agree_with_the_whole(tmp_path):
    """the global numbers the device cycle-init is handed (source particle weight, per-cell source counts, split factor)
    are the ones the host's own cycleInit uses: start + source + split - rr particles come out of it"""
    deck = decks.derive("CTS2_1", nx=6, ny=6, nz=6, lx=6, ly=6, lz=6, nParticles=4320, nSteps=2)
    mc = _mc(tmp_path, deck)
    mc.cycle_init()
    n = len(mc.processing())
    # 10 % of nParticles are sourced per cycle (src/MC_SourceNow.cc:59), the population is split up to the target
    w = mc.get_double("source_particle_weight")
    assert w > 0
    assert abs(n - 4320) <= 0.05 * 4320
