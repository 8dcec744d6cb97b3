"""Pins the CPU oracle (oracle/qs_oracle.c) and the host model (quicksilver_b200/csrc/host) against the
UNMODIFIED reference: fixtures under tests/golden/*.npz were dumped from the reference's own memory by
oracle/_ref/qs_dump (tests/golden/make_golden.py).  No GPU needed.

Per fixture and cycle:
  * the host model's flattened problem image equals the reference's mesh / nuclear data bit for bit,
  * the host model's cycleInit (source, population control, roulette) produces the reference's processing
    vault byte for byte,
  * the oracle's tracking of that vault gives the reference's census records byte for byte (sorted by
    identifier: vault order carries no physics), the integer balance counters exactly, and the scalar flux
    to 1e-12 relative (floating-point summation order over secondaries is the only difference),
  * cycleFinalize reproduces the row the reference binary prints.
"""
import os
import sys

import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import decks, host
from quicksilver_b200._capi import BAL, PARTICLE_DTYPE

sys.path.insert(0, H.GOLDEN)
import make_golden  # noqa: E402

CASES = sorted(make_golden.CASES)
FLUX_RTOL = 1e-12
TRACKING_COUNTERS = ("absorb", "census", "escape", "collision", "fission", "produce", "scatter", "num_segments")


def _model(case, tmp_path):
    return host.MonteCarlo(["-i", H.golden_deck(case, tmp_path)])


@pytest.mark.parametrize("case", CASES)
def test_host_image_equals_reference_memory(case, tmp_path):
    g = H.golden_case(case)
    mc = _model(case, tmp_path)
    im = mc.image
    n_dom, n_groups, n_iso, n_mat, nx, ny, nz = (int(v) for v in g["problem/header"])
    assert (im.n_domains, im.n_groups, im.n_isotopes, im.n_materials) == (n_dom, n_groups, n_iso, n_mat)
    assert (im.global_nx, im.global_ny, im.global_nz) == (nx, ny, nz)
    assert np.array_equal(im.array("energies"), g["problem/energies"])

    # mesh, domain by domain (flat cell = domain offset + local cell)
    off = im.array("domain_cell_offset")
    for d in range(n_dom):
        gid, n_cells, _, _ = (int(v) for v in g["problem/d%d_info" % d])
        assert int(im.array("domain_gid")[d]) == gid
        lo, hi = int(off[d]), int(off[d + 1])
        assert hi - lo == n_cells
        assert np.array_equal(im.array("nodes")[lo:hi].reshape(n_cells, 42), g["problem/d%d_nodes" % d])
        assert np.array_equal(im.array("planes")[lo:hi].reshape(n_cells, 96), g["problem/d%d_planes" % d])
        assert np.array_equal(im.array("cell_volume")[lo:hi], g["problem/d%d_volume" % d])
        assert np.array_equal(im.array("cell_material")[lo:hi], g["problem/d%d_material" % d])
        assert np.array_equal(im.array("cell_id")[lo:hi], g["problem/d%d_cell_id" % d])
        assert np.all(g["problem/d%d_density" % d] == 1.0)     # what lets xs_total be per material
        # facet -> points table is the constant of src/MC_Domain.cc:41-50 for every cell
        fpts = g["problem/d%d_fpts" % d].reshape(n_cells, 24, 3)
        assert np.array_equal(fpts, np.broadcast_to(np.array(FACET_POINTS), fpts.shape))
        # adjacency: [event, adj.domain, adj.cell, adj.facet, nbr index, nbr gid, foreman, current.facet]
        adj = g["problem/d%d_adj" % d].reshape(n_cells, 24, 8)
        assert np.array_equal(adj[:, :, 7], np.broadcast_to(np.arange(24), (n_cells, 24)))
        for face in range(6):
            sub = adj[:, 4 * face:4 * face + 4, :]
            assert np.all(sub[:, :, :3] == sub[:, :1, :3])      # the 4 facets of a face share event + neighbour
            event = sub[:, 0, 0]
            assert np.array_equal(im.array("face_event")[lo:hi, face], event)
            transit = event == 3
            want_flat = off[sub[:, 0, 1]] + sub[:, 0, 2]
            assert np.array_equal(im.array("face_adj_cell")[lo:hi, face][transit], want_flat[transit])
            assert np.array_equal(im.array("face_adj_domain")[lo:hi, face][transit], sub[:, 0, 1][transit])
            boundary = ~transit
            assert np.array_equal(im.array("face_adj_cell")[lo:hi, face][boundary], np.arange(lo, hi)[boundary])
            for k in range(4):
                f = 4 * face + k
                assert np.all(sub[:, k, 3][transit] == OPPOSING_FACET[f])
                assert np.all(sub[:, k, 3][boundary] == f)
        assert not np.any(adj[:, :, 0] == 4)                    # a single rank has no off-processor facets

    # nuclear data: every isotope of a material carries one reaction table; xs_react holds
    # atomFraction * density * sigma in the (isotope, reaction) scan order of src/CollisionEvent.cc:67-83
    sigma = g["problem/react_sigma"].reshape(-1, n_groups)
    rtype = g["problem/react_type"]
    iso_nreact = g["problem/iso_nreact"]
    iso_first = np.concatenate([[0], np.cumsum(iso_nreact)])
    iso_gid, iso_af = g["problem/mat_iso_gid"], g["problem/mat_iso_af"]
    mat_niso = g["problem/mat_niso"]
    assert np.array_equal(im.array("mat_n_isotopes"), mat_niso)
    assert np.array_equal(im.array("mat_mass"), g["problem/mat_mass"])
    k = 0
    for m in range(n_mat):
        column, types = [], []
        for _ in range(int(mat_niso[m])):
            iso, af = int(iso_gid[k]), float(iso_af[k])
            k += 1
            rows = sigma[iso_first[iso]:iso_first[iso + 1]]
            column.append(af * 1.0 * rows)
            types.append(rtype[iso_first[iso]:iso_first[iso + 1]])
        if not column:
            continue
        column = np.concatenate(column)                         # [n_iso*n_react][n_groups]
        types = np.concatenate(types)
        n = len(column)
        assert np.array_equal(im.array("xs_react")[m, :, :n], column.T)
        assert np.array_equal(im.array("mat_react_type")[m, :n], types)
        # weightedMacroscopicCrossSection: sum over isotopes of (af * density * sum over reactions of sigma)
        total = np.zeros(n_groups)
        kk = k - int(mat_niso[m])
        for j in range(int(mat_niso[m])):
            iso, af = int(iso_gid[kk + j]), float(iso_af[kk + j])
            s = np.zeros(n_groups)
            for row in sigma[iso_first[iso]:iso_first[iso + 1]]:
                s = s + row
            total = total + af * 1.0 * s
        assert np.array_equal(im.array("xs_total")[m], total)


@pytest.mark.parametrize("case", CASES)
def test_cycle_init_oracle_tracking_and_finalize_equal_reference(case, tmp_path):
    g = H.golden_case(case)
    _, _, cycles, stored = make_golden.CASES[case]
    mc = _model(case, tmp_path)
    dt = mc.get_double("dt")
    off = mc.image.array("domain_cell_offset")
    for c in range(cycles):
        mc.cycle_init()
        vault = mc.processing()
        assert mc.get_double("source_particle_weight") == float(g["cycle%d/source_particle_weight" % c][0])
        if c < stored:
            want_in = H.particles_from_bytes(g["cycle%d/tracking_input" % c])
            assert len(vault) == len(want_in)
            assert H.sort_particles(vault).tobytes() == H.sort_particles(want_in).tobytes(), "cycle %d processing vault" % c
        got = H.oracle_track(mc.image, dt, vault, strict=False, threads=1)
        want_bal = g["cycle%d/balance" % c]
        for key in TRACKING_COUNTERS:
            assert int(got.balance[BAL[key]]) == int(want_bal[BAL[key]]), "cycle %d %s" % (c, key)
        if c < stored:
            want_census = H.sort_particles(H.particles_from_bytes(g["cycle%d/census" % c]))
            have = H.sort_particles(got.census)
            for field in PARTICLE_DTYPE.names:
                assert np.array_equal(have[field], want_census[field]), "cycle %d census field %s" % (c, field)
            assert have.tobytes() == want_census.tobytes()
        if c == 0:
            want_flux = np.concatenate([g["cycle0/d%d_flux" % d].reshape(-1, mc.image.n_groups) for d in range(mc.image.n_domains)])
            assert want_flux.shape[0] == int(off[-1])
            assert np.allclose(got.flux, want_flux, rtol=FLUX_RTOL, atol=0.0)
        want_sum = float(g["cycle%d/scalar_flux_sum" % c][0])
        assert abs(got.flux.sum() - want_sum) <= 1e-11 * abs(want_sum)
        assert got.n_retry_moves == 0 and got.n_forced_collisions == 0

        mc.set_tracking_result(got.census, got.balance, got.flux.sum())
        row, flux = mc.cycle_finalize()
        assert np.array_equal(row, want_bal), "cycle %d balance row" % c
        ints, _ = host.table_row(row, flux)
        assert ints == [int(v) for v in g["table/ints"][c]]
        assert "%.6e" % flux == str(g["table/flux_text"][c]) or abs(flux - float(g["table/flux_text"][c])) <= 2e-6 * abs(flux)


def test_known_answer_structure_of_limit_decks(tmp_path):
    """AllAbsorb: every collision is an absorption and nothing reaches census; AllEscape: no collisions,
    everything escapes (SURVEY.md 8c known-answer structure, Examples/AllAbsorb, Examples/AllEscape)."""
    for case, key in (("allabsorb_voronoi", "absorb"), ("allescape_voronoi", "escape")):
        mc = _model(case, tmp_path)
        mc.cycle_init()
        vault = mc.processing()
        r = H.oracle_track(mc.image, mc.get_double("dt"), vault, strict=False, threads=1)
        b = {k: int(r.balance[BAL[k]]) for k in BAL}
        assert b["census"] == 0 and b["scatter"] == 0 and b["fission"] == 0
        assert b[key] == len(vault)
        if key == "absorb":
            assert b["collision"] == b["absorb"] and b["escape"] == 0
        else:
            assert b["collision"] == 0 and b["absorb"] == 0


def test_no_collisions_deck_only_streams(tmp_path):
    """Examples/NoCollisions (total cross section 1e-80, reflecting box): no collision ever happens, every particle is
    followed facet by facet -- reflections included -- until census; the segment count is facet crossings + census."""
    mc = _model("nocollisions_voronoi", tmp_path)
    mc.cycle_init()
    vault = mc.processing()
    r = H.oracle_track(mc.image, mc.get_double("dt"), vault, strict=False, threads=1)
    b = {k: int(r.balance[BAL[k]]) for k in BAL}
    assert b["collision"] == b["absorb"] == b["scatter"] == b["fission"] == b["produce"] == b["escape"] == 0
    assert b["census"] == len(vault) == len(r.census) and b["num_segments"] > b["census"]      # facet crossings = segments - census
    assert r.n_forced_collisions == 0 and r.n_retry_moves == 0


def test_strict_math_build_of_the_oracle_tracks_the_libm_build(tmp_path):
    """The device validation kernels use the portable log/sin/cos of csrc/qs_strict_math.h so that CPU and GPU
    agree bit for bit; that variant of the oracle must stay statistically indistinguishable from the libm one
    (in practice: identical integer tallies on these sizes, flux within 1e-9)."""
    mc = _model("cts2_tiny", tmp_path)
    mc.cycle_init()
    vault = mc.processing()
    a = H.oracle_track(mc.image, mc.get_double("dt"), vault, strict=False, threads=1)
    b = H.oracle_track(mc.image, mc.get_double("dt"), vault, strict=True, threads=1)
    for key in TRACKING_COUNTERS:
        assert abs(int(a.balance[BAL[key]]) - int(b.balance[BAL[key]])) <= 0.002 * max(int(a.balance[BAL[key]]), 1000)
    assert abs(a.flux.sum() - b.flux.sum()) <= 1e-3 * a.flux.sum()


def test_threaded_oracle_equals_serial_oracle(tmp_path):
    mc = _model("p2_tiny", tmp_path)
    mc.cycle_init()
    vault = mc.processing()
    a = H.oracle_track(mc.image, mc.get_double("dt"), vault, strict=True, threads=1)
    b = H.oracle_track(mc.image, mc.get_double("dt"), vault, strict=True, threads=4)
    assert np.array_equal(a.balance, b.balance)
    assert H.sort_particles(a.census).tobytes() == H.sort_particles(b.census).tobytes()
    assert np.allclose(a.flux, b.flux, rtol=1e-12, atol=0)


@pytest.mark.parametrize("name,cycles", [("AllAbsorb", 20), ("Coral2_P2_1", 2)])
def test_full_size_cycle_tables_of_baseline_md(name, cycles, tmp_path):
    """Full-size Examples decks through host model + oracle reproduce the cycle tables captured from the
    reference binary in BASELINE.md section 4 (tests/golden/balance_tables.json)."""
    from quicksilver_b200 import decks
    golden = H.golden_table(name)
    deck = decks.write_deck(decks.derive(name, nSteps=cycles), str(tmp_path / "full.inp"))
    mc = host.MonteCarlo(["-i", deck])
    dt = mc.get_double("dt")
    for c in range(cycles):
        mc.cycle_init()
        r = H.oracle_track(mc.image, dt, mc.processing(), strict=False, threads=os.cpu_count() or 1, want_flux=True)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        row, flux = mc.cycle_finalize()
        ints, _ = host.table_row(row, flux)
        assert ints == golden[c][0], "cycle %d" % c
        assert abs(flux - golden[c][1]) <= 1e-6 * abs(golden[c][1])


# src/MC_Domain.cc:41-50
FACET_POINTS = [[1, 3, 8], [3, 7, 8], [7, 5, 8], [5, 1, 8], [0, 4, 9], [4, 6, 9], [6, 2, 9], [2, 0, 9],
                [3, 2, 10], [2, 6, 10], [6, 7, 10], [7, 3, 10], [0, 1, 11], [1, 5, 11], [5, 4, 11], [4, 0, 11],
                [4, 5, 12], [5, 7, 12], [7, 6, 12], [6, 4, 12], [0, 2, 13], [2, 3, 13], [3, 1, 13], [1, 0, 13]]
OPPOSING_FACET = [7, 6, 5, 4, 3, 2, 1, 0, 12, 15, 14, 13, 8, 11, 10, 9, 20, 23, 22, 21, 16, 19, 18, 17]


@pytest.mark.skipif(not os.path.exists(H.REF_DUMP), reason="oracle/_ref/qs_dump not built (needs /root/reference)")
@pytest.mark.parametrize("deck_name,over", [
    ("CTS2_1", dict(nSteps=3)),                                                       # 40 960 particles: source x10 split, then splits x ~4
    ("Coral2_P1_1", dict(nSteps=3, nParticles=81920, nx=12, ny=12, nz=12, lx=12, ly=12, lz=12)),   # roulette (factor < 1) from cycle 1 on
])
def test_threaded_cycle_init_equals_the_reference_at_sizes_above_the_team_threshold(tmp_path, deck_name, over):
    """the host cycleInit runs on an OpenMP team above 32 768 particles; the committed fixtures are smaller, so this one
    dumps the reference's vaults live (oracle/_ref/qs_dump) and compares cycle by cycle"""
    deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / "deck.inp"))
    H.run_reference_dump(["-i", deck], str(tmp_path / "dump"), particle_cycles=3, threads=4)
    mc = host.MonteCarlo(["-i", deck])
    dt = mc.get_double("dt")
    exercised = dict(rr=0, split=0)
    for c in range(3):
        ref = H.read_qsd(str(tmp_path / "dump" / ("cycle_%03d.qsd" % c)))
        mc.cycle_init()
        vault = mc.processing()
        assert len(vault) > 32768
        want = H.particles_from_bytes(ref["tracking_input"])
        assert H.sort_particles(vault).tobytes() == H.sort_particles(want).tobytes(), "cycle %d processing vault" % c
        r = H.oracle_track(mc.image, dt, vault, strict=False, threads=os.cpu_count() or 1)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        row, _ = mc.cycle_finalize()
        assert np.array_equal(row, ref["balance"]), "cycle %d balance row" % c
        exercised["rr"] += int(row[BAL["rr"]]); exercised["split"] += int(row[BAL["split"]])
    assert exercised["split"] > 0 if deck_name == "CTS2_1" else exercised["rr"] > 0


_THREAD_SCRIPT = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from quicksilver_b200 import decks, host
mc = host.MonteCarlo(["-i", sys.argv[2]])
h = hashlib.sha256()
for cycle in range(4):
    mc.cycle_init()
    v = mc.processing()
    h.update(v.tobytes())                      # NOT sorted: the vault order itself must not depend on the team size
    # stand-in for tracking: cycle 1 keeps a third (forces splitting), cycle 2 triples the census (forces roulette)
    census = v[::3] if cycle == 1 else (np.concatenate([v, v, v]) if cycle == 2 else v)
    mc.set_tracking_result(census, np.zeros(13, np.uint64), 0.0)
    row, _ = mc.cycle_finalize()
    h.update(row.tobytes())
print(h.hexdigest())
"""


def test_cycle_init_vault_is_identical_for_any_team_size(tmp_path):
    import subprocess
    import sys
    deck = decks.write_deck(decks.derive("CTS2_1", nSteps=4, lowWeightCutoff=0.6), str(tmp_path / "deck.inp"))
    digests = set()
    for threads in ("1", "3", "8"):
        out = subprocess.run([sys.executable, "-c", _THREAD_SCRIPT, H.ROOT, deck], check=True, stdout=subprocess.PIPE, text=True,
                             env=dict(os.environ, QSB_HOST_THREADS=threads), timeout=600).stdout.strip()
        digests.add(out)
    assert len(digests) == 1, digests


@pytest.mark.parametrize("case", CASES)
def test_oracle_cycle_init_equals_reference_vaults(case, tmp_path):
    """oracle/qs_oracle.c: qso_cycle_init -- the stage-by-stage restatement of cycleInit (source, population control marching
    backwards with erase-swap, low-weight roulette) -- against the vaults dumped from the reference (libm mode, byte for
    byte after sorting by identifier) and against the host model's counts; the global numbers come from qsb_mc_source_plan."""
    g = H.golden_case(case)
    _, _, cycles, stored = make_golden.CASES[case]
    mc = _model(case, tmp_path)
    dt = mc.get_double("dt")
    census = np.zeros(0, PARTICLE_DTYPE)
    for c in range(stored):
        off, tally, weight, factor = mc.source_plan(len(census))
        vault, n_source, n_rr, n_split = H.oracle_cycle_init(mc.image, census, off, tally, weight, mc.get_double("eMin"), mc.get_double("eMax"), dt,
                                                             factor, mc.get_double("lowWeightCutoff"), strict=False)
        want_in = H.particles_from_bytes(g["cycle%d/tracking_input" % c])
        assert weight == float(g["cycle%d/source_particle_weight" % c][0])
        assert len(vault) == len(want_in)
        assert H.sort_particles(vault).tobytes() == H.sort_particles(want_in).tobytes(), "cycle %d processing vault" % c
        # the host model's own cycleInit: same vault, same counters
        mc.cycle_init()
        assert H.sort_particles(mc.processing()).tobytes() == H.sort_particles(vault).tobytes()
        r = H.oracle_track(mc.image, dt, vault, strict=False, threads=1)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        row, _ = mc.cycle_finalize()
        assert (int(row[BAL["source"]]), int(row[BAL["rr"]]), int(row[BAL["split"]])) == (n_source, n_rr, n_split)
        assert np.array_equal(row, g["cycle%d/balance" % c])
        census = r.census


def test_oracle_cycle_init_strict_mode_equals_host_strict_mode(tmp_path):
    """strict-math mode (portable log/sin/cos): the oracle's cycleInit and the host model's agree bit for bit on the decks of
    tests/test_gpu_resident.py -- the device cycle-init kernel is checked against the host model there, so it equals this
    oracle too"""
    from test_gpu_resident import CASES as RESIDENT_CASES
    for name, (deck_name, over, cycles) in sorted(RESIDENT_CASES.items()):
        deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / (name + ".inp")))
        mc = host.MonteCarlo(["-i", deck])
        mc.set_strict_math(True)
        dt = mc.get_double("dt")
        census = np.zeros(0, PARTICLE_DTYPE)
        for c in range(min(cycles, 3)):
            off, tally, weight, factor = mc.source_plan(len(census))
            vault, n_source, n_rr, n_split = H.oracle_cycle_init(mc.image, census, off, tally, weight, mc.get_double("eMin"), mc.get_double("eMax"),
                                                                 dt, factor, mc.get_double("lowWeightCutoff"), strict=True)
            mc.cycle_init()
            assert H.sort_particles(mc.processing()).tobytes() == H.sort_particles(vault).tobytes(), (name, c)
            r = H.oracle_track(mc.image, dt, vault, strict=True, threads=os.cpu_count() or 1)
            mc.set_tracking_result(r.census, r.balance, r.flux.sum())
            row, _ = mc.cycle_finalize()
            assert (int(row[BAL["source"]]), int(row[BAL["rr"]]), int(row[BAL["split"]])) == (n_source, n_rr, n_split), (name, c)
            census = r.census


# ---- the strict-math chain (the checker of the GPU validation build) against the REFERENCE BINARY's tables -----------------

HOMOGENEOUS_FLAGS = ["-X", "100", "-Y", "100", "-Z", "100", "-x", "16", "-y", "16", "-z", "16", "-I", "1", "-J", "1", "-K", "1", "-n", "40960"]
STRICT_TABLES = [
    ("CTS2_1", dict(nSteps=10), [], 10), ("Coral2_P1_1", dict(nSteps=10), [], 10), ("Coral2_P2_1", dict(nSteps=10), [], 10),
    ("Homogeneous_v5", dict(nSteps=10), HOMOGENEOUS_FLAGS, 10), ("Homogeneous_v7", dict(nSteps=10), HOMOGENEOUS_FLAGS, 10),
    ("NonFlatXC", dict(dt=5e-10, nParticles=100000, nSteps=5), [], 5),
]


@pytest.mark.parametrize("name,over,flags,cycles,strict_init",
                         [t + (False,) for t in STRICT_TABLES] + [t + (True,) for t in STRICT_TABLES[1:2] + STRICT_TABLES[4:]])
def test_strict_math_chain_reproduces_the_reference_binarys_tables(name, over, flags, cycles, strict_init, tmp_path):
    """The GPU validation kernels are checked bit for bit against the oracle in strict-math mode (portable log/sin/cos);
    the reference binary uses libm.  This closes the link between the two: at the Examples decks' LITERAL sizes the
    strict-math chain -- host cycleInit in either math mode + qso_track(strict=1) -- prints the very table the unmodified
    reference binary prints (tests/golden/balance_tables.json): all twelve integer columns of all
    cycles, flux to the 7 printed digits.  tests/test_gpu_literal.py then holds the device to the same tables."""
    golden = H.golden_table(name)
    deck = decks.write_deck(decks.derive(name, over), str(tmp_path / "d.inp"))
    mc = host.MonteCarlo(["-i", deck] + flags)
    mc.set_strict_math(strict_init)
    dt = mc.get_double("dt")
    for c in range(cycles):
        mc.cycle_init()
        r = H.oracle_track(mc.image, dt, mc.processing(), strict=True, threads=os.cpu_count() or 1, want_flux=True)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        row, flux = mc.cycle_finalize()
        ints, _ = host.table_row(row, flux)
        assert ints == golden[c][0], "%s cycle %d" % (name, c)
        assert abs(flux - golden[c][1]) <= 1e-6 * abs(golden[c][1])
    mc.close()


@pytest.mark.skipif(not os.path.exists(H.REF_DUMP_STRICT), reason="oracle/_ref/qs_dump_strict not built (make -C oracle ref; needs /root/reference)")
@pytest.mark.parametrize("deck_name,over,cycles", [
    ("CTS2_1", dict(nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=5120), 4),
    ("Coral2_P1_1", dict(nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=20480), 4),
    ("Coral2_P2_1", dict(nx=6, ny=6, nz=6, lx=6 / 11.0, ly=6 / 11.0, lz=6 / 11.0, nParticles=8640), 3),
    ("NonFlatXC", dict(nx=5, ny=5, nz=5, lx=50, ly=50, lz=50, nParticles=3000, dt=5e-10), 3),
    ("Homogeneous_v5", dict(nx=6, ny=6, nz=6, nParticles=4320), 3),
])
def test_strict_chain_equals_the_reference_in_strict_math_mode_record_for_record(tmp_path, deck_name, over, cycles):
    """The checker of the GPU validation build is the strict-math chain (host cycleInit in strict mode + qso_track(strict=1)).
    oracle/_ref/qs_dump_strict is the UNMODIFIED reference with log / sin / cos of its four tracking translation units mapped
    to the same portable functions (oracle/strict_math_map.h): cycle after cycle its processing vault, its census vault --
    every record, every byte -- and its balance row must equal the chain's.  With test_gpu_parity / test_gpu_literal (GPU ==
    chain) this ties the device to the reference itself, not only to its printed tables."""
    deck = decks.write_deck(decks.derive(deck_name, dict(over, nSteps=cycles)), str(tmp_path / "deck.inp"))
    H.run_reference_dump(["-i", deck], str(tmp_path / "dump"), particle_cycles=cycles, threads=1, exe=H.REF_DUMP_STRICT)
    mc = host.MonteCarlo(["-i", deck])
    mc.set_strict_math(True)
    dt = mc.get_double("dt")
    n_census = 0
    for c in range(cycles):
        ref = H.read_qsd(str(tmp_path / "dump" / ("cycle_%03d.qsd" % c)))
        mc.cycle_init()
        vault = mc.processing()
        want_in = H.particles_from_bytes(ref["tracking_input"])
        assert H.sort_particles(vault).tobytes() == H.sort_particles(want_in).tobytes(), "cycle %d processing vault" % c
        r = H.oracle_track(mc.image, dt, vault, strict=True, threads=1)
        want_census = H.particles_from_bytes(ref["census"])
        assert H.sort_particles(r.census).tobytes() == H.sort_particles(want_census).tobytes(), "cycle %d census records" % c
        n_census += len(want_census)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        row, _ = mc.cycle_finalize()
        assert np.array_equal(row, ref["balance"]), "cycle %d balance row" % c
    assert n_census > 0
    mc.close()
