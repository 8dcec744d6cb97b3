"""GPU parity tests of the device-resident cycle: cycleInit on the device -- MC_SourceNow +
PopulationControl + RouletteLowWeightParticles (src/main.cc:96-121) in one kernel, population never leaving HBM --
followed by the tracking kernel, against an ALL-CPU chain on the same deck: the oracle's cycleInit (oracle/qs_oracle.c:
qso_cycle_init, pinned against the reference's vault dumps by tests/test_oracle_golden.py) + the oracle's tracking, strict-math mode.

Bit-exact bar: every integer balance column of every cycle (start, source, rr, split and the tracking tallies), and
every census record byte for byte -- a census record is the end state of a history, so a single differing bit in a
particle the device sourced, split or re-weighted would show up in it.  Scalar flux: 1e-12 relative (atomic order)."""
import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import decks, device, host
from quicksilver_b200._capi import BAL

pytestmark = pytest.mark.gpu

FLUX_RTOL = 1e-11

CASES = {
    # name: (deck, overrides, cycles)                                      what cycleInit does on it
    "cts2_split": ("CTS2_1", dict(nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=5120, nSteps=4), 4),           # split, factor ~4
    "cts2_low_weight": ("CTS2_1", dict(nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=5120, nSteps=4, lowWeightCutoff=0.7), 4),
    "p1_roulette": ("Coral2_P1_1", dict(nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=20480, nSteps=4), 4),    # factor < 1
    "p2_small": ("Coral2_P2_1", dict(nx=6, ny=6, nz=6, lx=0.5454545454545454, ly=0.5454545454545454, lz=0.5454545454545454,
                                     nParticles=8640, nSteps=3), 3),
    "allabsorb_4dom": ("AllAbsorb", dict(nSteps=4), 4),                                                           # 4 domains on one rank, loadBalance 1
    "nofission_octant": ("NoFission", dict(nParticles=20000, nSteps=3), 3),
    "nonflat_supercritical": ("NonFlatXC", dict(nParticles=20000, nSteps=3, dt=5e-10), 3),                        # population x4 per cycle
}


def cpu_chain(deck, cycles):
    """The all-CPU chain: the oracle's cycleInit (oracle/qs_oracle.c: qso_cycle_init, the reference's three stages one after
    the other) + the oracle's tracking, both in strict-math mode.  The host model supplies the global numbers
    (qsb_mc_source_plan) and keeps the books; its own cycleInit must produce the oracle's vault.
    Returns [(global row, flux sum, sorted census, flux)] per cycle."""
    mc = host.MonteCarlo(["-i", deck])
    mc.set_strict_math(True)
    dt, e_min, e_max, cutoff = (mc.get_double(k) for k in ("dt", "eMin", "eMax", "lowWeightCutoff"))
    out = []
    census = np.zeros(0, H.PARTICLE_DTYPE)
    for _ in range(cycles):
        off, tally, weight, factor = mc.source_plan(len(census))
        vault, n_source, n_rr, n_split = H.oracle_cycle_init(mc.image, census, off, tally, weight, e_min, e_max, dt, factor, cutoff, strict=True)
        mc.cycle_init()
        assert H.sort_particles(mc.processing()).tobytes() == H.sort_particles(vault).tobytes()
        want = H.oracle_track(mc.image, dt, vault, strict=True, threads=1, want_flux=True)
        mc.set_tracking_result(want.census, want.balance, want.flux.sum())
        row, flux = mc.cycle_finalize()
        assert (int(row[BAL["source"]]), int(row[BAL["rr"]]), int(row[BAL["split"]])) == (n_source, n_rr, n_split)
        out.append((row.copy(), flux, H.sort_particles(want.census), want.flux))
        census = want.census
    mc.close()
    return out


@pytest.mark.parametrize("name", sorted(CASES))
def test_resident_cycles_match_cpu_chain_bit_for_bit(tmp_path, name):
    deck_name, over, cycles = CASES[name]
    deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / (name + ".inp")))
    want = cpu_chain(deck, cycles)
    mc = host.MonteCarlo(["-i", deck])
    ctx = device.DeviceContext(mc.image, mc.get_double("dt"), validation=True, particle_capacity=1 << 20)
    exercised = dict(rr=0, split=0, source=0)
    for cycle in range(cycles):
        res = mc.cycle_init_resident(ctx)
        stats = mc.cycle_tracking_resident(ctx)
        flux = ctx.get_scalar_flux()
        census = H.sort_particles(ctx.get_census())
        row, flux_sum = mc.cycle_finalize()
        w_row, w_flux_sum, w_census, w_flux = want[cycle]
        assert [int(v) for v in row] == [int(v) for v in w_row], "cycle %d balance row (%s)" % (cycle, ", ".join(H._capi.BAL_NAMES))
        assert int(res.n_processing) == int(row[BAL["start"]] + row[BAL["source"]] + row[BAL["split"]] - row[BAL["rr"]])
        assert len(census) == len(w_census) == int(stats.n_census)
        for field in H.PARTICLE_DTYPE.names:
            assert np.array_equal(census[field], w_census[field]), "cycle %d census field %s differs" % (cycle, field)
        assert census.tobytes() == w_census.tobytes()
        assert np.allclose(flux, w_flux, rtol=FLUX_RTOL, atol=0.0)
        assert abs(flux_sum - w_flux_sum) <= 1e-11 * abs(w_flux_sum)
        for k in exercised:
            exercised[k] += int(row[BAL[k]])
    assert exercised["source"] > 0
    if name in ("cts2_split", "cts2_low_weight"):
        assert exercised["split"] > 0
    if name in ("cts2_low_weight", "p1_roulette", "nonflat_supercritical"):
        assert exercised["rr"] > 0
    ctx.close()
    mc.close()


def test_host_and_resident_cycles_can_be_mixed(tmp_path):
    """host cycle -> resident cycles (the processed vault moves to the device) -> census back -> host cycle: same table
    as the CPU chain; the cells' running source counts stay in step on both sides."""
    deck_name, over, _ = CASES["cts2_split"]
    deck = decks.write_deck(decks.derive(deck_name, dict(over, nSteps=5)), str(tmp_path / "mixed.inp"))
    want = cpu_chain(deck, 5)
    mc = host.MonteCarlo(["-i", deck])
    mc.set_strict_math(True)
    ctx = device.DeviceContext(mc.image, mc.get_double("dt"), validation=True, particle_capacity=1 << 20)
    rows = []
    for cycle in range(5):
        if cycle in (0, 4):                       # host-staged cycles: host cycleInit + the drop-in tracking call
            if cycle == 4:
                with pytest.raises(host.QsbError):
                    mc.cycle_init()               # the census is still on the device: refused, not silently empty
                mc.census_to_host(ctx)
            mc.cycle_init()
            mc.cycle_tracking(ctx)
        else:
            mc.cycle_init_resident(ctx)
            mc.cycle_tracking_resident(ctx)
        row, _ = mc.cycle_finalize()
        rows.append([int(v) for v in row])
    assert rows == [[int(v) for v in w[0]] for w in want]
    census = H.sort_particles(mc.processed())
    assert census.tobytes() == want[4][2].tobytes()
    ctx.close()
    mc.close()


def test_resident_cycle_reports_vault_overflow(tmp_path):
    deck_name, over, _ = CASES["cts2_split"]
    deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / "overflow.inp"))
    mc = host.MonteCarlo(["-i", deck])
    ctx = device.DeviceContext(mc.image, mc.get_double("dt"), particle_capacity=256)      # the cycle sources 512 and splits them x10
    with pytest.raises(host.QsbError) as err:
        mc.cycle_init_resident(ctx)
    assert err.value.code == -4
    ctx.close()
    mc.close()


def test_resident_init_refuses_a_streamed_census(tmp_path):
    """after the host-buffer drop-in call the census lives in the host model, not in the device vault: the device-level
    call must say so instead of carrying over stale records"""
    deck_name, over, _ = CASES["cts2_split"]
    deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / "streamed.inp"))
    mc = host.MonteCarlo(["-i", deck])
    ctx = device.DeviceContext(mc.image, mc.get_double("dt"), particle_capacity=1 << 18)
    mc.cycle_init()
    mc.cycle_tracking(ctx)
    mc.cycle_finalize()
    n_cells = mc.image.n_cells
    with pytest.raises(host.QsbError) as err:
        ctx.cycle_init_resident(1, np.zeros(n_cells + 1, np.int32), np.zeros(n_cells, np.uint64), 1.0, 1e-9, 20.0)
    assert err.value.code == -5
    # the host-model call repairs it by itself: it hands the processed vault back to the device first
    res = mc.cycle_init_resident(ctx)
    assert int(res.n_start) > 0
    ctx.close()
    mc.close()


def _reference_groups(edges, energy):
    """NuclearData::getEnergyGroup (src/NuclearData.cc:208-227) for an array of energies"""
    n = len(edges)
    g = np.clip(np.searchsorted(edges, energy, side="right") - 1, 0, n - 2)
    g = np.where(energy <= edges[0], 0, g)
    return np.where(energy > edges[-1], n - 1, g)


def test_energy_spectrum_of_a_resident_census(tmp_path):
    """EnergySpectrum (src/EnergySpectrum.cc:12-35) when the census never comes to the host: one histogram kernel over the
    census energies per cycle; checked against the same census downloaded and binned with the reference's group search,
    and against the all-CPU chain's spectrum file."""
    deck_name, over, cycles = CASES["cts2_split"]
    deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / "spectrum.inp"))
    cpu = host.MonteCarlo(["-i", deck, "-e", str(tmp_path / "cpu")])
    cpu.set_strict_math(True)
    mc = host.MonteCarlo(["-i", deck, "-e", str(tmp_path / "gpu")])
    dt = mc.get_double("dt")
    ctx = device.DeviceContext(mc.image, dt, validation=True, particle_capacity=1 << 20)
    edges = mc.image.array("energies").copy()
    running = np.zeros(len(edges), dtype=np.uint64)
    for _ in range(cycles):
        cpu.cycle_init()
        want = H.oracle_track(cpu.image, dt, cpu.processing(), strict=True, threads=1, want_flux=True)
        cpu.set_tracking_result(want.census, want.balance, want.flux.sum())
        cpu.cycle_finalize()
        mc.cycle_init_resident(ctx)
        mc.cycle_tracking_resident(ctx)
        census = ctx.get_census()
        hist = np.bincount(_reference_groups(edges, census["kinetic_energy"]), minlength=len(edges)).astype(np.uint64)
        assert np.array_equal(ctx.census_energy_spectrum(), hist)
        running += hist
        mc.cycle_finalize()
        assert np.array_equal(mc.energy_spectrum(), running)
    assert int(running.sum()) > 0 and np.array_equal(mc.energy_spectrum(), cpu.energy_spectrum())
    mc.write_energy_spectrum(), cpu.write_energy_spectrum()
    assert (tmp_path / "gpu.dat").read_text() == (tmp_path / "cpu.dat").read_text()
    # a streamed census (136-byte records in device memory) is binned through the same kernel with a record stride
    mc.census_to_host(ctx)
    mc.cycle_init()
    mc.cycle_tracking(ctx)
    census = mc.processed()
    hist = np.bincount(_reference_groups(edges, census["kinetic_energy"]), minlength=len(edges)).astype(np.uint64)
    assert np.array_equal(ctx.census_energy_spectrum(), hist)
    ctx.close()
    mc.close()
    cpu.close()
