"""The reference's closing report (coralBenchmarkCorrectness, src/CoralBenchmark.cc, and the figure-of-merit line of
src/MC_Fast_Timer.cc:97-104) reproduced by the host model: text compared with what the unmodified reference printed for the
same decks (tests/golden/coral_reports.json, captured from oracle/_ref/qs by tests/golden/make_golden.py).  Tracking is done
by the CPU oracle here; the per-cell fluence is accumulated the way Tallies::CycleFinalize does (src/Tallies.cc:90-121)."""
import json
import os

import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import decks, host

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "coral_reports.json")))


def _run(tmp_path, case):
    deck_name = case.split(":")[0]
    over = GOLDEN[case]["overrides"]
    deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / "deck.inp"))
    mc = host.MonteCarlo(["-i", deck])
    dt = mc.get_double("dt")
    fluence = np.zeros(mc.image.n_cells)
    for _ in range(mc.get_int("nSteps")):
        mc.cycle_init()
        r = H.oracle_track(mc.image, dt, mc.processing(), strict=False, threads=os.cpu_count() or 1)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        for g in range(mc.image.n_groups):            # fluence[cell] += flux[cell][group], group after group
            fluence += r.flux[:, g]
        mc.cycle_finalize()
    return mc, fluence


@pytest.mark.parametrize("case", ["Coral2_P1_1:short", "Coral2_P2_1"])
def test_coral_benchmark_report_matches_reference_text(tmp_path, case):
    mc, fluence = _run(tmp_path, case)
    text, passed = mc.coral_benchmark_report(fluence)
    want = GOLDEN[case]["report"]
    assert text == want, "\n--- got ---\n%s\n--- reference ---\n%s" % (text, want)
    assert passed == want.count("PASS::")
    fom = mc.format_figure_of_merit(2.0)
    segs = float(mc.cumulative_balance()[host.BAL["num_segments"]])
    assert fom == "%-25s %12.3e %-25s\n" % ("Figure Of Merit", segs / 2.0, "[Num Segments / Cycle Tracking Time]")


def test_report_is_empty_without_coral_benchmark(tmp_path):
    deck = decks.write_deck(decks.derive("AllAbsorb", nSteps=1), str(tmp_path / "c.inp"))
    mc = host.MonteCarlo(["-i", deck])
    text, passed = mc.coral_benchmark_report(np.ones(mc.image.n_cells))
    assert text == "" and passed == 0


def test_report_flags_broken_ratios_and_lost_particles(tmp_path):
    """FAIL branches: a cumulative balance that violates the expected reaction ratios / the conservation identity."""
    deck = decks.write_deck(decks.derive("Coral2_P1_1", nx=4, ny=4, nz=4, lx=4, ly=4, lz=4, nParticles=2560, nSteps=1), str(tmp_path / "p.inp"))
    mc = host.MonteCarlo(["-i", deck])
    mc.cycle_init()
    bal = np.zeros(13, np.uint64)
    bal[host.BAL["absorb"]], bal[host.BAL["scatter"]], bal[host.BAL["fission"]] = 400, 1000, 50      # fission ratio off by 10x
    bal[host.BAL["collision"]], bal[host.BAL["num_segments"]], bal[host.BAL["census"]] = 1450, 4000, 100
    mc.set_tracking_result(np.zeros(0, H.PARTICLE_DTYPE), bal, 1.0)
    mc.cycle_finalize()
    text, passed = mc.coral_benchmark_report(np.ones(mc.image.n_cells))
    assert "FAIL:: Absorption / Fission / Scatter Ratios NOT maintained with 1% tolerance" in text
    assert "Relative Absorb to Fission:  " in text and "Relative Scatter to Fission: " in text
    assert " FAIL:: Collision to Facet Crossing Ratio balanced NOT maintained within 1% tolerance" in text
    assert "FAIL:: Particles Were Lost During Run, test for done should have failed" in text
    assert "PASS:: Fluence is homogenous across cells with 6% tolerance" in text
    assert passed == 1
