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


def _skeleton(text):
    """a report with every number replaced by '#': layout, names and wording only"""
    import re
    return re.sub(r" +", " ", re.sub(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", "#", text))      # (column widths: exact lines below)


@pytest.mark.skipif(not os.path.exists(H.REF_QS), reason="oracle/_ref/qs not built (needs /root/reference)")
def test_timer_report_has_the_reference_layout(tmp_path):
    """Cumulative_Report / Last_Cycle_Report (src/MC_Fast_Timer.cc:58-152): same headings, timer names, column formats and
    Figure Of Merit line as the reference binary prints (numbers differ: they are wall-clock times)."""
    import subprocess
    deck = decks.write_deck(decks.derive("CTS2_1", nx=4, ny=4, nz=4, lx=4, ly=4, lz=4, nParticles=640, nSteps=2, cycleTimers=1), str(tmp_path / "t.inp"))
    out = subprocess.run([H.REF_QS, "-i", deck], check=True, stdout=subprocess.PIPE, text=True, env=dict(os.environ, OMP_NUM_THREADS="1")).stdout
    import re
    ref_cumulative = out[re.search(r"\nTimer +Cumulative", out).start():]
    ref_cumulative = ref_cumulative[:ref_cumulative.index("[Num Segments / Cycle Tracking Time]") + len("[Num Segments / Cycle Tracking Time]") + 1]
    first = re.search(r"\nTimer +Last Cycle", out).start()
    ref_last = out[first:out.index("\n", out.index("cycleFinalize", first)) + 1]

    mc = host.MonteCarlo(["-i", deck])
    assert mc.get_int("cycleTimers") == 1
    dt = mc.get_double("dt")
    last_reports = []
    for _ in range(2):
        mc.cycle_init()
        r = H.oracle_track(mc.image, dt, mc.processing(), strict=False, threads=1)
        mc.set_tracking_result(r.census, r.balance, r.flux.sum())
        mc.timer_add("cycleTracking", 1234.0, 1)                 # the tracking section belongs to the caller here
        mc.timer_add("cycleTracking_Kernel", 1000.0, 3)
        mc.cycle_finalize()
        last_reports.append(mc.timer_report(last_cycle=True))
    ours = mc.timer_report()
    assert _skeleton(ours) == _skeleton(ref_cumulative), "\n--- got ---\n%s\n--- reference ---\n%s" % (ours, ref_cumulative)
    assert _skeleton(last_reports[0]) == _skeleton(ref_last)
    # the numbers: calls counted, last-cycle clocks cleared every cycle, FOM = segments / cycleTracking time
    us, calls = mc.timer("cycleTracking")
    assert (us, calls) == (2468.0, 2) and mc.timer("cycleInit")[1] == 2 and mc.timer("cycleFinalize")[1] == 2
    assert "cycleTracking                        2    2.468e+03    2.468e+03    2.468e+03    0.000e+00       100.00" in ours
    assert "cycleTracking                        2    1.234e+03" in last_reports[1]
    segs = float(mc.cumulative_balance()[host.BAL["num_segments"]])
    assert ours.endswith("%-25s %12.3e %-25s\n" % ("Figure Of Merit", segs / 2468e-6, "[Num Segments / Cycle Tracking Time]"))
