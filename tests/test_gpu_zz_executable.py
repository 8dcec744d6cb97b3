"""The `qs_b200` executable (quicksilver_b200/csrc/host/qs_main.cc): the reference's command line, cycle table and closing
report over the C ABI.  Runs last in the GPU suite (file name), after the library-level parity tests."""
import os
import subprocess

import pytest

import helpers as H
from quicksilver_b200 import decks

EXE = os.path.join(H.ROOT, "quicksilver_b200", "qs_b200")


def cycle_rows(stdout):
    """rows of the reference's cycle table in a program's output: ([12 integer columns], scalar flux) per cycle; the layout
    is the reference's (src/Tallies.hh:60-76), so the same parser reads oracle/_ref/qs and qs_b200"""
    rows = []
    for line in stdout.splitlines():
        f = line.split()
        if len(f) == 17 and f[0].isdigit():
            rows.append(([int(v) for v in f[1:13]], float(f[13])))
    return rows


@pytest.mark.skipif(not os.path.exists(H.REF_QS), reason="oracle/_ref/qs not built (needs /root/reference)")
def test_table_parser_reads_the_reference_binarys_output(tmp_path):
    deck = decks.write_deck("AllAbsorb", str(tmp_path / "aa.inp"))
    out = subprocess.run([H.REF_QS, "-i", deck], check=True, stdout=subprocess.PIPE, text=True, env=dict(os.environ, OMP_NUM_THREADS="2")).stdout
    rows, golden = cycle_rows(out), H.golden_table("AllAbsorb")
    assert len(rows) == len(golden) == 20
    for (ints, flux), (g_ints, g_flux) in zip(rows, golden):
        assert ints == g_ints and abs(flux - g_flux) <= 1e-6 * abs(g_flux)


@pytest.mark.gpu
def test_executable_reproduces_the_reference_table(tmp_path):
    """validation kernels + host cycleInit (the default): the table the reference binary prints, digit for digit"""
    deck = decks.write_deck("AllAbsorb", str(tmp_path / "aa.inp"))
    res = subprocess.run([EXE, "-i", deck], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    rows, golden = cycle_rows(res.stdout), H.golden_table("AllAbsorb")
    assert len(rows) == 20
    for cycle, ((ints, flux), (g_ints, g_flux)) in enumerate(zip(rows, golden)):
        assert ints == g_ints, "cycle %d: %s != %s" % (cycle, ints, g_ints)
        assert abs(flux - g_flux) <= 2e-6 * abs(g_flux)               # both sides are 7-digit prints
    assert "Simulation:" in res.stdout and "Figure Of Merit" in res.stdout and "cycleTracking_Kernel" in res.stdout


@pytest.mark.gpu
def test_executable_with_resident_cycles(tmp_path):
    """QSB_RESIDENT=1: cycleInit on the device too; the table must equal the all-CPU chain (strict-math source + oracle),
    and the closing report carries the CORAL self checks"""
    from test_gpu_resident import cpu_chain
    from quicksilver_b200 import host
    deck = decks.write_deck(decks.derive("Coral2_P1_1", nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=20480, nSteps=3), str(tmp_path / "p1.inp"))
    res = subprocess.run([EXE, "-i", deck], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300,
                         env=dict(os.environ, QSB_RESIDENT="1"))
    assert res.returncode == 0, res.stderr[-2000:]
    rows = cycle_rows(res.stdout)
    want = cpu_chain(deck, 3)
    assert len(rows) == 3
    for cycle, (ints, flux) in enumerate(rows):
        w_ints, w_flux = host.table_row(want[cycle][0], want[cycle][1])
        assert ints == w_ints, "cycle %d: %s != %s" % (cycle, ints, w_ints)
        assert abs(flux - w_flux) <= 1e-5 * abs(w_flux)            # printed with 7 digits
    assert "Test for lost / unaccounted for particles in this simulation\nPASS:: No Particles Lost During Run" in res.stdout


def test_executable_fails_loudly_without_a_gpu(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("checks the behaviour on a machine WITHOUT a GPU")
    except ImportError:
        pass
    deck = decks.write_deck(decks.derive("CTS2_1", nx=4, ny=4, nz=4, lx=4, ly=4, lz=4, nParticles=640, nSteps=1), str(tmp_path / "d.inp"))
    res = subprocess.run([EXE, "-i", deck], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert res.returncode == 3 and "no CPU path" in res.stderr
    assert "Simulation:" in res.stdout and cycle_rows(res.stdout) == []      # the echo is printed, no cycle is faked
