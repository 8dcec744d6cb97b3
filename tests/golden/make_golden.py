#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs oracle/_ref/qs_dump (the reference's own sources from /root/reference/src compiled by oracle/Makefile and
driven through its own cycleInit / cycleTracking / cycleFinalize, see oracle/ref_dump.cc) on small variants of the
Examples decks and stores what the reference held in memory:

    <case>.npz   problem image (nodes, planes, adjacency, cell state, nuclear data) and, per cycle, the
                 processing vault after cycleInit (= tracking input), the processed vault after
                 cycleTracking (= census), the 13 balance counters, the scalar flux (cycle 0) and its sum

The fixtures travel with the repository (the GPU box has no /root/reference); this script only runs in the build
container:   python tests/golden/make_golden.py            (rebuilds every fixture)

The reference is run with OMP_NUM_THREADS=1 so the flux summation order is the vault order.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from quicksilver_b200 import decks  # noqa: E402

P2L = 4.0 / 11.0

# name: (deck, Simulation overrides, cycles run, cycles whose vaults are stored)
CASES = {
    "cts2_tiny": ("CTS2_1", dict(nx=5, ny=5, nz=5, lx=5, ly=5, lz=5, nParticles=1250, nSteps=3), 3, 2),
    "p1_tiny": ("Coral2_P1_1", dict(nx=4, ny=4, nz=4, lx=4, ly=4, lz=4, nParticles=2560, nSteps=3), 3, 2),
    "p2_tiny": ("Coral2_P2_1", dict(nx=4, ny=4, nz=4, lx=P2L, ly=P2L, lz=P2L, nParticles=2560, nSteps=3), 3, 2),
    "allabsorb_voronoi": ("AllAbsorb", dict(nx=6, ny=6, nz=6, lx=6, ly=6, lz=6, nParticles=2160, nSteps=3), 3, 2),
    "allescape_voronoi": ("AllEscape", dict(nx=6, ny=6, nz=6, lx=6, ly=6, lz=6, nParticles=2160, nSteps=3), 3, 2),
    "nofission_octant": ("NoFission", dict(nx=5, ny=5, nz=5, lx=50, ly=50, lz=50, nParticles=2000, nSteps=3), 3, 2),
    "scattering_octant": ("AllScattering", dict(nx=5, ny=5, nz=5, lx=50, ly=50, lz=50, nParticles=2000, nSteps=2), 2, 1),
    "nonflat_two_materials": ("NonFlatXC", dict(nx=5, ny=5, nz=5, lx=50, ly=50, lz=50, nParticles=1500, nSteps=3, dt=5e-10), 3, 2),
    # total cross section 1e-80: no collision ever, particles stream and reflect until census (Examples/NoCollisions)
    "nocollisions_voronoi": ("NoCollisions", dict(nx=5, ny=5, nz=5, lx=50, ly=50, lz=50, nParticles=2000, nSteps=3), 3, 2),
    "homogeneous_v7": ("Homogeneous_v7", dict(nx=5, ny=5, nz=5, lx=100, ly=100, lz=100, xDom=1, yDom=1, zDom=1,
                                              nParticles=1250, nSteps=2), 2, 1),
}


def deck_of(case):
    name, over, _, _ = CASES[case]
    return decks.derive(name, over)


def main():
    import helpers as H
    if not os.path.exists(H.REF_DUMP):
        raise SystemExit("oracle/_ref/qs_dump missing: make -C oracle ref (needs /root/reference)")
    only = sys.argv[1:]
    for case, (name, over, cycles, stored) in CASES.items():
        if only and case not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            deck = decks.write_deck(deck_of(case), os.path.join(tmp, case + ".inp"))
            H.run_reference_dump(["-i", deck], os.path.join(tmp, "dump"), particle_cycles=stored, threads=1)
            out = {}
            for k, v in H.read_qsd(os.path.join(tmp, "dump", "problem.qsd")).items():
                out["problem/" + k] = v
            for c in range(cycles):
                for k, v in H.read_qsd(os.path.join(tmp, "dump", "cycle_%03d.qsd" % c)).items():
                    if k.endswith("_flux") and c > 0:
                        continue
                    out["cycle%d/%s" % (c, k)] = v
            # the cycle table the reference binary itself prints, for the same deck
            table = subprocess.run([H.REF_QS, "-i", deck], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                                   env=dict(os.environ, OMP_NUM_THREADS="1"), check=True).stdout
            rows = [line.split() for line in table.splitlines()]
            rows = [r for r in rows if len(r) == 17 and r[0].isdigit()]
            out["table/ints"] = np.array([[int(v) for v in r[1:13]] for r in rows], dtype=np.int64)
            out["table/flux_text"] = np.array([r[13] for r in rows])
            path = os.path.join(HERE, case + ".npz")
            np.savez_compressed(path, **out)
            print("%-24s %8.1f KB  %d cycles, input %d -> census %d" % (
                case, os.path.getsize(path) / 1e3, cycles, len(out["cycle0/tracking_input"]), len(out["cycle0/census"])))


# the reference's closing report (coralBenchmarkCorrectness, src/CoralBenchmark.cc) as its binary prints it
REPORT_CASES = [("Coral2_P1_1", dict(nSteps=10)), ("Coral2_P2_1", dict(nSteps=10)), ("Coral2_P1_1:short", dict(nSteps=2, nParticles=20480))]


def coral_reports():
    import json
    import helpers as H
    out = {}
    for name, over in REPORT_CASES:
        with tempfile.TemporaryDirectory() as tmp:
            deck = decks.write_deck(decks.derive(name.split(":")[0], over), os.path.join(tmp, "d.inp"))
            text = subprocess.run([H.REF_QS, "-i", deck], env=dict(os.environ, OMP_NUM_THREADS="8"), stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True, check=True).stdout
        lines = text.splitlines()
        first = [k for k, l in enumerate(lines) if l.startswith("Testing Ratios")][0] - 1
        last = [k for k, l in enumerate(lines) if "Fluence" in l and ("PASS" in l or "FAIL" in l)][0]
        while last + 1 < len(lines) and lines[last + 1].startswith("\t"):
            last += 1
        out[name] = {"overrides": over, "report": "\n".join(lines[first:last + 1]) + "\n"}
    with open(os.path.join(HERE, "coral_reports.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("coral_reports.json: %d reports" % len(out))


if __name__ == "__main__":
    if sys.argv[1:] == ["reports"]:
        coral_reports()
    else:
        main()
        if not sys.argv[1:]:
            coral_reports()
