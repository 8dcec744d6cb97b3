"""Input decks of the reference's Examples/ problems, generated rather than copied.

Each deck is a dict of blocks in the reference's deck grammar (parsed by
quicksilver_b200/csrc/host/Parameters.cc); `deck_text` renders it to the text the `-i` option reads.
Values are those of /root/reference/Examples/<dir>/<file> (cited per deck); `derive` applies the
documented deviations (nSteps shortened, NonFlatXC's dt).
"""
import copy
import os


def _xs(name, D=0, E=1, nuBar=None):
    x = {"name": name, "A": 0, "B": 0, "C": 0, "D": D, "E": E}
    if nuBar is not None:
        x["nuBar"] = nuBar
    return x


def _brick(material, hi, lo=0):
    return {"material": material, "shape": "brick", "xMax": hi, "xMin": lo, "yMax": hi, "yMin": lo, "zMax": hi, "zMin": lo}


def _material(name, xs=("flat", "flat", "flat"), **kw):
    m = {"name": name}
    m.update(kw)
    m["absorptionCrossSection"], m["fissionCrossSection"], m["scatteringCrossSection"] = xs
    return m


def _sizes(n, l, particles, dom=1):
    return {"lx": l, "ly": l, "lz": l, "nParticles": particles, "nx": n, "ny": n, "nz": n, "xDom": dom, "yDom": dom, "zDom": dom}


def _benchmark_sim(dt, sizes, e_min, coral):
    sim = {"dt": dt, "fMax": 0.1, "boundaryCondition": "reflect", "loadBalance": 0, "cycleTimers": 0, "debugThreads": 0}
    sim.update(sizes)
    sim.update({"nSteps": 100, "seed": 1029384756, "eMax": 20, "eMin": e_min, "nGroups": 230, "lowWeightCutoff": 0.001})
    if coral:
        sim["coralBenchmark"] = coral
    return sim


# energy-dependent tables shared by CTS2 and CORAL2 problem 2
_AFS = [_xs("absorb", D=-0.2, E=2), _xs("fission", D=-0.2, E=2, nuBar=2), _xs("scatter", D=0, E=97)]
_AFS_NAMES = ("absorb", "fission", "scatter")


def _p1(sizes):      # Examples/CORAL2_Benchmark/Problem1/Coral2_P1_1.inp | Coral2_P1.inp
    mat = _material("sourceMaterial", mass=12.011, nIsotopes=20, nReactions=9, sourceRate=1e10, totalCrossSection=1.5,
                    absorptionCrossSectionRatio=0.04, fissionCrossSectionRatio=0.05, scatteringCrossSectionRatio=1)
    return {"Simulation": _benchmark_sim(2e-09, sizes, 1e-09, 1), "Geometry": [_brick("sourceMaterial", 10000)],
            "Material": [mat], "CrossSection": [_xs("flat", nuBar=1.6)]}


def _p2(sizes):      # Examples/CORAL2_Benchmark/Problem2/Coral2_P2_1.inp | Coral2_P2.inp
    mat = _material("sourceMaterial", _AFS_NAMES, mass=1.5, nIsotopes=10, nReactions=3, sourceRate=1e10,
                    totalCrossSection=16.75, absorptionCrossSectionRatio=10, fissionCrossSectionRatio=8,
                    scatteringCrossSectionRatio=82)
    return {"Simulation": _benchmark_sim(1e-08, sizes, 1e-08, 2), "Geometry": [_brick("sourceMaterial", 10000)],
            "Material": [mat], "CrossSection": copy.deepcopy(_AFS)}


def _cts2(sizes):    # Examples/CTS2_Benchmark/CTS2_1.inp | CTS2.inp
    mat = _material("sourceMaterial", _AFS_NAMES, mass=1.5, nIsotopes=20, nReactions=9, sourceRate=1e10,
                    totalCrossSection=1.5227, absorptionCrossSectionRatio=10, fissionCrossSectionRatio=8,
                    scatteringCrossSectionRatio=82)
    return {"Simulation": _benchmark_sim(1.1e-07, sizes, 1e-08, 2), "Geometry": [_brick("sourceMaterial", 10000)],
            "Material": [mat], "CrossSection": copy.deepcopy(_AFS)}


def _limit_case(bc, total_xs, ratios, n_particles=9999):   # Examples/AllAbsorb/allAbsorb.inp, AllEscape/allEscape.inp
    sim = {"dt": 1e6, "fMax": 0.1, "loadBalance": 1, "lx": 10, "ly": 10, "lz": 10, "nParticles": n_particles,
           "nSteps": 20, "nx": 10, "ny": 10, "nz": 10, "seed": 1029384761, "xDom": 0, "yDom": 0, "zDom": 0,
           "eMax": 1, "eMin": 0.99999, "nGroups": 1}
    if bc:
        sim["boundaryCondition"] = bc
    a, f, s = ratios
    mats = [_material(n, nIsotopes=10, nReactions=9, sourceRate=r, totalCrossSection=total_xs,
                      absorptionCrossSectionRatio=a, fissionCrossSectionRatio=f, scatteringCrossSectionRatio=s)
            for n, r in (("boxMaterial", 0), ("sourceMaterial", 1e-2))]
    return {"Simulation": sim, "Geometry": [_brick("boxMaterial", 10), _brick("sourceMaterial", 1)],
            "Material": mats, "CrossSection": [_xs("flat", nuBar=2.4)]}


def _homogeneous(dt, total_xs, ratios, mass=None):   # Examples/Homogeneous/homogeneousProblem_v5_ts.inp, _v7_ts.inp
    sim = {"dt": dt, "fMax": 0.1, "boundaryCondition": "reflect", "loadBalance": 0, "cycleTimers": 0, "debugThreads": 0,
           "nSteps": 10, "seed": 1029384756, "eMax": 20, "eMin": 1e-09, "nGroups": 230, "lowWeightCutoff": 0.001}
    a, f, s = ratios
    kw = {"mass": mass} if mass is not None else {}
    mat = _material("sourceMaterial", nIsotopes=10, nReactions=9, sourceRate=1e10, totalCrossSection=total_xs,
                    absorptionCrossSectionRatio=a, fissionCrossSectionRatio=f, scatteringCrossSectionRatio=s, **kw)
    return {"Simulation": sim, "Geometry": [_brick("sourceMaterial", 1000)], "Material": [mat],
            "CrossSection": [_xs("flat", nuBar=1.0)]}


def _nonflat():      # Examples/NonFlatXC/NonFlatXC.inp (as shipped: explosively supercritical)
    sim = {"dt": 1e-08, "fMax": 0.1, "boundaryCondition": "reflect", "loadBalance": 0, "cycleTimers": 0, "debugThreads": 0,
           "lx": 100, "ly": 100, "lz": 100, "nParticles": 1000000, "batchSize": 0, "nBatches": 10, "nSteps": 10,
           "nx": 10, "ny": 10, "nz": 10, "seed": 1029384756, "xDom": 0, "yDom": 0, "zDom": 0, "eMax": 20, "eMin": 1e-08,
           "nGroups": 230, "lowWeightCutoff": 0.001, "bTally": 1, "fTally": 1, "cTally": 1, "coralBenchmark": 0}
    mats = [_material("sourceMaterial", ("absorb", "fission", "scatter"), mass=1000.0, nIsotopes=10, nReactions=9,
                      sourceRate=1e10, totalCrossSection=6, absorptionCrossSectionRatio=6e-3, fissionCrossSectionRatio=1,
                      scatteringCrossSectionRatio=5),
            _material("flatMaterial", nIsotopes=20, nReactions=9, sourceRate=1e10, totalCrossSection=1,
                      absorptionCrossSectionRatio=1, fissionCrossSectionRatio=1, scatteringCrossSectionRatio=1)]
    # the shipped "absorb" block gives D twice (-0.8446 then -0.5243); the last one wins
    xs = [_xs("flat", nuBar=2.4), _xs("absorb", D=-0.5243, E=-2.22), _xs("fission", D=-0.342, E=0, nuBar=2.4),
          _xs("scatter", D=0, E=0.7)]
    return {"Simulation": sim, "Geometry": [_brick("sourceMaterial", 100)], "Material": mats, "CrossSection": xs}


def _no_fission():   # Examples/NoFission/noFission.inp
    sim = {"dt": 1e-08, "fMax": 0.1, "boundaryCondition": "octant", "loadBalance": 1, "cycleTimers": 0, "debugThreads": 0,
           "lx": 100, "ly": 100, "lz": 100, "nParticles": 10000000, "nSteps": 10, "nx": 10, "ny": 10, "nz": 10,
           "seed": 1029384756, "xDom": 0, "yDom": 0, "zDom": 0, "eMax": 20, "eMin": 1e-9, "nGroups": 230}
    mat = _material("sourceMaterial", nIsotopes=10, nReactions=9, sourceRate=1e10, totalCrossSection=0.198,
                    absorptionCrossSectionRatio=0.494949495, fissionCrossSectionRatio=0, scatteringCrossSectionRatio=0.505050505)
    return {"Simulation": sim, "Geometry": [_brick("sourceMaterial", 100)], "Material": [mat],
            "CrossSection": [_xs("flat", nuBar=2.4)]}


def _no_collisions():   # Examples/NoCollisions/no.collisions.inp: total cross section 1e-80, particles only stream
    sim = {"dt": 1e-08, "fMax": 0.1, "loadBalance": 1, "lx": 100, "ly": 100, "lz": 100, "nParticles": 1000000, "nSteps": 10,
           "nx": 10, "ny": 10, "nz": 10, "seed": 1029384756, "xDom": 0, "yDom": 0, "zDom": 0, "eMax": 1.000001, "eMin": 1.000000,
           "nGroups": 230}
    mats = [_material("boxMaterial", nIsotopes=10, nReactions=9, sourceRate=0, totalCrossSection=1e-80,
                      absorptionCrossSectionRatio=1, fissionCrossSectionRatio=0, scatteringCrossSectionRatio=1),
            _material("sourceMaterial", nIsotopes=10, nReactions=9, sourceRate=1e10, totalCrossSection=1e-80,
                      absorptionCrossSectionRatio=1, fissionCrossSectionRatio=1, scatteringCrossSectionRatio=1)]
    return {"Simulation": sim, "Geometry": [_brick("boxMaterial", 100), _brick("sourceMaterial", 10)], "Material": mats,
            "CrossSection": [_xs("flat", nuBar=2.4)]}


DECKS = {
    "AllAbsorb": _limit_case(None, 1e10, (1, 0, 0)),
    "AllEscape": _limit_case("escape", 1e-20, (0, 0, 1)),
    "Coral2_P1_1": _p1(_sizes(16, 16, 163840)),
    "Coral2_P1": _p1({}),
    "Coral2_P2_1": _p2(_sizes(11, 1, 53240)),
    "Coral2_P2": _p2({}),
    "CTS2_1": _cts2(_sizes(16, 16, 40960)),
    "CTS2": _cts2({}),
    "Homogeneous_v5": _homogeneous(1e-08, 10, (0.04, 0.05, 1)),
    "Homogeneous_v7": _homogeneous(1e-06, 0.1, (0.1086, 0.0969, 0.7946), mass=12.011),
    "NonFlatXC": _nonflat(),
    "NoFission": _no_fission(),
    "NoCollisions": _no_collisions(),
}
# scattering-only variant of NoFission (Examples/AllScattering/scatteringOnly.inp)
DECKS["AllScattering"] = copy.deepcopy(DECKS["NoFission"])
DECKS["AllScattering"]["Material"][0].update(totalCrossSection=0.1, absorptionCrossSectionRatio=0, scatteringCrossSectionRatio=1)


def derive(name, simulation=None, **other):
    """A copy of DECKS[name] with Simulation keys overridden (e.g. nSteps=10)."""
    deck = copy.deepcopy(DECKS[name])
    deck["Simulation"].update(simulation or {})
    deck["Simulation"].update(other)
    return deck


def _fmt(v):
    return repr(v) if isinstance(v, float) else str(v)


def deck_text(deck):
    out = ["Simulation:"]
    out += ["   %s: %s" % (k, _fmt(v)) for k, v in deck["Simulation"].items()]
    for block in ("Geometry", "Material", "CrossSection"):
        for item in deck.get(block, []):
            out += ["", block + ":"]
            out += ["   %s: %s" % (k, _fmt(v)) for k, v in item.items()]
    return "\n".join(out) + "\n"


def write_deck(deck, path):
    if isinstance(deck, str):
        deck = DECKS[deck]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write(deck_text(deck))
    return path
