"""GPU parity tests: the sm_100a tracking path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Validation build (--fmad=false + strict log/sin/cos): integer balance tallies and every
census record bit-exact; scalar flux to 1e-12 relative (atomic summation order is the only difference)."""
import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import decks, device, host
from quicksilver_b200._capi import BAL

pytestmark = pytest.mark.gpu

FLUX_RTOL = 1e-12

CASES = {
    # name: (deck, overrides, cycles)
    "allabsorb_4dom": ("AllAbsorb", dict(nSteps=3), 3),
    "allescape_4dom": ("AllEscape", dict(nSteps=2), 2),
    "cts2_small": ("CTS2_1", dict(nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=5120, nSteps=3), 3),
    "p1_small": ("Coral2_P1_1", dict(nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=20480, nSteps=3), 3),
    "p2_small": ("Coral2_P2_1", dict(nx=6, ny=6, nz=6, lx=0.5454545454545454, ly=0.5454545454545454, lz=0.5454545454545454,
                                     nParticles=8640, nSteps=2), 2),
    "nofission_octant": ("NoFission", dict(nParticles=20000, nSteps=2), 2),
    # the two opposite event mixes of SURVEY 8(d) input 2: 97 % collisions (v5) / 71 % facet crossings (v7)
    "homogeneous_v5": ("Homogeneous_v5", dict(nx=8, ny=8, nz=8, lx=25, ly=25, lz=25, nParticles=5120, nSteps=2), 2),
    "homogeneous_v7": ("Homogeneous_v7", dict(nx=8, ny=8, nz=8, lx=25, ly=25, lz=25, nParticles=5120, nSteps=2), 2),
    "nonflat_supercritical": ("NonFlatXC", dict(nParticles=20000, nSteps=2, dt=5e-10), 2),
}


KERNELS = ["event", "history"]      # the two tracking kernels (tracking_mode bit 0): same results bit for bit


def _run_case(tmp_path, name, validation=True, check_geometry=False, check_reactions=False, kernel="event"):
    deck_name, over, cycles = CASES[name]
    deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / (name + ".inp")))
    mc = host.MonteCarlo(["-i", deck])
    dt = mc.get_double("dt")
    ctx = device.DeviceContext(mc.image, dt, validation=validation, particle_capacity=1 << 20, check_geometry=check_geometry,
                               check_reactions=check_reactions, event=(kernel == "event"))
    out = []
    for _ in range(cycles):
        mc.cycle_init()
        vault = mc.processing()
        ctx.cycle_begin()
        ctx.put_particles(vault)
        stats = ctx.track()
        census, balance, flux = ctx.get_census(), ctx.get_balance(), ctx.get_scalar_flux()
        flux_sum = ctx.scalar_flux_sum()
        want = H.oracle_track(mc.image, dt, vault, strict=True, threads=1)
        stats.diag = ctx.diagnostics()
        out.append((census, balance, flux, flux_sum, want, stats))
        # carry the ORACLE's census forward so every cycle starts from identical inputs
        mc.set_tracking_result(want.census, want.balance, want.flux.sum())
        mc.cycle_finalize()
    ctx.close()
    return out


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", sorted(CASES))
def test_validation_build_matches_oracle_bit_for_bit(tmp_path, name, kernel):
    for cycle, (census, balance, flux, flux_sum, want, stats) in enumerate(_run_case(tmp_path, name, kernel=kernel)):
        assert np.array_equal(balance, want.balance), "cycle %d balance %s != %s" % (cycle, balance, want.balance)
        got, ref = H.sort_particles(census), H.sort_particles(want.census)
        assert len(got) == len(ref)
        for field in H.PARTICLE_DTYPE.names:
            assert np.array_equal(got[field], ref[field]), "cycle %d census field %s differs" % (cycle, field)
        assert got.tobytes() == ref.tobytes()
        assert np.allclose(flux, want.flux, rtol=FLUX_RTOL, atol=0.0)
        assert abs(flux_sum - want.flux.sum()) <= 1e-11 * abs(want.flux.sum())
        assert stats.n_processed >= len(census)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", ["cts2_small", "p2_small", "allescape_4dom", "nofission_octant"])
def test_filtered_geometry_agrees_with_full_search(tmp_path, name, kernel):
    """check mode: every segment evaluates both the filtered single-facet path and the reference's full
    24-facet search; they must agree on facet, distance and coordinate bit for bit, and the filtered path
    must carry nearly all segments."""
    for census, balance, flux, flux_sum, want, stats in _run_case(tmp_path, name, check_geometry=True, kernel=kernel):
        assert stats.diag["compact_geometry"] == 1
        assert stats.diag["geometry_mismatch"] == 0
        assert stats.diag["slow_geometry"] <= 0.01 * float(balance[BAL["num_segments"]]) + 10
        assert np.array_equal(balance, want.balance)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", ["cts2_small", "p1_small", "p2_small", "nonflat_supercritical", "allabsorb_4dom"])
def test_direct_reaction_selection_agrees_with_subtraction_chain(tmp_path, name, kernel):
    """check mode: every collision evaluates both the one-division filtered selection and the reference's
    subtraction chain (src/CollisionEvent.cc:59-83); they must pick the same (isotope, reaction) every time."""
    for census, balance, flux, flux_sum, want, stats in _run_case(tmp_path, name, check_reactions=True, kernel=kernel):
        assert stats.diag["geometry_mismatch"] == 0          # the counter is shared by both check modes
        assert np.array_equal(balance, want.balance)
        assert int(balance[BAL["collision"]]) > 0


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", ["cts2_small", "p1_small"])
def test_fast_build_within_statistical_tolerance(tmp_path, name, kernel):
    """fast build (FMA contraction, Newton-refined rcp / rsqrt, the portable log / sin / cos contracted to FMAs, box-exit geometry):
    histories differ from the reference's in the last bits, tallies must agree statistically (small decks here: 1 %; the k-sigma
    gate at bench size is tests/test_gpu_literal.py)."""
    for census, balance, flux, flux_sum, want, stats in _run_case(tmp_path, name, validation=False, kernel=kernel):
        for key in ("num_segments", "collision", "scatter", "absorb", "fission", "census"):
            a, b = float(balance[BAL[key]]), float(want.balance[BAL[key]])
            assert abs(a - b) <= 0.01 * max(b, 100.0), (key, a, b)
        assert abs(flux_sum - want.flux.sum()) <= 1e-2 * abs(want.flux.sum())


def test_drop_in_cycle_tracking_reproduces_reference_table(tmp_path):
    """qsb_mc_cycle_tracking (host vault in, census + tallies out) over whole cycles reproduces the
    reference's cycle table (AllAbsorb; integer columns exact, flux to 7 digits)."""
    golden = H.golden_table("AllAbsorb")
    deck = decks.write_deck("AllAbsorb", str(tmp_path / "aa.inp"))
    mc = host.MonteCarlo(["-i", deck])
    ctx = device.DeviceContext(mc.image, mc.get_double("dt"), validation=True, particle_capacity=1 << 18)
    for cycle in range(20):
        mc.cycle_init()
        mc.cycle_tracking(ctx)
        row, flux = mc.cycle_finalize()
        ints, _ = host.table_row(row, flux)
        assert ints == golden[cycle][0], "cycle %d: %s != %s" % (cycle, ints, golden[cycle][0])
        assert abs(flux - golden[cycle][1]) <= 1e-6 * abs(golden[cycle][1])
    ctx.close()


def test_capacity_overflow_is_reported(tmp_path):
    deck = decks.write_deck(decks.derive("CTS2_1", nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=5120, nSteps=1), str(tmp_path / "c.inp"))
    mc = host.MonteCarlo(["-i", deck])
    mc.cycle_init()
    vault = mc.processing()
    ctx = device.DeviceContext(mc.image, mc.get_double("dt"), particle_capacity=len(vault) + 16)
    ctx.cycle_begin()
    ctx.put_particles(vault)
    with pytest.raises(host.QsbError) as err:
        ctx.track()
    assert err.value.code == -4
    ctx.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name,census_capacity", [("cts2_small", None), ("nonflat_supercritical", 1000), ("allescape_4dom", None)])
def test_streamed_host_buffers_match_oracle(tmp_path, name, census_capacity, kernel):
    """qsb_track_host: host vault streamed in (pageable numpy memory -> bounce-buffer staging), census streamed
    back in record form; with a small census buffer the excess is fetched with qsb_get_census_range."""
    deck_name, over, _ = CASES[name]
    deck = decks.write_deck(decks.derive(deck_name, over), str(tmp_path / (name + ".inp")))
    mc = host.MonteCarlo(["-i", deck])
    dt = mc.get_double("dt")
    ctx = device.DeviceContext(mc.image, dt, validation=True, particle_capacity=1 << 20, event=(kernel == "event"))
    mc.cycle_init()
    vault = mc.processing()
    ctx.cycle_begin()
    census, stats = ctx.track_host(vault, census_capacity=census_capacity)
    want = H.oracle_track(mc.image, dt, vault, strict=True, threads=1)
    assert np.array_equal(ctx.get_balance(), want.balance)
    assert H.sort_particles(census).tobytes() == H.sort_particles(want.census).tobytes()
    assert np.allclose(ctx.get_scalar_flux(), want.flux, rtol=FLUX_RTOL, atol=0.0)
    # the record-form census is also what qsb_get_census returns for this cycle
    assert H.sort_particles(ctx.get_census()).tobytes() == H.sort_particles(want.census).tobytes()
    ctx.close()


def test_drop_in_streams_whole_census_chunks(tmp_path, monkeypatch):
    """Coral2_P1_1 at its literal size (163 840 particles) with 16 384-record streaming chunks: the vault and the census span many chunks, so the
    kernel's chunk-complete flags and the overlapped D2H copies are exercised; two cycles through the drop-in call
    (page-locked host vaults) against a twin host model driven by the oracle."""
    deck = decks.write_deck(decks.derive("Coral2_P1_1", nSteps=2), str(tmp_path / "p1.inp"))
    gpu, cpu = host.MonteCarlo(["-i", deck]), host.MonteCarlo(["-i", deck])
    dt = gpu.get_double("dt")
    monkeypatch.setenv("QSB_STREAM_CHUNK_LOG2", "14")        # read when the context is created
    ctx = device.DeviceContext(gpu.image, dt, validation=True, particle_capacity=1 << 21)
    import os
    for cycle in range(2):
        gpu.cycle_init(), cpu.cycle_init()
        assert gpu.processing().tobytes() == cpu.processing().tobytes()
        stats = gpu.cycle_tracking(ctx)
        want = H.oracle_track(cpu.image, dt, cpu.processing(), strict=True, threads=os.cpu_count() or 1)
        assert stats.n_census == len(want.census) > 2 * 65536
        assert H.sort_particles(gpu.processed()).tobytes() == H.sort_particles(want.census).tobytes()
        cpu.set_tracking_result(want.census, want.balance, want.flux.sum())
        row_g, flux_g = gpu.cycle_finalize()
        row_c, flux_c = cpu.cycle_finalize()
        assert np.array_equal(row_g, row_c)
        assert abs(flux_g - flux_c) <= 1e-11 * abs(flux_c)
        # the next cycle must start from identical vaults: carry the oracle's census order into the GPU twin as well
        gpu.set_tracking_result(want.census, np.zeros(13, np.uint64), 0.0)
    ctx.close()


def test_device_fluence_accumulates_the_cycle_flux(tmp_path):
    """qsb_fluence_accumulate (Fluence::compute, src/Tallies.cc:100-121, done on the device by the drop-in call for the CORAL
    decks) against the per-cell sums of the scalar flux the device itself reports, cycle after cycle; then the reference's
    report text from the host model with that fluence."""
    deck = decks.write_deck(decks.derive("Coral2_P1_1", nx=8, ny=8, nz=8, lx=8, ly=8, lz=8, nParticles=20480, nSteps=3), str(tmp_path / "p1.inp"))
    mc = host.MonteCarlo(["-i", deck])
    ctx = device.DeviceContext(mc.image, mc.get_double("dt"), validation=True, particle_capacity=1 << 20)
    want = np.zeros(mc.image.n_cells)
    for _ in range(3):
        mc.cycle_init()
        mc.cycle_tracking(ctx)                       # coralBenchmark deck: accumulates the fluence on the device
        want += ctx.get_scalar_flux().sum(axis=1)
        mc.cycle_finalize()
    got = ctx.get_fluence()
    assert np.allclose(got, want, rtol=1e-12, atol=0.0)
    text, passed = mc.coral_benchmark_report(got)
    assert text.count("PASS::") == passed and "Test Fluence for homogeneity across cells" in text
    ctx.close()
