"""The reference-side binding (oracle/QsbBinding.cc -> oracle/_ref/qs_qsb): the UNMODIFIED reference with its cycleTracking
replaced by libqsb.so's device calls -- SURVEY 8(b), "the drop-in, dropped in".

CPU: the binding compiles and links against the reference's own objects, flattens the reference's MonteCarlo into a
qsb_image that equals the host model's image of the same deck bit for bit (so everything pinned on the host model's
image holds for it), and fails loudly without a GPU.  GPU: `qs_qsb -i <deck>` prints the very cycle table the reference
binary prints (cycleInit / cycleFinalize / report are the reference's code, tracking is the validation kernels)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import decks, host

QS_QSB = os.path.join(H.ORACLE_DIR, "_ref", "qs_qsb")
needs_binding = pytest.mark.skipif(not os.path.exists(QS_QSB), reason="oracle/_ref/qs_qsb not built (make -C oracle qsb_binding; needs /root/reference)")

FILES = {   # image array -> (binding dump, dtype)
    "domain_cell_offset": ("domainCellOffset", "<i4"), "domain_gid": ("domainGid", "<i4"), "planes": ("planes", "<f8"),
    "nodes": ("nodes", "<f8"), "cell_gid": ("cellGid", "<i4"), "cell_material": ("cellMaterial", "<i4"),
    "cell_volume": ("cellVolume", "<f8"), "cell_id": ("cellId", "<u8"), "face_event": ("faceEvent", "u1"),
    "face_adj_cell": ("faceAdjCell", "<i4"), "face_adj_domain": ("faceAdjDomain", "<i4"), "face_nbr_rank": ("faceNbrRank", "<i4"),
    "energies": ("energies", "<f8"), "mat_n_isotopes": ("matNIso", "<i4"), "mat_n_reactions": ("matNReact", "<i4"),
    "mat_mass": ("matMass", "<f8"), "mat_nu_bar": ("matNuBar", "<f8"), "mat_react_type": ("matReactType", "u1"),
    "xs_total": ("xsTotal", "<f8"), "xs_react": ("xsReact", "<f8"), "mat_periodic": ("matPeriodic", "u1"),
}


@needs_binding
@pytest.mark.parametrize("name,over", [
    ("CTS2_1", dict(nx=6, ny=6, nz=6, lx=6, ly=6, lz=6, nParticles=1000, nSteps=1)),
    ("NonFlatXC", dict(nParticles=1000, nSteps=1, dt=5e-10)),
    ("AllAbsorb", dict(nSteps=1)),                       # 4 Voronoi domains on one rank
    ("NoFission", dict(nParticles=1000, nSteps=1)),      # octant boundary conditions
])
def test_binding_flattens_the_reference_into_the_host_models_image(tmp_path, name, over):
    deck = decks.write_deck(decks.derive(name, over), str(tmp_path / "d.inp"))
    dump = tmp_path / "image"
    dump.mkdir()
    res = subprocess.run([QS_QSB, "-i", deck], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300,
                         env=dict(os.environ, QSB_BINDING_IMAGE_DUMP=str(dump), OMP_NUM_THREADS="2"))
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except ImportError:
        have_gpu = False
    if not have_gpu:        # no CPU path: the binding stops at qsb_create, after the reference's own echo of the parameters
        assert res.returncode == 3 and "no CPU path" in res.stderr and "Simulation:" in res.stdout
    mc = host.MonteCarlo(["-i", deck])
    im = mc.image
    hdr = np.fromfile(dump / "header.bin", "<i4")
    assert list(hdr[:11]) == [im.n_domains, im.n_cells, im.n_groups, im.n_materials, im.n_isotopes, im.max_reactions_per_material,
                              im.my_rank, im.n_ranks, im.global_nx, im.global_ny, im.global_nz]
    for field, (fname, dtype) in FILES.items():
        want = np.ascontiguousarray(im.array(field))
        got = np.fromfile(dump / (fname + ".bin"), dtype)
        assert got.size == want.size, field
        assert got.tobytes() == want.tobytes(), "image array %s differs between the binding and the host model" % field
    mc.close()


@needs_binding
@pytest.mark.gpu
@pytest.mark.parametrize("name,cycles", [("CTS2_1", 10), ("Coral2_P1_1", 10), ("AllAbsorb", 20)])
def test_reference_with_the_binding_prints_the_reference_table(tmp_path, name, cycles):
    from test_gpu_zz_executable import cycle_rows
    deck = decks.write_deck(decks.derive(name, nSteps=cycles), str(tmp_path / "d.inp"))
    res = subprocess.run([QS_QSB, "-i", deck], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600,
                         env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
    assert res.returncode == 0, res.stderr[-2000:]
    rows, golden = cycle_rows(res.stdout), H.golden_table(name)
    assert len(rows) == cycles
    for cycle, ((ints, flux), (g_ints, g_flux)) in enumerate(zip(rows, golden)):
        assert ints == g_ints, "cycle %d: %s != %s" % (cycle, ints, g_ints)
        assert abs(flux - g_flux) <= 2e-6 * abs(g_flux)
    assert "Figure Of Merit" in res.stdout and "cycleTracking_Kernel" in res.stdout
    if name == "Coral2_P1_1":       # the reference's own end-of-run self checks, fed by tallies that came from the device:
        # the same verdicts, word for word, as the unmodified reference binary prints for this deck (at 10 cycles of the
        # 16^3 deck that includes its "FAIL:: Fluence not homogenous ... Current Max Percent Diff: 16.2%")
        def verdicts(text):
            return [l.strip() for l in text.splitlines() if l.startswith(("PASS::", "FAIL::")) or "Current Max Percent Diff" in l]
        ref = subprocess.run([os.path.join(H.ROOT, "oracle", "_ref", "qs"), "-i", deck], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                             timeout=600, env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
        assert ref.returncode == 0, ref.stderr[-2000:]
        assert "PASS:: No Particles Lost During Run" in res.stdout
        assert len(verdicts(ref.stdout)) >= 4 and verdicts(res.stdout) == verdicts(ref.stdout)
