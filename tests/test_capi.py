"""The C-ABI library loads, exports every symbol include/qsb.h declares, and refuses (loudly, with an error
code -- never a CPU fallback) to create a device context when no B200-class GPU is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from quicksilver_b200 import _capi, decks, device, host

HEADER = os.path.join(H.ROOT, "include", "qsb.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsb_[a-z0-9_]+)\s*\(", text)) - {"qsb_allreduce_fn"})


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("qsb_create", "qsb_destroy", "qsb_cycle_begin", "qsb_put_particles", "qsb_track", "qsb_get_census",
                 "qsb_get_balance", "qsb_get_scalar_flux", "qsb_mc_create", "qsb_mc_cycle_init", "qsb_mc_cycle_tracking",
                 "qsb_mc_cycle_finalize", "qsb_last_error", "qsb_put_arrivals", "qsb_send_slab"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_capi.library_path())
    missing = [n for n in _declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding declares a signature for each of them
    _capi.lib()
    assert sorted(_capi.EXPORTS) == _declared_functions()


def test_record_layouts_match_the_reference():
    """MC_Base_Particle is 136 bytes (src/MC_Base_Particle.hh:75-92); the exchange record adds the direction cosine."""
    assert _capi.PARTICLE_DTYPE.itemsize == 136
    assert _capi.lib().qsb_exchange_record_bytes() == _capi.EXCHANGE_DTYPE.itemsize == 160
    assert _capi.lib().qsb_version().decode()


def test_error_convention_no_throw_no_abort(tmp_path):
    lib = _capi.lib()
    h = C.c_void_p()
    argv = (C.c_char_p * 3)(b"qs", b"-i", b"/nonexistent/deck.inp")
    rc = lib.qsb_mc_create(3, argv, 0, 1, C.byref(h))
    assert rc == -2 and not h.value
    assert b"deck" in lib.qsb_mc_last_error(None) or lib.qsb_mc_last_error(None)
    assert lib.qsb_mc_cycle_init(None) == -1
    assert lib.qsb_track(None, None) == -1
    assert lib.qsb_destroy(None) == -1


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the behaviour on a machine WITHOUT a GPU")
def test_device_context_fails_loudly_without_a_gpu(tmp_path):
    deck = decks.write_deck(decks.derive("CTS2_1", nx=4, ny=4, nz=4, lx=4, ly=4, lz=4, nParticles=640, nSteps=1), str(tmp_path / "d.inp"))
    mc = host.MonteCarlo(["-i", deck])
    with pytest.raises(host.QsbError) as err:
        device.DeviceContext(mc.image, mc.get_double("dt"))
    assert err.value.code == -3                       # QSB_ERR_CUDA
    assert "no CPU path" in str(err.value)
    # the drop-in call cannot run either: there is nothing behind it but the device
    mc.cycle_init()
    assert _capi.lib().qsb_mc_cycle_tracking(mc._h, None, None) == -1


def test_product_package_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's reference leg may use oracle/."""
    pkg = os.path.join(H.ROOT, "quicksilver_b200")
    for base, _, files in os.walk(pkg):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".h", ".hh")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in text and "qs_oracle" not in text and "qso_track" not in text, os.path.join(base, f)
