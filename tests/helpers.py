"""Shared test helpers: QSD reader (oracle/ref_dump.cc output), oracle binding, golden tables."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from quicksilver_b200 import _capi  # noqa: E402
from quicksilver_b200._capi import BAL, BAL_COUNT, EXCHANGE_DTYPE, PARTICLE_DTYPE  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_QS = os.path.join(ORACLE_DIR, "_ref", "qs")
REF_DUMP = os.path.join(ORACLE_DIR, "_ref", "qs_dump")
REF_DUMP_STRICT = os.path.join(ORACLE_DIR, "_ref", "qs_dump_strict")     # the reference with log / sin / cos mapped to qs_strict_math.h
GOLDEN = os.path.join(ROOT, "tests", "golden")


def read_qsd(path):
    """{name: ndarray} from a QSD1 file written by oracle/ref_dump.cc."""
    out = {}
    with open(path, "rb") as f:
        assert f.read(4) == b"QSD1"
        while True:
            head = f.read(4)
            if not head:
                break
            nlen = int(np.frombuffer(head, "<u4")[0])
            name = f.read(nlen).decode()
            dtype = f.read(1).decode()
            count, inner = (int(v) for v in np.frombuffer(f.read(16), "<u8"))
            np_dtype = {"d": "<f8", "i": "<i4", "u": "<u8", "b": "u1"}[dtype]
            data = np.frombuffer(f.read(count * inner * np.dtype(np_dtype).itemsize), np_dtype)
            out[name] = data.reshape(count, inner) if inner != 1 else data
    return out


def particles_from_bytes(arr):
    return np.ascontiguousarray(arr).view(PARTICLE_DTYPE).reshape(-1)


def run_reference_dump(argv, out_dir, particle_cycles=1, threads=4, exe=None):
    env = dict(os.environ, QS_DUMP_DIR=out_dir, QS_DUMP_PARTICLE_CYCLES=str(particle_cycles), OMP_NUM_THREADS=str(threads))
    subprocess.run([exe or REF_DUMP] + [str(a) for a in argv], check=True, env=env, stdout=subprocess.DEVNULL)


# ---- oracle binding -----------------------------------------------------------------------------------

class _OracleIO(C.Structure):
    _fields_ = [("initial", C.c_void_p), ("n_initial", C.c_uint64),
                ("arrivals", C.c_void_p), ("n_arrivals", C.c_uint64),
                ("census", C.c_void_p), ("census_cap", C.c_uint64), ("n_census", C.c_uint64),
                ("sends", C.c_void_p), ("send_rank", C.c_void_p), ("send_cap", C.c_uint64), ("n_sends", C.c_uint64),
                ("balance", C.c_uint64 * BAL_COUNT), ("flux", C.c_void_p), ("n_processed", C.c_uint64),
                ("n_retry_moves", C.c_uint64), ("n_forced_collisions", C.c_uint64), ("n_reaction_lookups", C.c_uint64)]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
        _oracle = C.CDLL(path)
        _oracle.qso_track.restype = C.c_int
        _oracle.qso_track.argtypes = [C.POINTER(_capi.Image), C.c_double, C.c_int, C.c_int, C.POINTER(_OracleIO)]
    return _oracle


class OracleResult:
    pass


class _OracleCycleInitIO(C.Structure):
    _fields_ = [("census", C.c_void_p), ("n_census", C.c_uint64), ("source_offsets", C.c_void_p), ("source_tally", C.c_void_p),
                ("source_weight", C.c_double), ("e_min", C.c_double), ("e_max", C.c_double), ("time_step", C.c_double),
                ("split_factor", C.c_double), ("low_weight_cutoff", C.c_double),
                ("out", C.c_void_p), ("out_cap", C.c_uint64), ("n_out", C.c_uint64),
                ("n_source", C.c_uint64), ("n_rr", C.c_uint64), ("n_split", C.c_uint64)]


def oracle_cycle_init(image, census, source_offsets, source_tally, source_weight, e_min, e_max, dt, split_factor, low_weight_cutoff,
                      strict=False):
    """the oracle's stage-by-stage restatement of cycleInit (oracle/qs_oracle.c: qso_cycle_init): returns
    (processing vault, n_source, n_rr, n_split)"""
    lib = oracle_lib()
    lib.qso_cycle_init.restype = C.c_int
    lib.qso_cycle_init.argtypes = [C.POINTER(_capi.Image), C.c_int, C.POINTER(_OracleCycleInitIO)]
    census = np.ascontiguousarray(census, dtype=PARTICLE_DTYPE)
    off = np.ascontiguousarray(source_offsets, dtype=np.int32)
    tally = np.ascontiguousarray(source_tally, dtype=np.uint64)
    cap = int((len(census) + int(off[-1])) * max(2.0, split_factor + 2.0)) + 1024
    out = np.zeros(cap, PARTICLE_DTYPE)
    io = _OracleCycleInitIO()
    io.census, io.n_census = census.ctypes.data, len(census)
    io.source_offsets, io.source_tally = off.ctypes.data, tally.ctypes.data
    io.source_weight, io.e_min, io.e_max, io.time_step = source_weight, e_min, e_max, dt
    io.split_factor, io.low_weight_cutoff = split_factor, low_weight_cutoff
    io.out, io.out_cap = out.ctypes.data, cap
    rc = lib.qso_cycle_init(C.byref(image), int(strict), C.byref(io))
    if rc != 0:
        raise RuntimeError("qso_cycle_init failed: %d" % rc)
    return out[:io.n_out].copy(), int(io.n_source), int(io.n_rr), int(io.n_split)


def oracle_track(image, dt, particles, arrivals=None, strict=False, threads=1, census_cap=None, send_cap=None, want_flux=True):
    """Run the CPU restatement on a processing vault.  Returns census, sends, balance, flux."""
    lib = oracle_lib()
    particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
    arrivals = np.zeros(0, EXCHANGE_DTYPE) if arrivals is None else np.ascontiguousarray(arrivals, dtype=EXCHANGE_DTYPE)
    n_in = len(particles) + len(arrivals)
    census_cap = census_cap or max(1024, 8 * n_in)
    send_cap = send_cap or max(1024, 8 * n_in)
    census = np.zeros(census_cap, PARTICLE_DTYPE)
    sends = np.zeros(send_cap, EXCHANGE_DTYPE)
    send_rank = np.zeros(send_cap, np.int32)
    flux = np.zeros((image.n_cells, image.n_groups)) if want_flux else None
    io = _OracleIO()
    io.initial, io.n_initial = particles.ctypes.data, len(particles)
    io.arrivals, io.n_arrivals = arrivals.ctypes.data, len(arrivals)
    io.census, io.census_cap = census.ctypes.data, census_cap
    io.sends, io.send_rank, io.send_cap = sends.ctypes.data, send_rank.ctypes.data, send_cap
    io.flux = flux.ctypes.data if want_flux else None
    rc = lib.qso_track(C.byref(image), dt, int(strict), int(threads), C.byref(io))
    if rc != 0:
        raise RuntimeError("qso_track failed: %d" % rc)
    r = OracleResult()
    r.census = census[:io.n_census].copy()
    r.sends = sends[:io.n_sends].copy()
    r.send_rank = send_rank[:io.n_sends].copy()
    r.balance = np.array(list(io.balance), dtype=np.uint64)
    r.flux = flux
    r.n_processed = io.n_processed
    r.n_retry_moves, r.n_forced_collisions, r.n_reaction_lookups = io.n_retry_moves, io.n_forced_collisions, io.n_reaction_lookups
    return r


def sort_particles(p):
    return p[np.argsort(p["identifier"], kind="stable")]


def golden_table(name):
    """rows of the reference's cycle table captured in BASELINE.md section 4: [(12 ints, flux), ...];
    columns start source rr split absorb scatter fission produce collisn escape census num_seg."""
    import json
    with open(os.path.join(GOLDEN, "balance_tables.json")) as f:
        return [(r[0], r[1]) for r in json.load(f)[name]["rows"]]


# ---- golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the reference) ------

def golden_case(case):
    return np.load(os.path.join(GOLDEN, case + ".npz"))


def golden_deck(case, tmp_path):
    """the deck a fixture was generated from (tests/golden/make_golden.py: CASES)."""
    sys.path.insert(0, GOLDEN)
    import make_golden
    from quicksilver_b200 import decks
    return decks.write_deck(make_golden.deck_of(case), os.path.join(str(tmp_path), case + ".inp"))


# ---- CPU stand-in for the device backend of quicksilver_b200.driver (tests of the exchange protocol) ------

class OracleBackend:
    """Same interface as driver.DeviceBackend, tracking done by the CPU oracle.  Lets the N-rank exchange /
    termination protocol of driver.exchange_rounds run under gloo on a machine without GPUs."""

    def __init__(self, image, dt, rank, world, strict=False):
        self.image, self.dt, self.rank, self.world, self.strict = image, dt, rank, world, strict
        self.torch_device = "cpu"

    def begin(self, vault):
        self.pending = np.ascontiguousarray(vault, dtype=PARTICLE_DTYPE)
        self.arrivals = np.zeros(0, EXCHANGE_DTYPE)
        self.census = []
        self.balance = np.zeros(BAL_COUNT, np.uint64)
        self.flux = np.zeros((self.image.n_cells, self.image.n_groups))
        self.sends = [np.zeros(0, EXCHANGE_DTYPE) for _ in range(self.world)]

    def track(self):
        r = oracle_track(self.image, self.dt, self.pending, arrivals=self.arrivals, strict=self.strict, threads=1)
        self.pending = np.zeros(0, PARTICLE_DTYPE)
        self.arrivals = np.zeros(0, EXCHANGE_DTYPE)
        self.census.append(r.census)
        self.balance += r.balance
        self.flux += r.flux
        for peer in range(self.world):
            self.sends[peer] = np.concatenate([self.sends[peer], r.sends[r.send_rank == peer]])
        return r

    def send_counts(self):
        return np.array([len(s) for s in self.sends], dtype=np.int64)

    def send_tensor(self, peer, n):
        import torch
        assert len(self.sends[peer]) == n
        return torch.from_numpy(np.ascontiguousarray(self.sends[peer]).view(np.uint8).reshape(-1).copy())

    def recv_tensor(self, n):
        import torch
        return torch.empty(n * EXCHANGE_DTYPE.itemsize, dtype=torch.uint8)

    def put_arrivals(self, tensor, n):
        self.arrivals = np.concatenate([self.arrivals, tensor.numpy().view(EXCHANGE_DTYPE).reshape(-1)[:n].copy()])

    def clear_sends(self):
        self.sends = [np.zeros(0, EXCHANGE_DTYPE) for _ in range(self.world)]

    def results(self):
        census = np.concatenate(self.census) if self.census else np.zeros(0, PARTICLE_DTYPE)
        return census, self.balance, float(self.flux.sum())
