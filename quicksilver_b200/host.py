"""Host model wrapper: the reference's MonteCarlo object + cycleInit / cycleTracking / cycleFinalize
(src/main.cc:38-121,310-324) over the qsb_mc_* C ABI.  Tracking itself is done by a DeviceContext
(quicksilver_b200.device) -- this module never computes a segment."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import BAL, BAL_COUNT, BAL_NAMES, PARTICLE_DTYPE, QsbError


class MonteCarlo:
    """mirror of `MonteCarlo* initMC(const Parameters&)` (src/initMC.cc:53-74) for one rank."""

    def __init__(self, argv, rank=0, n_ranks=1, allreduce=None):
        self._lib = _capi.lib()
        argv = ["qs"] + [str(a) for a in argv]
        arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
        self._h = C.c_void_p()
        rc = self._lib.qsb_mc_create(len(argv), arr, rank, n_ranks, C.byref(self._h))
        if rc != 0:
            raise QsbError(rc, (self._lib.qsb_mc_last_error(None) or b"").decode())
        self.rank, self.n_ranks = rank, n_ranks
        self._cb = None
        if allreduce is not None:
            self.set_allreduce(allreduce)
        self.image = _capi.Image()
        self._check(self._lib.qsb_mc_get_image(self._h, C.byref(self.image)))

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise QsbError(rc, (self._lib.qsb_mc_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.qsb_mc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_allreduce(self, fn):
        """fn(numpy array[, op]) must reduce the array over ranks in place (stand-in for mpiAllreduce); op is "sum"
        (default) or "max" (the fluence test of the benchmark report)."""
        def trampoline(_user, buf, count, dtype):
            ct = C.c_uint64 if dtype == 1 else C.c_double
            arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(ct)), (count,))
            if dtype == 2:
                fn(arr, "max")
            else:
                fn(arr)
        self._cb = _capi.ALLREDUCE_FN(trampoline)
        self._check(self._lib.qsb_mc_set_allreduce(self._h, self._cb, None))

    def get_int(self, key):
        v = C.c_int64()
        self._check(self._lib.qsb_mc_get_int(self._h, key.encode(), C.byref(v)))
        return v.value

    def get_double(self, key):
        v = C.c_double()
        self._check(self._lib.qsb_mc_get_double(self._h, key.encode(), C.byref(v)))
        return v.value

    def print_parameters(self):
        need = C.c_uint64()
        self._check(self._lib.qsb_mc_print_parameters(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        self._check(self._lib.qsb_mc_print_parameters(self._h, buf, need.value, None))
        return buf.value.decode()

    # -- the cycle ---------------------------------------------------------------------------------
    def cycle_init(self):
        self._check(self._lib.qsb_mc_cycle_init(self._h))

    def set_strict_math(self, on=True):
        """MC_SourceNow's log/sin/cos: libm (default, the reference's bits) or the portable functions the device uses."""
        self._check(self._lib.qsb_mc_set_strict_math(self._h, int(bool(on))))

    # -- the cycle with the population resident on the device (include/qsb.h, "device-resident cycles") --------------
    def cycle_init_resident(self, ctx):
        """cycleInit with the per-particle work on the device; returns the qsb_cycle_init_result."""
        res = _capi.CycleInitResult()
        self._check(self._lib.qsb_mc_cycle_init_resident(self._h, ctx._h, C.byref(res)))
        return res

    def cycle_tracking_resident(self, ctx):
        """cycleTracking on one rank, census left on the device."""
        stats = _capi.TrackStats()
        self._check(self._lib.qsb_mc_cycle_tracking_resident(self._h, ctx._h, C.byref(stats)))
        return stats

    def tracking_end_resident(self, ctx):
        self._check(self._lib.qsb_mc_tracking_end_resident(self._h, ctx._h))

    def source_plan(self, n_census):
        """(source_offsets[n_cells+1], source_tally[n_cells], source particle weight, split factor) of the coming cycle for a
        rank whose carried-over census holds n_census particles -- what cycle_init_resident hands to the device."""
        n = self.image.n_cells
        off, tally = np.zeros(n + 1, dtype=np.int32), np.zeros(n, dtype=np.uint64)
        w, f = C.c_double(), C.c_double()
        self._check(self._lib.qsb_mc_source_plan(self._h, off.ctypes.data_as(C.POINTER(C.c_int32)), tally.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                 C.byref(w), C.byref(f), int(n_census)))
        return off, tally, w.value, f.value

    def census_to_host(self, ctx):
        """bring the resident census back into the processed vault."""
        self._check(self._lib.qsb_mc_census_to_host(self._h, ctx._h))

    def processing(self):
        """the processing vault (tracking input) as a structured numpy array (copy)."""
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.qsb_mc_processing(self._h, C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=PARTICLE_DTYPE)
        raw = (C.c_char * (n.value * PARTICLE_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(raw, dtype=PARTICLE_DTYPE).copy()

    def set_tracking_result(self, census, balance, flux_sum):
        census = np.ascontiguousarray(census, dtype=PARTICLE_DTYPE)
        bal = np.ascontiguousarray(balance, dtype=np.uint64)
        assert bal.shape == (BAL_COUNT,)
        self._check(self._lib.qsb_mc_set_tracking_result(
            self._h, census.ctypes.data_as(C.c_void_p), len(census),
            bal.ctypes.data_as(C.POINTER(C.c_uint64)), float(flux_sum)))

    def cycle_tracking(self, ctx):
        """drop-in for cycleTracking(MonteCarlo*) on one rank: host vault -> device -> host."""
        stats = _capi.TrackStats()
        self._check(self._lib.qsb_mc_cycle_tracking(self._h, ctx._h, C.byref(stats)))
        return stats

    def tracking_begin(self, ctx):
        """first half of cycle_tracking: processing vault -> device (streamed), census -> processed vault."""
        self._check(self._lib.qsb_mc_tracking_begin(self._h, ctx._h))

    def tracking_end(self, ctx):
        """second half: the rest of the census, balance and flux sum into the host model's tallies."""
        self._check(self._lib.qsb_mc_tracking_end(self._h, ctx._h))

    def processed(self):
        """the processed vault (this cycle's census) as a structured numpy array (copy)."""
        n = self.get_int("nProcessed")
        ptr = C.c_void_p()
        self._check(self._lib.qsb_mc_processed(self._h, C.byref(ptr), None))
        if n == 0:
            return np.zeros(0, dtype=PARTICLE_DTYPE)
        raw = (C.c_char * (n * PARTICLE_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(raw, dtype=PARTICLE_DTYPE).copy()

    def cycle_finalize(self):
        row = np.zeros(BAL_COUNT, dtype=np.uint64)
        flux = C.c_double()
        self._check(self._lib.qsb_mc_cycle_finalize(self._h, row.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(flux)))
        return row, flux.value

    def cumulative_balance(self):
        row = np.zeros(BAL_COUNT, dtype=np.uint64)
        self._check(self._lib.qsb_mc_cumulative_balance(self._h, row.ctypes.data_as(C.POINTER(C.c_uint64))))
        return row

    def energy_spectrum(self):
        """EnergySpectrum: per group edge, census particles counted so far (summed over ranks; every rank calls it)."""
        n = C.c_uint64()
        self._check(self._lib.qsb_mc_energy_spectrum(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint64)
        self._check(self._lib.qsb_mc_energy_spectrum(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), n.value, None))
        return out

    def write_energy_spectrum(self):
        """PrintSpectrum (src/EnergySpectrum.cc:37-62): rank 0 writes <energySpectrum>.dat; nothing if no file is named."""
        self._check(self._lib.qsb_mc_write_energy_spectrum(self._h))

    def cross_sections_text(self):
        """checkCrossSections (src/initMC.cc:392-484): the text of <crossSectionsOut>.dat."""
        need = C.c_uint64()
        self._check(self._lib.qsb_mc_cross_sections_text(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        self._check(self._lib.qsb_mc_cross_sections_text(self._h, buf, need.value, None))
        return buf.value.decode()

    def coral_benchmark_report(self, fluence=None):
        """coralBenchmarkCorrectness (src/CoralBenchmark.cc): (report text -- empty unless the deck sets coralBenchmark,
        and on ranks other than 0 --, number of tests passed out of 4).  fluence: this rank's per-cell fluence."""
        f = None if fluence is None else np.ascontiguousarray(fluence, dtype=np.float64)
        ptr = None if f is None else f.ctypes.data_as(C.POINTER(C.c_double))
        n = 0 if f is None else len(f)
        need, passed = C.c_uint64(), C.c_int32()
        buf = C.create_string_buffer(8192)
        self._check(self._lib.qsb_mc_coral_benchmark_report(self._h, ptr, n, buf, 8192, C.byref(need), C.byref(passed)))
        return buf.value.decode(), passed.value

    TIMERS = ("main", "cycleInit", "cycleTracking", "cycleTracking_Kernel", "cycleTracking_MPI", "cycleTracking_Test_Done", "cycleFinalize")

    def timer_add(self, name, microseconds, calls=1):
        """add a caller-timed share to one of the reference's seven timers (MC_Fast_Timer)."""
        self._check(self._lib.qsb_mc_timer_add(self._h, self.TIMERS.index(name), float(microseconds), int(calls)))

    def timer(self, name):
        """(cumulative microseconds, number of calls) of one timer on this rank."""
        us, calls = C.c_double(), C.c_uint64()
        self._check(self._lib.qsb_mc_get_timer(self._h, self.TIMERS.index(name), C.byref(us), C.byref(calls)))
        return us.value, calls.value

    def timer_report(self, last_cycle=False):
        """Cumulative_Report (timer table + Figure Of Merit line) or Last_Cycle_Report; text on rank 0, every rank calls."""
        need = C.c_uint64()
        buf = C.create_string_buffer(4096)
        self._check(self._lib.qsb_mc_format_timer_report(self._h, int(bool(last_cycle)), buf, 4096, C.byref(need)))
        return buf.value.decode()

    def format_figure_of_merit(self, tracking_seconds):
        buf = C.create_string_buffer(256)
        self._check(self._lib.qsb_mc_format_figure_of_merit(self._h, float(tracking_seconds), buf, 256))
        return buf.value.decode()

    def format_cycle_row(self, cycle, row, flux, t_init=0.0, t_track=0.0, t_final=0.0):
        buf = C.create_string_buffer(2048)
        row = np.ascontiguousarray(row, dtype=np.uint64)
        self._check(self._lib.qsb_mc_format_cycle_row(self._h, cycle, row.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                      flux, t_init, t_track, t_final, buf, 2048))
        return buf.value.decode()


def table_row(row, flux):
    """[start source rr split absorb scatter fission produce collisn escape census num_seg] + flux:
    the column order of the reference's cycle table (src/Tallies.hh:60-76)."""
    order = ("start", "source", "rr", "split", "absorb", "scatter", "fission", "produce", "collision", "escape",
             "census", "num_segments")
    return [int(row[BAL[k]]) for k in order], flux
