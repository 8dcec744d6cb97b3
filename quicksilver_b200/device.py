"""Device context wrapper (qsb_ctx): one GPU's flattened problem image, SoA vaults and the sm_100a
tracking kernels behind the C ABI of include/qsb.h.  No CPU fallback: creation fails with QsbError
(QSB_ERR_CUDA) when no B200-class device is usable."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import BAL_COUNT, EXCHANGE_DTYPE, PARTICLE_DTYPE, QsbError


class DeviceContext:
    def __init__(self, image, dt, device=0, validation=True, particle_capacity=0, send_capacity=0,
                 threads_per_block=0, blocks_per_sm=0, check_geometry=False, check_reactions=False, event=None):
        """event: False = history-based tracking kernel (tracking_mode bit 0 set), True / None = the library default, the
        event-based kernel (QSB_TRACKING=history|event in the environment overrides either)"""
        self._lib = _capi.lib()
        self.image = image
        self.opt = _capi.Options(int(bool(validation)), (1 if event is False else 0) | (2 if check_geometry else 0) | (4 if check_reactions else 0), int(particle_capacity), int(send_capacity),
                                 int(threads_per_block), int(blocks_per_sm))
        self._h = C.c_void_p()
        rc = self._lib.qsb_create(int(device), C.byref(image), float(dt), C.byref(self.opt), C.byref(self._h))
        if rc != 0:
            raise QsbError(rc, (self._lib.qsb_last_error(None) or b"").decode())
        self.n_ranks = image.n_ranks

    def _check(self, rc):
        if rc != 0:
            raise QsbError(rc, (self._lib.qsb_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.qsb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cycle_begin(self, keep_census=False):
        self._check(self._lib.qsb_cycle_begin(self._h, int(keep_census)))

    def put_particles(self, particles):
        p = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        self._check(self._lib.qsb_put_particles(self._h, p.ctypes.data_as(C.c_void_p), len(p)))

    def put_census(self, particles):
        """host records -> census vault (the next cycle_init_resident carries them over)."""
        p = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        self._check(self._lib.qsb_put_census(self._h, p.ctypes.data_as(C.c_void_p), len(p)))

    def cycle_init_resident(self, plan_id, source_offsets, source_tally, source_weight, e_min, e_max, split_factor=1.0,
                            low_weight_cutoff=0.0):
        """qsb_cycle_init_resident: census vault + source particles -> population control -> roulette -> processing vault."""
        off = None if source_offsets is None else np.ascontiguousarray(source_offsets, dtype=np.int32)
        tal = None if source_tally is None else np.ascontiguousarray(source_tally, dtype=np.uint64)
        args = _capi.CycleInitArgs(int(plan_id),
                                   None if off is None else off.ctypes.data_as(C.POINTER(C.c_int32)),
                                   None if tal is None else tal.ctypes.data_as(C.POINTER(C.c_uint64)),
                                   float(source_weight), float(e_min), float(e_max), float(split_factor), float(low_weight_cutoff))
        res = _capi.CycleInitResult()
        self._check(self._lib.qsb_cycle_init_resident(self._h, C.byref(args), C.byref(res)))
        return res

    def track(self):
        stats = _capi.TrackStats()
        self._check(self._lib.qsb_track(self._h, C.byref(stats)))
        return stats

    def track_host(self, particles, census_capacity=None):
        """qsb_track_host: host vault in, census out, copies overlapped with tracking.  Returns (census, stats)."""
        p = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        cap = int(census_capacity if census_capacity is not None else len(p) + len(p) // 4 + 65536)
        out = np.empty(cap, dtype=PARTICLE_DTYPE)
        n = C.c_uint64()
        stats = _capi.TrackStats()
        self._check(self._lib.qsb_track_host(self._h, p.ctypes.data_as(C.c_void_p), len(p), out.ctypes.data_as(C.c_void_p), cap,
                                             C.byref(n), C.byref(stats)))
        if n.value > cap:
            more = np.empty(n.value - cap, dtype=PARTICLE_DTYPE)
            self._check(self._lib.qsb_get_census_range(self._h, cap, more.ctypes.data_as(C.c_void_p), len(more)))
            out = np.concatenate([out, more])
        return out[:n.value], stats

    def census_count(self):
        n = C.c_uint64()
        self._check(self._lib.qsb_census_count(self._h, C.byref(n)))
        return n.value

    def get_census(self):
        n = self.census_count()
        out = np.zeros(n, dtype=PARTICLE_DTYPE)
        got = C.c_uint64()
        self._check(self._lib.qsb_get_census(self._h, out.ctypes.data_as(C.c_void_p), n, C.byref(got)))
        return out[:got.value]

    def get_balance(self):
        out = np.zeros(BAL_COUNT, dtype=np.uint64)
        self._check(self._lib.qsb_get_balance(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def get_scalar_flux(self):
        out = np.zeros((self.image.n_cells, self.image.n_groups))
        self._check(self._lib.qsb_get_scalar_flux(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def scalar_flux_sum(self):
        s = C.c_double()
        self._check(self._lib.qsb_scalar_flux_sum(self._h, C.byref(s)))
        return s.value

    def fluence_accumulate(self):
        self._check(self._lib.qsb_fluence_accumulate(self._h))

    def census_energy_spectrum(self):
        """histogram of the current census over the energy groups (n_groups + 1 counts), computed on the device."""
        out = np.zeros(self.image.n_groups + 1, dtype=np.uint64)
        self._check(self._lib.qsb_census_energy_spectrum(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), len(out)))
        return out

    def get_fluence(self):
        out = np.zeros(self.image.n_cells)
        self._check(self._lib.qsb_get_fluence(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def send_counts(self):
        out = np.zeros(self.n_ranks, dtype=np.uint64)
        self._check(self._lib.qsb_send_counts(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def send_slab(self, peer):
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.qsb_send_slab(self._h, int(peer), C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def clear_sends(self):
        self._check(self._lib.qsb_clear_sends(self._h))

    def put_arrivals(self, device_ptr, n):
        self._check(self._lib.qsb_put_arrivals(self._h, C.c_void_p(device_ptr), int(n)))

    PEER_HANDLE_BYTES = 64
    MAX_PEERS = 8

    def peer_export(self):
        """qsb_peer_export: (64-byte CUDA IPC handle of this rank's exchange ring, ring capacity in records)."""
        buf = (C.c_ubyte * self.PEER_HANDLE_BYTES)()
        cap = C.c_uint64()
        self._check(self._lib.qsb_peer_export(self._h, buf, C.byref(cap)))
        return bytes(buf), cap.value

    def peer_connect(self, handles, watchdog_seconds=0.0):
        """qsb_peer_connect: `handles` = every rank's handle in rank order (bytes, n_ranks * 64)."""
        raw = bytes(handles)
        assert len(raw) == self.n_ranks * self.PEER_HANDLE_BYTES
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        self._check(self._lib.qsb_peer_connect(self._h, buf, self.n_ranks, float(watchdog_seconds)))
        self.peer_mode = True

    peer_mode = False

    def diagnostics(self):
        out = np.zeros(8, dtype=np.uint64)
        self._check(self._lib.qsb_get_diagnostics(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        keys = ("slow_geometry", "geometry_mismatch", "reaction_lookups", "compact_geometry", "registers",
                "blocks_per_sm", "grid", "vault_slots_used")
        return {k: int(v) for k, v in zip(keys, out)}

    def peer_diagnostics(self):
        """device-side timings of the last peer-mode launch (qsb_peer_diagnostics)"""
        out = np.zeros(8, dtype=np.uint64)
        self._check(self._lib.qsb_peer_diagnostics(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        keys = ("first_idle_ns", "done_ns", "send_cycles", "send_calls", "startup_wait_ns", "tickets", "deposited", "bulk_done_ns")
        return {k: int(v) for k, v in zip(keys, out)}

    def launch_count(self):
        return int(self._lib.qsb_launch_count(self._h))
