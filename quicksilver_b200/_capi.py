"""ctypes binding of include/qsb.h (the C ABI of libqsb.so).

The library holds the host model (qsb_mc_*) and the sm_100a tracking kernels (qsb_*).  There is no
Python or CPU fallback for the device entry points: if the shared library is missing the import
fails, and if no CUDA device is usable qsb_create returns QSB_ERR_CUDA.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

BAL_NAMES = ("absorb", "census", "escape", "collision", "end", "fission", "produce", "scatter",
             "start", "source", "rr", "split", "num_segments")
BAL = {n: i for i, n in enumerate(BAL_NAMES)}
BAL_COUNT = len(BAL_NAMES)

# MC_Base_Particle layout, 136 bytes (reference: src/MC_Base_Particle.hh:75-92)
PARTICLE_DTYPE = np.dtype([
    ("coordinate", "<f8", 3), ("velocity", "<f8", 3), ("kinetic_energy", "<f8"), ("weight", "<f8"),
    ("time_to_census", "<f8"), ("age", "<f8"), ("num_mean_free_paths", "<f8"), ("num_segments", "<f8"),
    ("random_number_seed", "<u8"), ("identifier", "<u8"), ("last_event", "<i4"), ("num_collisions", "<i4"),
    ("breed", "<i4"), ("species", "<i4"), ("domain", "<i4"), ("cell", "<i4")])
assert PARTICLE_DTYPE.itemsize == 136

# exchange record: base particle + direction cosine, 160 bytes
EXCHANGE_DTYPE = np.dtype([("p", PARTICLE_DTYPE), ("direction_cosine", "<f8", 3)])
assert EXCHANGE_DTYPE.itemsize == 160

_P = C.POINTER


class Image(C.Structure):
    """qsb_image: borrowed pointers into the host model's flat arrays."""
    _fields_ = [
        ("abi_version", C.c_int32), ("n_domains", C.c_int32), ("n_cells", C.c_int32), ("n_groups", C.c_int32),
        ("n_materials", C.c_int32), ("n_isotopes", C.c_int32), ("max_reactions_per_material", C.c_int32),
        ("my_rank", C.c_int32), ("n_ranks", C.c_int32),
        ("global_nx", C.c_int32), ("global_ny", C.c_int32), ("global_nz", C.c_int32),
        ("global_lx", C.c_double), ("global_ly", C.c_double), ("global_lz", C.c_double),
        ("domain_cell_offset", _P(C.c_int32)), ("domain_gid", _P(C.c_int32)),
        ("planes", _P(C.c_double)), ("nodes", _P(C.c_double)), ("cell_gid", _P(C.c_int32)),
        ("cell_material", _P(C.c_int32)), ("cell_volume", _P(C.c_double)), ("cell_id", _P(C.c_uint64)),
        ("face_event", _P(C.c_uint8)), ("face_adj_cell", _P(C.c_int32)), ("face_adj_domain", _P(C.c_int32)),
        ("face_nbr_rank", _P(C.c_int32)),
        ("energies", _P(C.c_double)), ("mat_n_isotopes", _P(C.c_int32)), ("mat_n_reactions", _P(C.c_int32)),
        ("mat_mass", _P(C.c_double)), ("mat_nu_bar", _P(C.c_double)), ("mat_react_type", _P(C.c_uint8)),
        ("xs_total", _P(C.c_double)), ("xs_react", _P(C.c_double)), ("mat_periodic", _P(C.c_uint8)),
    ]

    def array(self, name):
        """numpy view (no copy) of one of the image's arrays."""
        n_c, n_g, n_m, mr = self.n_cells, self.n_groups, self.n_materials, self.max_reactions_per_material
        shapes = {
            "domain_cell_offset": (self.n_domains + 1,), "domain_gid": (self.n_domains,),
            "planes": (n_c, 24, 4), "nodes": (n_c, 14, 3), "cell_gid": (n_c,), "cell_material": (n_c,),
            "cell_volume": (n_c,), "cell_id": (n_c,), "face_event": (n_c, 6), "face_adj_cell": (n_c, 6),
            "face_adj_domain": (n_c, 6), "face_nbr_rank": (n_c, 6), "energies": (n_g + 1,),
            "mat_n_isotopes": (n_m,), "mat_n_reactions": (n_m,), "mat_mass": (n_m,), "mat_nu_bar": (n_m,),
            "mat_react_type": (n_m, mr), "xs_total": (n_m, n_g), "xs_react": (n_m, n_g, mr), "mat_periodic": (n_m,),
        }
        ptr = getattr(self, name)
        shape = shapes[name]
        if int(np.prod(shape)) == 0:
            return np.zeros(shape, dtype=np.ctypeslib.as_array(ptr, (1,)).dtype if ptr else np.float64)
        return np.ctypeslib.as_array(ptr, shape)


class Options(C.Structure):
    _fields_ = [("validation", C.c_int32), ("tracking_mode", C.c_int32), ("particle_capacity", C.c_uint64),
                ("send_capacity", C.c_uint64), ("threads_per_block", C.c_int32), ("blocks_per_sm", C.c_int32)]


class TrackStats(C.Structure):
    _fields_ = [("n_processed", C.c_uint64), ("n_census", C.c_uint64), ("n_sent", C.c_uint64),
                ("n_launches", C.c_uint32), ("device_ms", C.c_float)]


class CycleInitArgs(C.Structure):
    """qsb_cycle_init_args"""
    _fields_ = [("plan_id", C.c_uint64), ("source_offsets", _P(C.c_int32)), ("source_tally", _P(C.c_uint64)),
                ("source_weight", C.c_double), ("e_min", C.c_double), ("e_max", C.c_double),
                ("split_factor", C.c_double), ("low_weight_cutoff", C.c_double)]


class CycleInitResult(C.Structure):
    """qsb_cycle_init_result"""
    _fields_ = [("n_start", C.c_uint64), ("n_source", C.c_uint64), ("n_rr", C.c_uint64), ("n_split", C.c_uint64),
                ("n_processing", C.c_uint64), ("device_ms", C.c_float), ("n_launches", C.c_uint32)]


ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32)


class QsbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("qsb error %d: %s" % (code, message))
        self.code = code


def library_path():
    return os.environ.get("QSB_LIBRARY", os.path.join(_HERE, "libqsb.so"))


def _declare(lib):
    vp, u64p = C.c_void_p, _P(C.c_uint64)
    sig = {
        "qsb_version": (C.c_char_p, []),
        "qsb_mc_create": (C.c_int, [C.c_int, _P(C.c_char_p), C.c_int, C.c_int, _P(vp)]),
        "qsb_mc_destroy": (C.c_int, [vp]),
        "qsb_mc_set_allreduce": (C.c_int, [vp, ALLREDUCE_FN, vp]),
        "qsb_mc_print_parameters": (C.c_int, [vp, C.c_char_p, C.c_uint64, u64p]),
        "qsb_mc_get_image": (C.c_int, [vp, _P(Image)]),
        "qsb_mc_get_int": (C.c_int, [vp, C.c_char_p, _P(C.c_int64)]),
        "qsb_mc_get_double": (C.c_int, [vp, C.c_char_p, _P(C.c_double)]),
        "qsb_mc_cycle_init": (C.c_int, [vp]),
        "qsb_mc_set_strict_math": (C.c_int, [vp, C.c_int]),
        "qsb_mc_cycle_init_resident": (C.c_int, [vp, vp, _P(CycleInitResult)]),
        "qsb_mc_cycle_tracking_resident": (C.c_int, [vp, vp, _P(TrackStats)]),
        "qsb_mc_tracking_end_resident": (C.c_int, [vp, vp]),
        "qsb_mc_census_to_host": (C.c_int, [vp, vp]),
        "qsb_mc_source_plan": (C.c_int, [vp, _P(C.c_int32), u64p, _P(C.c_double), _P(C.c_double), C.c_uint64]),
        "qsb_cycle_init_resident": (C.c_int, [vp, _P(CycleInitArgs), _P(CycleInitResult)]),
        "qsb_put_census": (C.c_int, [vp, vp, C.c_uint64]),
        "qsb_mc_processing": (C.c_int, [vp, _P(vp), u64p]),
        "qsb_mc_processed": (C.c_int, [vp, _P(vp), u64p]),
        "qsb_mc_set_tracking_result": (C.c_int, [vp, vp, C.c_uint64, u64p, C.c_double]),
        "qsb_mc_cycle_finalize": (C.c_int, [vp, u64p, _P(C.c_double)]),
        "qsb_mc_cumulative_balance": (C.c_int, [vp, u64p]),
        "qsb_mc_format_cycle_row": (C.c_int, [vp, C.c_int, u64p, C.c_double, C.c_double, C.c_double, C.c_double,
                                               C.c_char_p, C.c_uint64]),
        "qsb_mc_coral_benchmark_report": (C.c_int, [vp, _P(C.c_double), C.c_uint64, C.c_char_p, C.c_uint64, u64p, _P(C.c_int32)]),
        "qsb_mc_format_figure_of_merit": (C.c_int, [vp, C.c_double, C.c_char_p, C.c_uint64]),
        "qsb_mc_timer_add": (C.c_int, [vp, C.c_int, C.c_double, C.c_uint64]),
        "qsb_mc_get_timer": (C.c_int, [vp, C.c_int, _P(C.c_double), u64p]),
        "qsb_mc_format_timer_report": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_uint64, u64p]),
        "qsb_mc_last_error": (C.c_char_p, [vp]),
        "qsb_create": (C.c_int, [C.c_int, _P(Image), C.c_double, _P(Options), _P(vp)]),
        "qsb_destroy": (C.c_int, [vp]),
        "qsb_cycle_begin": (C.c_int, [vp, C.c_int]),
        "qsb_put_particles": (C.c_int, [vp, vp, C.c_uint64]),
        "qsb_track": (C.c_int, [vp, _P(TrackStats)]),
        "qsb_stream_begin": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64]),
        "qsb_stream_end": (C.c_int, [vp, u64p]),
        "qsb_get_census_range": (C.c_int, [vp, C.c_uint64, vp, C.c_uint64]),
        "qsb_track_host": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, u64p, _P(TrackStats)]),
        "qsb_census_count": (C.c_int, [vp, u64p]),
        "qsb_get_census": (C.c_int, [vp, vp, C.c_uint64, u64p]),
        "qsb_get_balance": (C.c_int, [vp, u64p]),
        "qsb_get_scalar_flux": (C.c_int, [vp, _P(C.c_double)]),
        "qsb_scalar_flux_sum": (C.c_int, [vp, _P(C.c_double)]),
        "qsb_fluence_accumulate": (C.c_int, [vp]),
        "qsb_census_energy_spectrum": (C.c_int, [vp, u64p, C.c_uint64]),
        "qsb_mc_energy_spectrum": (C.c_int, [vp, u64p, C.c_uint64, u64p]),
        "qsb_mc_write_energy_spectrum": (C.c_int, [vp]),
        "qsb_mc_cross_sections_text": (C.c_int, [vp, C.c_char_p, C.c_uint64, u64p]),
        "qsb_get_fluence": (C.c_int, [vp, _P(C.c_double)]),
        "qsb_send_counts": (C.c_int, [vp, u64p]),
        "qsb_send_slab": (C.c_int, [vp, C.c_int, _P(vp), u64p]),
        "qsb_clear_sends": (C.c_int, [vp]),
        "qsb_put_arrivals": (C.c_int, [vp, vp, C.c_uint64]),
        "qsb_exchange_record_bytes": (C.c_uint64, []),
        "qsb_peer_export": (C.c_int, [vp, vp, u64p]),
        "qsb_peer_connect": (C.c_int, [vp, vp, C.c_int, C.c_double]),
        "qsb_peer_disconnect": (C.c_int, [vp]),
        "qsb_last_error": (C.c_char_p, [vp]),
        "qsb_launch_count": (C.c_uint64, [vp]),
        "qsb_get_diagnostics": (C.c_int, [vp, u64p]),
        "qsb_peer_diagnostics": (C.c_int, [vp, u64p]),
        "qsb_kernel_hash": (C.c_char_p, []),
        "qsb_mc_cycle_tracking": (C.c_int, [vp, vp, _P(TrackStats)]),
        "qsb_mc_tracking_begin": (C.c_int, [vp, vp]),
        "qsb_mc_tracking_end": (C.c_int, [vp, vp]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype, fn.argtypes = res, args
    return sig, missing


_lib = None
EXPORTS = ()


def lib():
    """Load libqsb.so once.  Raises (never falls back) when the library has not been built."""
    global _lib, EXPORTS
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C quicksilver_b200/csrc`" % path)
        handle = C.CDLL(path)
        sig, missing = _declare(handle)
        if missing and not os.environ.get("QSB_LIBRARY"):
            raise ImportError("libqsb.so lacks symbols declared in include/qsb.h: %s" % ", ".join(missing))
        EXPORTS = tuple(sig)
        _lib = handle
    return _lib
