"""ctypes loader for csrc/libcrowd_b200.so -- the C ABI declared in include/crowd_b200.h.

Fails loudly: if the shared library is missing or cannot be loaded, ``ExtensionMissing`` is raised.  There is no CPU
fallback anywhere in this package.
"""
import ctypes as C
import os
import subprocess

from .exceptions import ExtensionMissing, InvalidType, InvalidValue, DeviceError, CrowdDynamicsException

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CROWD_B200_LIB') or os.path.join(_HERE, 'csrc', 'libcrowd_b200.so')
HEADER_PATH = os.path.abspath(os.path.join(_HERE, '..', 'include', 'crowd_b200.h'))

CDB_OK, CDB_ERR_INVALID_TYPE, CDB_ERR_INVALID_VALUE, CDB_ERR_CUDA, CDB_ERR_CAPACITY, CDB_ERR_STATE = range(6)

F_POSITION, F_VELOCITY, F_TARGET_DIRECTION, F_FORCE, F_FORCE_PREV, F_SHOULDERS, F_ORIENTATION, F_ANGULAR_VELOCITY, \
    F_TARGET_ORIENTATION, F_TORQUE, F_TORQUE_PREV = (1 << k for k in range(11))
F_ALL_MUTABLE = (1 << 11) - 1
F_WHOLE_RECORD = 1 << 31

STEP_NAVIGATION, STEP_ORIENTATION, STEP_ADJUSTING, STEP_AGENT_AGENT, STEP_AGENT_OBSTACLE, STEP_INTEGRATOR, \
    STEP_RESET = (1 << k for k in range(7))
STEP_ALL = (1 << 7) - 1
STEP_FLUCTUATION = 1 << 7
KNN_MAX = 32
POLY_DOMAIN, POLY_TARGETS = 0, 1

_lib = None


def build(force=False):
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, 'csrc')
    cmd = ['make', '-C', src_dir, '-s'] + (['-B'] if force else [])
    subprocess.check_call(cmd)
    return LIB_PATH


def _signatures(L):
    i64, f64, vp, i32, u32 = C.c_int64, C.c_double, C.c_void_p, C.c_int, C.c_uint32
    pf64, pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    sig = {
        'cdb_last_error': (C.c_char_p, []),
        'cdb_version': (i32, []),
        'cdb_device_count': (i32, [C.POINTER(i32)]),
        'cdb_measure_fp64_peak': (i32, [i32, pf64]),
        'cdb_create': (i32, [i32, i32, i64, C.POINTER(vp)]),
        'cdb_destroy': (i32, [vp]),
        'cdb_set_stream': (i32, [vp, vp]),
        'cdb_synchronize': (i32, [vp]),
        'cdb_num_agents': (i64, [vp]),
        'cdb_upload_agents_aos': (i32, [vp, vp, i64, i64]),
        'cdb_download_agents_aos': (i32, [vp, vp, i64, i64, u32]),
        'cdb_host_register': (i32, [vp, vp, i64, i64]),
        'cdb_host_unregister': (i32, [vp, vp]),
        'cdb_upload_agents_fields': (i32, [vp, vp, i64, i64, u32]),
        'cdb_transfer_stats': (i32, [vp, pi64, pi64, i32]),
        'cdb_set_obstacles': (i32, [vp, vp, i64]),
        'cdb_set_navigation_field': (i32, [vp, i64, vp, vp, i64, i64, f64, f64, f64]),
        'cdb_clear_navigation': (i32, [vp]),
        'cdb_build_navigation_field': (i32, [vp, i64, vp, i64, vp, i64, i64, i64, f64, f64, f64, f64, f64, vp, vp, vp, pi64]),
        'cdb_reset': (i32, [vp]),
        'cdb_set_seed': (i32, [vp, C.c_uint64]),
        'cdb_fluctuation': (i32, [vp]),
        'cdb_navigation': (i32, [vp]),
        'cdb_orientation': (i32, [vp]),
        'cdb_adjust': (i32, [vp]),
        'cdb_agent_agent': (i32, [vp, f64]),
        'cdb_agent_obstacle': (i32, [vp]),
        'cdb_integrate': (i32, [vp, f64, f64, pf64]),
        'cdb_step': (i32, [vp, u32, f64, f64, f64, i64, vp]),
        'cdb_get_time': (i32, [vp, pf64, pi64]),
        'cdb_set_graphs': (i32, [vp, i32]),
        'cdb_set_small_crowd_max': (i32, [vp, i64]),
        'cdb_snapshot_begin': (i32, [vp, pi64]),
        'cdb_snapshot_wait': (i32, [vp, i64, C.POINTER(vp), pi64]),
        'cdb_scalars_begin': (i32, [vp, pi64]),
        'cdb_scalars_wait': (i32, [vp, i64, pf64, pf64, pi64, vp, i64]),
        'cdb_set_deferred_sync': (i32, [vp, i32]),
        'cdb_sync_count': (i64, [vp]),
        'cdb_set_rebuild_policy': (i32, [vp, f64, i64, i64]),
        'cdb_strip_set_kind': (i32, [vp, i32]),
        'cdb_strip_set_global_agents': (i32, [vp, i64]),
        'cdb_strip_drift': (i32, [vp, pf64, pf64, pf64]),
        'cdb_get_rebuild_stats': (i32, [vp, pi64, pi64, pi64, pi64]),
        'cdb_set_variant': (i32, [vp, i32]),
        'cdb_launch_count': (i64, [vp]),
        'cdb_profile_enable': (i32, [vp, i32]),
        'cdb_profile_read': (i32, [vp, pf64, pi64]),
        'cdb_profile_read_phases': (i32, [vp, pf64, pi64]),
        'cdb_set_pair_capacity': (i32, [vp, i64]),
        'cdb_set_search_refinement': (i32, [vp, i32]),
        'cdb_get_ext_max': (i32, [vp, pf64]),
        'cdb_get_pair_stats': (i32, [vp, pi64, pi64, pi64]),
        'cdb_build_block_list': (i32, [vp, f64]),
        'cdb_get_grid': (i32, [vp, pi64]),
        'cdb_get_cell_ids': (i32, [vp, vp, i64]),
        'cdb_get_cell_tables': (i32, [vp, vp, i64, vp, vp, i64]),
        'cdb_get_neighbor_pairs': (i32, [vp, vp, i64, pi64]),
        'cdb_set_lattice': (i32, [vp, i64, i64, i64, i64]),
        'cdb_clear_lattice': (i32, [vp]),
        'cdb_set_strip': (i32, [vp, i64, i64, i64, i64, i32, i32, i64, i64]),
        'cdb_set_agent_ids': (i32, [vp, vp, i64]),
        'cdb_halo_buffer_doubles': (i64, [vp]),
        'cdb_migrant_buffer_doubles': (i64, [vp]),
        'cdb_strip_begin': (i32, [vp, u32, f64, vp, vp]),
        'cdb_strip_export_vmax': (i32, [vp, vp]),
        'cdb_strip_import_vmax': (i32, [vp, vp]),
        'cdb_strip_finish': (i32, [vp, u32, f64, f64, vp, vp, vp, vp]),
        'cdb_strip_absorb': (i32, [vp, vp, vp, pi64]),
        'cdb_strip_count': (i32, [vp, pi64]),
        'cdb_strip_exchange_alloc': (i32, [vp]),
        'cdb_strip_exchange_handle_bytes': (i64, []),
        'cdb_strip_exchange_handles': (i32, [vp, vp]),
        'cdb_strip_exchange_connect_ipc': (i32, [vp, vp, vp]),
        'cdb_strip_exchange_connect_local': (i32, [vp, vp, vp]),
        'cdb_strip_begin_direct': (i32, [vp, u32, f64, i32]),
        'cdb_strip_finish_direct': (i32, [vp, u32, f64, f64, i32]),
        'cdb_strip_absorb_direct': (i32, [vp, pi64]),
        'cdb_export_agents': (i32, [vp, vp, vp, i64, pi64]),
        'cdb_set_states': (i32, [vp, vp, vp, vp, vp, vp, i64]),
        'cdb_get_states': (i32, [vp, vp, vp, vp, i64]),
        'cdb_exit_detection': (i32, [vp, vp, i64, f64, i32]),
        'cdb_get_exit_detection': (i32, [vp, vp, vp, i64]),
        'cdb_nearest_neighbors': (i32, [vp, f64, i64, vp]),
        'cdb_leader_follower': (i32, [vp, f64, f64, f64]),
        'cdb_leader_follower_with_herding': (i32, [vp, f64, i64, f64, f64, f64, f64]),
        'cdb_get_direction': (i32, [vp, vp, i64]),
        'cdb_set_polygons': (i32, [vp, i32, vp, vp, i64]),
        'cdb_set_active': (i32, [vp, vp, i64]),
        'cdb_get_active': (i32, [vp, vp, i64]),
        'cdb_inside_domain': (i32, [vp, pi64]),
        'cdb_target_reached': (i32, [vp, vp, i64]),
        'cdb_get_target_reached': (i32, [vp, vp, i64, i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return sig


def load():
    """Load libcrowd_b200.so (once).  Raises ExtensionMissing when it is not there -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                                   'or `make -C crowddynamics_b200/csrc`; there is no CPU fallback' % LIB_PATH)
        try:
            L = C.CDLL(LIB_PATH)
        except OSError as e:
            raise ExtensionMissing('cannot load %s: %s' % (LIB_PATH, e))
        L._signatures = _signatures(L)
        _lib = L
    return _lib


def check(rc):
    if rc == CDB_OK:
        return
    msg = load().cdb_last_error().decode('utf-8', 'replace')
    if rc == CDB_ERR_INVALID_TYPE:
        raise InvalidType(msg)
    if rc == CDB_ERR_INVALID_VALUE:
        raise InvalidValue(msg)
    if rc == CDB_ERR_CUDA:
        raise DeviceError(msg)
    raise CrowdDynamicsException('libcrowd_b200 error %d: %s' % (rc, msg))
