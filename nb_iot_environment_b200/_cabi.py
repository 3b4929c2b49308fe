"""ctypes binding of libnbiot_b200.so (C ABI in include/nbiot.h).  There is no CPU fallback: if the
CUDA library is missing or no GPU is visible, every compute entry point raises."""
import ctypes
import os

from .controller import NbiotCtrl
from .record import NbiotConf

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NBIOT_LIB') or os.path.join(_HERE, 'libnbiot_b200.so')   # NBIOT_LIB: an alternative build (tuning experiments)

# every symbol include/nbiot.h declares
SYMBOLS = ['nbiot_abi_version', 'nbiot_last_error', 'nbiot_state_bytes', 'nbiot_hdr_bytes', 'nbiot_hdr_task_bytes', 'nbiot_create', 'nbiot_destroy',
           'nbiot_set_controller', 'nbiot_reset', 'nbiot_advance', 'nbiot_advance_masked', 'nbiot_records_ptr', 'nbiot_hist_ptrs', 'nbiot_gather_obs',
           'nbiot_reduce_stats', 'nbiot_step_host', 'nbiot_stat_ptrs', 'nbiot_trace_ptrs', 'nbiot_spill_ptr', 'nbiot_trace_counts', 'nbiot_stat_counts', 'nbiot_stat_histogram', 'nbiot_set_event_log', 'nbiot_launch_count', 'nbiot_launch_geometry', 'nbiot_kernel_name', 'nbiot_instance_times',
           'nbiot_rng_u01', 'nbiot_rng_pair', 'nbiot_rng_normal', 'nbiot_rng_bounded',
           # include/nbiot_agent.h
           'nbiot_mb_agent_create', 'nbiot_mb_agent_destroy', 'nbiot_mb_agent_act', 'nbiot_mb_agent_faults']

N_STATS = 24
STAT_NAMES = {0: 'ue_arrivals', 3: 'ue_departures', 6: 'backoffs', 9: 'ue_connections', 10: 'error_count', 11: 'attempts_count',
              12: 'total_bits', 13: 'subframes', 14: 'events', 15: 'draws', 16: 'faulted', 17: 'live_ues', 18: 'sum_delay_x1000',
              19: 'truncated', 20: 'max_lookahead', 21: 'max_live_ues'}
N_SUM_STATS = 20      # entries [0, 20) are sums, the rest maxima


class ObsItem(ctypes.Structure):
    _fields_ = [('offset', ctypes.c_int32), ('kind', ctypes.c_int32), ('length', ctypes.c_int32),
                ('state_mask', ctypes.c_uint32), ('norm', ctypes.c_double)]


class MbTables(ctypes.Structure):
    """nbiot_mb_tables_t (include/nbiot_agent.h)"""
    _fields_ = [('ml_est', ctypes.c_void_p), ('cfg_rate', ctypes.c_void_p), ('cfg_conf', ctypes.c_void_p), ('cfg_n', ctypes.c_int32 * 3),
                ('m3_det', ctypes.c_double * 15), ('m3_ce', ctypes.c_uint8 * 15), ('pad0', ctypes.c_uint8), ('rate_conf', ctypes.c_void_p),
                ('min_value', ctypes.c_double), ('max_value', ctypes.c_double), ('step', ctypes.c_double),
                ('period_list', ctypes.c_int32 * 6), ('th_default', ctypes.c_uint8 * 2), ('pad1', ctypes.c_uint8 * 6)]


MB_STATE_BYTES = 56       # sizeof(nbiot_mb_state_t)


class NbiotError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NbiotError(f'{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback')
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32
    L.nbiot_abi_version.restype = i32
    L.nbiot_last_error.restype = ctypes.c_char_p
    L.nbiot_state_bytes.restype = ctypes.c_size_t
    L.nbiot_state_bytes.argtypes = [ctypes.POINTER(NbiotConf), i32]
    L.nbiot_hdr_bytes.restype = ctypes.c_size_t
    if hasattr(L, 'nbiot_hdr_task_bytes'):
        L.nbiot_hdr_task_bytes.restype = ctypes.c_size_t
    L.nbiot_create.argtypes = [ctypes.POINTER(NbiotConf), i32, i32, vp, ctypes.c_size_t, vp, vp, vp, ctypes.POINTER(NbiotCtrl), vp, ctypes.POINTER(vp)]
    L.nbiot_destroy.argtypes = [vp]
    L.nbiot_set_controller.argtypes = [vp, ctypes.POINTER(NbiotCtrl), vp]
    L.nbiot_reset.argtypes = [vp, vp]
    L.nbiot_advance.argtypes = [vp, vp, i32, i32, vp]
    L.nbiot_advance_masked.argtypes = [vp, vp, vp, i32, i32, vp]
    L.nbiot_records_ptr.argtypes = [vp, ctypes.POINTER(vp)]
    L.nbiot_hist_ptrs.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i32)]
    L.nbiot_stat_ptrs.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i32)]
    L.nbiot_stat_counts.argtypes = [vp, vp, vp]
    L.nbiot_trace_ptrs.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i32)]
    L.nbiot_trace_counts.argtypes = [vp, vp, vp]
    L.nbiot_spill_ptr.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.nbiot_stat_histogram.argtypes = [vp, i32, i32, i32, vp, vp]
    L.nbiot_gather_obs.argtypes = [vp, ctypes.POINTER(ObsItem), i32, vp, vp, vp]
    L.nbiot_reduce_stats.argtypes = [vp, vp, vp]
    L.nbiot_step_host.argtypes = [vp, vp, i32, i32, ctypes.POINTER(ObsItem), i32, vp, vp, vp]
    L.nbiot_set_event_log.argtypes = [vp, vp, i32]
    L.nbiot_launch_count.restype = ctypes.c_longlong
    if hasattr(L, 'nbiot_instance_times'):
        L.nbiot_instance_times.argtypes = [vp, vp]
    L.nbiot_launch_count.argtypes = [vp]
    L.nbiot_launch_geometry.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
    if hasattr(L, 'nbiot_kernel_name'):
        L.nbiot_kernel_name.restype = ctypes.c_char_p
        L.nbiot_kernel_name.argtypes = [vp]
    L.nbiot_rng_u01.restype = ctypes.c_double
    L.nbiot_rng_u01.argtypes = [u64, u64, u32]
    L.nbiot_rng_normal.restype = ctypes.c_double
    L.nbiot_rng_normal.argtypes = [u64, u64, u32]
    L.nbiot_rng_bounded.restype = u64
    L.nbiot_rng_bounded.argtypes = [u64, u64, u32, u64]
    L.nbiot_rng_pair.argtypes = [u64, u64, u32, vp]
    L.nbiot_mb_agent_create.argtypes = [vp, ctypes.POINTER(MbTables), i32, i32, ctypes.c_double, vp, vp, ctypes.POINTER(vp)]
    L.nbiot_mb_agent_destroy.argtypes = [vp]
    L.nbiot_mb_agent_act.argtypes = [vp, vp, vp, vp, vp]
    L.nbiot_mb_agent_faults.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong)]
    if L.nbiot_abi_version() != 1:
        raise NbiotError('libnbiot_b200.so ABI version mismatch: rebuild')
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise NbiotError(f'libnbiot_b200 error {rc}: {load().nbiot_last_error().decode()}')
