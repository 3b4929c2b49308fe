"""Host-side mirrors of the C structs in include/nbiot_record.h and include/nbiot_conf.h.

`RECORD_DTYPE` is the per-instance decision record: the reward plus the `info` dict the
reference assembles in NodeB.get_node_info (reference system/node_b.py:278-418; schema in
SURVEY.md App. D).  `make_conf` turns the reference's `conf` dict
(system/system_creator.py:24-44) into the C struct.
"""
import ctypes

import numpy as np

MAX_OCC = 48
STEP_LIST = 16
N_ACTIONS = 23

STATE_NAMES = ['Scheduling', 'NPRACH_update', 'RAR_window', 'RAR_window_end', 'Idle']   # NB_ST_* order
STATE_IDS = {n: i for i, n in enumerate(STATE_NAMES)}

FAULT_UE_SLOTS, FAULT_LOOKAHEAD, FAULT_ACTION, FAULT_MCS_KEY, FAULT_EVENT_SLOTS, FAULT_LIST_TRUNC, FAULT_STALL = 1, 2, 4, 8, 16, 32, 64
FAULT_FATAL_MASK = 0x5F

# reward_criteria names of perf_monitor.py:29-38 -> NB_RW_*
REWARD_IDS = {'throughput': 0, 'average_delay': 1, 'accumulated_delay': 2, 'users': 3,
              'invd_users': 4, 'log_invd_users': 5}

# NPRACH_items order (parameters.py:136-153)
NPRACH_ITEMS = ['backoff', 'rar_window', 'mac_timer', 'transmax', 'panchor', 'th_C1', 'th_C0',
                'period_C0', 'rep_C0', 'sc_C0', 'period_C1', 'rep_C1', 'sc_C1', 'period_C2', 'rep_C2', 'sc_C2']

RECORD_DTYPE = np.dtype([
    ('reward', '<f8'), ('loss', '<f8', 4), ('sinr', '<f8', 4),
    ('detection_ratios', '<f8', 3), ('colision_ratios', '<f8', 3), ('msg3_detection', '<f8', 3),
    ('msg3_sent', '<f8', 3), ('msg3_received', '<f8', 3), ('distribution', '<f8', 15),
    ('RA_occupation', '<f8'), ('NPUSCH_occupation', '<f8'), ('NPRACH_occupation', '<f8'), ('av_delay', '<f8'),
    ('time', '<i4'), ('state', '<i4'), ('total_ues', '<i4'), ('fault', '<i4'), ('n_sel', '<i4'),
    ('ues', '<i4', 4), ('connection_time', '<i4', 4), ('buffer', '<i4', 4),
    ('n_rx', '<i4'), ('n_err', '<i4'), ('n_unfit', '<i4'), ('n_acc', '<i4'),
    ('rx_id', '<i4', STEP_LIST), ('rx_delay', '<i4', STEP_LIST), ('err_id', '<i4', STEP_LIST),
    ('unfit_id', '<i4', STEP_LIST), ('acc_id', '<i4', STEP_LIST), ('acc_delay', '<i4', STEP_LIST),
    ('ues_per_CE', '<i4', 3), ('RAR_in', '<i4', 3), ('RAR_sent', '<i4', 3), ('RAR_ids', '<i4', 3),
    ('RAR_detected', '<i4', 3), ('RAR_failed', '<i4', 3),
    ('valid_CE_control', '<i4'), ('NPDCCH_sf_left', '<i4'), ('total_departures', '<i4'),
    ('action', '<i4', 3), ('nprach_conf', '<i4', 16), ('preambles', '<i4', 3),
    ('departures', '<i4'), ('n_incoming', '<i4'), ('n_occ', '<i4', 3),
    ('draws_lo', '<i4'), ('draws_hi', '<i4'), ('events', '<i4'),
    ('det_hist', '<i2', (3, MAX_OCC)), ('col_hist', '<i2', (3, MAX_OCC)),
    ('carrier_busy', 'u1', 40),
])
assert RECORD_DTYPE.itemsize == 1632

FLOAT_FIELDS = [n for n in RECORD_DTYPE.names if RECORD_DTYPE[n].base == np.dtype('<f8')]
INT_FIELDS = [n for n in RECORD_DTYPE.names if n not in FLOAT_FIELDS]
# diagnostics that the Python reference cannot report the same way
NON_PARITY_FIELDS = ('events', 'fault')


class NbiotConf(ctypes.Structure):
    _fields_ = [('ratio', ctypes.c_double), ('max_d', ctypes.c_double),
                ('M', ctypes.c_int32), ('buffer_lo', ctypes.c_int32), ('buffer_hi', ctypes.c_int32),
                ('reward_criteria', ctypes.c_int32), ('sort_by_loss', ctypes.c_int32),
                ('sc_adjustment', ctypes.c_int32), ('mcs_automatic', ctypes.c_int32),
                ('ce_mcs_automatic', ctypes.c_int32), ('tx_all_buffer', ctypes.c_int32),
                ('random_traffic', ctypes.c_int32), ('n_carriers', ctypes.c_int32),
                ('ue_cap', ctypes.c_int32), ('grid_w', ctypes.c_int32), ('hist_cap', ctypes.c_int32),
                ('stat_cap', ctypes.c_int32), ('clustered', ctypes.c_int32), ('cluster_lo', ctypes.c_int32), ('cluster_hi', ctypes.c_int32),
                ('cluster_prob', ctypes.c_double), ('cluster_range', ctypes.c_double),
                ('trace_cap', ctypes.c_int32), ('trace_mask', ctypes.c_int32), ('capture_effect', ctypes.c_int32), ('kernel', ctypes.c_int32)]


# nbiot_trace_rec_t (include/nbiot_conf.h): one sample of PerfMonitor's lists
TRACE_DTYPE = np.dtype([('t', '<i4'), ('a', '<i4'), ('b', '<i4'), ('type', 'u1'), ('ce', 'u1'), ('pad', '<u2')])
assert TRACE_DTYPE.itemsize == 16
TR_NPRACH, TR_MSG3, TR_RARWIN, TR_RES_MSG3, TR_RES_NPUSCH, TR_RES_NPRACH, TR_DEPART = 1, 2, 4, 8, 16, 32, 64
TRACE_MASK_TRACES = TR_NPRACH | TR_MSG3                                            # conf['traces'] (perf_monitor.py:61-69)
TRACE_MASK_STATISTICS = TR_RARWIN | TR_RES_MSG3 | TR_RES_NPUSCH | TR_RES_NPRACH | TR_DEPART   # conf['statistics'] (:71-79)
DEFAULT_TRACE_CAP = 8192


def samples_to_lists(samples):
    """samples of one cell in order of occurrence (TRACE_DTYPE) -> PerfMonitor's lists under the reference's attribute names
    (perf_monitor.py:61-79): `traces` lists hold (value, t) tuples, the RAR / resource histories plain values, npusch_delivered
    (count, t) with the departures of one subframe and CE level merged (:164-168)."""
    out = {k: [[], [], []] for k in ('rar_detections', 'rar_errors', 'rar_collisions', 'nprach_arrivals', 'nprach_detections',
                                     'RAR_in', 'RAR_attmp', 'RAR_sent', 'RAR_detected', 'npusch_delivered')}
    out.update({k: [] for k in ('t_rar_detections', 't_nprach_arrivals', 't_nprach_detections',
                                'msg3_resources_history', 'NPUSCH_resources_history', 'NPRACH_resources_history')})
    for r in samples:
        ty, ce, t, a, b = int(r['type']), int(r['ce']), int(r['t']), int(r['a']), int(r['b'])
        if ty == TR_NPRACH:
            out['nprach_arrivals'][ce].append((a, t)); out['nprach_detections'][ce].append((b, t))
            out['t_nprach_arrivals'].append((a, t)); out['t_nprach_detections'].append((b, t))
        elif ty == TR_MSG3:
            out['rar_detections'][ce].append((a, t)); out['rar_errors'][ce].append((b & 1, t)); out['rar_collisions'][ce].append((b >> 1, t))
            out['t_rar_detections'].append((a, t))
        elif ty == TR_RARWIN:
            out['RAR_in'][ce].append(a & 0xFFFF); out['RAR_attmp'][ce].append(a >> 16)
            out['RAR_sent'][ce].append(b & 0xFFFF); out['RAR_detected'][ce].append(b >> 16)
        elif ty == TR_RES_MSG3:
            out['msg3_resources_history'].append(a)
        elif ty == TR_RES_NPUSCH:
            out['NPUSCH_resources_history'].append(a)
        elif ty == TR_RES_NPRACH:
            out['NPRACH_resources_history'].append(a)
        elif ty == TR_DEPART:
            lst = out['npusch_delivered'][ce]
            if lst and lst[-1][1] == t:
                lst[-1] = (lst[-1][0] + 1, t)
            else:
                lst.append((1, t))
    return out


_UNSUPPORTED_TRUE = ('animate_carrier', 'animate_stats', 'resource_monitor')
_KNOWN = {'ratio', 'M', 'buffer_range', 'reward_criteria', 'sort_criterium', 'levels', 'max_d', 'sc_adjustment',
          'mcs_automatic', 'ce_mcs_automatic', 'tx_all_buffer', 'animate_carrier', 'statistics', 'animate_stats',
          'traces', 'resource_monitor', 'random', 'clustered', 'cluster_ues', 'cluster_prob', 'cluster_range'}


DEFAULT_STAT_CAP = 2048


KERNEL_IDS = {None: 0, 'auto': 0, 'warp': 1, 'thread': 2, 'thread_voted': 3}      # NBIOT_KERNEL_* (include/nbiot_conf.h)


def make_conf(conf, n_carriers=1, ue_cap=None, grid_w=1024, hist_cap=256, stat_cap=None, trace_cap=None, capture_effect=False, kernel=None):
    """conf dict with the reference's keys and defaults (system_creator.py:24-44) -> NbiotConf."""
    unknown = set(conf) - _KNOWN
    if unknown:
        raise KeyError(f'unknown conf keys {sorted(unknown)}')
    for k in _UNSUPPORTED_TRUE:
        if conf.get(k, False):
            raise NotImplementedError(f"conf['{k}']=True is outside the batched hot path (SURVEY.md section 8)")
    crit = conf['reward_criteria']
    if crit not in REWARD_IDS:
        # 'average_total_delay' (perf_monitor.py:264-274) is the one criterion of the reference that is absent here: the reference itself
        # fails on it -- it looks every served UE up in ue_access_times, which clear_info (:96-103) empties at each scheduling step, so the
        # first reception of a UE that connected in an earlier step raises KeyError (step 2 of configs[0] with the default agents)
        raise NotImplementedError(f"reward_criteria '{crit}' is not supported")
    sort = conf.get('sort_criterium', 't_connection')
    if sort not in ('t_connection', 'loss'):
        raise NotImplementedError(f"sort_criterium '{sort}' is not supported")
    if n_carriers not in (1, 2):
        raise ValueError('n_carriers must be 1 or 2')
    if grid_w & (grid_w - 1) or grid_w < 256 or grid_w > 4096:
        raise ValueError('grid_w must be a power of two in 256..4096')
    M = int(conf['M'])
    lo_b, hi_b = int(conf['buffer_range'][0]), int(conf['buffer_range'][1])
    if M < 1:
        raise ValueError("conf['M'] must be >= 1")
    if not 0 <= float(conf['ratio']):
        raise ValueError("conf['ratio'] must be >= 0 (values above 1 act as 1, ue_generator.py:39)")
    if not (0 < lo_b <= 65535 and hi_b <= 65535):        # hi <= lo is legal in the reference: every buffer is lo (ue_generator.py:98-104)
        raise ValueError("conf['buffer_range'] must lie in 1..65535 (a UE's buffer is held in 16 bits)")
    c = NbiotConf()
    c.ratio = float(conf['ratio'])
    c.max_d = float(conf.get('max_d') or 0.0)
    c.M = M
    c.buffer_lo, c.buffer_hi = int(conf['buffer_range'][0]), int(conf['buffer_range'][1])
    c.reward_criteria = REWARD_IDS[crit]
    c.sort_by_loss = int(sort == 'loss')
    c.sc_adjustment = int(conf.get('sc_adjustment', True))
    c.mcs_automatic = int(conf.get('mcs_automatic', True))
    c.ce_mcs_automatic = int(conf.get('ce_mcs_automatic', True))
    c.tx_all_buffer = int(conf.get('tx_all_buffer', True))
    c.random_traffic = int(conf.get('random', False))
    c.n_carriers = n_carriers
    c.ue_cap = int(ue_cap if ue_cap is not None else M)
    c.grid_w = int(grid_w)
    c.hist_cap = int(hist_cap)
    # conf['statistics'] keeps the per-departure histories (perf_monitor.py:61-84): stat_cap entries per cell
    c.stat_cap = int(stat_cap) if stat_cap is not None else (DEFAULT_STAT_CAP if conf.get('statistics', False) else 0)
    if c.stat_cap < 0:
        raise ValueError('stat_cap must be >= 0')
    # conf['traces'] / conf['statistics']: PerfMonitor's sample lists (perf_monitor.py:61-79) as a ring of trace_cap samples per cell
    c.trace_mask = (TRACE_MASK_TRACES if conf.get('traces', False) else 0) | (TRACE_MASK_STATISTICS if conf.get('statistics', False) else 0)
    c.trace_cap = (int(trace_cap) if trace_cap is not None else DEFAULT_TRACE_CAP) if c.trace_mask else 0
    if c.trace_cap < 0:
        raise ValueError('trace_cap must be >= 0')
    if not c.trace_cap:
        c.trace_mask = 0
    c.capture_effect = int(bool(capture_effect))        # Channel(capture_effect=True), channel.py:124-126
    if kernel not in KERNEL_IDS:
        raise ValueError(f'kernel must be one of {sorted(k for k in KERNEL_IDS if k)} or None')
    c.kernel = KERNEL_IDS[kernel]
    # clustered placement (system_creator.py:41-44, population.py:57-63)
    c.clustered = int(bool(conf.get('clustered', False)))
    lo, hi = conf.get('cluster_ues', [10, 30])
    c.cluster_lo, c.cluster_hi = int(lo), int(hi)
    if c.clustered and not c.cluster_hi > c.cluster_lo:
        raise ValueError('cluster_ues must be [MIN, MAX] with MAX > MIN (rng.integers(MIN, MAX))')
    c.cluster_prob = float(conf.get('cluster_prob', 0.5))
    c.cluster_range = float(conf.get('cluster_range', 0.5))
    return c




def record_to_info(rec, hist=None, spill=None):
    """One record -> the reference's `info` dict (same keys and value types as node_b.py:299-408)."""
    state = STATE_NAMES[int(rec['state'])]
    info = {'time': int(rec['time']), 'state': state, 'total_ues': int(rec['total_ues']),
            'carrier_state': 1.0 - rec['carrier_busy'].astype(np.float64) / 12.0}
    if state in ('Scheduling', 'RAR_window_end'):
        n = int(rec['n_sel'])
        info['ues'] = [int(v) for v in rec['ues'][:n]]
        info['connection_time'] = [int(v) if i < n else 0.0 for i, v in enumerate(rec['connection_time'])]
        info['loss'] = [float(v) for v in rec['loss']]
        info['sinr'] = [float(v) for v in rec['sinr']]
        info['buffer'] = [int(v) if i < n else 0.0 for i, v in enumerate(rec['buffer'])]
        def full(field, count, row):
            """the record's 16 entries, then the side buffer's (spill: int32[6][spill cap] of this cell) when the list is longer"""
            n_ = int(rec[count])
            v = [int(x) for x in rec[field][:min(n_, STEP_LIST)]]
            if n_ > STEP_LIST and spill is not None:
                v += [int(x) for x in spill[row][:n_ - STEP_LIST]]
            return v
        info['receptions'] = dict(zip(full('rx_id', 'n_rx', 0), full('rx_delay', 'n_rx', 1)))
        info['errors'] = full('err_id', 'n_err', 2)
        info['unfit'] = full('unfit_id', 'n_unfit', 3)
        info['access'] = dict(zip(full('acc_id', 'n_acc', 4), full('acc_delay', 'n_acc', 5)))
    elif state == 'RAR_window':
        for key in ('ues_per_CE', 'RAR_in', 'RAR_sent', 'RAR_ids', 'RAR_detected', 'RAR_failed'):
            info[key] = [int(v) for v in rec[key]]
        info['valid_CE_control'] = bool(rec['valid_CE_control'])
        info['NPDCCH_sf_left'] = int(rec['NPDCCH_sf_left'])
        info['total_departures'] = int(rec['total_departures'])
        conf = dict(zip(NPRACH_ITEMS, (int(v) for v in rec['nprach_conf'])))
        for key in ('sc_C0', 'sc_C1', 'sc_C2', 'period_C0', 'period_C1', 'period_C2', 'th_C1', 'th_C0'):
            info[key] = conf[key]
        info['action'] = [int(v) for v in rec['action']]
        info['final_action'] = info['action']
    else:
        for key in ('detection_ratios', 'colision_ratios', 'msg3_detection', 'msg3_sent', 'msg3_received', 'distribution'):
            info[key] = [float(v) for v in rec[key]]
        info['preambles'] = [int(v) for v in rec['preambles']]
        info['NPRACH_detection'] = [[int(v) for v in rec['det_hist'][ce][:min(int(rec['n_occ'][ce]), MAX_OCC)]] for ce in range(3)]
        info['NPRACH_collision'] = [[int(v) for v in rec['col_hist'][ce][:min(int(rec['n_occ'][ce]), MAX_OCC)]] for ce in range(3)]
        for key in ('RA_occupation', 'NPUSCH_occupation', 'NPRACH_occupation', 'av_delay'):
            info[key] = float(rec[key])
        info['departures'] = int(rec['departures'])
        if hist is not None:
            info['incoming'], info['delays'], info['service_times'] = hist
        info['NPRACH conf'] = dict(zip(NPRACH_ITEMS, (int(v) for v in rec['nprach_conf'])))
    return info
