"""Control and observation vocabulary of the NB-IoT cell, as the reference defines it in
system/parameters.py (line numbers cited per block).  These names are the API surface agents and
wrappers are written against; the numeric 3GPP tables used by the kernels live in
include/nbiot_params.h."""

N_users = 4            # parameters.py:14
Horizon = 40           # parameters.py:15
NPDCCH_sf = 8          # parameters.py:59
NPDCCH_period = 30     # parameters.py:251
max_p = 5              # parameters.py:33
N_actions = 23         # parameters.py:187

threshold_list = [-115, -114, -113, -112, -111, -110, -109, -108, -106, -104, -102, -100, -98, -96, 0]   # :37
max_th = len(threshold_list) - 1
th_values = threshold_list[:-1]
values_ = len(th_values)
th_pairs = [(th_values[i], th_values[j]) for i in range(values_) for j in range(i, values_)]          # :49-57
th_indexes = {(th_values[i], th_values[j]): [i, j] for i in range(values_) for j in range(i, values_)}

# parameters.py:106-134  name -> [length, normalisation constant]
obs_dict = {
    'total_ues': [1, 100], 'connection_time': [N_users, 1000], 'loss': [N_users, 50], 'sinr': [N_users, 70],
    'buffer': [N_users, 100], 'detection_ratios': [3, 1], 'colision_ratios': [3, 1], 'msg3_detection': [3, 1],
    'NPUSCH_occupation': [1, 1], 'NPRACH_occupation': [1, 1], 'av_delay': [1, 10000], 'ues_per_CE': [3, 20],
    'RAR_in': [3, 20], 'RAR_sent': [3, 20], 'RAR_detected': [3, 20], 'RAR_ids': [3, 20],
    'NPDCCH_sf_left': [1, NPDCCH_sf], 'sc_C0': [1, 3], 'sc_C1': [1, 3], 'sc_C2': [1, 3],
    'period_C0': [1, 7], 'period_C1': [1, 7], 'period_C2': [1, 7], 'th_C0': [1, values_], 'th_C1': [1, values_],
    'distribution': [values_ + 1, 1], 'carrier_state': [Horizon, 1],
}

# parameters.py:136-153
NPRACH_items = ['backoff', 'rar_window', 'mac_timer', 'transmax', 'panchor', 'th_C1', 'th_C0', 'period_C0', 'rep_C0',
                'sc_C0', 'period_C1', 'rep_C1', 'sc_C1', 'period_C2', 'rep_C2', 'sc_C2']

# parameters.py:159-185  (indices 1 and 2 are shared between the Scheduling and RAR_window states)
control_items = {
    'carrier': 0, 'id': 1, 'Imcs': 2, 'ce_level': 1, 'rar_Imcs': 2, 'N_ru': 3, 'delay': 4, 'sc': 5, 'Nrep': 6,
    'backoff': 7, 'rar_window': 8, 'mac_timer': 9, 'transmax': 10, 'panchor': 11, 'period_C0': 12, 'rep_C0': 13,
    'sc_C0': 14, 'period_C1': 15, 'rep_C1': 16, 'sc_C1': 17, 'period_C2': 18, 'rep_C2': 19, 'sc_C2': 20,
    'th_C1': 21, 'th_C0': 22,
}

# parameters.py:191-218  name -> [max with 1 carrier, max with 2 carriers]
control_max_values = {
    'carrier': [0, 1], 'id': [N_users - 1, N_users - 1], 'Imcs': [6, 6], 'ce_level': [2, 2], 'rar_Imcs': [2, 2],
    'N_ru': [7, 7], 'delay': [3, 3], 'sc': [3, 3], 'Nrep': [4, 4], 'backoff': [12, 12], 'rar_window': [7, 7],
    'mac_timer': [7, 7], 'transmax': [10, 10], 'panchor': [15, 15], 'period_C0': [max_p, max_p], 'rep_C0': [7, 7],
    'sc_C0': [3, 7], 'period_C1': [max_p, max_p], 'rep_C1': [7, 7], 'sc_C1': [3, 7], 'period_C2': [max_p, max_p],
    'rep_C2': [7, 7], 'sc_C2': [3, 7], 'th_C1': [max_th, max_th], 'th_C0': [max_th, max_th],
}

# parameters.py:221-247
control_default_values = {
    'carrier': 0, 'id': 0, 'Imcs': 2, 'ce_level': 1, 'rar_Imcs': 1, 'N_ru': 4, 'delay': 0, 'sc': 3, 'Nrep': 1,
    'backoff': 3, 'rar_window': 6, 'mac_timer': 2, 'transmax': 4, 'panchor': 12, 'period_C0': 3, 'rep_C0': 0,
    'sc_C0': 1, 'period_C1': 2, 'rep_C1': 0, 'sc_C1': 1, 'period_C2': 3, 'rep_C2': 2, 'sc_C2': 1, 'th_C1': 8, 'th_C0': 11,
}


def default_action_vector():
    """control_default_values scattered by control_items in dict order."""
    a = [0] * N_actions
    for name, idx in control_items.items():
        a[idx] = control_default_values[name]
    return a


# control_agents.py:29-71 (the reference's default six-agent split)
nprach_actions = ['rar_window', 'mac_timer', 'transmax', 'panchor', 'period_C0', 'rep_C0', 'sc_C0', 'period_C1', 'rep_C1',
                  'sc_C1', 'period_C2', 'rep_C2', 'sc_C2', 'th_C1', 'th_C0']
agents_conf = [
    {'id': 0, 'action_items': ['id'], 'obs_items': ['total_ues', 'connection_time', 'loss', 'sinr', 'buffer', 'carrier_state'],
     'next': 1, 'states': ['Scheduling']},
    {'id': 1, 'action_items': ['Imcs', 'Nrep'], 'obs_items': [], 'next': 2, 'states': ['Scheduling']},
    {'id': 2, 'action_items': ['carrier', 'delay', 'sc'], 'obs_items': [], 'next': -1, 'states': ['Scheduling']},
    {'id': 3, 'action_items': ['carrier', 'ce_level', 'rar_Imcs', 'delay', 'sc', 'Nrep'], 'obs_items': [], 'next': -1, 'states': ['RAR_window']},
    {'id': 4, 'action_items': ['backoff'], 'obs_items': [], 'next': -1, 'states': ['RAR_window_end']},
    {'id': 5, 'action_items': nprach_actions, 'obs_items': [], 'next': -1, 'states': ['NPRACH_update']},
]
