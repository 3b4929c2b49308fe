"""CPU restatement of the reference's model-based NPRACH agent, one cell at a time -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the package never does.
It follows NPRACH_THAgent.get_action (reference agent_nprach.py:62-146) over the tables that
tools/make_agent_tables.py derives from the agent's model (nb_iot_environment_b200/data/mb_agent.npz), with the
floating-point operations in the reference's order. Pinned by tests/golden/mb_agent.npz: traces of the unmodified
reference agent (tools/gen_golden_agent.py), closed loop on the reference environment and on synthetic info dicts.
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'nb_iot_environment_b200', 'data', 'mb_agent.npz')
_T = None


def tables():
    global _T
    if _T is None:
        z = np.load(_DATA)
        _T = {k: z[k] for k in z.files}
    return _T


class OracleNPRACHAgent:
    """NPRACH_THAgent (agent_nprach.py:41-146); `info` needs NPRACH_detection, NPRACH_collision (3 lists each),
    incoming (only its length is used) and distribution (15 floats)."""

    def __init__(self, buffer_size=100, margin=1.8, no_th=False):
        t = tables()
        self.t = t
        self.n = 0
        self.buffer_size = buffer_size
        self.k = 0
        self.no_th = no_th
        self.margin = margin
        self.lambda_estimate = 0.01
        self.msg3_detection = [0.9, 0.9, 0.9]
        self.conf = [2, 2, 2, 2, 2, 2]
        self.th_conf = [int(t['th_default'][0]), int(t['th_default'][1])]

    def set_parameter(self, margin):
        self.margin = margin

    def _estimate(self, d_list, c_list, nsc_index):          # detection_model.estimate_incoming_traffic (:45-55)
        if not len(d_list):
            return 0
        ml = self.t['ml_est'][nsc_index]
        est = [int(ml[d, c]) for d, c in zip(d_list, c_list)]
        return sum(est) / len(est)

    def _msg3_detection(self, dist):                         # random_access_model.msg3_detection_estimation (:86-131), th_1 != th_0
        acc, mar = [0, 0, 0], [0, 0, 0]
        det, cls = self.t['m3_det'], self.t['m3_ce']
        for i in range(14):
            acc[cls[i]] += dist[i] * float(det[i])
            mar[cls[i]] += dist[i]
        acc[0] += dist[14] * float(det[14])
        mar[0] += dist[14]
        default = [0.92, 0.9, 0.8]
        return [acc[ce] / mar[ce] if mar[ce] > 0 else default[ce] for ce in range(3)]

    def _configurator(self, msg3_detection, rates):          # configurator.configurator / find_conf (:161-181)
        conf = [0] * 6
        for ce in range(3):
            target = rates[ce] / msg3_detection[ce]
            n = int(self.t['cfg_n'][ce])
            j = n - 1
            for q in range(n):
                if self.t['cfg_rate'][ce, q] > target:
                    j = q
                    break
            conf[ce], conf[ce + 3] = int(self.t['cfg_conf'][ce, j, 0]), int(self.t['cfg_conf'][ce, j, 1])
        return conf

    def get_action(self, info):
        t = self.t
        min_value, max_value, step = (float(x) for x in t['consts'])
        self.n += 1
        rates = [0.0, 0.0, 0.0]
        self.lambda_estimate = 0.0
        for ce in range(3):
            period = int(t['period_list'][self.conf[ce + 3]])
            a_ = self._estimate(info['NPRACH_detection'][ce], info['NPRACH_collision'][ce], self.conf[ce])
            lam = a_ / period
            self.lambda_estimate += lam
            rates[ce] = self.margin * lam
        arrival_rate = sum(rates)
        if self.k < self.buffer_size:
            self.k += len(info['incoming'])
            self.msg3_detection = self._msg3_detection(info['distribution'])
            conf = self._configurator(self.msg3_detection, rates)
            th_conf = self.th_conf
        elif self.no_th:
            conf = self._configurator(self.msg3_detection, rates)
            th_conf = self.th_conf
        else:
            rc = t['rate_conf']
            if arrival_rate >= min_value and arrival_rate <= max_value:
                index = min(int((arrival_rate - min_value) // step) + 1, len(rc) - 1)
                row = rc[index]
            elif arrival_rate < min_value:
                row = rc[0]
            else:
                row = rc[-1]
            th_conf = [int(row[0]), int(row[1])]
            conf = [int(x) for x in row[2:]]
            if not (arrival_rate <= max_value):              # only this branch stores its choice (agent_nprach.py:109-113)
                self.conf = conf
                self.th_conf = th_conf
        return list(th_conf) + list(conf)
