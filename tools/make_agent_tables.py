#!/usr/bin/env python3
"""Tables of the model-based NPRACH agent (reference agent_nprach.py:41-146) as flat arrays:
nb_iot_environment_b200/data/mb_agent.npz.  Build container only (reads /root/reference to CHECK itself).

The agent's model is evaluated at import time from three data files and a handful of constants; everything it
consults at run time is a function of small integers, so it becomes tables:

  ml_est   int16[4][D][C]   ml_based_estimator(detections, collisions, N) for N = N_sc_list = 3, 6, 9, 12
                            (detection_model.py:25-43).  Computed here from the conditional probability
                            P(n, s, c, r) of Vogt / "Efficient Random Access Channel Evaluation and Load Estimation
                            in LTE-A With Massive MTC" by its recursion (detection_model.py:84-107) -- NOT read from
                            model/p_detections_collisions.pickle; the pickle is only used to check the values.
                            Only r in {12, 24, 36, 48} has probabilities, and the agent passes N_sc (3..12), not the
                            preamble count, so N = 12 is the only row that consults the model (reference behaviour).
  cfg_rate float64[3][K], cfg_conf uint8[3][K][2], cfg_n int32[3]
                            the rate-ordered Pareto lists of (nsc index, period index) per CE level that
                            configurator.configurator searches (configurator.py:104-181) for the agent's fixed
                            thresholds (agent_nprach.py:28-34: th_C1 = -106, th_C0 = -100 -> msg3 RUs 1,4,16,
                            NPRACH subframes 6,6,45)
  m3_det   float64[15], m3_ce uint8[15]
                            msg3_detection_estimation (random_access_model.py:86-131) for those thresholds is a
                            weighted mean of per-threshold detection rates: coefficient and CE class per histogram bin
                            (model/th_detection_rates.pickle is reference data)
  rate_conf uint8[186][8]   model/rate_to_conf.json (reference data): arrival-rate bucket -> (th_C1 idx, th_C0 idx,
                            nsc0..2, period0..2)
"""
import json, math, os, pickle, sys
from functools import lru_cache
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
REF = os.environ.get('NBIOT_REFERENCE', '/root/reference')
OUT = os.path.join(ROOT, 'nb_iot_environment_b200', 'data', 'mb_agent.npz')
D_MAX, C_MAX = 128, 64

# constants of system/parameters.py used by the model (cited where they are used)
THRESHOLD_LIST = [-115, -114, -113, -112, -111, -110, -109, -108, -106, -104, -102, -100, -98, -96, 0]   # :37
TH_VALUES = THRESHOLD_LIST[:-1]                                                                            # :46
EXT_TH_VALUES = [THRESHOLD_LIST[0] - 1] + THRESHOLD_LIST[:-1]                                              # random_access_model.py:26
N_SC_LIST = [3, 6, 9, 12]                                                                                  # :340
PERIOD_LIST = [80, 160, 240, 320, 640, 1280]                                                               # :339
NPDCCH_PERIOD, NPDCCH_SF, NPDCCH_SF_PER_CE = 30, 8, [1, 2, 4]                                              # :251-252, :58
RAR_WINDOW_SIZE = 8     # RAR_WindowSize_list[control_default_values["rar_window"] = 6] (parameters.py:232, :321; checked below)
TH_C1, TH_C0 = -106, -100
MSG3_RUS, N_SFS = [1, 4, 16], [6, 6, 45]


@lru_cache(maxsize=None)
def feasible(n, s, c, r):                    # detection_model.py:61-81
    if c < 0 or s < 0 or s + c > r:
        return False
    return any(s + beta * c == n for beta in range(2, 2 * r + 1))


@lru_cache(maxsize=None)
def P(n, s, c, r):                           # detection_model.py:84-107 (same three terms, same order)
    if n == 0 and s == 0 and c == 0:
        return 1
    if not feasible(n, s, c, r):
        return 0
    return (r - (s - 1 + c)) / r * P(n - 1, s - 1, c, r) + (s + 1) / r * P(n - 1, s + 1, c - 1, r) + c / r * P(n - 1, s, c, r)


def tables_P():
    sys.setrecursionlimit(10000)
    pdc, pd = {}, {}
    for r in (12, 24, 36, 48):               # detection_model.py:110-123
        for n in range(2 * r + 1):
            for s in range(r + 1):
                acc = 0
                for c in range(r + 1):
                    p = P(n, s, c, r)
                    if p > 0:
                        acc += p
                        pdc[(n, s, c, r)] = p
                pd[(s, n, r)] = acc
    return pdc, pd


def ml_estimator(pdc, d, c, N):             # detection_model.py:25-43
    n_opt = max(0, d - 1)
    p_max = pdc.get((n_opt, d, c, N), 0.0)
    p_prev = p_max
    for n in range(n_opt, 2 * N + 1):
        p = pdc.get((n, d, c, N), 0.0)
        if p > p_max:
            p_max, n_opt = p, n
        if p < p_prev:
            return n_opt
        p_prev = p
    return n_opt


def rate_lists(pd):
    """configurator.create_dictionaries / remove_items / create_ordered_dict (configurator.py:104-158)"""
    out = []
    for ce in range(3):
        RUs, N_sf = MSG3_RUS[ce], N_SFS[ce]
        rate, res = {}, {}
        for p_i, period in enumerate(PERIOD_LIST):
            for n_i, nsc in enumerate(N_SC_LIST):
                max_sfs = period - N_sf - 4                                      # :55-63
                RAR_p = min(np.floor(max_sfs / NPDCCH_PERIOD), RAR_WINDOW_SIZE)
                DCIs = NPDCCH_SF * RAR_p
                RAR_window_sf = RAR_p * NPDCCH_PERIOD
                k_max = min(np.floor(max(0, RAR_window_sf) / RUs), np.floor(max(0, DCIs) / NPDCCH_SF_PER_CE[ce]))   # :36-40
                r = 4 * nsc                                                      # :68-86
                best = 0
                for n in range(1, r + 1):
                    avg = 0
                    for k in range(1, n + 1):
                        avg += min(k, k_max) * pd[(k, n, r)]
                    if avg > best:
                        best = avg
                v = best / period
                rate[(n_i, p_i)] = v
                res[(n_i, p_i)] = (nsc * N_sf) / (12 * period) + v * RUs         # :97-102
        drop = set()
        for k1 in rate:
            for k2 in rate:
                if rate[k2] > rate[k1] and res[k2] <= res[k1]:
                    drop.add(k1)
                    break
        for k in drop:
            del rate[k]
        out.append(sorted(rate.items(), key=lambda x: x[1]))
    return out


def msg3_coefficients(th_det):
    """per histogram bin i: the CE class and the detection rate msg3_detection_estimation multiplies its probability with"""
    msg3_conf = {-116: (2, 1, 16), -115: (2, 1, 16), -114: (2, 1, 16), -113: (2, 1, 16), -112: (2, 1, 16), -111: (2, 1, 8), -110: (2, 1, 8),
                 -109: (2, 1, 8), -108: (2, 1, 8), -107: (2, 1, 4), -106: (2, 1, 4), -105: (2, 1, 4), -104: (2, 1, 4), -103: (2, 1, 2),
                 -102: (2, 1, 2), -101: (2, 1, 2)}                                # parameters.py:74-91
    no_rep_msg3_th = -100
    det, cls = np.zeros(15), np.zeros(15, np.uint8)
    ce, conf, th = 0, None, None
    for i, th in enumerate(TH_VALUES):                                           # random_access_model.py:92-114
        if th <= TH_C1:
            ce, conf = 2, msg3_conf[TH_VALUES[0]]
        elif th > TH_C1 and th <= TH_C0:
            ce, conf = 1, ((2, 1, 1) if TH_C1 >= no_rep_msg3_th else msg3_conf[TH_C1])
        if th > TH_C0:
            ce, conf = 0, ((2, 1, 1) if TH_C0 >= no_rep_msg3_th else msg3_conf[TH_C0])
        det[i], cls[i] = th_det[EXT_TH_VALUES[i]][conf], ce
    det[14], cls[14] = th_det[th][conf], 0                                       # :116-117: the marginal bin, last th / conf
    return det, cls


def main():
    pdc, pd = tables_P()
    ml = np.zeros((4, D_MAX, C_MAX), np.int16)
    for ni, N in enumerate(N_SC_LIST):
        for d in range(D_MAX):
            for c in range(C_MAX):
                ml[ni, d, c] = ml_estimator(pdc, d, c, N)
    lists = rate_lists(pd)
    K = max(len(l) for l in lists)
    cfg_rate = np.full((3, K), np.inf); cfg_conf = np.zeros((3, K, 2), np.uint8); cfg_n = np.zeros(3, np.int32)
    for ce, l in enumerate(lists):
        cfg_n[ce] = len(l)
        for j, ((n_i, p_i), v) in enumerate(l):
            cfg_rate[ce, j] = v; cfg_conf[ce, j] = (n_i, p_i)

    # ---- reference data files + self-check against the reference's own code
    th_det = pickle.load(open(os.path.join(REF, 'model', 'th_detection_rates.pickle'), 'rb'))
    m3_det, m3_ce = msg3_coefficients(th_det)
    rtc = json.load(open(os.path.join(REF, 'model', 'rate_to_conf.json')))
    th_index = {}
    for i, t1 in enumerate(TH_VALUES):
        for j in range(i, len(TH_VALUES)):
            th_index[(t1, TH_VALUES[j])] = (i, j)                                # parameters.py:47-57
    rate_conf = np.asarray([list(th_index[tuple(th)]) + list(conf) for th, conf in rtc], np.uint8)
    assert rate_conf.shape == (186, 8)

    ref_pdc = pickle.load(open(os.path.join(REF, 'model', 'p_detections_collisions.pickle'), 'rb'))
    ref_pd = pickle.load(open(os.path.join(REF, 'model', 'p_detections.pickle'), 'rb'))
    assert set(ref_pdc) == set(pdc) and all(ref_pdc[k] == pdc[k] for k in pdc), 'P(n,s,c,r) differs from the reference pickle'
    assert all(ref_pd[k] == pd[k] for k in ref_pd), 'p_detections differs from the reference pickle'
    sys.path.insert(0, ROOT)
    from tools.refharness import load_reference
    load_reference(1)
    cwd = os.getcwd(); os.chdir(REF)
    import agent_nprach as A, model.configurator as con, model.random_access_model as ra, model.detection_model as dm, system.parameters as par
    os.chdir(cwd)
    assert (A.th_C1, A.th_C0, A.msg3_RUs, A.N_sfs_list) == (TH_C1, TH_C0, MSG3_RUS, N_SFS)
    assert con.RAR_window_size == RAR_WINDOW_SIZE and par.NPDCCH_period == NPDCCH_PERIOD and par.NPDCCH_sf == NPDCCH_SF
    assert par.N_sc_list == N_SC_LIST and par.period_list == PERIOD_LIST and (A.min_value, A.max_value, A.step) == (0.0035, 0.374, 0.002)
    con.update_parameters(A.msg3_RUs, A.N_sfs_list, [0.9, 0.9, 0.9])
    for ce in range(3):
        assert [(k, v) for k, v in con.rate_dictionaries[ce]] == lists[ce], f'rate list of CE {ce} differs from configurator.py'
    for ni, N in enumerate(N_SC_LIST):
        for d in range(0, D_MAX, 1):
            for c in range(0, C_MAX, 3):
                assert dm.ml_based_estimator(d, c, N) == ml[ni, d, c]
    rng = np.random.default_rng(5)
    for _ in range(200):
        dist = rng.random(15); dist /= dist.sum()
        if rng.random() < 0.3:
            dist[rng.integers(0, 15, size=6)] = 0.0
        got = ra.msg3_detection_estimation(TH_C1, TH_C0, list(dist))
        acc, mar = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
        for i in range(14):
            acc[m3_ce[i]] += dist[i] * m3_det[i]; mar[m3_ce[i]] += dist[i]
        acc[0] += dist[14] * m3_det[14]; mar[0] += dist[14]
        mine = [acc[ce] / mar[ce] if mar[ce] > 0 else [0.92, 0.9, 0.8][ce] for ce in range(3)]
        assert mine == got, (mine, got)
    np.savez_compressed(OUT, ml_est=ml, cfg_rate=cfg_rate, cfg_conf=cfg_conf, cfg_n=cfg_n, m3_det=m3_det, m3_ce=m3_ce, rate_conf=rate_conf,
                        consts=np.asarray([0.0035, 0.374, 0.002]), n_sc_list=np.asarray(N_SC_LIST), period_list=np.asarray(PERIOD_LIST),
                        th_default=np.asarray([th_index[(TH_C1, TH_C0)][0], th_index[(TH_C1, TH_C0)][1]], np.uint8))
    print('wrote', OUT, 'ml_est', ml.shape, 'rate lists', cfg_n.tolist(), 'P entries', len(pdc))


if __name__ == '__main__':
    main()
