#!/usr/bin/env python3
"""The LIVE reference's model-based NPRACH agent (agent_nprach.NPRACH_THAgent with model/*.py and its pickles) against its CPU restatement
(oracle/mb_agent_oracle.py) on seeded random inputs (tests/fuzz_agent.py): the 8-item action, the arrival estimate, the sample counter and the
msg3 detection estimate of every call, floats bit for bit.
usage: python tools/fuzz_ref_agent.py <first case> <last case>      (build container only: needs /root/reference)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from fuzz_agent import agent_sequence  # noqa: E402
from test_gpu_parity import NPRACH_AGENTS  # noqa: E402
from tools.refharness import load_reference  # noqa: E402
from oracle.mb_agent_oracle import OracleNPRACHAgent  # noqa: E402

load_reference(1)
cwd = os.getcwd(); os.chdir(os.environ.get('NBIOT_REFERENCE', '/root/reference'))      # the agent's modules open ./model/*.pickle at import
import agent_nprach as A  # noqa: E402
os.chdir(cwd)
lo, hi = int(sys.argv[1]), int(sys.argv[2])
branches = set()
for k in range(lo, hi):
    no_th, seq = agent_sequence(k)
    ref = A.NPRACH_THAgent(dict(NPRACH_AGENTS[3]), ['departures', 'NPRACH_occupation', 'service_times', 'beta'], no_th=no_th)
    orc = OracleNPRACHAgent(no_th=no_th)
    for s, info in enumerate(seq):
        ref.set_parameter(info['margin']); orc.set_parameter(info['margin'])
        k_before = orc.k
        want = [int(x) for x in ref.get_action(None, 0.0, info, None)]
        got = orc.get_action(info)
        assert got == want, (k, s, got, want)
        assert orc.lambda_estimate == ref.lambda_estimate and orc.k == ref.k, (k, s, orc.lambda_estimate, ref.lambda_estimate)
        assert [float(x) for x in orc.msg3_detection] == [float(x) for x in ref.msg3_detection], (k, s)
        branches.add('gather' if k_before < 100 else 'no_th' if no_th else 'table')
print('OK agent cases', lo, hi, 'branches', sorted(branches))
