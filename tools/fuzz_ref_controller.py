#!/usr/bin/env python3
"""Controller-level fuzz against the LIVE unmodified reference: a random configuration (tests/fuzz_confs.py), one of the agent sets of
the reference's scripts, a random external agent, random fixed actions for the internal agents, random external actions; then
Controller.run_until with every agent acting (controller.py:156-195). The kernel source runs on the CPU (tests/_hostemu, the write
tables of nb_iot_environment_b200/controller.py applied by the core); records and rewards are compared at every external decision.
usage: python tools/fuzz_ref_controller.py <case> [ext steps] [lanes]      (build container only: needs /root/reference)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import conftest  # noqa: E402,F401
from fuzz_confs import controller_case  # noqa: E402
from golden_util import record_diff, REF_RTOL  # noqa: E402
from test_gpu_parity import NPRACH_AGENTS, NPUSCH_AGENTS  # noqa: E402
from nb_iot_environment_b200 import parameters as par  # noqa: E402
from nb_iot_environment_b200.controller import ControllerTables  # noqa: E402
from tools.refharness import load_reference, PhiloxShim, info_to_record  # noqa: E402
from hostemu_py import EmuBatch  # noqa: E402

k = int(sys.argv[1]); STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
lanes = sys.argv[3] if len(sys.argv) > 3 else 1
lanes = int(lanes) if str(lanes).isdigit() else lanes
conf, nc, cap, agents, ext, fixed, draw = controller_case(k)
seed = 9000 + k

m = load_reference(nc)
rng = PhiloxShim(seed)
node, perf, _, _ = m['create_system'](rng, dict(conf))
if cap:
    node.m.channel.capture_effect = True
ref_agents = [m['agents'].DummyAgent(dict(a)) for a in agents]
for a, f in zip(ref_agents, fixed):
    a.set_action(f)
ctrl = m['controller'].Controller(node, agents=ref_agents)
ctrl.reset()
ctrl.set_ext_agent(ext)

tabs = ControllerTables(agents, ext, nc)
for a, f in zip(tabs.agents, fixed):
    a.set_action(f)
b = EmuBatch(conf, [seed], n_carriers=nc, lanes=lanes, ctrl=tabs.build(), capture_effect=cap, ue_cap=conf['M'], grid_w=4096, hist_cap=1024)
b.reset()
recs, _ = b.advance(None, max_steps=0) if False else (b.records(), None)
names = agents[ext]['action_items']
t_last = 0
for s in range(STEPS):
    a = [draw(n) for n in names]
    obs, rew, _, _, info = ctrl.step(list(a))
    recs, _ = b.advance(np.asarray([a], dtype=np.uint8))
    want = info_to_record(dict(info, _full_conf=getattr(node, 'NPRACH_conf', None)) if info['state'] == 'RAR_window' else info, rew, rng)
    diff = record_diff(want, recs[0], rtol=REF_RTOL)
    assert not diff, f'case {k} agents {agents[ext]["states"]} ext {ext} step {s}: {diff[:3]}'
    assert not (recs[0]['fault'] & 0x1F), f'case {k}: fault {recs[0]["fault"]}'
    t_last = int(info['time'])
T = t_last + 4000
ctrl.run_until(T)
recs, _ = b.advance(None, until_time=T)
want = info_to_record(dict(ctrl.info, _full_conf=getattr(node, 'NPRACH_conf', None)) if ctrl.info['state'] == 'RAR_window' else ctrl.info, ctrl.reward, rng)
diff = record_diff(want, recs[0], rtol=REF_RTOL)
assert not diff, f'case {k} run_until: {diff[:3]}'
print('OK case', k, 'agent set', [a['states'] for a in agents][ext], 'ext', ext, 'carriers', nc, 'capture', cap, 'to t =', int(recs[0]['time']))
