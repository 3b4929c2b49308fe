"""The reference's CPU environment as a multiprocess vector env, for TIMING (bench.py's cpu_baseline / --impl reference legs).

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under nb_iot_environment_b200/ imports this module.

What runs is the UNMODIFIED reference: its system/, controller.py, control_agents.py and gym-system/ as staged by
`make -C oracle ref` into the git-ignored oracle/_ref/ (this repository holds no reference source; the GPU box gets
oracle/_ref/ with the snapshot, like the built .so files). One environment per process is mandatory -- the event list and the
subscriber table of the reference are module globals (system/event_manager.py:24, :88) -- which is also how the reference's own
scripts parallelise (scripts/nprach/evaluate_RL.py:169-170: a ProcessPoolExecutor over runs). Each worker builds the
scripts/nprach set-up of evaluate_RL.py:105-131 (create_system, four DummyAgents, Controller, set_ext_agent(3),
gym.make('gym_system:System-v1')) with the Philox shim as its random generator, waits at a barrier, and steps the external
actions of the GPU benchmark for its cell. Imports, construction and reset are outside the timed loop.

Missing from the image and replaced by import-only stand-ins (tools/ref_stubs): gymnasium (spaces, Env, make/register) and
matplotlib (imported by system/carrier.py:14 and system/perf_monitor.py:14, used by their animation code only)."""
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_STAGED = os.path.join(HERE, '_ref')


def reference_dir():
    """the staged copy (travels to the GPU box), else the read-only reference of the build container, else None"""
    for d in (REF_STAGED, os.environ.get('NBIOT_REFERENCE', '/root/reference')):
        if d and os.path.isfile(os.path.join(d, 'system', 'node_b.py')) and os.path.isfile(os.path.join(d, 'controller.py')):
            return d
    return None


def _worker(ref, conf, agents, ext_agent, seed, actions, warmup, barrier, out, idx):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, 'tools', 'ref_stubs'))
        sys.path.insert(0, os.path.join(ref, 'gym-system'))
        sys.path.insert(0, ref)
        from tools.refharness import PhiloxShim
        import gymnasium as gym
        from system.system_creator import create_system
        from controller import Controller
        from control_agents import DummyAgent
        node, _, _, _ = create_system(PhiloxShim(seed), dict(conf))
        ctrl = Controller(node, agents=[DummyAgent(dict(a)) for a in agents])
        ctrl.reset()
        ctrl.set_ext_agent(ext_agent)
        env = gym.make('gym_system:System-v1', system=ctrl)
        env.reset()
        trace = []                                   # (observation, reward) of every step: compared with the GPU's by the caller
        for a in actions[:warmup]:
            o, r, _, _, _ = env.step([int(x) for x in a])
            trace.append(([float(v) for v in o], float(r)))
        t_start = node.time
        barrier.wait(timeout=600)
        t0 = time.perf_counter()
        for a in actions[warmup:]:
            o, r, _, _, _ = env.step([int(x) for x in a])
            trace.append((o, r))
        dt = time.perf_counter() - t0
        trace = [([float(v) for v in o], float(r)) for o, r in trace]
        out.put((idx, dt, int(node.time - t_start), None, trace))
    except Exception as e:          # a worker that dies must not leave the others at the barrier forever
        try:
            barrier.abort()
        except Exception:
            pass
        out.put((idx, 0.0, 0, repr(e), None))


def run(conf, agents, ext_agent, seeds, actions, warmup, processes=None):
    """actions: uint8[warmup + steps][len(seeds)][k]. Returns dict(value = sum of simulated subframes / slowest worker's loop time,
    seconds, subframes, cores, per_core)."""
    ref = reference_dir()
    if ref is None:
        raise RuntimeError('no reference: run `make -C oracle ref` where /root/reference exists')
    P = len(seeds)
    ctx = mp.get_context('spawn')
    barrier, out = ctx.Barrier(P), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(ref, conf, agents, ext_agent, int(seeds[i]), actions[:, i].tolist(), warmup, barrier, out, i), daemon=True)
             for i in range(P)]
    for p in procs:
        p.start()
    import queue
    res, deadline = [], time.time() + 1200
    while len(res) < P:
        try:
            res.append(out.get(timeout=2))
        except queue.Empty:
            dead = [p for p in procs if not p.is_alive() and p.exitcode not in (0, None)]
            if dead or time.time() > deadline:       # a worker died before it could report (import error, spawn failure) or hangs
                for p in procs:
                    if p.is_alive():
                        p.kill()
                raise RuntimeError(f'reference worker exited with code {dead[0].exitcode}' if dead else 'reference workers timed out')
    for p in procs:
        p.join(timeout=30)
    errs = [r[3] for r in res if r[3]]
    if errs:
        raise RuntimeError('reference worker failed: ' + errs[0])
    seconds = max(r[1] for r in res)
    subframes = sum(r[2] for r in res)
    traces = [None] * P
    for r in res:
        traces[r[0]] = r[4]
    return {'traces': traces, 'value': subframes / max(seconds, 1e-9), 'seconds': seconds, 'subframes': subframes, 'cores': P,
            'per_core': sum(r[2] / max(r[1], 1e-9) for r in res) / P, 'reference_dir': 'oracle/_ref' if ref == REF_STAGED else ref}
