#!/usr/bin/env python3
"""One-off wider fuzz on the GPU: tests/test_fuzz.py's CUDA-vs-oracle drive over many more cases.  usage: python tools/gpu_fuzz.py lo hi [kernel] [decisions]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
lo, hi = int(sys.argv[1]), int(sys.argv[2])
os.environ['NBIOT_KERNEL'] = sys.argv[3] if len(sys.argv) > 3 else 'warp'
D = int(sys.argv[4]) if len(sys.argv) > 4 else 500
import __graft_entry__ as g  # noqa: E402
g.build_cuda()
import subprocess  # noqa: E402
subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
import test_fuzz as tf  # noqa: E402
bad, t0 = 0, time.time()
for k in range(lo, hi):
    try:
        _, b = tf._drive(tf._Oracles, tf._Cuda(), k, [8000 + k, 8500 + k, 8800 + k], D, 0.0, 'cuda')
        assert b.env.stats()['faulted'] == 0
        b.env.close()
    except AssertionError as e:
        bad += 1
        print('FAIL', k, str(e)[:600], flush=True)
print('gpu fuzz', os.environ['NBIOT_KERNEL'], 'cases', lo, hi, 'decisions', D, 'bad', bad, 'seconds', round(time.time() - t0))
