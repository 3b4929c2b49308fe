#!/usr/bin/env python3
"""One fuzz case (tests/fuzz_confs.py) on the LIVE unmodified reference and on the C oracle, decision by decision.
usage: python tools/fuzz_ref_oracle.py <case> [decisions] [seed,seed...]      (build container only: needs /root/reference)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from fuzz_confs import fuzz_case  # noqa: E402
from golden_util import REF_RTOL  # noqa: E402
import test_fuzz  # noqa: E402
from tools.refharness import RefCell, load_reference  # noqa: E402


class References:
    def __init__(self, conf, nc, cap, seeds):
        load_reference(nc)
        self.cells = [RefCell(conf, s, n_carriers=nc, capture_effect=cap) for s in seeds]

    def reset(self):
        return [c.reset() for c in self.cells]

    def step(self, acts):
        return [c.step([int(x) for x in a]) for c, a in zip(self.cells, acts)]


k = int(sys.argv[1]); D = int(sys.argv[2]) if len(sys.argv) > 2 else 500
seeds = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [7000 + k]
a, b = test_fuzz._drive(References, test_fuzz._Oracles, k, seeds, D, REF_RTOL, 'reference vs oracle')
print('OK case', k, fuzz_case(k))
