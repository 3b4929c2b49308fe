"""PhiloxShim: a duck-typed stand-in for numpy.random.Generator that can be passed as the `rng`
argument of the reference's create_system (system/system_creator.py:22). It implements the calls the
reference makes with the counter-based law of include/nbiot_rng.h, evaluated by the
host twins exported from libnbiot_b200.so -- the same compiled arithmetic the kernels use -- so that a
reference cell driven by PhiloxShim(seed) and instance `seed` of the batched simulator consume the same
random stream."""
import ctypes

from . import _cabi


class PhiloxShim:
    def __init__(self, seed, stream=0):
        self.L = _cabi.load()
        self.seed, self.stream, self.ctr = int(seed), int(stream), 0
        self._pair = (ctypes.c_double * 2)()

    def _next(self):
        c = self.ctr
        self.ctr += 1
        return c

    def random(self, size=None):
        if size is None:
            return self.L.nbiot_rng_u01(self.seed, self._next(), self.stream)
        if size == 2:
            self.L.nbiot_rng_pair(self.seed, self._next(), self.stream, self._pair)
            return [self._pair[0], self._pair[1]]
        raise NotImplementedError('random(size): only size=2 is drawn by the reference (system/channel.py:112)')

    def uniform(self, a=0.0, b=1.0):
        return a + (b - a) * self.L.nbiot_rng_u01(self.seed, self._next(), self.stream)

    def normal(self, mu=0.0, sigma=1.0):
        return mu + sigma * self.L.nbiot_rng_normal(self.seed, self._next(), self.stream)

    def binomial(self, n, p):
        return sum(1 for _ in range(int(n)) if self.L.nbiot_rng_u01(self.seed, self._next(), self.stream) < p)

    def integers(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        low, high = int(low), int(high)
        if size is None:
            return low + int(self.L.nbiot_rng_bounded(self.seed, self._next(), self.stream, high - low))
        return [low + int(self.L.nbiot_rng_bounded(self.seed, self._next(), self.stream, high - low)) for _ in range(int(size))]
