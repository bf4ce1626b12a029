"""TEST INFRASTRUCTURE: wraps an engine (the real one on the GPU box, or tests/oracle_engine.OracleEngine) and multiplies
every LML and gradient entry it returns by (1 + eps * N(0,1)), eps = 1e-13 -- the size of the difference between two correct
FP64 implementations (another BLAS, another summation order).  An optimiser run whose result changes under this noise is
not reproducible by the reference itself; the replay tests use the wrapper to tell such runs from the stable ones."""
import numpy as np


class NoisyEngine(object):
    def __init__(self, engine, eps=1e-13, seed=0):
        self._engine = engine
        self._eps = eps
        self._rng = np.random.RandomState(seed)
        self._peer = None

    def __getattr__(self, name):
        return getattr(self._engine, name)

    def peer(self):
        if self._peer is None:
            self._peer = NoisyEngine(self._engine.peer(), self._eps, int(self._rng.randint(1 << 30)))
        return self._peer

    def lml_grad(self, batch, theta, ids=None, with_grad=True):
        lml, dlml, status, tries = self._engine.lml_grad(batch, theta, ids=ids, with_grad=with_grad)
        lml *= 1.0 + self._eps * self._rng.randn(lml.size)
        if with_grad:
            dlml *= 1.0 + self._eps * self._rng.randn(dlml.size)
        return lml, dlml, status, tries
