"""TEST INFRASTRUCTURE: an `Engine` look-alike whose numbers come from the CPU oracle.

It exists so that the HOST side of the product (gpy_compat objects, fit.fit_models, GPModel, BatchedEvaluator, parallel)
can be driven without a GPU -- by the reference's own selectors in tests/test_dropin_reference.py and by
tools/make_dropin_golden.py, which records the fixtures the GPU tests replay against the real engine.  It is never
importable from the product: ms_project_b200 has no CPU path, and `install()` below is a test-only monkeypatch of
`ms_project_b200.engine.get_engine`."""
import numpy as np

from oracle import gp as ogp
from oracle.adapter import to_oracle_expr


class OracleEngine(object):
    def __init__(self):
        self.N = self.D = 0
        self.slots = 0
        self.max_params = 0
        self._X = self._y = None
        self.n_calls = 0
        self.n_items = 0
        self.quirks = False

    # -- geometry: everything fits
    def peer(self):
        return self

    def slots_for(self, N, D, n_items, max_params, mem_fraction=0.8, extra_bytes=0):
        return int(n_items)

    def bind(self, X, y, slots, max_params=32):
        self._X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        self._y = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).reshape(-1, 1)
        self.N, self.D = self._X.shape
        self.slots, self.max_params = int(slots), int(max_params)

    def ensure_bound(self, X, y, n_items, max_params):
        self.bind(X, y, max(int(n_items), 1), max(int(max_params), 1))

    def upload_programs(self, batch):
        n, nt = batch.n_items, int(batch.theta_off[-1])
        batch._o_lml = np.full(n, np.nan)
        batch._o_dlml = np.zeros(nt)
        batch._o_status = np.zeros(n, dtype=np.int32)
        batch._o_tries = np.zeros(n, dtype=np.int32)
        batch._o_expr = [to_oracle_expr(k) for k in batch.kernels]
        if getattr(batch, 'lookup', None) is not None:      # OP_LOOKUP leaves: the builder's distance table (engine.upload_programs)
            from oracle import kernels as ok
            builder, n_models = batch.lookup
            ok.set_lookup_table(np.asarray(builder.get_kernel(n_models), dtype=np.float64).copy())
        batch.engine = self
        return batch

    def launch_count(self):
        return 0

    # -- numerics
    def lml_grad(self, batch, theta, ids=None, with_grad=True):
        theta = np.asarray(theta, dtype=np.float64)
        ids = range(batch.n_items) if ids is None else [int(i) for i in ids]
        self.n_calls += 1
        for i in ids:
            a, b = int(batch.theta_off[i]), int(batch.theta_off[i + 1])
            self.n_items += 1
            try:
                with np.errstate(all='ignore'):
                    lml, g, inf = ogp.lml_and_grad(batch._o_expr[i], theta[a:b - 1], theta[b - 1], self._X, self._y,
                                                   quirks=self.quirks)
                if not np.isfinite(lml):
                    batch._o_lml[i], batch._o_status[i] = lml, 3
                    continue
                batch._o_lml[i] = lml
                batch._o_dlml[a:b] = g
                batch._o_tries[i] = inf['tries']
                batch._o_status[i] = 1 if inf['tries'] else 0
            except (np.linalg.LinAlgError, ValueError):
                batch._o_lml[i] = np.nan
                batch._o_status[i] = 2
                batch._o_tries[i] = 5
        return batch._o_lml, batch._o_dlml, batch._o_status, batch._o_tries

    def build_K(self, batch, theta, ids=None, add_diag=False):
        from oracle import kernels as ok
        ids = range(batch.n_items) if ids is None else ids
        out = []
        for i in ids:
            a, b = int(batch.theta_off[i]), int(batch.theta_off[i + 1])
            K = ok.K(batch._o_expr[i], theta[a:b - 1], self._X, None, 'gpy')
            if add_diag:
                K = K + (theta[b - 1] + ogp.EXACT_INFERENCE_JITTER) * np.eye(self.N)
            out.append(K)
        return np.stack(out)

    def posterior(self, batch, theta, item, Xs, include_likelihood=True):
        a, b = int(batch.theta_off[item]), int(batch.theta_off[item + 1])
        mu, var = ogp.predict(batch._o_expr[item], theta[a:b - 1], theta[b - 1], self._X, self._y, np.asarray(Xs, dtype=np.float64),
                              include_likelihood=include_likelihood)
        lml = ogp.lml_and_grad(batch._o_expr[item], theta[a:b - 1], theta[b - 1], self._X, self._y)[0]
        return mu.ravel(), var.ravel(), lml, 0


class installed(object):
    """Context manager: every `get_engine(...)` of the product returns one OracleEngine per role."""

    def __enter__(self):
        from ms_project_b200 import engine as _engine
        self._mod = _engine
        self._saved = (_engine.get_engine, dict(_engine._ENGINES))
        self.engines = {}

        def get_engine(device=0, role='search'):
            if role not in self.engines:
                self.engines[role] = OracleEngine()
            return self.engines[role]
        _engine.get_engine = get_engine
        return self

    def __exit__(self, *exc):
        self._mod.get_engine = self._saved[0]
        return False
