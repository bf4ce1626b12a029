import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


from oracle.adapter import to_oracle_expr  # noqa: E402,F401


def from_string(s):
    """'SE_1 * PER_1 + RQ_2' -> product kernel object (through the product's own operators)."""
    from ms_project_b200.gpy_compat import kern as K
    import re
    cls = {'SE': K.RBF, 'RQ': K.RationalQuadratic, 'PER': K.StandardPeriodic, 'LIN': K.LinScaleShift, 'CONST': K.Bias}
    env = {}
    def leaf(m):
        fam, d = m.group(1), int(m.group(2)) - 1
        name = '_k%d' % len(env)
        env[name] = cls[fam](1, active_dims=[d])
        return name
    expr = re.sub(r'([A-Z]+)_(\d+)', leaf, s)
    return eval(expr, {}, env)


@pytest.fixture(scope='session')
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from ms_project_b200.engine import get_engine
    return get_engine(0)


def make_data(n, d, seed):
    rng = np.random.RandomState(seed)
    X = rng.randn(n, d)
    y = np.sin(X[:, 0] * 2.0) + 0.3 * rng.randn(n)
    X = (X - X.mean(0)) / X.std(0)
    y = (y - y.mean()) / y.std()
    return X, y.reshape(-1, 1)
