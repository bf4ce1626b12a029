"""Minimal parameter hierarchy with the paramz / GPy surface the autoks backend shim touches
(src/autoks/backend/kernel.py:60,96-115,155,172,313,322,335-377; src/autoks/distance/distance.py:186):
`param_array`, `parameter_names()`, `parameters`, `priors.{size,items()}`, `size`, `copy()`,
`traverse(fn)`, `obj[:] = values`, `getattr(obj, name).set_prior(prior, warning=False)`.

Every parameter of the scoring path is a positive scalar, so a `Param` is one float with a Logexp
constraint and an optional prior."""
import copy as _copy

import numpy as np

from .transformations import Logexp, Logistic, Fixed


class Param(object):
    def __init__(self, name, value, constraint=None):
        self.name = name
        v = np.atleast_1d(np.asarray(value, dtype=np.float64)).ravel()
        if v.size != 1:
            raise ValueError('only scalar (non-ARD) parameters are supported on the scoring path; got %r' % (value,))
        self._v = v.copy()
        self.constraint = Logexp() if constraint is None else constraint
        self.prior = None
        self.gradient = np.zeros(1)
        self._parent = None

    # -- value access
    @property
    def values(self):
        return self._v

    @property
    def param_array(self):
        return self._v

    @property
    def size(self):
        return 1

    def __float__(self):
        return float(self._v[0])

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._v, dtype=dtype)

    def __getitem__(self, item):
        return self._v[item]

    def __setitem__(self, item, value):
        self._v[item] = value

    # -- priors / constraints
    def set_prior(self, prior, warning=True):
        self.prior = prior

    def unset_priors(self):
        self.prior = None

    def constrain_positive(self, warning=True):
        self.constraint = Logexp()

    def constrain_bounded(self, lower, upper, warning=True):
        self.constraint = Logistic(lower, upper)
        self._v[:] = np.clip(self._v, lower, upper)

    def constrain_fixed(self, value=None, warning=True):
        if value is not None:
            self._v[:] = value
        self.constraint = Fixed()

    fix = constrain_fixed

    def unconstrain_fixed(self):
        self.constraint = Logexp()

    unfix = unconstrain_fixed

    @property
    def is_fixed(self):
        return isinstance(self.constraint, Fixed)

    def parameter_names(self):
        return [self.name]

    @property
    def priors(self):
        return PriorIndex([self])

    def flattened_parameters(self):
        return [self]

    def traverse(self, fn, *a, **kw):
        fn(self, *a, **kw)

    def __repr__(self):
        return '%s=%r' % (self.name, float(self._v[0]))


class PriorIndex(object):
    """`obj.priors`: prior -> indices into `param_array`.  `.items()` yields distinct prior objects in
    order of first appearance in `param_array` (GPy's ParameterIndexOperations keeps insertion order;
    this is the documented stand-in, shared with the restart sampler)."""

    def __init__(self, params):
        self._map = []
        for i, p in enumerate(params):
            if p.prior is None:
                continue
            for entry in self._map:
                if entry[0] is p.prior:
                    entry[1].append(i)
                    break
            else:
                self._map.append((p.prior, [i]))

    @property
    def size(self):
        return sum(len(ix) for _, ix in self._map)

    def items(self):
        return [(p, np.asarray(ix, dtype=int)) for p, ix in self._map]

    def properties(self):
        return [p for p, _ in self._map]

    def __len__(self):
        return len(self._map)


class Parameterized(object):
    def __init__(self, name):
        self.name = name
        self.parameters = []
        self._parent = None

    def link_parameter(self, p):
        p._parent = self
        self.parameters.append(p)
        if isinstance(p, Param):
            object.__setattr__(self, p.name, p)

    def link_parameters(self, *ps):
        for p in ps:
            self.link_parameter(p)

    def unlink_parameter(self, p):
        self.parameters = [q for q in self.parameters if q is not p]
        p._parent = None

    # -- flattened views
    def flattened_parameters(self):
        out = []
        for p in self.parameters:
            if isinstance(p, Param):
                out.append(p)
            else:
                out.extend(p.flattened_parameters())
        return out

    @property
    def param_array(self):
        ps = self.flattened_parameters()
        return np.array([float(p) for p in ps], dtype=np.float64)

    @property
    def size(self):
        return len(self.flattened_parameters())

    @property
    def priors(self):
        return PriorIndex(self.flattened_parameters())

    @property
    def gradient(self):
        return np.array([float(p.gradient[0]) for p in self.flattened_parameters()])

    def __setitem__(self, item, value):
        ps = self.flattened_parameters()
        arr = np.array([float(p) for p in ps], dtype=np.float64)
        arr[item] = value
        for p, v in zip(ps, arr):
            p._v[0] = v

    def __getitem__(self, item):
        if isinstance(item, str):
            return self._by_name(item)
        return self.param_array[item]

    def _by_name(self, pattern):
        """paramz's `obj['regexp']`: the parameter(s) whose hierarchical name matches (test_backend_kernel.py:70 reads
        `kernel['variance'].priors`).  One match -> that Param, several -> the list."""
        import re
        rx = re.compile(pattern)
        names = self.parameter_names()
        hits = [p for n, p in zip(names, self.flattened_parameters()) if rx.match(n) or rx.match(n.split('.')[-1])]
        if not hits:
            raise AttributeError('no parameter matches %r' % pattern)
        return hits[0] if len(hits) == 1 else hits

    def parameter_names(self, add_self=False, adjust_for_printing=False, recursive=True):
        names = []
        for p in self.parameters:
            if isinstance(p, Param):
                names.append(p.name)
            elif recursive:
                names.extend(p.name + '.' + n for n in p.parameter_names())
        if add_self:
            names = [self.name + '.' + n for n in names]
        return names

    def traverse(self, fn, *a, **kw):
        fn(self, *a, **kw)
        for p in self.parameters:
            p.traverse(fn, *a, **kw)

    def copy(self):
        c = _copy.deepcopy(self)
        c._parent = None
        return c

    def _size_transformed(self):
        """Number of optimisation variables: fixed parameters do not count (BIC's k, backend/model.py:101-104)."""
        return sum(1 for p in self.flattened_parameters() if not p.is_fixed)

    def constrain_bounded(self, lower, upper, warning=True):
        for p in self.flattened_parameters():
            p.constrain_bounded(lower, upper, warning=warning)

    def constrain_fixed(self, value=None, warning=True):
        for p in self.flattened_parameters():
            p.constrain_fixed(value, warning=warning)

    def constrain_positive(self, warning=True):
        """paramz `constrain_positive` on a whole object (KernelKernelGPModel calls it on the likelihood,
        src/autoks/gp_regression_models.py:100): every parameter gets the Logexp transform."""
        for p in self.flattened_parameters():
            p.constrain_positive(warning=warning)

    def set_prior(self, prior, warning=True):
        for p in self.flattened_parameters():
            p.set_prior(prior, warning=warning)

    def unset_priors(self):
        for p in self.flattened_parameters():
            p.unset_priors()

    def randomize(self, rand_gen=None, *args, **kwargs):
        """Priorizable.randomize as paramz/GPy run it: N(0,1) in optimiser space, then prior draws where a
        prior exists (one rvs(count) per distinct prior).  Uses the global np.random like the reference."""
        if rand_gen is None:
            rand_gen = np.random.normal
        ps = self.flattened_parameters()
        free = [p for p in ps if not p.is_fixed]
        x = rand_gen(size=len(free), *args, **kwargs)          # paramz draws `_size_transformed()` numbers
        vals = np.array([float(p) for p in ps])
        xi = iter(x)
        for i, p in enumerate(ps):
            if not p.is_fixed:
                vals[i] = float(p.constraint.f(next(xi)))
        for prior, ind in self.priors.items():
            vals[ind] = prior.rvs(ind.size)
        for p, v in zip(ps, vals):
            if not p.is_fixed:
                p._v[0] = v
