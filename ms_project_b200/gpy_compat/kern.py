"""Kernel objects with the GPy-fork surface the reference binds (src/autoks/backend/kernel.py:5-22):
`RBF`, `RationalQuadratic`, `StandardPeriodic`, `LinScaleShift`, `Bias`, `Add`, `Prod`.

These hold structure and hyperparameters only.  All arithmetic (`K`, the LML and its gradients, the
Hellinger Gram matrices) runs in libb200gp on the GPU; nothing here evaluates a kernel on the CPU.
`to_dict` keeps the fork's dictionary shape (rational-quadratic.ipynb#2, periodic.ipynb#2,
linear.ipynb#2) because `encode_kernel` / `decode_kernel` round-trip through `str(dict)` / `eval`
(backend/kernel.py:473-498)."""
import copy as _copy

import numpy as np

from .param import Param, Parameterized


class Kern(Parameterized):
    _registry = {}
    op_name = None           # device opcode family for leaves ('SE', 'RQ', 'PER', 'LIN', 'CONST')

    def __init__(self, input_dim, active_dims, name, useGPU=False):
        super(Kern, self).__init__(name)
        self.input_dim = int(input_dim)
        if active_dims is None:
            active_dims = np.arange(self.input_dim)
        self.active_dims = np.atleast_1d(np.asarray(active_dims, dtype=int))
        self.useGPU = useGPU

    # -- structure
    @property
    def parts(self):
        return [self]

    def __add__(self, other):
        return self.add(other)

    def add(self, other, name='sum'):
        assert isinstance(other, Kern), 'only kernels can be added to kernels...'
        return Add([self, other], name=name)

    def __mul__(self, other):
        return self.prod(other)

    def prod(self, other, name='mul'):
        assert isinstance(other, Kern), 'only kernels can be multiplied to kernels...'
        return Prod([self, other], name)

    # -- numerics (device)
    def K(self, X, X2=None):
        from .. import engine
        return engine.kernel_matrix(self, X, X2)

    def Kdiag(self, X):
        from .. import engine
        return engine.kernel_diag(self, X)

    # -- serialisation
    def _save_to_input_dict(self):
        return {'input_dim': self.input_dim, 'active_dims': self.active_dims.tolist(), 'name': self.name,
                'useGPU': self.useGPU}

    def to_dict(self):
        raise NotImplementedError

    @staticmethod
    def from_dict(input_dict):
        input_dict = _copy.deepcopy(input_dict)
        cls_name = input_dict.pop('class')
        input_dict['name'] = str(input_dict['name'])
        cls = Kern._registry[cls_name]
        return cls._build_from_input_dict(cls, input_dict)

    @staticmethod
    def _build_from_input_dict(kernel_class, input_dict):
        input_dict.pop('useGPU', None)
        return kernel_class(**input_dict)

    def __str__(self):
        return '%s(%s)' % (self.name, ', '.join('%s=%.6g' % (n, v) for n, v in zip(self.parameter_names(), self.param_array)))


def _register(cls_string):
    def deco(cls):
        Kern._registry[cls_string] = cls
        cls._class_string = cls_string
        return cls
    return deco


def _scalar(x, default=1.0):
    if x is None:
        return default
    x = np.asarray(x, dtype=np.float64).ravel()
    assert x.size == 1, 'Only one value needed for a non-ARD kernel'
    return float(x[0])


class Stationary(Kern):
    def __init__(self, input_dim, variance, lengthscale, ARD, active_dims, name, useGPU=False):
        super(Stationary, self).__init__(input_dim, active_dims, name, useGPU=useGPU)
        if ARD:
            raise NotImplementedError('ARD kernels are outside the autoks grammar (1-D base kernels only)')
        self.ARD = False
        self.link_parameters(Param('variance', _scalar(variance)), Param('lengthscale', _scalar(lengthscale)))

    def _save_to_input_dict(self):
        d = super(Stationary, self)._save_to_input_dict()
        d['variance'] = self.variance.values.tolist()
        d['lengthscale'] = self.lengthscale.values.tolist()
        d['ARD'] = self.ARD
        return d


@_register('GPy.kern.RBF')
class RBF(Stationary):
    """k = variance * exp(-r^2 / 2), r = |x - x'| / lengthscale   (rbf.ipynb#1-2)."""
    op_name = 'SE'

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False, active_dims=None, name='rbf', useGPU=False,
                 inv_l=False):
        super(RBF, self).__init__(input_dim, variance, lengthscale, ARD, active_dims, name, useGPU=useGPU)
        self.use_invLengthscale = False

    def to_dict(self):
        d = super(RBF, self)._save_to_input_dict()
        d['class'] = 'GPy.kern.RBF'
        d['inv_l'] = self.use_invLengthscale
        return d


@_register('GPy.kern.RBFDistanceBuilderKernelKernel')
class RBFDistanceBuilderKernelKernel(RBF):
    """The BOMS "kernel kernel" (fork class used at src/autoks/core/strategies/bayes_opt_gp_strategy.py:27-28; prototype
    in notebooks/exploratory/test custom kernels/kernel-kernel.ipynb#2).  A subclass of RBF: the reference wraps it in a
    `Covariance`, whose tokeniser only knows the classes of KERNEL_DICT (backend/kernel.py:141-150), so the fork class
    must be one of them -- it prints as SE_1.  An RBF whose unscaled distance
    between two inputs -- model INDICES -- is looked up in the distance builder's matrix,
        k(i, j) = variance * exp(-1/2 (D[i, j] / lengthscale)^2),   D = distance_builder.get_kernel(n_models).
    On the device this is the OP_LOOKUP leaf; the table is uploaded whenever a program batch holding the leaf is."""
    op_name = 'LOOKUP'

    def __init__(self, distance_builder, n_models, input_dim=1, variance=1., lengthscale=None, ARD=False, active_dims=None,
                 name='kernel_kernel', useGPU=False):
        super(RBFDistanceBuilderKernelKernel, self).__init__(input_dim, variance, lengthscale, ARD, active_dims, name,
                                                             useGPU=useGPU)
        self.distance_builder = distance_builder
        self.n_models = int(n_models)

    def __deepcopy__(self, memo):
        # copies share the builder (it owns device state and the NaN-initialised distance matrix)
        import copy as _c
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for key, val in self.__dict__.items():
            new.__dict__[key] = val if key == 'distance_builder' else _c.deepcopy(val, memo)
        return new

    def to_dict(self):
        d = super(RBFDistanceBuilderKernelKernel, self)._save_to_input_dict()
        d['class'] = 'GPy.kern.RBFDistanceBuilderKernelKernel'
        d['n_models'] = self.n_models
        return d


@_register('GPy.kern.RationalQuadratic')
class RationalQuadratic(Stationary):
    """k = variance * (1 + r^2 / (2 power))^-power   (rational-quadratic.ipynb#2)."""
    op_name = 'RQ'

    def __init__(self, input_dim, variance=1., lengthscale=None, power=2., ARD=False, active_dims=None, name='rat_quad'):
        super(RationalQuadratic, self).__init__(input_dim, variance, lengthscale, ARD, active_dims, name)
        self.link_parameters(Param('power', _scalar(power)))

    def to_dict(self):
        d = super(RationalQuadratic, self)._save_to_input_dict()
        d['class'] = 'GPy.kern.RationalQuadratic'
        d['power'] = self.power.values.tolist()
        return d


@_register('GPy.kern.StandardPeriodic')
class StandardPeriodic(Kern):
    """k = variance * exp(-2 sin^2(pi (x - x') / period) / lengthscale^2)   (periodic.ipynb#2).
    Parameter order is [variance, period, lengthscale] (the notebook's link_parameters order)."""
    op_name = 'PER'

    def __init__(self, input_dim, variance=1., period=None, lengthscale=None, ARD1=False, ARD2=False, active_dims=None,
                 name='standard_periodic', useGPU=False):
        super(StandardPeriodic, self).__init__(input_dim, active_dims, name, useGPU=useGPU)
        if ARD1 or ARD2:
            raise NotImplementedError('ARD kernels are outside the autoks grammar (1-D base kernels only)')
        self.ARD1, self.ARD2 = False, False
        self.link_parameters(Param('variance', _scalar(variance)), Param('period', _scalar(period)),
                             Param('lengthscale', _scalar(lengthscale)))

    def to_dict(self):
        d = super(StandardPeriodic, self)._save_to_input_dict()
        d['class'] = 'GPy.kern.StandardPeriodic'
        d['variance'] = self.variance.values.tolist()
        d['period'] = self.period.values.tolist()
        d['lengthscale'] = self.lengthscale.values.tolist()
        d['ARD1'] = self.ARD1
        d['ARD2'] = self.ARD2
        return d


@_register('GPy.kern.LinScaleShift')
class LinScaleShift(Kern):
    """k = variances * (x - shifts)(x' - shifts)   (linear.ipynb#2)."""
    op_name = 'LIN'

    def __init__(self, input_dim, variances=None, shifts=None, ARD=False, active_dims=None, name='lin'):
        super(LinScaleShift, self).__init__(input_dim, active_dims, name)
        if ARD:
            raise NotImplementedError('ARD kernels are outside the autoks grammar (1-D base kernels only)')
        self.ARD = False
        self.link_parameter(Param('variances', _scalar(variances)))
        self.link_parameter(Param('shifts', _scalar(shifts)))

    def to_dict(self):
        d = super(LinScaleShift, self)._save_to_input_dict()
        d['class'] = 'GPy.kern.LinScaleShift'
        d['variances'] = self.variances.values.tolist()
        d['shifts'] = self.shifts.values.tolist()
        d['ARD'] = self.ARD
        return d


@_register('GPy.kern.Bias')
class Bias(Kern):
    """k = variance (GPy Bias; used by linear.ipynb#4 and named `constant` in the north star)."""
    op_name = 'CONST'

    def __init__(self, input_dim, variance=1., active_dims=None, name='bias'):
        super(Bias, self).__init__(input_dim, active_dims, name)
        self.link_parameter(Param('variance', _scalar(variance)))

    def to_dict(self):
        d = super(Bias, self)._save_to_input_dict()
        d['class'] = 'GPy.kern.Bias'
        d['variance'] = self.variance.values.tolist()
        return d


class CombinationKernel(Kern):
    def __init__(self, kernels, name, extra_dims=()):
        assert all(isinstance(k, Kern) for k in kernels)
        active = np.array(sorted(set(int(d) for k in kernels for d in np.atleast_1d(k.active_dims))), dtype=int)
        input_dim = int(active.max()) + 1 if active.size else 0
        super(CombinationKernel, self).__init__(input_dim, active, name)
        # unique names for duplicated parts, as paramz does (rbf, rbf_1, ...)
        seen = {}
        for k in kernels:
            base = k.name.rsplit('_', 1)[0] if k.name.rsplit('_', 1)[-1].isdigit() else k.name
            n = seen.get(base, 0)
            seen[base] = n + 1
            if n > 0:
                k.name = '%s_%d' % (base, n)
            self.link_parameter(k)

    @property
    def parts(self):
        return self.parameters

    def _save_to_input_dict(self):
        d = super(CombinationKernel, self)._save_to_input_dict()
        d['parts'] = {ii: p.to_dict() for ii, p in enumerate(self.parts)}
        return d

    @staticmethod
    def _build_from_input_dict(kernel_class, input_dict):
        parts = input_dict.pop('parts', None)
        subkerns = [Kern.from_dict(parts[p]) for p in sorted(parts, key=lambda q: int(q))]
        return kernel_class(subkerns)


@_register('GPy.kern.Add')
class Add(CombinationKernel):
    """Sum of kernels; nested Adds are flattened (their parts are copied to the front, GPy add.py)."""

    def __init__(self, subkerns, name='sum'):
        newkerns = []
        for kern in subkerns:
            if isinstance(kern, Add):
                for part in kern.parts[::-1]:
                    newkerns.insert(0, part.copy())
            else:
                newkerns.append(kern.copy())
        super(Add, self).__init__(newkerns, name)

    def to_dict(self):
        d = super(Add, self)._save_to_input_dict()
        d['class'] = 'GPy.kern.Add'
        return d


@_register('GPy.kern.Prod')
class Prod(CombinationKernel):
    """Product of kernels; nested Prods are flattened (GPy prod.py)."""

    def __init__(self, kernels, name='mul'):
        newkerns = []
        for kern in kernels:
            if isinstance(kern, Prod):
                for part in kern.parts[::-1]:
                    newkerns.insert(0, part.copy())
            else:
                newkerns.append(kern.copy())
        super(Prod, self).__init__(newkerns, name)

    def to_dict(self):
        d = super(Prod, self)._save_to_input_dict()
        d['class'] = 'GPy.kern.Prod'
        return d
