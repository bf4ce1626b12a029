"""`Covariance` with the reference's surface (src/autoks/core/covariance.py:22-155): wraps a kernel object and, on
every `raw_kernel` assignment, recomputes infix / postfix tokens and the symbolic expressions the search uses as
identity keys -- `symbolic_expr` (candidate de-duplication, covariance.py:232-238) and `symbolic_expr_expanded`
(visited set / populations, base.py:337,385; gp_model_population.py:36-38)."""
from . import program
from .gpy_compat.kern import Kern


class Covariance(object):
    def __init__(self, kernel):
        self.raw_kernel = kernel

    @property
    def raw_kernel(self):
        return self._raw_kernel

    @raw_kernel.setter
    def raw_kernel(self, new_kernel):
        if not isinstance(new_kernel, Kern):
            raise TypeError('kernel must be %s. Found type %s.' % (Kern.__name__, new_kernel.__class__.__name__))
        self._raw_kernel = new_kernel
        self.infix_tokens = program.infix_tokens(new_kernel)
        self.postfix_tokens = program.postfix_tokens(new_kernel)
        self.infix = _tokens_to_str(self.infix_tokens)
        self.infix_full = _tokens_to_str(self.infix_tokens, show_params=True)
        self.postfix = _tokens_to_str(self.postfix_tokens)
        self.symbolic_expr = _postfix_to_symbol(self.postfix_tokens)
        self.symbolic_expr_expanded = self.symbolic_expr.expand()

    def to_dict(self):
        return {'__class__': 'Covariance', '__module__': 'src.autoks.core.covariance', 'kernel': self.raw_kernel.to_dict()}

    @classmethod
    def from_dict(cls, input_dict):
        return cls(Kern.from_dict(input_dict['kernel']))

    def is_base(self):
        return len(self.postfix_tokens) == 1

    def symbolically_equals(self, other):
        return self.symbolic_expr == other.symbolic_expr

    def symbolic_expanded_equals(self, other):
        return self.symbolic_expr_expanded == other.symbolic_expr_expanded

    def infix_equals(self, other):
        return isinstance(other, Covariance) and other.infix == self.infix

    def __add__(self, other):
        return Covariance(self.raw_kernel + other.raw_kernel)

    def __mul__(self, other):
        return Covariance(self.raw_kernel * other.raw_kernel)

    def __str__(self):
        return str(self.symbolic_expr)

    def __repr__(self):
        return 'Covariance(kernel=%r)' % self.infix_full


def _tokens_to_str(tokens, show_params=False):
    """tokens_to_str / subkernel_expression (src/autoks/backend/kernel.py:129-150,228-247)."""
    out = []
    for t in tokens:
        if isinstance(t, str):
            out.append(t)
        else:
            s = program.leaf_name(t)
            if show_params:
                s += '(' + ', '.join('%s=%.6f' % (n, v) for n, v in zip(t.parameter_names(), t.param_array)) + ')'
            out.append(s)
    return ' '.join(out)


def _postfix_to_symbol(postfix_tokens):
    """postfix_tokens_to_symbol (src/autoks/symbolic/util.py:13-25) over sympy Symbols named FAMILY_<dim>."""
    from sympy import Symbol
    stack = []
    for t in postfix_tokens:
        if isinstance(t, str):
            b = stack.pop()
            a = stack.pop()
            stack.append(a * b if t == '*' else a + b)
        else:
            stack.append(Symbol(program.leaf_name(t)))
    return stack.pop()


# ---------------------------------------------------------------------------------------------------------------------
# Kernel-matrix diversity metrics (src/autoks/core/covariance.py:278-338).  The reference centres
# each N x N matrix with two N^3 matmuls per pair; here one device call handles the whole population (csrc/align.cu).
# ---------------------------------------------------------------------------------------------------------------------
def pairwise_centered_alignments(covariances, x, engine=None):
    """Alignment of all pairs of covariances (covariance.py:320-338).  Like the reference, only i < j is computed and
    the result is `(dists + dists.T) / 2`: off-diagonal entries hold HALF the alignment, the diagonal is 0."""
    import numpy as np
    from .program import ProgramBatch
    n = len(covariances)
    dists = np.zeros((n, n))
    if n < 2:
        return dists
    if engine is None:
        from .engine import get_engine
        engine = get_engine(0, 'metrics')
    x = np.asarray(x, dtype=np.float64)
    batch = ProgramBatch([c.raw_kernel for c in covariances])
    need = engine.lib.b200gp_alignment_min_slots(int(x.shape[0]), n)
    slots = engine.slots_for(x.shape[0], x.shape[1], need, batch.max_params, mem_fraction=0.85,
                             extra_bytes=0 if engine._ws is None else engine._ws.numel())
    if slots < need:
        raise MemoryError('%d kernel matrices of order %d do not fit on the device at once (%d of %d slots)'
                          % (n, x.shape[0], slots, need))
    engine.bind(x, np.zeros(x.shape[0]), need, batch.max_params)
    engine.upload_programs(batch)
    A = engine.centered_alignments(batch, batch.initial_theta())
    iu = np.triu_indices(n, 1)
    dists[iu] = A[iu]
    return (dists + dists.T) / 2.


def centered_alignment_of_kernels(cov1, cov2, x, engine=None):
    """centered_alignment(compute_kernel(k1, x), compute_kernel(k2, x)) for two covariances (covariance.py:301-308)."""
    return 2.0 * float(pairwise_centered_alignments([cov1, cov2], x, engine)[0, 1])
