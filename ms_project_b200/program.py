"""Host-side compiler: kernel object -> device "kernel program" (include/b200gp.h).

Follows the reference's own tokenisation: `kernel_to_infix_tokens` (src/autoks/backend/kernel.py:214-225)
then `infix_tokens_to_postfix_tokens` (src/evalg/encoding.py:415-447) give the postfix list that
`Covariance.postfix_tokens` holds (src/autoks/core/covariance.py:32-45); n-ary Add / Prod nodes become
left-associated binary chains, and theta offsets follow GPy's `param_array` order (parts in order)."""
import numpy as np

from .gpy_compat.kern import Kern, Add, Prod, CombinationKernel

OPCODES = {'SE': 0, 'RQ': 1, 'PER': 2, 'LIN': 3, 'CONST': 4, '+': 5, '*': 6, 'LOOKUP': 7}


def _leaf_rows(t):
    return (t[:, 0] <= 4) | (t[:, 0] == 7)
MAX_TOKENS = 64
MAX_PARAMS = 128
MAX_STACK = 8


def infix_tokens(kernel):
    """In-order token list with parentheses, e.g. [k1, '*', '(', k2, '+', k3, ')'] (leaf tokens are the
    kernel objects themselves, like the reference's)."""
    if not isinstance(kernel, CombinationKernel):
        return [kernel]
    op = '+' if isinstance(kernel, Add) else '*'
    out = []
    for i, part in enumerate(kernel.parts):
        if i:
            out.append(op)
        if isinstance(part, CombinationKernel):
            out += ['('] + infix_tokens(part) + [')']
        else:
            out.append(part)
    return out


def postfix_tokens(kernel):
    """Shunting-yard over the infix tokens ('*' binds tighter than '+')."""
    prec = {'*': 3, '+': 1, '(': 0}
    stack, out = [], []
    for tok in infix_tokens(kernel):
        if tok in ('+', '*'):
            while stack and prec[stack[-1]] >= prec[tok]:
                out.append(stack.pop())
            stack.append(tok)
        elif tok == '(':
            stack.append(tok)
        elif tok == ')':
            top = stack.pop()
            while top != '(':
                out.append(top)
                top = stack.pop()
        else:
            out.append(tok)
    while stack:
        out.append(stack.pop())
    return out


def leaf_name(k):
    """`subkernel_expression` (backend/kernel.py:141-150): FAMILY_<active_dim + 1>."""
    return '%s_%d' % (k.op_name, int(k.active_dims[0]) + 1)


def postfix_string(kernel):
    return ' '.join(t if isinstance(t, str) else leaf_name(t) for t in postfix_tokens(kernel))


def infix_string(kernel):
    return ' '.join(t if isinstance(t, str) else leaf_name(t) for t in infix_tokens(kernel))


def compile_kernel(kernel):
    """-> (tokens int32 [ntok, 4], n_params).  Raises ValueError for programs the device cannot hold."""
    if not isinstance(kernel, Kern):
        raise TypeError('kernel must be a gpy_compat Kern, got %s' % type(kernel).__name__)
    # theta offsets in param_array order (parts in order, depth first)
    offsets = {}
    off = 0
    order = []

    def walk(k):
        nonlocal off
        if isinstance(k, CombinationKernel):
            for p in k.parts:
                walk(p)
        else:
            offsets[id(k)] = off
            order.append(k)
            off += k.size

    walk(kernel)
    toks = []
    stack = []
    depth = 0
    for t in postfix_tokens(kernel):
        if isinstance(t, str):
            rhs = stack.pop()
            lhs = stack.pop()
            toks.append((OPCODES[t], lhs, rhs, 0))
        else:
            if t.op_name is None:
                raise ValueError('kernel %s has no device opcode' % type(t).__name__)
            if t.active_dims.size != 1:
                raise ValueError('leaf kernels must have exactly one active dimension')
            toks.append((OPCODES[t.op_name], int(t.active_dims[0]), offsets[id(t)], 0))
        stack.append(len(toks) - 1)
        depth = max(depth, len(stack))
    assert len(stack) == 1
    if len(toks) > MAX_TOKENS:
        raise ValueError('kernel program has %d tokens (max %d)' % (len(toks), MAX_TOKENS))
    if off > MAX_PARAMS:
        raise ValueError('kernel has %d parameters (max %d)' % (off, MAX_PARAMS))
    return np.asarray(toks, dtype=np.int32).reshape(-1, 4), off


class ProgramBatch(object):
    """Programs of M candidate kernels packed for the device: tokens, prog_off, theta_off."""

    def __init__(self, kernels):
        self.kernels = list(kernels)
        compiled = {}                    # the same kernel object repeated (restarts, prior samples) is compiled once
        progs = []
        for k in self.kernels:
            if id(k) not in compiled:
                compiled[id(k)] = compile_kernel(k)
            progs.append(compiled[id(k)])
        self.n_items = len(progs)
        self.n_params = np.array([p for _, p in progs], dtype=np.int64)      # kernel params (noise excluded)
        self.tokens = np.concatenate([t for t, _ in progs], axis=0) if progs else np.zeros((0, 4), np.int32)
        self.prog_off = np.zeros(self.n_items + 1, dtype=np.int32)
        self.prog_off[1:] = np.cumsum([t.shape[0] for t, _ in progs])
        self.theta_off = np.zeros(self.n_items + 1, dtype=np.int32)
        self.theta_off[1:] = np.cumsum(self.n_params + 1)
        self.max_params = int(self.n_params.max()) if self.n_items else 1
        self.max_dim = int(max((t[_leaf_rows(t), 1].max() for t, _ in progs), default=0))
        # OP_LOOKUP leaves (BOMS surrogate): every such leaf of a batch reads ONE distance table, the builder's
        self.lookup = None
        seen = set()
        for k in self.kernels:
            if id(k) not in seen:
                seen.add(id(k))
                k.traverse(self._note_lookup)

    def _note_lookup(self, k):
        if getattr(k, 'op_name', None) == 'LOOKUP':
            if self.lookup is not None and self.lookup[0] is not k.distance_builder:
                raise ValueError('all table look-up leaves of one batch must share the distance builder')
            self.lookup = (k.distance_builder, max(int(k.n_models), self.lookup[1] if self.lookup else 0))

    def pack_theta(self, kernel_thetas, noise_vars):
        """Flat theta in device layout: per item [kernel params..., noise variance]."""
        out = np.empty(int(self.theta_off[-1]), dtype=np.float64)
        for i in range(self.n_items):
            a, b = self.theta_off[i], self.theta_off[i + 1]
            out[a:b - 1] = kernel_thetas[i]
            out[b - 1] = noise_vars[i]
        return out

    def initial_theta(self, noise_vars=None):
        if noise_vars is None:
            noise_vars = np.ones(self.n_items)
        return self.pack_theta([k.param_array for k in self.kernels], noise_vars)
