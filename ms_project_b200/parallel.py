"""Multi-GPU scoring: candidates shard across ranks (one process per GPU), no traffic while fitting, one
all_gather of fixed-width result records per batch.  Works over any torch.distributed backend:
NCCL on the B200 node (records travel as device tensors over NVLink), gloo in the CPU tests."""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def world():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(costs, world_size):
    """Contiguous index ranges with balanced total cost (greedy prefix split): rank r owns
    [bounds[r], bounds[r+1]).  Contiguity keeps results re-assembling in candidate order."""
    costs = np.asarray(costs, dtype=np.float64)
    n = costs.size
    bounds = [0]
    csum = np.concatenate([[0.0], np.cumsum(costs)])
    total = csum[-1]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(csum, target, side='left'))
        # pick the split point closest to the target, never before the previous bound
        if k > 0 and abs(csum[k - 1] - target) <= abs(csum[min(k, n)] - target):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return bounds


def lpt_assign(costs, world_size):
    """Longest-processing-time-first: items in descending cost order, each to the least-loaded rank (ties: lowest rank).
    Returns `owner[item]`.  Deterministic, so every rank computes the same assignment without talking."""
    costs = np.asarray(costs, dtype=np.float64)
    load = np.zeros(world_size)
    owner = np.zeros(costs.size, dtype=np.int64)
    for i in np.argsort(-costs, kind='stable'):
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return owner


def all_gather_records(local, counts, device=None):
    """local: [n_local, width] float64 records of this rank; counts: records per rank (known to all).
    Returns the concatenation over ranks, identical on every rank."""
    import torch
    dist = _dist()
    rank, ws = world()
    local = np.ascontiguousarray(np.asarray(local, dtype=np.float64))
    if ws == 1:
        return local.copy()
    width = local.shape[1]
    nmax = int(max(counts))
    backend = dist.get_backend()
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    buf = torch.zeros((nmax, width), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[:local.shape[0]].copy_(torch.from_numpy(local))
    out = torch.empty((ws * nmax, width), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(ws, nmax, width)
    return np.concatenate([out[r, :counts[r]] for r in range(ws)], axis=0)


class Comm(object):
    """What gp_model.score_models needs from the process group: who fits what, who failed, and the record exchange."""

    def __init__(self, policy='lpt'):
        self.rank, self.ws = world()
        self.policy = policy
        self._owner = None

    def shard(self, items, costs):
        """The members of `items` (indices into the model list) this rank fits.  'lpt': balanced by the a-priori cost;
        'contiguous': balanced contiguous ranges."""
        items = list(items)
        if self.policy == 'contiguous':
            bounds = shard_bounds(costs, self.ws)
            owner = np.zeros(len(items), dtype=np.int64)
            for r in range(self.ws):
                owner[bounds[r]:bounds[r + 1]] = r
        else:
            owner = lpt_assign(costs, self.ws)
        self._owner = owner
        return [it for it, o in zip(items, owner) if o == self.rank]

    def union(self, local_items):
        """Union over ranks of small index lists (the models that need a retry), sorted."""
        counts_local = np.array([[float(len(local_items))]])
        counts = [int(c) for c in all_gather_records(counts_local, [1] * self.ws)[:, 0]]
        rec = np.asarray(local_items, dtype=np.float64).reshape(-1, 1)
        return sorted(int(v) for v in all_gather_records(rec, counts)[:, 0])

    def exchange(self, models, items, mine):
        """All-gather {index, score, lml, failed, n_evals, theta...} of the models each rank fitted and write them into
        every rank's model objects."""
        from . import fit as _fit
        from .gpy_compat.likelihoods import Gaussian
        pmax = max(models[i].covariance.raw_kernel.size for i in items) + 1
        rec = np.zeros((len(mine), 6 + pmax))
        for k, i in enumerate(mine):
            m = models[i]
            th = np.concatenate([m.covariance.raw_kernel.param_array, m.likelihood.param_array])
            r = m.fit_result
            rec[k, :6] = [i, m.score, r.lml, float(m.failed_fitting), r.n_evals, th.size]
            rec[k, 6:6 + th.size] = th
        counts = [int(np.sum(self._owner == r)) for r in range(self.ws)]
        mine_set = set(mine)
        for row in all_gather_records(rec, counts):
            i, p = int(row[0]), int(row[5])
            if i in mine_set:
                continue
            m = models[i]
            m.covariance.raw_kernel[:] = row[6:6 + p - 1]
            m.covariance.raw_kernel = m.covariance.raw_kernel            # re-tokenise like gp_model.py:79
            if m.likelihood is None:
                m.likelihood = Gaussian()
            m.likelihood[:] = row[6 + p - 1:6 + p]
            m.failed_fitting = bool(row[3])
            m.score = float(row[1])
            r = _fit.FitResult()
            r.lml, r.n_evals, r.theta = float(row[2]), int(row[4]), np.array(row[6:6 + p])
            m.fit_result = r


def score_models_sharded(models, x, y, scoring_func, optimizer=None, n_restarts=10, engine=None, model_dict=None,
                         policy='lpt', max_iters=1000):
    """`gp_model.score_models` across all ranks: every rank holds the full candidate list, fits only its share, then the
    records are all-gathered and written into EVERY rank's model objects, so the search loop continues identically
    everywhere.  Restart draws are made for the whole list on every rank (same global seed => same draws as a single-GPU
    run); results are bit-identical to the single-GPU fit because every (model, restart) instance is an independent
    optimisation and no device reduction depends on the batch composition."""
    from .gp_model import score_models
    return score_models(models, x, y, scoring_func, optimizer=optimizer, n_restarts=n_restarts, engine=engine,
                        model_dict=model_dict, comm=Comm(policy) if world()[1] > 1 else None, max_iters=max_iters)


def hellinger_matrix_sharded(builder, active_models, n_models):
    """All-pairs Hellinger matrix with the pair list split by rows across ranks; row blocks are all-gathered."""
    rank, ws = world()
    rows = np.arange(n_models)
    costs = (n_models - 1 - rows).astype(np.float64) + 1e-9          # pairs (i, j > i) per row
    bounds = shard_bounds(costs, ws)
    lo, hi = bounds[rank], bounds[rank + 1]
    ii, jj = [], []
    for i in range(lo, hi):
        for j in range(i + 1, n_models):
            ii.append(i)
            jj.append(j)
    dist, _ = builder.pair_distances(ii, jj)
    local = np.zeros((hi - lo, n_models))
    k = 0
    for i in range(lo, hi):
        m = n_models - 1 - i
        local[i - lo, i + 1:] = dist[k:k + m]
        k += m
    counts = [bounds[r + 1] - bounds[r] for r in range(ws)]
    upper = all_gather_records(local, counts)
    D = upper + upper.T
    builder._average_distance[:n_models, :n_models] = D
    np.fill_diagonal(builder._average_distance, 0)
    return D
