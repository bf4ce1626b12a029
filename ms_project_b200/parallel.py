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


def model_cost(kernel, n_data):
    """Relative cost of one LML+grad evaluation: N^3 for the factorisation + inverse, plus the element-wise
    passes that scale with the number of leaves."""
    n_leaves = sum(1 for _ in kernel.flattened_parameters()) / 2.5 + 1.0
    return float(n_data) ** 3 + 150.0 * n_leaves * float(n_data) ** 2


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


def score_models_sharded(models, x, y, scoring_func, optimizer=None, n_restarts=10, engine=None, model_dict=None,
                         score_fn=None):
    """`score_models` across all ranks: every rank holds the full candidate list, fits only its shard, then the
    records {score, lml, failed, n_evals, n_params, theta...} are all-gathered and written into EVERY rank's model
    objects, so the search loop continues identically everywhere.  Restart draws are made for the whole list on
    every rank (same global seed => same draws as a single-GPU run), each rank using its own slice."""
    from . import fit as _fit
    from .gp_model import score_models
    rank, ws = world()
    n = len(models)
    if n == 0:
        return []
    if score_fn is None:
        score_fn = score_models
    costs = [model_cost(m.covariance.raw_kernel, np.asarray(x).shape[0]) for m in models]
    bounds = shard_bounds(costs, ws)
    lo, hi = bounds[rank], bounds[rank + 1]
    counts = [bounds[r + 1] - bounds[r] for r in range(ws)]
    mine = models[lo:hi]
    score_fn(mine, x, y, scoring_func, optimizer=optimizer, n_restarts=n_restarts, engine=engine, model_dict=model_dict,
             all_models=models, shard=(lo, hi))
    pmax = max(m.covariance.raw_kernel.size for m in models) + 1
    rec = np.zeros((len(mine), 5 + pmax))
    for i, m in enumerate(mine):
        th = np.concatenate([m.covariance.raw_kernel.param_array, m.likelihood.param_array])
        r = m.fit_result
        rec[i, 0] = m.score
        rec[i, 1] = np.nan if r is None else r.lml
        rec[i, 2] = float(m.failed_fitting)
        rec[i, 3] = 0 if r is None else r.n_evals
        rec[i, 4] = th.size
        rec[i, 5:5 + th.size] = th
    full = all_gather_records(rec, counts)
    for m, row in zip(models, full):
        p = int(row[4])
        m.covariance.raw_kernel[:] = row[5:5 + p - 1]
        m.covariance.raw_kernel = m.covariance.raw_kernel            # re-tokenise like gp_model.py:79
        if m.likelihood is None:
            from .gpy_compat.likelihoods import Gaussian
            m.likelihood = Gaussian()
        m.likelihood[:] = row[5 + p - 1:5 + p]
        m.failed_fitting = bool(row[2])
        m.score = float(row[0])
    return [m.score for m in models]


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
