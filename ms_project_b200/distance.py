"""BOMS kernel-kernel distances with the reference's builder interface
(src/autoks/distance/distance.py:22-194): `DistanceBuilder` / `HellingerDistanceBuilder` with the same
constructor arguments, `precompute_information`, `compute_distance`, `update`, `get_kernel`,
`create_precomputed_info`, `hellinger_distance`, and the NaN-initialised symmetric `_average_distance`.

The arithmetic -- 20 Gram matrices and log-determinants per model, one small Cholesky per (pair, sample) -- runs in
libb200gp (csrc/hellinger.cu).  Mini-Gram matrices stay resident in HBM (500 models x 20 x 100^2 doubles = 800 MB)
and pairs are evaluated in one launch per `compute_distance` call instead of the reference's Python double loop."""
import ctypes

import numpy as np
from scipy.special import ndtri

from . import _lib
from .gpy_compat.priors import Gaussian, LogGaussian
from .program import ProgramBatch

_PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101, 103, 107,
           109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179, 181, 191, 193, 197, 199, 211, 223, 227, 229,
           233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283, 293, 307, 311, 313, 317, 331, 337, 347, 349, 353, 359,
           367, 373, 379, 383, 389, 397, 401, 409, 419, 421, 431, 433, 439, 443, 449, 457, 461, 463, 467, 479, 487, 491,
           499, 503, 509, 521, 523, 541]


def scrambled_halton(n_samples, n_dims, seed=0):
    """Permuted (generalised) Halton points in (0, 1).  The reference draws its probability samples from
    `ghalton.GeneralizedHalton(ghalton.EA_PERMS[:d])` (src/autoks/distance/sampling/halton.py:13-15); that C++
    extension and its permutation table are not available here, so the permutations are seeded instead.  The sample
    matrix is an explicit INPUT of the distance builders, so any other matrix (e.g. the reference's own) can be
    passed in unchanged."""
    if n_dims > len(_PRIMES):
        raise ValueError('supports up to %d dimensions' % len(_PRIMES))
    rng = np.random.RandomState(seed)
    out = np.empty((n_samples, n_dims))
    for j in range(n_dims):
        b = _PRIMES[j]
        perm = np.concatenate([[0], 1 + rng.permutation(b - 1)])     # 0 stays fixed so points stay inside (0, 1)
        for i in range(n_samples):
            k, f, r = i + 1, 1.0, 0.0
            while k > 0:
                f /= b
                r += f * perm[k % b]
                k //= b
            out[i, j] = r
    return out


def probability_samples(max_num_hyperparameters=40, num_samples=20, sampling_method='generalized_halton', seed=0):
    """src/autoks/distance/util.py:8-18 -> a num_samples x max_num_hyperparameters array in (0, 1)."""
    if sampling_method in ('generalized_halton', 'halton'):
        return scrambled_halton(num_samples, max_num_hyperparameters, seed)
    if sampling_method == 'random':
        return np.random.RandomState(seed).rand(num_samples, max_num_hyperparameters)
    raise ValueError('Sampler type %s not found.' % sampling_method)


def prior_sample(priors, prob_samples, z=None):
    """src/autoks/distance/util.py:21-71 (inverse-CDF samples; each prior family indexes the probability
    columns from 0, exactly as the reference does).  `norm.ppf(u, mu, sd)` is evaluated as `ndtri(u) * sd + mu` -- the
    operation scipy performs inside its generic wrapper, whose per-call overhead was 80 % of the C3 precompute (500 kernels);
    `z` = ndtri(prob_samples), if the caller has it cached."""
    priors = list(priors)
    S = prob_samples.shape[0]
    g_ind = [i for i, p in enumerate(priors) if type(p) is Gaussian]
    lg_ind = [i for i, p in enumerate(priors) if type(p) is LogGaussian]
    if len(g_ind) + len(lg_ind) < len(priors):
        raise ValueError('Unsupported prior found in priors')
    hyps = np.full((S, len(priors)), np.nan)
    if z is None:
        z = ndtri(prob_samples)

    def ppf(sub):
        mu = np.array([p.mu for p in sub])
        sd = np.array([p.sigma for p in sub])
        return z[:, :len(sub)] * sd[None, :] + mu[None, :]

    if g_ind:
        hyps[:, g_ind] = ppf([priors[i] for i in g_ind])
    if lg_ind:
        hyps[:, lg_ind] = np.exp(ppf([priors[i] for i in lg_ind]))
    return hyps


def get_priors(kernel):
    """src/autoks/backend/kernel.py:92-102: the prior of every parameter, in `param_array` order (what filling
    `priors[ind] = prior` from `kernel.priors.items()` yields)."""
    params = kernel.flattened_parameters()
    if any(p.prior is None for p in params):
        raise ValueError('All kernel parameters must have priors')
    priors = np.empty(len(params), dtype=object)
    for i, p in enumerate(params):
        priors[i] = p.prior
    return priors


def cov_nearest(cov, method='clipped', threshold=1e-15, n_fact=100, return_all=False):
    """`statsmodels.stats.correlation_tools.cov_nearest` as the reference calls it (`cov_nearest(k, threshold=tolerance)`,
    distance.py:296: the default 'clipped' method; statsmodels is not installable here, so its published algorithm is
    restated): covariance -> correlation, eigenvalues below `threshold` clipped to it (corr_clipped / clip_evals, one
    symmetric eigendecomposition), re-normalised to unit diagonal, scaled back by the original standard deviations.
    The rare repair path of `chol_safe`: host NumPy, like the reference."""
    if method != 'clipped':
        raise NotImplementedError("only method='clipped' (the default the reference uses) is restated")
    cov = np.asanyarray(cov, dtype=np.float64)
    std_ = np.sqrt(np.diag(cov))
    corr = cov / np.outer(std_, std_)
    evals, evecs = np.linalg.eigh(corr)                # raises LinAlgError for a non-square input (test_distance.py:119-121)
    if np.any(evals < threshold):
        corr = np.dot(evecs * np.maximum(evals, threshold), evecs.T)
        x_std = np.sqrt(np.diag(corr))
        corr = corr / x_std / x_std[:, None]
    cov_ = corr * np.outer(std_, std_)
    return (cov_, corr, std_) if return_all else cov_


def fix_numerical_problem(k, tolerance):
    """distance.py:287-298."""
    k = cov_nearest(k, threshold=tolerance)
    return np.linalg.cholesky(k).T


def chol_safe(k, tolerance):
    """distance.py:300-312: plain Cholesky (upper factor), repaired through cov_nearest when it fails."""
    try:
        return np.linalg.cholesky(k).T
    except np.linalg.LinAlgError:
        return fix_numerical_problem(k, tolerance)


def _repaired_logdet(k, tolerance):
    """log-determinant the reference ends up with when its np.linalg.cholesky has failed (the device reported status 1):
    2 sum log diag(chol(cov_nearest(k)))."""
    with np.errstate(all='ignore'):
        c = fix_numerical_problem(np.asarray(k, dtype=np.float64), tolerance)
        return float(2.0 * np.sum(np.log(np.diag(c)), axis=0))


class ActiveModels(object):
    """Minimal stand-in for ActiveSet (src/autoks/core/active_set.py): an indexable model store."""

    def __init__(self, max_n_models=1000):
        self.max_n_models = max_n_models
        self.models = []

    def add(self, model):
        self.models.append(model)
        return len(self.models) - 1

    def __len__(self):
        return len(self.models)


class HellingerInfo(object):
    """What `model.info` holds: the log-determinants on the host and the slot of the model's mini-Gram stack in HBM.
    Unpacks like the reference's tuple `(log_det, mini_gram_matrices[n, n, S])`, fetching the matrices on demand."""

    def __init__(self, builder, slot, log_det, status):
        self.builder, self.slot, self.log_det, self.status = builder, slot, log_det, status

    def __iter__(self):
        yield self.log_det
        yield self.builder.mini_gram_matrices(self.slot)


class DistanceBuilder(object):
    def __init__(self, noise_prior, num_samples, max_num_hyperparameters, max_num_kernels, active_models,
                 initial_model_indices, data_X, sampling_method='generalized_halton', probability_samples_matrix=None,
                 engine=None):
        self.num_samples = num_samples
        self.max_num_hyperparameters = max_num_hyperparameters
        self.max_num_kernels = max_num_kernels
        self._sampling_method = sampling_method
        if probability_samples_matrix is None:
            probability_samples_matrix = probability_samples(max_num_hyperparameters, num_samples, sampling_method)
        self.probability_samples = np.asarray(probability_samples_matrix, dtype=np.float64)
        assert self.probability_samples.shape == (num_samples, max_num_hyperparameters)
        self._z = ndtri(self.probability_samples)          # standard-normal quantiles of the samples, shared by every model
        assert type(noise_prior) is Gaussian
        self.hyperparameter_data_noise_samples = np.exp(prior_sample([noise_prior], self.probability_samples))
        self._average_distance = np.full((max_num_kernels, max_num_kernels), np.nan)
        np.fill_diagonal(self._average_distance, 0)
        if engine is None:
            from .engine import get_engine
            engine = get_engine(0, 'distance')
        self.engine = engine
        self._init_device(np.asarray(data_X, dtype=np.float64))
        self.precompute_information(active_models, initial_model_indices, data_X)

    def _init_device(self, data_X):
        pass

    def precompute_information(self, active_models, new_candidates_indices, data_X):
        raise NotImplementedError

    def update(self, active_models, new_candidates_indices, all_candidates_indices, selected_indices, data_X):
        """distance.py:69-96."""
        self.precompute_information(active_models, new_candidates_indices, data_X)
        new_evaluated_models = selected_indices[-1]
        all_old = np.setdiff1d(all_candidates_indices, new_candidates_indices)
        self.compute_distance(active_models, [new_evaluated_models], list(all_old.tolist()))
        self.compute_distance(active_models, selected_indices, new_candidates_indices)

    def get_kernel(self, index):
        return self._average_distance[:index, :index]


class HellingerDistanceBuilder(DistanceBuilder):
    """Hellinger distance between the models' prior Gram matrices (distance.py:126-194)."""
    tol_logdet = 0.02          # default `tol` of hellinger_distance (distance.py:140)

    def _init_device(self, data_X):
        torch = self.engine.torch
        n, D = data_X.shape
        # up to 128 points every Gram matrix is one CTA (in-register Cholesky, hellinger.cu); beyond that -- the reference
        # hands the builders the whole training set (smbo_model_selector.py:91-97) -- the Gram matrices go through the
        # scoring path's workspace slots (fused K-build + blocked DMMA Cholesky): b200gp_gram_logdet / _pairs_large
        self._large = n > 128
        self._n, self._D = n, D
        self._X_host = np.ascontiguousarray(data_X)
        self._d_X = torch.from_numpy(self._X_host).to(self.engine.device)
        self._d_lambda = torch.from_numpy(np.ascontiguousarray(self.hyperparameter_data_noise_samples[:, 0])).to(self.engine.device)
        S = self.num_samples
        self._d_logdet = torch.full((self.max_num_kernels, S), float('nan'), dtype=torch.float64, device=self.engine.device)
        self._d_G = None            # allocated on first use: [max_num_kernels, S, n, n]
        self._capacity = 0

    def _ensure_store(self, upto):
        torch = self.engine.torch
        if self._capacity >= upto:
            return
        cap = min(self.max_num_kernels, max(upto, 2 * self._capacity, 64))
        new = torch.empty((cap, self.num_samples, self._n, self._n), dtype=torch.float64, device=self.engine.device)
        if self._d_G is not None:
            new[:self._capacity].copy_(self._d_G)
        self._d_G, self._capacity = new, cap

    def mini_gram_matrices(self, slot):
        """n x n x S host array of model `slot` (the reference's layout)."""
        return self._d_G[slot].permute(1, 2, 0).contiguous().cpu().numpy()

    def precompute_information(self, active_models, new_candidates_indices, data_X):
        idx = [int(i) for i in new_candidates_indices]
        if not idx:
            return
        data_X = np.asarray(data_X, dtype=np.float64)
        if data_X.shape != self._X_host.shape or not np.array_equal(data_X, self._X_host):
            raise ValueError('data_X differs from the subset this builder was created with')
        torch, eng = self.engine.torch, self.engine
        self._ensure_store(max(idx) + 1)
        kernels = [active_models.models[i].covariance.raw_kernel for i in idx]
        S = self.num_samples
        hyps = [prior_sample(get_priors(k), self.probability_samples, self._z) for k in kernels]        # each S x P_i
        if self._large:
            self._precompute_large(active_models, idx, kernels, hyps)
            return
        batch = ProgramBatch(kernels)
        theta_off = np.zeros(len(idx) + 1, dtype=np.int32)
        theta_off[1:] = np.cumsum([h.shape[1] for h in hyps])
        packed = np.concatenate([np.ascontiguousarray(h).reshape(-1) for h in hyps])
        d_tok = torch.from_numpy(batch.tokens.reshape(-1)).to(eng.device)
        d_poff = torch.from_numpy(batch.prog_off).to(eng.device)
        d_toff = torch.from_numpy(theta_off).to(eng.device)
        d_theta = torch.from_numpy(packed).to(eng.device)
        M = len(idx)
        d_ld = torch.empty((M, S), dtype=torch.float64, device=eng.device)
        d_G = torch.empty((M, S, self._n, self._n), dtype=torch.float64, device=eng.device)
        d_st = torch.empty((M, S), dtype=torch.int32, device=eng.device)
        rc = eng.lib.b200gp_hellinger_precompute(eng.ctx, self._d_X.data_ptr(), self._n, self._D, d_tok.data_ptr(),
                                                 d_poff.data_ptr(), d_theta.data_ptr(), d_toff.data_ptr(),
                                                 self._d_lambda.data_ptr(), M, S, d_ld.data_ptr(), d_G.data_ptr(),
                                                 d_st.data_ptr())
        _lib.check(eng.ctx, rc, 'b200gp_hellinger_precompute')
        sel = torch.tensor(idx, dtype=torch.long, device=eng.device)
        self._d_logdet.index_copy_(0, sel, d_ld)
        self._d_G.index_copy_(0, sel, d_G)
        ld_host, st_host = d_ld.cpu().numpy(), d_st.cpu().numpy()
        self._finish_precompute(active_models, idx, kernels, hyps, ld_host, st_host)

    precompute_tolerance = 1e-6       # `tolerance` of create_precomputed_info (distance.py:174)

    def _finish_precompute(self, active_models, idx, kernels, hyps, ld_host, st_host):
        """Host epilogue of a precompute: (1) a Gram matrix whose plain Cholesky failed on the device (status 1) gets the
        log-determinant `chol_safe` arrives at through `cov_nearest` (distance.py:190-191, 287-312); (2) the reference's
        loop leaves every kernel at its LAST prior sample (`covariance.raw_kernel[:] = hyp`, distance.py:186, never
        restored), which is where a later fit's restart 0 starts from -- reproduced."""
        torch = self.engine.torch
        if self._needs_logdet and np.any(st_host != 0):
            for j, s in zip(*np.nonzero(st_host)):
                G = self._d_G[idx[j], s].cpu().numpy()
                try:
                    ld_host[j, s] = _repaired_logdet(G, self.precompute_tolerance)
                    st_host[j, s] = 0
                except np.linalg.LinAlgError:
                    ld_host[j, s] = np.nan
            sel = torch.tensor(idx, dtype=torch.long, device=self.engine.device)
            self._d_logdet.index_copy_(0, sel, torch.from_numpy(np.ascontiguousarray(ld_host)).to(self.engine.device))
        for j, i in enumerate(idx):
            kernels[j][:] = hyps[j][-1]
            active_models.models[i].info = HellingerInfo(self, i, ld_host[j].copy(), st_host[j].copy())

    _needs_logdet = True        # the vector builders (Frobenius / correlation) only need the Gram matrices

    def _precompute_large(self, active_models, idx, kernels, hyps):
        """n > 128: one item per (model, sample), theta = [hyperparameter sample..., lambda_s], through the engine's
        workspace slots (b200gp_gram_logdet, diag_mode 2 = K + lambda I)."""
        torch, eng = self.engine.torch, self.engine
        S, M, n = self.num_samples, len(idx), self._n
        lam = np.asarray(self.hyperparameter_data_noise_samples[:, 0], dtype=np.float64)
        batch = ProgramBatch([k for k in kernels for _ in range(S)])
        eng.ensure_bound(self._X_host, np.zeros(n), batch.n_items, batch.max_params)
        eng.upload_programs(batch)
        theta = batch.pack_theta([hyps[m][s] for m in range(M) for s in range(S)], [lam[s] for _ in range(M) for s in range(S)])
        d_G = torch.empty((M * S, n, n), dtype=torch.float64, device=eng.device)
        d_ld = torch.full((M * S,), float('nan'), dtype=torch.float64, device=eng.device)
        d_st = torch.zeros((M * S,), dtype=torch.int32, device=eng.device)
        if self._needs_logdet:
            eng.gram_logdet(batch, theta, 2, K_out=d_G, logdet=d_ld, status=d_st)
        else:
            eng.gram_logdet(batch, theta, 2, K_out=d_G)
        sel = torch.tensor(idx, dtype=torch.long, device=eng.device)
        self._d_logdet.index_copy_(0, sel, d_ld.view(M, S))
        self._d_G.index_copy_(0, sel, d_G.view(M, S, n, n))
        ld_host, st_host = d_ld.view(M, S).cpu().numpy().copy(), d_st.view(M, S).cpu().numpy().copy()
        self._finish_precompute(active_models, idx, kernels, hyps, ld_host, st_host)

    def _pair_distances_large(self, pi, pj):
        """n > 128: the host lists the (pair, sample) cells hellinger_distance refactorises (|ld_i - ld_j| > tol,
        distance.py:145-153), the device factorises 1/2 (G_i + G_j) for exactly those in the workspace slots."""
        torch, eng = self.engine.torch, self.engine
        S = self.num_samples
        ld = self._d_logdet.cpu().numpy()
        with np.errstate(invalid='ignore'):
            different = np.abs(ld[pi] - ld[pj]) > self.tol_logdet
        wp, ws = np.nonzero(different)
        nwork = int(wp.size)
        eng.ensure_bound(self._X_host, np.zeros(self._n), max(nwork, 1), max(eng.max_params, 1))
        d_pi, d_pj = torch.from_numpy(pi).to(eng.device), torch.from_numpy(pj).to(eng.device)
        d_wp = torch.from_numpy(np.ascontiguousarray(wp, dtype=np.int32)).to(eng.device)
        d_ws = torch.from_numpy(np.ascontiguousarray(ws, dtype=np.int32)).to(eng.device)
        npairs = int(pi.size)
        work = torch.empty(eng.lib.b200gp_hellinger_pairs_large_work_bytes(npairs, S), dtype=torch.uint8, device=eng.device)
        d_dist = torch.empty(npairs, dtype=torch.float64, device=eng.device)
        d_st = torch.empty(npairs, dtype=torch.int32, device=eng.device)
        with eng.on_stream():
            rc = eng.lib.b200gp_hellinger_pairs_large(eng.ctx, self._d_logdet.data_ptr(), self._d_G.data_ptr(), self._capacity, S,
                                                      self._n, d_pi.data_ptr(), d_pj.data_ptr(), npairs, d_wp.data_ptr(),
                                                      d_ws.data_ptr(), nwork, work.data_ptr(), d_dist.data_ptr(), d_st.data_ptr())
            _lib.check(eng.ctx, rc, 'b200gp_hellinger_pairs_large')
            return d_dist.cpu().numpy(), d_st.cpu().numpy()

    def create_precomputed_info(self, covariance, data_X):
        """distance.py:170-194 for one covariance -> (log_det[S], mini_gram[n, n, S]) host arrays."""
        class _M(object):
            pass
        m = _M()
        m.covariance = covariance
        store = ActiveModels()
        store.models = [m]
        tmp = HellingerDistanceBuilder.__new__(HellingerDistanceBuilder)
        tmp.__dict__.update(self.__dict__)
        tmp.max_num_kernels = 1
        tmp._init_device(np.asarray(data_X, dtype=np.float64))
        tmp.precompute_information(store, [0], data_X)
        return tuple(m.info)

    def pair_distances(self, pair_i, pair_j):
        """Distances of model pairs (indices into the active set), one launch."""
        torch, eng = self.engine.torch, self.engine
        pi = np.ascontiguousarray(np.asarray(pair_i, dtype=np.int32))
        pj = np.ascontiguousarray(np.asarray(pair_j, dtype=np.int32))
        npairs = int(pi.size)
        if npairs == 0:
            return np.zeros(0), np.zeros(0, dtype=np.int32)
        if max(pi.max(), pj.max()) >= self._capacity:
            raise ValueError('pair refers to a model without precomputed information')
        if self._large:
            return self._pair_distances_large(pi, pj)
        d_pi, d_pj = torch.from_numpy(pi).to(eng.device), torch.from_numpy(pj).to(eng.device)
        S = self.num_samples
        work = torch.empty(eng.lib.b200gp_hellinger_pairs_work_bytes(npairs, S), dtype=torch.uint8, device=eng.device)
        d_dist = torch.empty(npairs, dtype=torch.float64, device=eng.device)
        d_st = torch.empty(npairs, dtype=torch.int32, device=eng.device)
        rc = eng.lib.b200gp_hellinger_pairs(eng.ctx, self._d_logdet.data_ptr(), self._d_G.data_ptr(), self._capacity, S,
                                            self._n, d_pi.data_ptr(), d_pj.data_ptr(), npairs,
                                            ctypes.c_double(self.tol_logdet), work.data_ptr(), d_dist.data_ptr(),
                                            d_st.data_ptr())
        _lib.check(eng.ctx, rc, 'b200gp_hellinger_pairs')
        return d_dist.cpu().numpy(), d_st.cpu().numpy()

    def compute_distance(self, active_models, indices_i, indices_j):
        """distance.py:110-118: every (i, j) of indices_i x indices_j, written symmetrically."""
        ii, jj = [], []
        for i in indices_i:
            for j in indices_j:
                ii.append(int(i))
                jj.append(int(j))
        dist, status = self.pair_distances(ii, jj)
        dist, status = self._repair_pairs(ii, jj, dist, status)
        self.last_status = status
        for i, j, d in zip(ii, jj, dist):
            self._average_distance[i, j] = d
            self._average_distance[j, i] = d

    def _repair_pairs(self, ii, jj, dist, status):
        """Pairs for which some 1/2 (G_i + G_j) failed its plain Cholesky on the device (status 1, distance NaN): the
        reference's chol_safe repairs such a matrix through cov_nearest (distance.py:153, 300-312) -- redone here on the
        host for exactly those pairs (rare; eigendecompositions stay on the CPU)."""
        bad = np.nonzero(np.asarray(status) != 0)[0]
        if bad.size == 0:
            return dist, status
        dist, status = np.array(dist, dtype=np.float64), np.array(status)
        ld = self._d_logdet.cpu().numpy()
        for k in bad:
            i, j = int(ii[k]), int(jj[k])
            try:
                dist[k] = _hellinger_distance_host(ld[i], self.mini_gram_matrices(i), ld[j], self.mini_gram_matrices(j),
                                                   self.tol_logdet)
                status[k] = 0
            except np.linalg.LinAlgError:
                dist[k] = np.nan
        return dist, status

    @staticmethod
    def metric(data_i, data_j, **kwargs):
        return HellingerDistanceBuilder.hellinger_distance(*data_i, *data_j, **kwargs)

    @staticmethod
    def hellinger_distance(log_det_i, mini_gram_matrices_i, log_det_j, mini_gram_matrices_j, tol=0.02, engine=None):
        """distance.py:135-168 for explicit host arrays (n x n x S), evaluated on the device."""
        if engine is None:
            from .engine import get_engine
            engine = get_engine(0, 'distance')
        torch = engine.torch
        Gi = np.ascontiguousarray(np.transpose(np.asarray(mini_gram_matrices_i, dtype=np.float64), (2, 0, 1)))
        Gj = np.ascontiguousarray(np.transpose(np.asarray(mini_gram_matrices_j, dtype=np.float64), (2, 0, 1)))
        S, n = Gi.shape[0], Gi.shape[1]
        d_G = torch.from_numpy(np.stack([Gi, Gj])).to(engine.device)
        ld2 = np.stack([np.asarray(log_det_i, dtype=np.float64), np.asarray(log_det_j, dtype=np.float64)])
        d_ld = torch.from_numpy(ld2).to(engine.device)
        d_pi = torch.zeros(1, dtype=torch.int32, device=engine.device)
        d_pj = torch.ones(1, dtype=torch.int32, device=engine.device)
        if n > 128:       # blocked path: only the order of the matrices matters for the binding
            with np.errstate(invalid='ignore'):
                ws = np.nonzero(np.abs(ld2[0] - ld2[1]) > tol)[0].astype(np.int32)
            engine.bind(np.zeros((n, 1)), np.zeros(n), slots=max(int(ws.size), 1), max_params=1)
            d_wp = torch.zeros(max(int(ws.size), 1), dtype=torch.int32, device=engine.device)
            d_ws = torch.from_numpy(ws if ws.size else np.zeros(1, np.int32)).to(engine.device)
            work = torch.empty(engine.lib.b200gp_hellinger_pairs_large_work_bytes(1, S), dtype=torch.uint8, device=engine.device)
            d_dist = torch.empty(1, dtype=torch.float64, device=engine.device)
            d_st = torch.empty(1, dtype=torch.int32, device=engine.device)
            with engine.on_stream():
                rc = engine.lib.b200gp_hellinger_pairs_large(engine.ctx, d_ld.data_ptr(), d_G.data_ptr(), 2, S, n, d_pi.data_ptr(),
                                                             d_pj.data_ptr(), 1, d_wp.data_ptr(), d_ws.data_ptr(), int(ws.size),
                                                             work.data_ptr(), d_dist.data_ptr(), d_st.data_ptr())
                _lib.check(engine.ctx, rc, 'b200gp_hellinger_pairs_large')
                if int(d_st.cpu().numpy()[0]) != 0:
                    return _hellinger_distance_host(log_det_i, mini_gram_matrices_i, log_det_j, mini_gram_matrices_j, tol)
                return float(d_dist.cpu().numpy()[0])
        work = torch.empty(engine.lib.b200gp_hellinger_pairs_work_bytes(1, S), dtype=torch.uint8, device=engine.device)
        d_dist = torch.empty(1, dtype=torch.float64, device=engine.device)
        d_st = torch.empty(1, dtype=torch.int32, device=engine.device)
        rc = engine.lib.b200gp_hellinger_pairs(engine.ctx, d_ld.data_ptr(), d_G.data_ptr(), 2, S, n, d_pi.data_ptr(),
                                               d_pj.data_ptr(), 1, ctypes.c_double(tol), work.data_ptr(), d_dist.data_ptr(),
                                               d_st.data_ptr())
        _lib.check(engine.ctx, rc, 'b200gp_hellinger_pairs')
        if int(d_st.cpu().numpy()[0]) != 0:
            return _hellinger_distance_host(log_det_i, mini_gram_matrices_i, log_det_j, mini_gram_matrices_j, tol)
        return float(d_dist.cpu().numpy()[0])


def _hellinger_distance_host(log_det_i, G_i, log_det_j, G_j, tol):
    """hellinger_distance (distance.py:135-168) for ONE pair whose device evaluation reported a failed factorisation:
    every factorisation goes through chol_safe, so the failing ones get the cov_nearest repair."""
    log_det_i, log_det_j = np.asarray(log_det_i, dtype=np.float64), np.asarray(log_det_j, dtype=np.float64)
    with np.errstate(invalid='ignore'):
        different = np.abs(log_det_i - log_det_j) > tol
    ldpq = log_det_i.copy()
    for s in np.nonzero(different)[0]:
        c = chol_safe(0.5 * (G_i[:, :, s] + G_j[:, :, s]), tol)
        ldpq[s] = 2 * np.sum(np.log(np.diag(c)), axis=0)
    hellinger = 1 - np.exp(0.25 * (log_det_i + log_det_j) - 0.5 * ldpq)
    return float(np.sqrt(np.clip(np.mean(hellinger, axis=0), 0, 1)))


class _VectorInfo(object):
    """`model.info` of the Frobenius / correlation builders: the reference's (n^2 x S) float32 matrix of vec(K_s + lambda_s I)
    columns (distance.py:214-232), fetched from HBM on demand."""

    def __init__(self, builder, slot):
        self.builder, self.slot = builder, slot

    def __array__(self, dtype=None, copy=None):
        G = self.builder._d_G[self.slot]                         # [S][n][n] float64 on the device
        v = G.reshape(G.shape[0], -1).t().contiguous().cpu().numpy().astype(np.float32)
        return v if dtype is None else v.astype(dtype)


class _VectorDistanceBuilder(HellingerDistanceBuilder):
    """Shared machinery of FrobeniusDistanceBuilder / CorrelationDistanceBuilder (distance.py:197-284): the sampled prior
    Gram matrices are the ones the Hellinger precompute builds; only the pair metric differs."""
    _metric_code = 0
    _needs_logdet = False

    def precompute_information(self, active_models, new_candidates_indices, data_X):
        super(_VectorDistanceBuilder, self).precompute_information(active_models, new_candidates_indices, data_X)
        for i in [int(i) for i in new_candidates_indices]:
            active_models.models[i].info = _VectorInfo(self, i)

    def create_precomputed_info(self, covariance, data_X):
        log_det, G = super(_VectorDistanceBuilder, self).create_precomputed_info(covariance, data_X)
        n = G.shape[0]
        return G.reshape(n * n, G.shape[2]).astype(np.float32)

    def pair_distances(self, pair_i, pair_j):
        torch, eng = self.engine.torch, self.engine
        pi = np.ascontiguousarray(np.asarray(pair_i, dtype=np.int32))
        pj = np.ascontiguousarray(np.asarray(pair_j, dtype=np.int32))
        npairs = int(pi.size)
        if npairs == 0:
            return np.zeros(0), np.zeros(0, dtype=np.int32)
        if max(pi.max(), pj.max()) >= self._capacity:
            raise ValueError('pair refers to a model without precomputed information')
        d_pi, d_pj = torch.from_numpy(pi).to(eng.device), torch.from_numpy(pj).to(eng.device)
        d_dist = torch.empty(npairs, dtype=torch.float64, device=eng.device)
        rc = eng.lib.b200gp_vector_pairs(eng.ctx, self._d_G.data_ptr(), self._capacity, self.num_samples, self._n,
                                         d_pi.data_ptr(), d_pj.data_ptr(), npairs, self._metric_code, d_dist.data_ptr())
        _lib.check(eng.ctx, rc, 'b200gp_vector_pairs')
        return d_dist.cpu().numpy(), np.zeros(npairs, dtype=np.int32)

    @classmethod
    def _host_pair(cls, a, b, engine=None):
        """Distance of two explicit (n^2 x S) vector matrices, evaluated on the device."""
        if engine is None:
            from .engine import get_engine
            engine = get_engine(0, 'distance')
        torch = engine.torch
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        n2, S = a.shape
        n = int(round(np.sqrt(n2)))
        G = np.ascontiguousarray(np.stack([a.T, b.T]))           # [2][S][n^2]
        d_G = torch.from_numpy(G).to(engine.device)
        d_pi = torch.zeros(1, dtype=torch.int32, device=engine.device)
        d_pj = torch.ones(1, dtype=torch.int32, device=engine.device)
        d_dist = torch.empty(1, dtype=torch.float64, device=engine.device)
        rc = engine.lib.b200gp_vector_pairs(engine.ctx, d_G.data_ptr(), 2, S, n, d_pi.data_ptr(), d_pj.data_ptr(), 1,
                                            cls._metric_code, d_dist.data_ptr())
        _lib.check(engine.ctx, rc, 'b200gp_vector_pairs')
        return float(d_dist.cpu().numpy()[0])


class FrobeniusDistanceBuilder(_VectorDistanceBuilder):
    """distance.py:197-232: average Frobenius distance between the sampled prior Gram matrices."""
    _metric_code = 0

    @staticmethod
    def metric(data_i, data_j, **kwargs):
        return FrobeniusDistanceBuilder.frobenius_distance(data_i, data_j)

    @staticmethod
    def frobenius_distance(a, b):
        return FrobeniusDistanceBuilder._host_pair(a, b)


class CorrelationDistanceBuilder(_VectorDistanceBuilder):
    """distance.py:235-284: average correlation distance sqrt(1/2 (1 - corr)); what SMBOModelSelector wires
    (src/autoks/core/model_selection/smbo_model_selector.py:91)."""
    _metric_code = 1

    @staticmethod
    def metric(data_i, data_j, **kwargs):
        return CorrelationDistanceBuilder.correlation_distance(data_i, data_j)

    @staticmethod
    def correlation_distance(a, b):
        return CorrelationDistanceBuilder._host_pair(a, b)
