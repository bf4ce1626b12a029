"""Measure BASELINE.json's configurations C1, C3, C4, C5 on one B200 (C2 is bench.py's headline line), each next
to a bounded sample of the CPU oracle on the box's host cores.  One JSON object per config on stdout and, with
--out, in a file (committed as profiles/configs_<tag>.json).

  python tools/bench_configs.py [--configs c1,c3,c4,c5] [--cpu-budget 15] [--out gpurun_out/configs.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def host_cores():
    n = len(os.sched_getaffinity(0))
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def c1(eng, budget):
    """CKS greedy search, depth 3, SE/RQ/LIN/PER, N = 500, 3 restarts, BIC (cks_model_selector.py:28-52)."""
    from ms_project_b200 import fitness, workloads, program
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, BatchedEvaluator
    from ms_project_b200.gp_models import gp_regression
    from oracle import model as omodel
    from oracle.adapter import to_oracle_expr, prior_tuple
    X, Y = workloads.config_c1()
    hp = workloads.boms_hyperpriors()
    noise_prior = hp['GP']['variance']
    bases = workloads.get_all_1d_kernels(['SE', 'RQ', 'LIN', 'PER'], 1, hp)

    def expand(best):
        out, seen = [], set()
        for k in workloads.expand_single_kernel(best, bases):
            key = workloads.polynomial_key(k)
            if key not in seen:
                seen.add(key)
                out.append(k)
        return out
    def search():
        np.random.seed(0)
        ev = BatchedEvaluator(fitness.negative_bic, n_restarts_optimizer=3, engine=eng,
                              model_dict=gp_regression(likelihood_hyperprior={'variance': noise_prior}))
        cands = [GPModel(Covariance(k.copy())) for k in bases]
        seq, n_scored, evals = [], 0, 0
        t0 = time.perf_counter()
        for d in range(3):
            done = ev.evaluate_models(cands, X, Y, eval_budget=10 ** 6)
            n_scored += len(done)
            evals += sum(m.fit_result.n_evals for m in done)
            best = max((m for m in cands if m.evaluated), key=lambda m: m.score)
            seq.append(program.infix_string(best.covariance.raw_kernel))
            cands = [GPModel(Covariance(k)) for k in expand(best.covariance.raw_kernel)]
        return seq, n_scored, evals, time.perf_counter() - t0

    first = [k.copy() for k in bases]
    _, _, _, dt_cold = search()           # first search of the process: second engine / workspaces are created inside it
    seq, n_scored, evals, dt = search()   # the timed one (same seed, same result)
    # CPU: the same first-level candidates, sequentially, until the budget is spent
    rng = np.random.RandomState(0)
    c0, n_cpu, cpu_evals = time.perf_counter(), 0, 0
    for k in first:
        kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
        m, _failed = omodel.fit(to_oracle_expr(k), X, Y, n_restarts=3, rng=rng, theta=k.param_array.copy(), kern_priors=kp,
                                noise_prior=prior_tuple(noise_prior))
        n_cpu += 1
        cpu_evals += sum(r.get('funcalls', 0) for r in m.optimization_runs)
        if time.perf_counter() - c0 > budget:
            break
    cdt = time.perf_counter() - c0
    return {'config': 'C1 CKS depth 3, SE/RQ/LIN/PER, N=500, 3 restarts, nBIC', 'selected_sequence': seq,
            'candidates_fitted': n_scored, 'nlml_grad_evaluations': int(evals), 'seconds': dt, 'seconds_first_search_of_process': dt_cold,
            'candidates_per_s': n_scored / dt, 'evaluations_per_s': evals / dt,
            'cpu': {'kind': 'port', 'candidates': n_cpu, 'seconds': cdt, 'candidates_per_s': n_cpu / cdt,
                    'evaluations_per_s': cpu_evals / cdt if cpu_evals else None}}


class _Model(object):
    def __init__(self, covariance):
        self.covariance = covariance
        self.info = None


def c3(eng, budget):
    """BOMS Hellinger: 500 candidates, 100-point subset, D = 4, 20 samples: precompute + all 124,750 pairs."""
    import torch
    from ms_project_b200 import workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.distance import ActiveModels, HellingerDistanceBuilder, probability_samples
    from ms_project_b200.gpy_compat.priors import Gaussian
    from oracle import hellinger as ohell
    from oracle.adapter import to_oracle_expr, prior_tuple
    X, Y, pool, xsub = workloads.config_c3()
    n = len(pool)
    am = ActiveModels(max_n_models=n)
    for k in pool:
        am.add(_Model(Covariance(k)))
    prob = probability_samples(40, 20)
    noise_prior = Gaussian(np.log(0.01), 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    builder = HellingerDistanceBuilder(noise_prior, 20, 40, n, am, list(range(n)), xsub, probability_samples_matrix=prob, engine=eng)
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t0
    ii, jj = np.triu_indices(n, 1)
    builder.pair_distances(ii[:1000], jj[:1000])          # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dist, status = builder.pair_distances(ii, jj)
    torch.cuda.synchronize()
    t_pairs = time.perf_counter() - t0
    # the distance SMBO actually wires (CorrelationDistanceBuilder, smbo_model_selector.py:91): same sampled Gram stack
    from ms_project_b200.distance import CorrelationDistanceBuilder
    cb = CorrelationDistanceBuilder(noise_prior, 20, 40, n, am, list(range(n)), xsub, probability_samples_matrix=prob, engine=eng)
    cb.pair_distances(ii[:1000], jj[:1000])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cdist_, _ = cb.pair_distances(ii, jj)
    torch.cuda.synchronize()
    t_corr = time.perf_counter() - t0
    v0, v1 = np.asarray(am.models[0].info), np.asarray(am.models[7].info)
    c0 = time.perf_counter()
    for _ in range(20):
        ohell.correlation_distance(v0, v1)
    c_corr = (time.perf_counter() - c0) / 20
    # CPU sample
    lam = ohell.noise_samples(prior_tuple(noise_prior), prob)
    c0 = time.perf_counter()
    infos = []
    for k in pool[:12]:
        kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
        infos.append(ohell.create_precomputed_info(to_oracle_expr(k), kp, xsub, prob, lam))
    c_pre = (time.perf_counter() - c0) / len(infos)
    c0, npair = time.perf_counter(), 0
    for a in range(len(infos)):
        for b_ in range(a + 1, len(infos)):
            ohell.hellinger_distance(infos[a][0], infos[a][1], infos[b_][0], infos[b_][1])
            npair += 1
        if time.perf_counter() - c0 > budget:
            break
    c_pair = (time.perf_counter() - c0) / max(npair, 1)
    return {'config': 'C3 BOMS Hellinger: 500 candidates (SE/RQ per dim, level 2), subset n=100, D=4, S=20',
            'models': n, 'pairs': int(ii.size), 'precompute_seconds': t_pre, 'pairs_seconds': t_pairs,
            'pairs_per_s': ii.size / t_pairs, 'pair_sample_choleskys_per_s': ii.size * 20 / t_pairs,
            'fp64_gflops_pairs': ii.size * 20 * (100 ** 3 / 3.0) / t_pairs / 1e9,
            'nan_distances': int(np.isnan(dist).sum()), 'bad_status': int((status != 0).sum()),
            'correlation_pairs_seconds': t_corr, 'correlation_pairs_per_s': ii.size / t_corr,
            'correlation_effective_GBps': ii.size * 2 * 20 * 100 * 100 * 8 / t_corr / 1e9,
            'correlation_cpu_seconds_per_pair': c_corr, 'correlation_nan': int(np.isnan(cdist_).sum()),
            'cpu': {'kind': 'port', 'precompute_seconds_per_model': c_pre, 'seconds_per_pair': c_pair, 'pairs_timed': npair,
                    'pairs_per_s': 1.0 / c_pair, 'projected_seconds_all_pairs': c_pair * ii.size + c_pre * n}}


def c3full(eng, budget):
    """The same BOMS Hellinger distance with the builders' data_X = the WHOLE C3 training set (n = 1024), which is what the
    reference passes (smbo_model_selector.py:91-97): 100 candidates, 4950 pairs x 20 samples through the blocked path."""
    import torch
    from ms_project_b200 import workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.distance import ActiveModels, HellingerDistanceBuilder, probability_samples
    from ms_project_b200.gpy_compat.priors import Gaussian
    from oracle import hellinger as ohell
    from oracle.adapter import to_oracle_expr, prior_tuple
    X, Y, pool, _ = workloads.config_c3()
    pool = pool[:100]
    n, npts = len(pool), X.shape[0]
    am = ActiveModels(max_n_models=n)
    for k in pool:
        am.add(_Model(Covariance(k)))
    prob = probability_samples(40, 20)
    noise_prior = Gaussian(np.log(0.01), 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    builder = HellingerDistanceBuilder(noise_prior, 20, 40, n, am, list(range(n)), X, probability_samples_matrix=prob, engine=eng)
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t0
    ii, jj = np.triu_indices(n, 1)
    t0 = time.perf_counter()
    dist, status = builder.pair_distances(ii, jj)
    torch.cuda.synchronize()
    t_pairs = time.perf_counter() - t0
    ld = builder._d_logdet.cpu().numpy()
    nwork = int((np.abs(ld[ii] - ld[jj]) > builder.tol_logdet).sum())
    lam = ohell.noise_samples(prior_tuple(noise_prior), prob)
    c0 = time.perf_counter()
    infos = []
    for k in pool[:2]:
        kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
        infos.append(ohell.create_precomputed_info(to_oracle_expr(k), kp, X, prob, lam))
    c_pre = (time.perf_counter() - c0) / len(infos)
    c0 = time.perf_counter()
    d_ref = ohell.hellinger_distance(infos[0][0], infos[0][1], infos[1][0], infos[1][1])
    c_pair = time.perf_counter() - c0
    return {'config': 'C3 with data_X = the whole training set: 100 candidates, n=%d, D=4, S=20 (blocked DMMA path)' % npts,
            'models': n, 'pairs': int(ii.size), 'precompute_seconds': t_pre, 'pairs_seconds': t_pairs,
            'pair_sample_choleskys': nwork, 'fp64_tflops_pairs': nwork * (float(npts) ** 3 / 3.0) / t_pairs / 1e12,
            'nan_distances': int(np.isnan(dist).sum()), 'bad_status': int((status != 0).sum()),
            'parity_pair01_abs_err_d2': float(abs(dist[0] ** 2 - d_ref ** 2)),
            'cpu': {'kind': 'port', 'precompute_seconds_per_model': c_pre, 'seconds_per_pair': c_pair,
                    'projected_seconds_all_pairs': c_pair * ii.size + c_pre * n}}


def c4(eng, budget):
    """evalg generation: 512 candidates, D = 8, N = 4096: one NLML+grad evaluation of the whole population."""
    import torch
    from ms_project_b200 import workloads
    from ms_project_b200.program import ProgramBatch
    from oracle import gp as ogp
    from oracle.adapter import to_oracle_expr
    X, Y, kernels = workloads.config_c4()
    batch = ProgramBatch(kernels)
    slots = eng.slots_for(X.shape[0], X.shape[1], batch.n_items, batch.max_params, mem_fraction=0.85)
    eng.bind(X, Y, slots, batch.max_params)
    eng.upload_programs(batch)
    rng = np.random.default_rng(0)
    thetas = [np.exp(0.3 * rng.standard_normal(int(p))) for p in batch.n_params]
    noises = np.full(batch.n_items, 0.1)
    theta = batch.pack_theta(thetas, noises)
    eng.lml_grad(batch, theta)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lml, dlml, status, tries = eng.lml_grad(batch, theta)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    N = X.shape[0]
    c0, n_cpu = time.perf_counter(), 0
    for i in range(batch.n_items):
        ref, g, _ = ogp.lml_and_grad(to_oracle_expr(kernels[i]), thetas[i], noises[i], X, Y)
        a, b_ = batch.theta_off[i], batch.theta_off[i + 1]
        assert abs(lml[i] - ref) <= 1e-8 * max(1.0, abs(ref)), (i, lml[i], ref)
        assert np.max(np.abs(dlml[a:b_] - g)) <= 1e-6 * max(1.0, np.max(np.abs(g)))
        n_cpu += 1
        if time.perf_counter() - c0 > budget:
            break
    cdt = time.perf_counter() - c0
    # population diversity statistic of every generation (pairwise centred alignment, covariance.py:320-338)
    from ms_project_b200.covariance import Covariance, pairwise_centered_alignments
    from oracle import kernels as okern, scoring as osc
    covs = [Covariance(k) for k in kernels]
    pairwise_centered_alignments(covs[:8], X, engine=eng)
    torch.cuda.synchronize()
    a0 = time.perf_counter()
    A = pairwise_centered_alignments(covs, X, engine=eng)
    torch.cuda.synchronize()
    t_align = time.perf_counter() - a0
    K0 = okern.K(to_oracle_expr(kernels[0]), kernels[0].param_array, X, None, 'direct')
    K1 = okern.K(to_oracle_expr(kernels[1]), kernels[1].param_array, X, None, 'direct')
    a0 = time.perf_counter()
    ref01 = osc.centered_alignment(K0, K1)
    c_align = time.perf_counter() - a0
    npairs = batch.n_items * (batch.n_items - 1) // 2
    align = {'seconds_all_pairs': t_align, 'pairs': npairs, 'fp64_tflops_gram': npairs * float(N) ** 2 / t_align / 1e12,
             'parity_pair01_abs_err': abs(2 * A[0, 1] - ref01), 'cpu_seconds_per_pair': c_align,
             'cpu_projected_seconds_all_pairs': c_align * npairs}
    return {'config': 'C4 evalg population 512, D=8, N=4096: NLML+grad of every candidate (one generation step)',
            'items': batch.n_items, 'slots': slots, 'seconds': dt, 'evaluations_per_s': batch.n_items / dt,
            'fp64_tflops': batch.n_items * float(N) ** 3 / dt / 1e12, 'status_counts': np.bincount(status, minlength=4).tolist(),
            'parity_checked_items': n_cpu, 'centered_alignments': align,
            'cpu': {'kind': 'port', 'items': n_cpu, 'seconds': cdt, 'evaluations_per_s': n_cpu / cdt}}


def c5(eng, budget):
    """Single large fit N = 16384, D = 2, SE*PER + RQ: blocked DMMA Cholesky alone and the full LML + gradient."""
    import torch
    import scipy.linalg
    from ms_project_b200 import workloads, _lib
    from ms_project_b200.program import ProgramBatch
    X, Y, k, noise = workloads.config_c5()
    n = X.shape[0]
    batch = ProgramBatch([k])
    eng.bind(X, Y, 1, batch.max_params)
    eng.upload_programs(batch)
    theta = batch.pack_theta([k.param_array], [noise])
    eng.set_profiling(True)
    best = None
    for r in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lml, dlml, st, tr = eng.lml_grad(batch, theta)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ph = eng.last_phase_ms()
        if best is None or dt < best[0]:
            best = (dt, ph)
    K = torch.empty((n, n), dtype=torch.float64, device='cuda')
    rc = eng.lib.b200gp_build_K(eng.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(), batch.d_theta.data_ptr(),
                                batch.d_theta_off.data_ptr(), batch.d_all_ids.data_ptr(), 1, 1, K.data_ptr())
    _lib.check(eng.ctx, rc, 'build_K')
    K0 = K.clone()
    pot = []
    for r in range(3):
        K.copy_(K0)
        torch.cuda.synchronize()
        ld, pst = eng.potrf(K)
        pot.append(eng.last_phase_ms()[0])
    eng.set_profiling(False)
    # CPU: LAPACK dpotrf of the same matrix on the host cores
    Kh = K0.cpu().numpy()
    c0 = time.perf_counter()
    L = scipy.linalg.cholesky(Kh, lower=True, overwrite_a=True, check_finite=False)
    cdt = time.perf_counter() - c0
    ld_cpu = 2.0 * float(np.sum(np.log(np.diag(L))))
    dt, ph = best
    return {'config': 'C5 single fit N=16384, D=2, SE_1*PER_1+RQ_2', 'lml': float(lml[0]), 'status': int(st[0]),
            'lml_grad_seconds': dt, 'lml_grad_fp64_tflops': float(n) ** 3 / dt / 1e12,
            'phases_ms': dict(zip(['build_K', 'cholesky', 'inverse', 'solve', 'Kinv', 'grad'], ph)),
            'potrf_ms': min(pot), 'potrf_fp64_tflops': float(n) ** 3 / 3 / (min(pot) * 1e-3) / 1e12,
            'potrf_frac_of_dmma_peak': float(n) ** 3 / 3 / (min(pot) * 1e-3) / 1e12 / 36.94,
            'logdet': float(ld), 'logdet_cpu': ld_cpu, 'logdet_rel_err': abs(ld - ld_cpu) / abs(ld_cpu),
            'cpu': {'kind': 'LAPACK dpotrf (scipy) on the same K_y', 'seconds': cdt, 'fp64_tflops': float(n) ** 3 / 3 / cdt / 1e12}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='c1,c3,c4,c5')
    ap.add_argument('--cpu-budget', type=float, default=15.0)
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    from ms_project_b200.engine import get_engine
    eng = get_engine(0)
    cores = host_cores()
    out = {'host_cores': cores}
    for name in args.configs.split(','):
        fn = {'c1': c1, 'c3': c3, 'c3full': c3full, 'c4': c4, 'c5': c5}[name.strip()]
        res = fn(eng, args.cpu_budget)
        out[name.strip()] = res
        print(json.dumps({name.strip(): res}))
        sys.stdout.flush()
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
