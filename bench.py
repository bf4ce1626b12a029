#!/usr/bin/env python
"""bench.py -- headline benchmark of the candidate-kernel scoring path (BASELINE.json).

Metric: GP candidate kernels scored / sec (N=2048, NLML+grad); FP64 TFLOP/s = value * N^3.
Workload (config C2): 256 random CKS-grammar kernels x 3 restart draws = 768
(kernel, theta) items over one synthetic 1-D data set with N = 2048.  One STEP = one NLML+gradient
evaluation of every item (the unit an L-BFGS-B iteration of `optimize_restarts` consumes).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); candidates shard across ranks with no traffic during
scoring and one NCCL all_gather of the scores per step (weak scaling: 768 items per GPU).  The same line carries the
STRONG-scaling legs (fixed total work sharded over the N ranks through parallel.score_models_sharded): `fit` = the whole
C2 configuration fitted (256 candidates x 3 restarts), `strong_c4` = one C4 generation (512 candidates, N = 4096, D = 8):
one NLML+grad evaluation of the population and a 3-restart fit with a stated optimiser budget.
`--impl reference` times the CPU oracle (the NumPy/SciPy-LAPACK restatement of the reference's GPy path --
GPy itself is not installable) on ALL host cores -- a process pool, one candidate per core, BLAS threads = 1,
the analogue of src/training/run_experiment_from_file.py:81-84 -- on a stratified sample of the same 768 items.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'candidate_kernels_scored_per_sec_N2048_nlml_grad'
UNIT = 'kernels/s'
N_DATA = 2048
N_KERNELS = 256
N_RESTARTS = 3
C4_MAX_ITERS = 10          # optimiser budget of the C4 strong-scaling fit (stated in the line; paramz's default is 1000)


def shared_config(world):
    """`config` of BOTH arms (ours and --impl reference): same keys, same values."""
    n_items = N_KERNELS * N_RESTARTS
    return {'workload': 'C2: 256 random CKS-grammar kernels x 3 restart draws = 768 (kernel, theta) items per GPU, '
                        'N=2048, D=1; one step = one NLML+grad evaluation of every item',
            'N': N_DATA, 'items_per_gpu': n_items, 'parallelism': 'candidate-shard x%d' % world,
            'l2': 'working set %.1f GiB per step >> 126 MB L2 (inputs larger than L2)' % (n_items * 2 * 2048 * 2048 * 8 / 2 ** 30)}


def sample_indices(n_items, count):
    """Stratified sample of the item list: `count` indices evenly spread over all items (every kernel family and every
    restart draw is represented), the same in every process."""
    count = max(1, min(count, n_items))
    return [int(i) for i in np.linspace(0, n_items - 1, count).round()]


def build_items(seed_offset=0, n_kernels=N_KERNELS, n_restarts=N_RESTARTS, n=N_DATA):
    """C2 items: kernels (each repeated n_restarts times) and their hyperparameters.  Restart 0 uses the GPy
    defaults; restarts > 0 draw from the BOMS hyperpriors exactly like paramz `randomize` does."""
    from ms_project_b200 import workloads
    X, y, kernels = workloads.config_c2(n=n, n_kernels=n_kernels, seed=1 + seed_offset)
    rng = np.random.RandomState(100 + seed_offset)
    noise_prior = workloads.boms_hyperpriors()['GP']['variance']
    items, thetas, noises = [], [], []
    for k in kernels:
        for r in range(n_restarts):
            kk = k.copy()
            if r == 0:
                nz = 1.0
            else:
                state = np.random.get_state()
                np.random.set_state(rng.get_state())
                kk.randomize()
                nz = float(noise_prior.rvs(1)[0])
                rng.set_state(np.random.get_state())
                np.random.set_state(state)
            items.append(kk)
            thetas.append(kk.param_array.copy())
            noises.append(nz)
    return X, y, items, thetas, noises


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons, 'samples': len(sm)}


def fp64_peak():
    """FP64 tensor-pipe peak.  MEASURED_PEAKS.json carries HBM and bf16 only, so the denominator is the
    DMMA m8n8k4 issue rate measured on this pool's B200 by tools/fp64_peaks.cu (profiles/fp64_peaks_r01.json)."""
    try:
        p = json.load(open(os.path.join(ROOT, 'profiles', 'fp64_peaks_r01.json')))
        return float(p['dmma_tflops']['best']), 'profiles/fp64_peaks_r01.json (DMMA m8n8k4 register-resident, measured on B200)'
    except Exception:
        return 36.9, 'fallback: 36.9 TFLOP/s DMMA measured earlier on this pool'


def host_cores():
    return len(os.sched_getaffinity(0))


def cpu_reference_rate(X, y, items, thetas, noises, budget_s, indices):
    """Variant A: NLML+grad evaluations / s of the CPU oracle, sequential over items like
    ModelSelector._evaluate_models (src/autoks/core/model_selection/base.py:330), every BLAS thread on each item.
    Items that raise LinAlgError are not counted as completed."""
    from oracle import gp as ogp
    from oracle.adapter import to_oracle_expr
    ncpu = host_cores()
    try:        # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core it can use
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=ncpu)
        cores = max([p.get('num_threads', 1) for p in threadpool_info()] or [1])
    except Exception:
        cores = ncpu
    done, failed, t0 = 0, 0, time.time()
    for i in indices:
        try:
            ogp.lml_and_grad(to_oracle_expr(items[i]), thetas[i], noises[i], X, y)
            done += 1
        except np.linalg.LinAlgError:
            failed += 1
        if time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    return done / dt, cores, done, failed, dt


_POOL_STATE = {}


def _pool_init(X, y, exprs, thetas, noises):
    try:
        from threadpoolctl import threadpool_limits
        _POOL_STATE['limit'] = threadpool_limits(limits=1)      # one BLAS thread per worker
    except Exception:
        pass
    _POOL_STATE.update(X=X, y=y, exprs=exprs, thetas=thetas, noises=noises)


def _pool_eval(i):
    from oracle import gp as ogp
    st = _POOL_STATE
    try:
        ogp.lml_and_grad(st['exprs'][i], st['thetas'][i], st['noises'][i], st['X'], st['y'])
        return 1
    except np.linalg.LinAlgError:
        return 0


class CpuPool(object):
    """Variant B, the strongest CPU baseline: a process pool with one candidate per core and one BLAS thread per process
    (how the reference parallelises experiments: src/training/run_experiment_from_file.py:81-84)."""

    def __init__(self, X, y, items, thetas, noises):
        import multiprocessing as mp
        from oracle.adapter import to_oracle_expr
        self.cores = host_cores()
        exprs = [to_oracle_expr(k) for k in items]
        ctx = mp.get_context('fork' if 'torch' not in sys.modules else 'spawn')
        os.environ['OMP_NUM_THREADS'] = '1'
        self.pool = ctx.Pool(self.cores, initializer=_pool_init, initargs=(X, y, exprs, thetas, noises))

    def rate(self, indices):
        t0 = time.time()
        done = sum(self.pool.map(_pool_eval, indices, chunksize=1))
        dt = time.time() - t0
        return done / dt, done, len(indices) - done, dt

    def close(self):
        self.pool.close()
        self.pool.join()


def _timed_collective(fn, world, device):
    """Wall time of a host-driven, multi-rank leg: barrier + synchronize on both sides, MAX over ranks."""
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return out, dt


def fit_rate(eng, X, y, items, world, n_candidates=N_KERNELS):
    """STRONG scaling leg 1: the whole C2 configuration FULLY FITTED (optimize_restarts with 3 restarts, L-BFGS-B maxfun
    1000, BOMS hyperpriors) through the public scoring API; with N ranks the 256 candidates are sharded (LPT on an a-priori
    cost) and the records all-gathered, every rank ends with all 256 scores."""
    from ms_project_b200 import fitness, workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel
    from ms_project_b200.gp_models import gp_regression
    from ms_project_b200.parallel import score_models_sharded
    kernels = [items[3 * i].copy() for i in range(n_candidates)]
    models = [GPModel(Covariance(k)) for k in kernels]
    md = gp_regression(likelihood_hyperprior={'variance': workloads.boms_hyperpriors()['GP']['variance']})
    np.random.seed(0)
    scores, dt = _timed_collective(lambda: score_models_sharded(models, X, y, fitness.negative_bic, n_restarts=N_RESTARTS,
                                                                engine=eng, model_dict=md), world, eng.device)
    evals = sum(m.fit_result.n_evals for m in models)
    return {'config': 'C2 fitted: 256 candidates x 3 restarts, N=2048, L-BFGS-B maxfun 1000; total work fixed, sharded over the ranks',
            'scaling': 'strong', 'n_gpus': world, 'candidates': n_candidates, 'restarts': N_RESTARTS, 'seconds': dt,
            'candidates_per_s': n_candidates / dt, 'nlml_grad_evaluations': int(evals), 'evaluations_per_s': evals / dt,
            'failed': int(sum(m.failed_fitting for m in models)), 'best_nbic': float(np.nanmax(scores))}


def strong_c4(eng, world, rank, generations=1):
    """STRONG scaling leg 2: one generation of configuration C4 (512 candidates, N = 4096, D = 8) sharded over the ranks:
    (i) one NLML+grad evaluation of the whole population, (ii) a 3-restart fit with max_iters = C4_MAX_ITERS.
    `generations` > 1 (--c4-generations; BASELINE.json's C4 names 20) repeats (ii) with a fresh population of 512 random trees
    per generation over the same data -- the work of the evolutionary search's scoring path without its (reference-side,
    host) variation operators."""
    import torch
    import torch.distributed as dist
    from ms_project_b200 import fitness, workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel
    from ms_project_b200.gp_models import gp_regression
    from ms_project_b200.parallel import score_models_sharded, lpt_assign
    from ms_project_b200.gp_model import model_cost
    from ms_project_b200.program import ProgramBatch
    X, Y, kernels = workloads.config_c4()
    n = len(kernels)
    owner = lpt_assign([model_cost(k, X.shape[0]) for k in kernels], world)
    mine = [i for i in range(n) if owner[i] == rank]
    batch = ProgramBatch([kernels[i] for i in mine])
    slots = eng.slots_for(X.shape[0], X.shape[1], batch.n_items, batch.max_params, mem_fraction=0.8,
                          extra_bytes=0 if eng._ws is None else eng._ws.numel())
    eng.bind(X, Y, slots, batch.max_params)
    eng.upload_programs(batch)
    rng = np.random.default_rng(0)
    thetas = [np.exp(0.3 * rng.standard_normal(int(k.size))) for k in kernels]
    theta = batch.pack_theta([thetas[i] for i in mine], np.full(len(mine), 0.1))
    gathered = torch.empty(world * ((n + world - 1) // world + 64), dtype=torch.float64, device=eng.device)

    def eval_step():
        out = eng.lml_grad(batch, theta)
        if world > 1:
            pad = torch.zeros(gathered.numel() // world, dtype=torch.float64, device=eng.device)
            pad[:len(mine)] = batch.d_out[:len(mine)]
            dist.all_gather_into_tensor(gathered, pad)
        return out
    eval_step()
    out, dt_eval = _timed_collective(eval_step, world, eng.device)
    status = np.asarray(out[2])[:len(mine)]
    models = [GPModel(Covariance(k.copy())) for k in kernels]
    md = gp_regression(likelihood_hyperprior={'variance': workloads.boms_hyperpriors()['GP']['variance']})
    np.random.seed(1)
    scores, dt_fit = _timed_collective(lambda: score_models_sharded(models, X, Y, fitness.log_likelihood_normalized, n_restarts=3,
                                                                    engine=eng, model_dict=md, max_iters=C4_MAX_ITERS),
                                       world, eng.device)
    evals = sum(m.fit_result.n_evals for m in models)
    N = float(X.shape[0])
    out = {'config': 'C4 evalg generation: 512 candidates, N=4096, D=8; total work fixed, sharded over the ranks (LPT)',
           'scaling': 'strong', 'n_gpus': world, 'candidates': n,
           'eval_step_seconds': dt_eval, 'eval_step_evaluations_per_s': n / dt_eval, 'eval_step_fp64_tflops': n * N ** 3 / dt_eval / 1e12,
           'eval_status_counts_rank0': np.bincount(status, minlength=4).tolist(),
           'fit_restarts': 3, 'fit_max_iters': C4_MAX_ITERS, 'fit_seconds': dt_fit, 'fit_candidates_per_s': n / dt_fit,
           'fit_nlml_grad_evaluations': int(evals), 'fit_evaluations_per_s': evals / dt_fit,
           'fit_failed': int(sum(m.failed_fitting for m in models)), 'fit_best_loglikn': float(np.nanmax(scores))}
    if generations > 1:
        total, total_evals, best = dt_fit, int(evals), float(np.nanmax(scores))
        for g in range(1, generations):
            _x, _y, ks = workloads.config_c4(seed=3 + g)               # a fresh population; the data stay those of seed 3
            pop = [GPModel(Covariance(k)) for k in ks]
            np.random.seed(1 + g)
            sc, dt_g = _timed_collective(lambda: score_models_sharded(pop, X, Y, fitness.log_likelihood_normalized, n_restarts=3,
                                                                      engine=eng, model_dict=md, max_iters=C4_MAX_ITERS),
                                         world, eng.device)
            total += dt_g
            total_evals += sum(m.fit_result.n_evals for m in pop)
            best = max(best, float(np.nanmax(sc)))
        out['generations'] = {'count': generations, 'candidates_per_generation': n, 'fit_seconds_total': total,
                              'nlml_grad_evaluations': total_evals, 'evaluations_per_s': total_evals / total,
                              'best_loglikn': best}
    return out


def c5_potrf(eng):
    """The north star's C5 figure, driver-visible: blocked FP64 DMMA Cholesky of the N = 16384 K_y (SE_1*PER_1 + RQ_2, D = 2),
    timed by CUDA events around the factorisation inside the library."""
    import torch
    from ms_project_b200 import workloads, _lib
    from ms_project_b200.program import ProgramBatch
    X, Y, k, noise = workloads.config_c5()
    n = X.shape[0]
    batch = ProgramBatch([k])
    eng.bind(X, Y, 1, batch.max_params)
    eng.upload_programs(batch)
    theta = batch.pack_theta([k.param_array], [noise])
    batch.h_theta.numpy()[:] = theta
    batch.d_theta.copy_(batch.h_theta)
    K0 = torch.empty((n, n), dtype=torch.float64, device=eng.device)
    rc = eng.lib.b200gp_build_K(eng.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(), batch.d_theta.data_ptr(),
                                batch.d_theta_off.data_ptr(), batch.d_all_ids.data_ptr(), 1, 1, K0.data_ptr())
    _lib.check(eng.ctx, rc, 'b200gp_build_K')
    K = torch.empty_like(K0)
    eng.set_profiling(True)
    ms, ld, st = [], 0.0, 0
    for _ in range(4):
        K.copy_(K0)
        torch.cuda.synchronize()
        ld, st = eng.potrf(K)
        ms.append(eng.last_phase_ms()[0])
    t0 = time.perf_counter()
    lml, _g, st2, _t = eng.lml_grad(batch, theta)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    eng.set_profiling(False)
    peak, _ = fp64_peak()
    best = min(ms[1:])
    return {'config': 'C5: N=16384, D=2, SE_1*PER_1+RQ_2', 'potrf_ms': best, 'potrf_fp64_tflops': float(n) ** 3 / 3 / (best * 1e-3) / 1e12,
            'potrf_frac_of_dmma_peak': float(n) ** 3 / 3 / (best * 1e-3) / 1e12 / peak, 'logdet': float(ld), 'status': int(st),
            'lml_grad_seconds': dt, 'lml_grad_fp64_tflops': float(n) ** 3 / dt / 1e12, 'lml': float(lml[0]), 'lml_status': int(st2[0])}


def run_reference(args, rank, world):
    """The reference's CPU path on ALL host cores of the box, whatever N is (rank 0 alone; the other ranks exit): every step
    pool-evaluates one stratified sample of the 768 items, one candidate per core, BLAS threads = 1."""
    if rank != 0:
        return
    X, y, items, thetas, noises = build_items()
    pool = CpuPool(X, y, items, thetas, noises)
    per_step = pool.cores
    n_steps = args.warmup + args.steps
    idx = sample_indices(len(items), per_step * n_steps)
    for w in range(args.warmup):
        pool.rate(idx[w::n_steps])
    t0 = time.time()
    n_done = n_failed = 0
    for st in range(args.warmup, n_steps):
        _, d, f, _ = pool.rate(idx[st::n_steps])
        n_done += d
        n_failed += f
    dt = time.time() - t0
    pool.close()
    val = n_done / dt
    # variant A beside it: sequential over items, every BLAS thread on each
    rate_a, cores_a, done_a, failed_a, dt_a = cpu_reference_rate(X, y, items, thetas, noises, 10.0, sample_indices(len(items), 12))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / max(args.steps, 1) * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': shared_config(args.gpus),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': pool.cores, 'kind': 'port',
                         'sample': 'process pool, one candidate per core, BLAS threads = 1 (variant B): %d items per step x %d steps, a '
                                   'stratified sample (evenly spaced indices) of the 768 C2 items, %d evaluated, %d not PD; NumPy/SciPy-LAPACK '
                                   'oracle (GPy fork not installable; the oracle restates its algorithm)'
                                   % (per_step, args.steps, n_done, n_failed)},
        'cpu_baseline_sequential': {'value': rate_a, 'unit': UNIT, 'cores': cores_a, 'kind': 'port',
                                    'sample': 'sequential over %d evenly spaced items, all BLAS threads per item (variant A), %.1f s, %d not PD'
                                              % (done_a, dt_a, failed_a)},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'fp64_tflops': val * N_DATA ** 3 / 1e12,
    }
    print(json.dumps(line))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ms_project_b200.engine import Engine
    from ms_project_b200.program import ProgramBatch

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    X, y, items, thetas, noises = build_items(seed_offset=rank)
    eng = Engine(local_rank)
    batch = ProgramBatch(items)
    n_items = batch.n_items
    slots = eng.slots_for(N_DATA, 1, n_items, batch.max_params, mem_fraction=0.85)
    eng.bind(X, y, slots, batch.max_params)
    eng.upload_programs(batch)
    theta = batch.pack_theta(thetas, noises)
    batch.h_theta.numpy()[:] = theta
    batch.d_theta.copy_(batch.h_theta)
    gathered = torch.empty(world * n_items, dtype=torch.float64, device=eng.device) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        eng.lml_grad_device(batch)
        if world > 1:
            dist.all_gather_into_tensor(gathered, batch.d_out[:n_items])

    def step_e2e():
        out = eng.lml_grad(batch, theta)
        if world > 1:
            dist.all_gather_into_tensor(gathered, batch.d_out[:n_items])
        return out

    # ---- device-resident timing (value); phase events (CUDA events on the context stream) stay on: they are what
    #      the roofline of the dominant kernel is computed from, inside the timed region ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.set_profiling(True)
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(eng.stream)
    phase_sum = np.zeros(6)
    for _ in range(args.steps):
        step_device()
        phase_sum += np.asarray(eng.last_phase_ms()[:6])      # lml_grad_device has synchronised the stream already
    ev1.record(eng.stream)
    barrier()
    eng.set_profiling(False)
    launches = eng.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    status = batch.d_int[:n_items].cpu().numpy()

    # ---- end-to-end through the host API (pinned host buffers, H2D + D2H inside the timed region) ----
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    # ---- roofline of the dominant kernel: kinv64_kernel (one launch per step: K^-1 = V V^T on the DMMA pipe,
    #      N^3/3 flops per item), timed by the library's CUDA events around that launch in every timed step ----
    phases = (phase_sum / max(args.steps, 1)).tolist()
    chunk_items = n_items - ((n_items + slots - 1) // slots - 1) * slots     # items of the (last) chunk the events cover
    peak, peak_src = fp64_peak()
    kinv_tflops = chunk_items * float(N_DATA) ** 3 / 3.0 / (phases[4] * 1e-3) / 1e12
    dmma_ms = phases[1] + phases[2] + phases[4]
    dmma_tflops = chunk_items * float(N_DATA) ** 3 / (dmma_ms * 1e-3) / 1e12
    traffic, traffic_note = None, None
    try:
        cands = sorted(f for f in os.listdir(os.path.join(ROOT, 'profiles')) if f.startswith('ncu_traffic_') and f.endswith('.json'))
        tr = json.load(open(os.path.join(ROOT, 'profiles', cands[-1])))            # the latest capture (r02... sorts after r01)
        per_item = (tr['kinv64_kernel']['dram_bytes_read'] + tr['kinv64_kernel']['dram_bytes_write']) / tr['items_per_launch']
        traffic = per_item * chunk_items
        traffic_note = ('dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of the kinv64 kernel at %d items per '
                        'launch (%s), scaled to the %d items of this launch (algorithmic: 8 N^2 bytes per item = read V upper + '
                        'write K^-1 lower)' % (tr['items_per_launch'], cands[-1], chunk_items))
    except Exception:
        pass

    # ---- legs every rank takes part in (collectives inside): the strong-scaling fits; their workspaces replace the headline's
    if not args.no_fit and rank != 0:        # the strong legs share ONE candidate list (rank 0's); the headline's differ per rank
        X, y, items, _, _ = build_items(seed_offset=0)
    fit = None if args.no_fit else fit_rate(eng, X, y, items, world)
    c4 = None if args.no_fit else strong_c4(eng, world, rank, args.c4_generations)

    if rank == 0:
        total_items = world * n_items
        value = total_items * args.steps / (ms_max * 1e-3)
        flops_item = float(N_DATA) ** 3 / 3.0

        def phase(ms_, flops, kernels):
            tf = chunk_items * flops / (ms_ * 1e-3) / 1e12
            return {'ms': ms_, 'achieved': tf, 'peak': peak, 'unit': 'TFLOP/s', 'frac': tf / peak, 'kernels': kernels}
        step_ms = ms_max / args.steps
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': shared_config(world),
            'details': {'slots': slots, 'tma': 'TMA-fed DMMA tile GEMMs (UTMALDG + mbarrier ring); cp.async kernels when the tensor maps cannot be encoded'},
            'fp64_tflops': value * float(N_DATA) ** 3 / 1e12,
            'e2e': {'value': total_items * args.steps / e2e_s, 'unit': UNIT,
                    'h2d_bytes_per_step': int(theta.nbytes), 'd2h_bytes_per_step': int(batch.h_out.numel() * 8 + batch.h_int.numel() * 4),
                    'api': 'Engine.lml_grad (host numpy theta in, host LML / gradients / status out); the fully fitted rate through '
                           'GPModel / score_models is the `fit` object'},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
            'roofline': {'bound': 'tensor', 'achieved': kinv_tflops, 'peak': peak, 'unit': 'TFLOP/s', 'frac': kinv_tflops / peak,
                         'traffic': traffic, 'traffic_note': traffic_note, 'algorithmic_bytes': 8.0 * N_DATA ** 2 * chunk_items,
                         'kernel': 'kinv64 kernel (FP64 DMMA tile GEMM on 64 x 64 CTA tiles, K^-1 = V V^T; 1 launch per step)',
                         'algorithmic': 'N^3/3 flops per item x %d items per launch' % chunk_items,
                         'launch_ms': phases[4], 'peak_source': peak_src},
            # every DMMA phase and the whole step against the same peak (not only the best kernel)
            'roofline_phases': {
                'cholesky': phase(phases[1], flops_item, 'potrf_diag + trsm + chol_update64 (right-looking blocked Cholesky)'),
                'inverse': phase(phases[2], flops_item, 'trsm + inv_update64 (L^-T)'),
                'kinv': phase(phases[4], flops_item, 'kinv64 (K^-1 = V V^T)'),
                'dmma_phases': phase(dmma_ms, 3 * flops_item, 'the three above, N^3 flops per item'),
                'step': {'ms': step_ms, 'achieved': value / world * float(N_DATA) ** 3 / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': value / world * float(N_DATA) ** 3 / 1e12 / peak,
                         'kernels': 'whole NLML+grad step per GPU incl. K-build, solves, gradient pass: N^3 algorithmic flops per item'}},
            'phases_ms': {'build_K': phases[0], 'cholesky': phases[1], 'inverse': phases[2], 'solve': phases[3],
                          'Kinv': phases[4], 'grad': phases[5], 'items': chunk_items},
            'status_counts': np.bincount(status, minlength=4).tolist(),
        }
        if fit is not None:
            line['fit'] = fit
            line['strong_c4'] = c4
        if world == 1 and not args.no_fit:
            line['c5'] = c5_potrf(eng)
        if world == 1 and not args.no_cpu:
            # the CPU baselines on this box's host cores, on a stratified sample of the SAME item list:
            # B = process pool, one candidate per core, BLAS threads 1 (the strongest); A = sequential, all BLAS threads per item
            pool = CpuPool(X, y, items, thetas, noises)
            idx = sample_indices(n_items, 5 * pool.cores)
            pool.rate(idx[:pool.cores])                    # warm the workers
            rounds = [pool.rate(idx[(1 + 2 * r) * pool.cores:(3 + 2 * r) * pool.cores]) for r in range(2)]
            pool.close()
            rate_b, done_b, failed_b, dt_b = max(rounds, key=lambda t: t[0])       # the FASTER of two rounds: a baseline errs high
            line['cpu_baseline'] = {'value': rate_b, 'unit': UNIT, 'cores': pool.cores, 'kind': 'port',
                                    'sample': 'process pool, one candidate per core, BLAS threads = 1: the faster of two rounds of %d '
                                              'evenly spaced items of the 768 (%.1f s, %d not PD; the other round %.2f kernels/s); '
                                              'NumPy/SciPy-LAPACK oracle' % (done_b, dt_b, failed_b, min(r[0] for r in rounds))}
            rate_a, cores_a, done_a, failed_a, dt_a = cpu_reference_rate(X, y, items, thetas, noises, 12.0, sample_indices(n_items, 16))
            line['cpu_baseline_sequential'] = {'value': rate_a, 'unit': UNIT, 'cores': cores_a, 'kind': 'port',
                                               'sample': 'sequential over %d evenly spaced items, all BLAS threads per item, %.1f s, %d not PD'
                                                         % (done_a, dt_a, failed_a)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-fit', action='store_true', help='skip the fitted-candidates/s leg')
    ap.add_argument('--c4-generations', type=int, default=1, help='generations of the C4 strong-scaling fit leg (BASELINE.json names 20)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
