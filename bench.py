#!/usr/bin/env python
"""bench.py -- headline benchmark of the candidate-kernel scoring path (BASELINE.json).

Metric: GP candidate kernels scored / sec (N=2048, NLML+grad); FP64 TFLOP/s = value * N^3.
Workload (config C2): 256 random CKS-grammar kernels x 3 restart draws = 768
(kernel, theta) items over one synthetic 1-D data set with N = 2048.  One STEP = one NLML+gradient
evaluation of every item (the unit an L-BFGS-B iteration of `optimize_restarts` consumes).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); candidates shard across ranks with no traffic during
scoring and one NCCL all_gather of the scores per step (weak scaling: 768 items per GPU).
`--impl reference` times the CPU oracle (the NumPy/SciPy-LAPACK restatement of the reference's GPy path --
GPy itself is not installable) on the host cores, on a bounded sample of the same items.
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


def cpu_reference_rate(X, y, items, thetas, noises, budget_s, max_items):
    """NLML+grad evaluations / s of the CPU oracle (all BLAS threads), sequential over items like
    ModelSelector._evaluate_models (src/autoks/core/model_selection/base.py:330)."""
    from oracle import gp as ogp
    from oracle.adapter import to_oracle_expr
    ncpu = len(os.sched_getaffinity(0))
    try:        # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core it can use
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=ncpu)
        cores = max([p.get('num_threads', 1) for p in threadpool_info()] or [1])
    except Exception:
        cores = ncpu
    done, t0 = 0, time.time()
    for i in range(min(max_items, len(items))):
        try:
            ogp.lml_and_grad(to_oracle_expr(items[i]), thetas[i], noises[i], X, y)
        except np.linalg.LinAlgError:
            pass
        done += 1
        if time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    return done / dt, cores, done, dt


def fit_rate(eng, X, y, items, n_candidates=N_KERNELS):
    """Secondary figure: FULLY FITTED candidates / s (optimize_restarts with 3 restarts, L-BFGS-B maxfun 1000) of the
    whole C2 configuration (256 kernels), through the public scoring API (GPModel batch seam), BOMS hyperpriors."""
    from ms_project_b200 import fitness, workloads
    from ms_project_b200.covariance import Covariance
    from ms_project_b200.gp_model import GPModel, score_models
    from ms_project_b200.gp_models import gp_regression
    kernels = [items[3 * i].copy() for i in range(n_candidates)]
    models = [GPModel(Covariance(k)) for k in kernels]
    md = gp_regression(likelihood_hyperprior={'variance': workloads.boms_hyperpriors()['GP']['variance']})
    np.random.seed(0)
    t0 = time.perf_counter()
    scores = score_models(models, X, y, fitness.negative_bic, n_restarts=N_RESTARTS, engine=eng, model_dict=md)
    dt = time.perf_counter() - t0
    evals = sum(m.fit_result.n_evals for m in models)
    return {'candidates': n_candidates, 'restarts': N_RESTARTS, 'seconds': dt, 'candidates_per_s': n_candidates / dt,
            'nlml_grad_evaluations': int(evals), 'evaluations_per_s': evals / dt,
            'failed': int(sum(m.failed_fitting for m in models)), 'best_nbic': float(np.nanmax(scores))}


def run_reference(args, rank, world):
    if rank != 0:
        return
    X, y, items, thetas, noises = build_items()
    per_step = 4
    # warm-up + timed steps, each a bounded sample of the same workload
    for w in range(args.warmup):
        cpu_reference_rate(X, y, items[w:], thetas[w:], noises[w:], 1e9, 1)
    t0 = time.time()
    n_done = 0
    cores = 1
    for s in range(args.steps):
        lo = (s * per_step) % len(items)
        _, cores, d, _ = cpu_reference_rate(X, y, items[lo:], thetas[lo:], noises[lo:], 1e9, per_step)
        n_done += d
    dt = time.time() - t0
    val = n_done / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / max(args.steps, 1) * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C2: 256 random CKS-grammar kernels x 3 restart draws, N=2048, D=1; NLML+grad per item',
                   'N': N_DATA, 'items': len(items)},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d items per step x %d steps of the C2 item list, sequential, NumPy/SciPy-LAPACK oracle '
                                   '(GPy fork not installable; oracle restates its algorithm)' % (per_step, args.steps)},
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
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic_r01.json')))
        per_item = (tr['kinv64_kernel']['dram_bytes_read'] + tr['kinv64_kernel']['dram_bytes_write']) / tr['items_per_launch']
        traffic = per_item * chunk_items
        traffic_note = ('dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of kinv64_kernel at %d items per '
                        'launch, scaled to the %d items of this launch (algorithmic: 8 N^2 bytes per item = read V upper + '
                        'write K^-1 lower)' % (tr['items_per_launch'], chunk_items))
    except Exception:
        pass

    if rank == 0:
        total_items = world * n_items
        value = total_items * args.steps / (ms_max * 1e-3)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C2: 256 random CKS-grammar kernels x 3 restart draws = 768 (kernel, theta) items per GPU, '
                                   'N=2048, D=1; one step = one NLML+grad evaluation of every item',
                       'N': N_DATA, 'items_per_gpu': n_items, 'slots': slots, 'parallelism': 'candidate-shard x%d' % world,
                       'l2': 'working set %.1f GiB per step >> 126 MB L2 (inputs larger than L2)' % (n_items * 2 * 2048 * 2048 * 8 / 2 ** 30)},
            'fp64_tflops': value * float(N_DATA) ** 3 / 1e12,
            'e2e': {'value': total_items * args.steps / e2e_s, 'unit': UNIT,
                    'h2d_bytes_per_step': int(theta.nbytes), 'd2h_bytes_per_step': int(batch.h_out.numel() * 8 + batch.h_int.numel() * 4)},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
            'roofline': {'bound': 'tensor', 'achieved': kinv_tflops, 'peak': peak, 'unit': 'TFLOP/s', 'frac': kinv_tflops / peak,
                         'traffic': traffic, 'traffic_note': traffic_note, 'algorithmic_bytes': 8.0 * N_DATA ** 2 * chunk_items,
                         'kernel': 'kinv64_kernel (FP64 DMMA tile GEMM on 64 x 64 CTA tiles, K^-1 = V V^T; 1 launch per step)',
                         'algorithmic': 'N^3/3 flops per item x %d items per launch' % chunk_items,
                         'launch_ms': phases[4], 'peak_source': peak_src},
            'roofline_dmma_phases': {'achieved': dmma_tflops, 'peak': peak, 'unit': 'TFLOP/s', 'frac': dmma_tflops / peak,
                                     'kernels': 'chol_update64/chol_trsm/potrf_diag + inv_update64/inv_trsm + kinv64 (Cholesky, L^-T, K^-1)',
                                     'algorithmic': 'N^3 flops per item', 'ms': dmma_ms},
            'phases_ms': {'build_K': phases[0], 'cholesky': phases[1], 'inverse': phases[2], 'solve': phases[3],
                          'Kinv': phases[4], 'grad': phases[5], 'items': chunk_items},
            'status_counts': np.bincount(status, minlength=4).tolist(),
        }
        if world == 1 and not args.no_fit:
            line['fit'] = fit_rate(eng, X, y, items)
        if world == 1 and not args.no_cpu:
            rate, cores, done, dt = cpu_reference_rate(X, y, items, thetas, noises, budget_s=20.0, max_items=64)
            line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': 'first %d of the 768 items, sequential, %.1f s (NumPy/SciPy-LAPACK oracle)' % (done, dt)}
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
