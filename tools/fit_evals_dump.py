"""Per-model evaluation counts of the C2 fit (256 candidates x 3 restarts) -> gpurun_out/fit_evals.json, for studying how
well an a-priori cost predicts the work of a model (shard balance of parallel.Comm)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, '.')
import bench
from ms_project_b200 import fitness, workloads, program
from ms_project_b200.covariance import Covariance
from ms_project_b200.engine import Engine
from ms_project_b200.gp_model import GPModel, score_models, model_cost
from ms_project_b200.gp_models import gp_regression

X, y, items, _, _ = bench.build_items()
kernels = [items[3 * i].copy() for i in range(256)]
models = [GPModel(Covariance(k)) for k in kernels]
md = gp_regression(likelihood_hyperprior={'variance': workloads.boms_hyperpriors()['GP']['variance']})
eng = Engine(0)
np.random.seed(0)
t0 = time.time()
score_models(models, X, y, fitness.negative_bic, n_restarts=3, engine=eng, model_dict=md)
dt = time.time() - t0
out = []
for m in models:
    k = m.covariance.raw_kernel
    fams = []
    k.traverse(lambda q: fams.append(q.op_name) if getattr(q, 'op_name', None) in ('SE', 'RQ', 'PER', 'LIN') else None)
    out.append({'expr': program.infix_string(k), 'n_params': int(k.size), 'families': fams, 'n_evals': int(m.fit_result.n_evals),
                'runs': [int(r['funcalls']) for r in m.fit_result.runs], 'cost': model_cost(k, X.shape[0])})
os.makedirs('gpurun_out', exist_ok=True)
json.dump({'seconds': dt, 'models': out}, open('gpurun_out/fit_evals.json', 'w'))
print('fit', dt, 's; evals', sum(o['n_evals'] for o in out))
