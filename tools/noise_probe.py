import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from noisy_engine import NoisyEngine
from ms_project_b200 import workloads, fitness
from ms_project_b200.covariance import Covariance
from ms_project_b200.gp_model import GPModel, score_models
from ms_project_b200.gp_models import gp_regression
from ms_project_b200.engine import get_engine
from test_gpu_fit import make_data, _with_priors, from_string
engine = get_engine(0)
X, y = make_data(160, 1, seed=17)
names = ['PER_1', 'PER_1 + SE_1', 'PER_1 * SE_1', 'PER_1 + LIN_1', 'PER_1 * RQ_1', 'PER_1 * SE_1 + RQ_1', 'PER_1 + PER_1',
         'PER_1 * LIN_1 + SE_1', 'SE_1 + RQ_1 * PER_1', 'PER_1 * PER_1']
noise_prior = workloads.boms_hyperpriors()['GP']['variance']
def device_fit(eng):
    models = [GPModel(Covariance(_with_priors(from_string(s)))) for s in names]
    np.random.seed(23)
    score_models(models, X, y, fitness.negative_bic, n_restarts=3, engine=eng,
                 model_dict=gp_regression(likelihood_hyperprior={'variance': noise_prior}))
    return models
base = device_fit(engine)
for lvl in (1e-13, 1e-12, 1e-11, 1e-10):
    for seed in (1, 2, 3):
        m = device_fit(NoisyEngine(engine, lvl, seed))
        print(lvl, seed, [round(r['f_opt'], 4) for r in m[7].fit_result.runs], 'n differing restarts overall',
              sum(abs(a['f_opt'] - b['f_opt']) > 1e-5 * max(1, abs(b['f_opt'])) for i in range(10) for a, b in zip(m[i].fit_result.runs, base[i].fit_result.runs)))
