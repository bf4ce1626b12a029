import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from conftest import from_string, make_data
from test_gpu_fit import _with_priors, _oracle_fit
from ms_project_b200 import workloads, fitness
from ms_project_b200.covariance import Covariance
from ms_project_b200.gp_model import GPModel, score_models
from ms_project_b200.gp_models import gp_regression
from ms_project_b200.engine import get_engine
from oracle import model as omodel
from oracle.adapter import to_oracle_expr, prior_tuple
eng = get_engine(0)
X, y = make_data(120, 2, seed=7)
names = ['SE_1', 'RQ_1 + SE_2', 'SE_1 * SE_2', 'LIN_1 + SE_2', 'RQ_2 * LIN_1']
kernels = [_with_priors(from_string(s)) for s in names]
noise_prior = workloads.boms_hyperpriors()['GP']['variance']
# oracle with per-restart info
rng = np.random.RandomState(11)
for k, name in zip([k.copy() for k in kernels], names):
    kp = [prior_tuple(p.prior) for p in k.flattened_parameters()]
    m, failed = omodel.fit(to_oracle_expr(k), X, y, n_restarts=3, rng=rng, theta=k.param_array.copy(), kern_priors=kp, noise_prior=prior_tuple(noise_prior))
    print('ORACLE', name, 'lml', m.log_likelihood(), 'runs', getattr(m, 'optimization_runs', None) and [(r['f_opt'], r.get('funcalls')) for r in m.optimization_runs], 'theta', m.param_array)
md = gp_regression(likelihood_hyperprior={'variance': noise_prior})
models = [GPModel(Covariance(k)) for k in kernels]
np.random.seed(11)
scores = score_models(models, X, y, fitness.negative_bic, n_restarts=3, engine=eng, model_dict=md)
for m, name in zip(models, names):
    print('OURS  ', name, 'lml', m.fit_result.lml, 'runs', [(r['f_opt'], r['funcalls'], r['warnflag']) for r in m.fit_result.runs], 'theta', m.fit_result.theta)
