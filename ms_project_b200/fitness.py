"""Scores of a fitted model -- identical floats to src/autoks/core/fitness_functions.py:7-27 and
src/autoks/model_selection_criteria.py:9-46 (they only read n_data / n_params / log_likelihood through
src/autoks/backend/model.py:95-108, which the GP stand-in provides)."""
import numpy as np


def n_data(model):
    return model.num_data


def n_params(model):
    return model._size_transformed()


def log_likelihood(model):
    return model.log_likelihood()


def BIC(model):
    n, k = n_data(model), n_params(model)
    return np.log(n) * k - 2 * log_likelihood(model)


def AIC(model):
    return 2 * n_params(model) - 2 * log_likelihood(model)


def pl2(model):
    n, k = n_data(model), n_params(model)
    return -log_likelihood(model) / n + k / (2 * n)


def log_likelihood_normalized(model):
    return log_likelihood(model) / n_data(model)


def likelihood_normalized(model):
    return np.exp(log_likelihood(model)) / n_data(model)


def negative_bic(model):
    return -BIC(model)


# alias table of ModelSelector.__init__ (base.py:58-76 resolves names through _FITNESS_FUNCTION_ALIAS)
FITNESS_FUNCTIONS = {'loglikn': log_likelihood_normalized, 'lmlnorm': log_likelihood_normalized,
                     'nbic': negative_bic, 'negative_bic': negative_bic,
                     'log_likelihood_normalized': log_likelihood_normalized}
