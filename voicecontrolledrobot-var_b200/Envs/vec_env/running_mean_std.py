"""Running mean / variance of a stream of batches, float64 (interface of
Envs/vec_env/running_mean_std.py::RunningMeanStd; used by the reward wrapper for the discounted
return normalisation, vec_pretext_normalize.py:31-32,56-58).  Batches are merged with the
pairwise (Chan et al.) formula, the same arithmetic `var_reward_normalize` performs on the device."""
import numpy as np


def merge_moments(mean_a, var_a, n_a, mean_b, var_b, n_b):
    """Moments of the union of two samples given each sample's (mean, biased variance, count)."""
    n = n_a + n_b
    delta = mean_b - mean_a
    mean = mean_a + delta * n_b / n
    m2 = var_a * n_a + var_b * n_b + np.square(delta) * n_a * n_b / n
    return mean, m2 / n, n


class RunningMeanStd(object):
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr):
        self.update_from_moments(np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        self.mean, self.var, self.count = merge_moments(self.mean, self.var, self.count, batch_mean, batch_var,
                                                        batch_count)
