"""ORACLE (test infrastructure only -- never imported by the product path).

Host arithmetic of the batched VAR reward query:
  Envs/vec_env/vec_pretext_normalize.py:96-101  calcReward
  Envs/vec_env/vec_pretext_normalize.py:47-61   step_wait reward normalisation
  Envs/vec_env/running_mean_std.py:16-35        RunningMeanStd (float64)
PINNED by oracle/make_golden.py against the imported reference classes
(tests/golden/reward.npz).
"""
import numpy as np


def calc_reward(env_reward, image_feat, goal_sound_feat, current_sound_feat=0.0, rep_dim=3,
                sound_sound=False):
    img_sound_dot = np.sum(image_feat[:, :rep_dim] * goal_sound_feat, axis=1)
    sound_sound_dot = np.sum(current_sound_feat * goal_sound_feat, axis=1)
    emb = img_sound_dot + sound_sound_dot * sound_sound
    return emb + env_reward, img_sound_dot, sound_sound_dot


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr):
        bm, bv, bc = np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0]
        delta = bm - self.mean
        tot = self.count + bc
        new_mean = self.mean + delta * bc / tot
        m2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class ReturnNormalizer:
    """State carried by VecPretextNormalize.step_wait (lines 53-59)."""

    def __init__(self, num_envs, gamma=0.99, epsilon=1e-8, cliprew=10.0):
        self.ret = np.zeros(num_envs)
        self.rms = RunningMeanStd(shape=())
        self.gamma, self.epsilon, self.cliprew = gamma, epsilon, cliprew

    def step(self, rews, news):
        orig = rews.copy()
        self.ret = self.ret * self.gamma + rews
        self.rms.update(self.ret)
        out = np.clip(rews / np.sqrt(self.rms.var + self.epsilon), -self.cliprew, self.cliprew)
        self.ret[news] = 0.0
        return out, orig
