"""Host mirror of Envs/vec_env/vec_pretext_normalize.py::VecPretextNormalize.

Same constructor, attributes (`origStepReward`, `ret_rms`, ...) and methods as the reference
wrapper; `getEmbeddings` + `calcReward` (lines 82-101) are served by ONE batched device query
(`var_net_reward`: image branch, goal-sound branch or its cached embedding, normalisation, dot
product and env-reward add) instead of a model call, two D2H copies and a numpy dot.  The
uint8 observation is uploaded as uint8 (4x fewer H2D bytes than the reference's float64->float32
path); the 1/255 scale is applied inside the first conv's loader."""
import numpy as np
import torch

from ..._lib import check, lib, ptr, stream_ptr
from .running_mean_std import RunningMeanStd


class VecEnvWrapper(object):
    """Minimal stand-in for Envs/vec_env/vec_env.py:142-177 (no gym dependency)."""

    def __init__(self, venv, observation_space=None, action_space=None):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space = observation_space or getattr(venv, "observation_space", None)
        self.action_space = action_space or getattr(venv, "action_space", None)

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        return self.venv.close()

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.venv, name)


class VecPretextNormalize(VecEnvWrapper):
    def __init__(self, venv, ob=True, ret=True, clipob=10., cliprew=10., gamma=0.99, epsilon=1e-8, config=None,
                 pretextObj=None):
        VecEnvWrapper.__init__(self, venv)
        self.config = config
        self.pretextObj = pretextObj
        self.pretextModel = self.pretextObj.pretextModel
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 VAR reward path needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.ob_rms = RunningMeanStd(shape=self.observation_space.shape) if ob else None
        self.ret_rms = RunningMeanStd(shape=()) if ret else None
        self.clipob, self.cliprew = clipob, cliprew
        self.ret = np.zeros(self.num_envs)
        self.gamma, self.epsilon = gamma, epsilon
        self.origStepReward = np.zeros(self.num_envs)
        self.rl_obs_space = None
        self.processing_func = {'ArmConfig': self.processArm, 'AI2ThorConfig': self.processAI2Thor}
        self._cached_goal_feat = None  # device copy of pretextModel.cached_sound
        self._last = None

    # ------------------------------------------------------------------ reward query
    def _query(self, O, envReward):
        """-> image_feat, goal_sound_feat (numpy [N, D]), img_sound_dot, reward (numpy [N])."""
        model = self.pretextModel
        eng = model._get_engine(self.device)
        img = np.ascontiguousarray(O['image'][:, :3])
        if img.dtype != np.uint8:
            img = img.astype(np.float32) / np.float32(255.)  # non-uint8 observations keep the reference scaling
        image = torch.from_numpy(img).to(self.device, non_blocking=True)
        goal = O['goal_sound']
        F = self.config.sound_dim[1]
        # pretext_base.py:29-32: an all-inf goal sound means "reuse the cached embedding"
        fresh = goal is not None and not bool(np.isinf(goal.flat[0]) and np.isinf(goal).all())
        env_r = torch.from_numpy(np.asarray(envReward, dtype=np.float32)).to(self.device, non_blocking=True)
        if fresh:
            snd = torch.from_numpy(np.ascontiguousarray(goal, dtype=np.float32).reshape(-1, F, 40)).to(
                self.device, non_blocking=True)
            img_feat, goal_feat, dot, rew = eng.reward(image, goal_sounds=snd, env_reward=env_r)
            self._cached_goal_feat = goal_feat
            model.cached_sound = goal_feat
        else:
            cached = model.cached_sound if model.cached_sound is not None else self._cached_goal_feat
            if cached is None:
                raise RuntimeError("all-inf goal sound before any goal sound was encoded (no cached embedding)")
            cached = cached.to(self.device).float().contiguous()
            img_feat, goal_feat, dot, rew = eng.reward(image, goal_feat_cached=cached, env_reward=env_r)
        out = torch.cat([img_feat, goal_feat, dot[:, None], rew[:, None]], dim=1).cpu().numpy()  # one D2H
        D = img_feat.shape[1]
        return out[:, :D], out[:, D:2 * D], out[:, 2 * D], out[:, 2 * D + 1]

    def getEmbeddings(self, O):
        image_feat, goal_sound_feat, _, _ = self._query(O, np.zeros(len(O['image']), np.float32))
        if self.config.RLRewardSoundSound:
            with torch.no_grad():
                cur = torch.from_numpy(O['current_sound']).float().to(self.device)
                current_sound_feat = self.pretextModel(None, None, cur)['sound_feat_negative'].cpu().numpy()
        else:
            current_sound_feat = 0.
        return image_feat, goal_sound_feat, current_sound_feat

    def calcReward(self, envReward, image_feat=None, goal_sound_feat=None, current_sound_feat=None):
        """Host form kept for callers that already hold embeddings (vec_pretext_normalize.py:96-101)."""
        img_sound_dot = np.sum(image_feat[:, :self.config.representationDim] * goal_sound_feat, axis=1)
        sound_sound_dot = np.sum(current_sound_feat * goal_sound_feat, axis=1)
        reward = img_sound_dot + sound_sound_dot * self.config.RLRewardSoundSound + envReward
        return reward, img_sound_dot, sound_sound_dot

    def _process(self, O, envReward, extra_key, extra_scale):
        if self.pretextModel is None:
            return O, envReward
        if self.config.RLRewardSoundSound:
            image_feat, goal_sound_feat, current_sound_feat = self.getEmbeddings(O)
            reward, _, _ = self.calcReward(envReward, image_feat, goal_sound_feat, current_sound_feat)
        else:
            image_feat, goal_sound_feat, _, dev_reward = self._query(O, envReward)
            # the device adds envReward in fp32; keep the reference's float64 host add for the sum
            reward = (dev_reward - np.asarray(envReward, dtype=np.float32)).astype(np.float64) + envReward
        s = {extra_key: O[extra_key] / extra_scale if extra_scale else O[extra_key],
             'goal_sound_feat': goal_sound_feat, 'image': O['image'] / 255., 'image_feat': image_feat}
        return self._obfilt(s), reward

    def processArm(self, O, envReward, done, infos):
        return self._process(O, envReward, 'robot_pose', None)

    def processAI2Thor(self, O, envReward, done, infos):
        return self._process(O, envReward, 'occupancy', 255.)

    # ------------------------------------------------------------------ wrapper protocol
    def step_wait(self):
        obs, env_rews, news, infos = self.venv.step_wait()
        obs, rews = self.processing_func[self.config.name](obs, env_rews, news, infos)
        self.origStepReward = rews.copy()
        self.ret = self.ret * self.gamma + rews
        if self.ret_rms:
            self.ret_rms.update(self.ret)
            rews = np.clip(rews / np.sqrt(self.ret_rms.var + self.epsilon), -self.cliprew, self.cliprew)
        self.ret[news] = 0.
        return obs, rews, news, infos

    def _obfilt(self, obs):
        if self.ob_rms and self.config.RLTrain:
            self.ob_rms.update(obs)
            return np.clip((obs - self.ob_rms.mean) / np.sqrt(self.ob_rms.var + self.epsilon), -self.clipob,
                           self.clipob)
        return obs

    # ------------------------------------------------------------------ device-resident variant
    def step_wait_device(self):
        """`step_wait` without the host round trip of the reference (GPU -> numpy dot -> numpy RMS ->
        torch -> GPU, vec_pretext_normalize.py:47-61 + envs.py:90-98): the reward query, the
        discounted-return RunningMeanStd update and the clipping all stay on the device
        (`var_net_reward` + `var_reward_normalize`).  Returns what `VecPyTorch.step_wait` would hand
        to the policy: a dict of float32 device tensors, rewards [N, 1] on the device, `news`, `infos`.
        `origStepReward` is refreshed lazily from the device copy."""
        O, env_rews, news, infos = self.venv.step_wait()
        dev, N = self.device, self.num_envs
        model = self.pretextModel
        eng = model._get_engine(dev)
        img_u8 = torch.from_numpy(np.ascontiguousarray(O['image'][:, :3], dtype=np.uint8)).to(dev, non_blocking=True)
        goal = O['goal_sound']
        F = self.config.sound_dim[1]
        fresh = goal is not None and not bool(np.isinf(goal.flat[0]) and np.isinf(goal).all())
        env_r = torch.from_numpy(np.asarray(env_rews, dtype=np.float32)).to(dev, non_blocking=True)
        if fresh:
            snd = torch.from_numpy(np.ascontiguousarray(goal, dtype=np.float32).reshape(-1, F, 40)).to(dev, non_blocking=True)
            img_feat, goal_feat, _, rew = eng.reward(img_u8, goal_sounds=snd, env_reward=env_r)
            model.cached_sound = goal_feat
        else:
            img_feat, goal_feat, _, rew = eng.reward(img_u8, goal_feat_cached=model.cached_sound.float().contiguous(),
                                                     env_reward=env_r)
        if not hasattr(self, "_ret_dev"):
            self._ret_dev = torch.from_numpy(self.ret.astype(np.float64)).to(dev)
            rms = self.ret_rms
            init = [float(rms.mean), float(rms.var), float(rms.count)] if rms else [0.0, 1.0, 1e-4]
            self._rms_dev = torch.tensor(init, dtype=torch.float64, device=dev)
        done = torch.from_numpy(np.asarray(news, dtype=np.uint8)).to(dev, non_blocking=True)
        orig = torch.empty(N, dtype=torch.float32, device=dev)
        out = torch.empty(N, dtype=torch.float32, device=dev)
        check(lib.var_reward_normalize(ptr(rew), ptr(done), N, ptr(self._ret_dev), ptr(self._rms_dev), float(self.gamma),
                                       float(self.epsilon), float(self.cliprew), 1 if self.ret_rms else 0, ptr(orig),
                                       ptr(out), stream_ptr()), "var_reward_normalize")
        self._orig_dev = orig
        extra = 'robot_pose' if self.config.name == 'ArmConfig' else 'occupancy'
        ex = torch.from_numpy(np.asarray(O[extra], dtype=np.float32)).to(dev, non_blocking=True)
        obs = {extra: ex if extra == 'robot_pose' else ex / 255., 'goal_sound_feat': goal_feat,
               'image': img_u8.float() / 255., 'image_feat': img_feat}
        return obs, out.unsqueeze(1), news, infos

    def sync_device_stats(self):
        """Pull the device-side return / RMS state back into the numpy attributes of the reference."""
        if hasattr(self, "_ret_dev"):
            self.ret = self._ret_dev.cpu().numpy()
            m, v, c = self._rms_dev.cpu().tolist()
            if self.ret_rms:
                self.ret_rms.mean, self.ret_rms.var, self.ret_rms.count = np.float64(m), np.float64(v), c
            self.origStepReward = self._orig_dev.cpu().numpy().astype(np.float64)

    def reset(self):
        if hasattr(self, "_ret_dev"):
            self._ret_dev.zero_()
        self.ret = np.zeros(self.num_envs)
        obs = self.venv.reset()
        obs, _ = self.processing_func[self.config.name](obs, np.zeros((self.num_envs,)),
                                                        np.array([True] * self.num_envs), ({},) * self.num_envs)
        return obs
