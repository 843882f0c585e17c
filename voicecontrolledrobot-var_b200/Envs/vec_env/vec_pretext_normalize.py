"""Host mirror of Envs/vec_env/vec_pretext_normalize.py::VecPretextNormalize.

Same constructor, attributes (`origStepReward`, `ret_rms`, ...) and methods as the reference
wrapper; `getEmbeddings` + `calcReward` (lines 82-101) are served by ONE batched device query
(`var_net_reward`: image branch, goal-sound branch or its cached embedding, normalisation, dot
product and env-reward add) instead of a model call, two D2H copies and a numpy dot.  The
uint8 observation is uploaded as uint8 (4x fewer H2D bytes than the reference's float64->float32
path); the 1/255 scale is applied inside the first conv's loader.

Multi-GPU: reward queries shard by env index with NO collective -- one process per GPU, rank r of G
owns envs [r*N/G, (r+1)*N/G) (VAR/RL_VAR.py::shard_envs), builds its venv with that many workers and
wraps it in its own VecPretextNormalize: weights are replicated, `cached_sound` holds the rank's own
envs, and a query is row-wise independent, so the gathered result equals the single-GPU query bit
for bit (tests/test_gpu_parity.py::test_reward_query_sharded_by_env_is_bit_identical).  The one piece
of state the reference shares across envs is the discounted-return RunningMeanStd (`ret_rms`); each
rank keeps the statistics of its own envs unless `merge_ret_rms()` is called (an optional 3-double
all_gather, off the query path)."""
import numpy as np
import torch

from ..._lib import check, lib, ptr, stream_ptr
from .running_mean_std import RunningMeanStd, merge_moments


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
        """-> image_feat, goal_sound_feat (numpy [N, D]), img_sound_dot, reward (numpy [N]).
        The query for this wrapper's fixed N is a captured CUDA graph (engine.RewardGraph): host
        observations are copied straight into its static input buffers, one graph launch runs the image
        branch (+ the goal-sound branch on steps that carry a real goal sound) and the fused
        normalise / dot / reward tail, and one D2H copy brings back [img_feat | goal_feat | dot | reward]."""
        model = self.pretextModel
        eng = model._get_engine(self.device)
        img = np.ascontiguousarray(O['image'][:, :3])
        if img.dtype != np.uint8:
            img = img.astype(np.float32) / np.float32(255.)  # non-uint8 observations keep the reference scaling
        goal = O['goal_sound']
        N, F, D = img.shape[0], self.config.sound_dim[1], self.config.representationDim
        # pretext_base.py:29-32: an all-inf goal sound means "reuse the cached embedding"
        fresh = goal is not None and not bool(np.isinf(goal.flat[0]) and np.isinf(goal).all())
        g = eng.reward_graph(N, torch.uint8 if img.dtype == np.uint8 else torch.float32, fresh)
        g.images.copy_(torch.from_numpy(img), non_blocking=True)
        g.env_reward.copy_(torch.from_numpy(np.asarray(envReward, dtype=np.float32)), non_blocking=True)
        if fresh:
            g.goal_sounds.copy_(torch.from_numpy(np.ascontiguousarray(goal, dtype=np.float32).reshape(-1, F, 40)),
                                non_blocking=True)
        else:
            cached = model.cached_sound if model.cached_sound is not None else self._cached_goal_feat
            if cached is None:
                raise RuntimeError("all-inf goal sound before any goal sound was encoded (no cached embedding)")
            g.goal_feat_cached.copy_(cached.to(self.device).float(), non_blocking=True)
        g.launch()
        if fresh:
            self._cached_goal_feat = g.goal_feat.clone()  # the graph's output buffer is overwritten by the next query
            model.cached_sound = self._cached_goal_feat
        out = g.out.cpu().numpy()  # one D2H (synchronises)
        return (out[:N * D].reshape(N, D), out[N * D:2 * N * D].reshape(N, D), out[2 * N * D:2 * N * D + N],
                out[2 * N * D + N:])

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
            image_feat, goal_sound_feat, dot, _ = self._query(O, envReward)
            # float64 host add as in the reference (vec_pretext_normalize.py:99); the device-side fp32 sum is what
            # step_wait_device() uses
            reward = dot.astype(np.float64) + envReward
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

    def merge_ret_rms(self):
        """Optional: make `ret_rms` the statistic over ALL ranks' envs (what a single-process run of the
        reference would hold) by merging the per-rank (mean, var, count) with the parallel-variance rule of
        running_mean_std.py:22-35, in rank order on every rank.  Not called by step_wait."""
        import torch.distributed as dist
        if not (self.ret_rms and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        mine = torch.tensor([float(self.ret_rms.mean), float(self.ret_rms.var), float(self.ret_rms.count)],
                            dtype=torch.float64, device=self.device if dist.get_backend() == "nccl" else "cpu")
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, mine)
        mean, var, count = [float(v) for v in parts[0].cpu()]
        for p in parts[1:]:
            mean, var, count = merge_moments(mean, var, count, *[float(v) for v in p.cpu()])
        self.ret_rms.mean, self.ret_rms.var, self.ret_rms.count = np.float64(mean), np.float64(var), count

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
