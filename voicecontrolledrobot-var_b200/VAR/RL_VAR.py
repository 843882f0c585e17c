"""Host mirror of VAR/RL_VAR.py::RL_VAR (and of the part of RL.py::RLBase it stands on).

`RL_VAR().run()` is the reference's RL entry point (RL.py:286-291): it builds the `Pretext` object,
loads the trained VAR weights (RL.py:256-257), creates the vectorised envs wrapped in
`VecPretextNormalize` (Envs/vec_env/envs.py:45-49) and then trains or tests the policy; every
`envs.step(action)` of those loops (RL.py:164, VAR/RL_VAR.py:44) is one batched VAR reward query --
the hot path this package serves.  PPO, the policy networks and the simulators are callers of the path
and stay with the reference: when this module is imported inside the reference
tree, `RL_VAR` subclasses the reference's own `RLBase`, so `trainRL` / `loadPolicy` / `manualControl`
are the reference's, while the VAR model and the reward wrapper are the B200 ones; outside the tree a
local `RLBase` keeps the same constructor / `run()` / `testRL()` contract and asks the caller for the
two out-of-scope pieces (an env factory and a policy loader).

Multi-GPU: reward queries shard by env index, one process per GPU and no collective -- rank r of G
simulates and queries envs [r*N/G, (r+1)*N/G) (`shard_envs`, `make_vec_envs(num_processes=...)`)."""
import os

import numpy as np
import torch

from ..Envs.vec_env.vec_pretext_normalize import VecPretextNormalize
from ..pretext import Pretext

try:  # inside the reference tree (RL.py pulls in gym, the PPO package and the simulators' configs)
    from RL import RLBase as _ReferenceRLBase
except Exception:  # noqa: BLE001 -- any missing dependency of RL.py means "not inside the reference tree"
    _ReferenceRLBase = None


def shard_envs(num_envs, rank=None, world_size=None):
    """Env-index range [lo, hi) of this rank: reward queries shard by env with no collective."""
    if rank is None or world_size is None:
        import torch.distributed as dist
        on = dist.is_available() and dist.is_initialized()
        rank, world_size = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    return (num_envs * rank) // world_size, (num_envs * (rank + 1)) // world_size


def wrap_vec_env(venv, gamma, config, pretextObj):
    """envs.py:45-49: `VecPretextNormalize(envs, ob=False, ret=..., config, pretextObj)`."""
    if gamma is None:
        return VecPretextNormalize(venv, ob=False, ret=False, config=config, pretextObj=pretextObj)
    return VecPretextNormalize(venv, ob=False, gamma=gamma, config=config, pretextObj=pretextObj)


class _LocalRLBase(object):
    """RL.py:17-24, :251-284 without the PPO / simulator imports.  `make_vec_envs` and `loadPolicy`
    are the two hooks the reference implements with out-of-scope code; pass them to the constructor
    (or override them) when running outside the reference tree."""

    def __init__(self, config, make_vec_envs=None, load_policy=None):
        self.config = config
        self.device = torch.device(f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu")
        print("Using device:", self.device)
        self.pretextObj = Pretext(self.config)
        self._make_vec_envs, self._load_policy = make_vec_envs, load_policy

    def make_vec_envs(self, **kw):
        if self._make_vec_envs is None:
            raise NotImplementedError("the vectorised simulators are not part of this package: pass make_vec_envs= "
                                      "(Envs/vec_env/envs.py:25-53) or run inside the reference tree")
        return self._make_vec_envs(**kw)

    def loadPolicy(self, envs):
        if self._load_policy is None:
            raise NotImplementedError("the PPO policy is not part of this package: pass load_policy= (RL.py:42-72)")
        return self._load_policy(envs)

    def trainRL(self):
        raise NotImplementedError("PPO training stays with the reference (RL.py:74-245); its envs.step() calls are "
                                  "served by VecPretextNormalize.step_wait of this package")

    def testRL(self, eval_envs):
        raise NotImplementedError("Please Implement this method")

    def _envs(self, num_processes):
        cfg = self.config
        kw = dict(env_name=cfg.RLEnvName, seed=cfg.RLEnvSeed, num_processes=num_processes, gamma=cfg.RLGamma,
                  device=self.device, randomCollect=False, config=cfg, pretextObj=self.pretextObj)
        return self.make_vec_envs(**kw)

    def run(self):
        cfg = self.config
        if not (cfg.RLManualControl and not cfg.RLManualControlLoaded):
            self.pretextObj.loadPretextModel()
        if cfg.RLManualControl:
            raise NotImplementedError("manual control drives the simulator GUI (RL.py:27-40)")
        if cfg.RLTrain:
            self.trainRL()
        else:
            self.testRL(self._envs(1))


RLBase = _ReferenceRLBase if _ReferenceRLBase is not None else _LocalRLBase


class RL_VAR(RLBase):
    """VAR/RL_VAR.py:8-76.  `config` defaults to the reference's `cfg.main_config()` (line 10)."""

    def __init__(self, config=None, **hooks):
        if config is None:
            from cfg import main_config  # inside the reference tree, as VAR/RL_VAR.py:5
            config = main_config()
        if RLBase is _LocalRLBase:
            super().__init__(config, **hooks)
        else:
            super().__init__(config)
            # the reference base class built its own Pretext: the reward path must use the B200 one
            self.pretextObj = Pretext(self.config)

    def testRL(self, eval_envs):
        """Roll the trained policy until every test episode is done; per episode the un-normalised VAR
        reward `eval_envs.venv.origStepReward` (vec_pretext_normalize.py:54) is accumulated
        (VAR/RL_VAR.py:49-50) and a success flag is derived from `goal_area_count`; results go to
        `test_<policy>.csv` beside the policy file (VAR/RL_VAR.py:62-75)."""
        cfg = self.config
        baseEnv = eval_envs.venv.unwrapped.envs[0]
        skillList = self.loadPolicy(eval_envs)
        policy = skillList[0]
        eval_episode_rewards, results, goal_area_count_list = [], [], []
        eval_env_rewards = 0.
        obs = eval_envs.reset()
        hidden = torch.zeros(1, policy.recurrent_hidden_state_size, device=self.device)
        masks = torch.zeros(1, 1, device=self.device)
        episode_num = baseEnv.size_per_class_cumsum[-1]
        objs = np.repeat(np.arange(cfg.taskNum, dtype=np.int64), baseEnv.size_per_class)
        while baseEnv.episodeCounter < episode_num:
            with torch.no_grad():
                _, action, _, hidden = policy.act(obs, hidden, masks, deterministic=cfg.RLDeterministic)
            obs, _, done, infos = eval_envs.step(action)  # one batched VAR reward query
            if cfg.render:
                eval_envs.render()
                print('step reward', eval_envs.venv.origStepReward)
            eval_env_rewards = eval_env_rewards + eval_envs.venv.origStepReward
            masks = torch.tensor([[0.0] if d else [1.0] for d in done], dtype=torch.float32, device=self.device)
            if done:
                goal_area_count = infos[0]['goal_area_count']
                goal_area_count_list.append(goal_area_count)
                results.append(int(goal_area_count >= cfg.success_threshold))
                eval_episode_rewards.append(float(np.asarray(eval_env_rewards).reshape(-1)[0]))
                eval_env_rewards = 0.
        if not cfg.render:
            import pandas as pd
            path = cfg.skillInfos[0]['path']
            save_path = os.path.join(os.path.dirname(path), 'test_' + os.path.splitext(os.path.basename(path))[0] + '.csv')
            pd.DataFrame({'objIdx': objs, 'goal area count': goal_area_count_list, 'rewards': eval_episode_rewards,
                          'results': results}).to_csv(save_path, mode='w', header=True, index=False)
            print('results saved to', save_path)
            print('success rate', sum(results) * 1. / episode_num)
        eval_envs.close()
