"""VAR/RL_VAR.py / RL.py call site of the reward path (RL.py:164, VAR/RL_VAR.py:49-50).

PPO, the policy networks and the vectorised simulators are out of scope;
what the RL driver needs from this package is the wrapper factory below, which mirrors the
`VecPretextNormalize` construction of Envs/vec_env/envs.py:45-49."""
from ..Envs.vec_env.vec_pretext_normalize import VecPretextNormalize


def wrap_vec_env(venv, gamma, config, pretextObj):
    """envs.py:45-49: `VecPretextNormalize(envs, ob=False, ret=..., config, pretextObj)`."""
    if gamma is None:
        return VecPretextNormalize(venv, ob=False, ret=False, config=config, pretextObj=pretextObj)
    return VecPretextNormalize(venv, ob=False, gamma=gamma, config=config, pretextObj=pretextObj)
