import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    return load


@pytest.fixture(scope="session")
def vb():
    """The product package (fails loudly if libvar_b200.so is not built)."""
    import var_b200
    return var_b200


def rel_to_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# the reference's iTHOR task / synonym tables (Envs/ai2thor/env_config.py:18-45, Envs/ai2thor/config.py:121-133),
# restated as data: /root/reference does not exist on the GPU box
ITHOR_ALL_TASKS = {"livingRoom": {"FloorLamp": ["ToggleObjectOn", "ToggleObjectOff"],
                                  "Television": ["ToggleObjectOn", "ToggleObjectOff"]}}
ITHOR_SYNONYM = {"livingRoom": ["none"], "FloorLamp": ["lights", "lamp"], "Television": ["music"],
                 "ToggleObjectOn": ["increase", "activate"], "ToggleObjectOff": ["decrease", "deactivate"]}
ITHOR_OBJ_ACT = {"lights": ["activate", "deactivate"], "music": ["activate", "deactivate"],
                 "lamp": ["activate", "deactivate"]}


class _Cfg:
    pass


def ithor_config():
    """AI2ThorConfig + EnvConfig attributes the triplet path reads (Envs/ai2thor/config.py, env_config.py)."""
    c = _Cfg()
    c.name = "AI2ThorConfig"
    c.img_dim = (3, 96, 96); c.sound_dim = (1, 600, 40); c.representationDim = 3
    c.envFolder = "ai2thor"; c.tripletMargin = 1.0
    c.allTasks = ITHOR_ALL_TASKS; c.synonym = ITHOR_SYNONYM
    c.taskNum = sum(len(a) for o in c.allTasks.values() for a in o.values())
    c.soundSource = {"dataset": "FSC", "train_test": "train", "FSC_max_sound_dur": 6., "size": 1000,
                     "FSC_obj_act": ITHOR_OBJ_ACT, "FSC_locations": ["none"], "FSC_csv": "train_data.csv"}
    c.RLRewardSoundSound = False; c.realTimeVec = False; c.RLTrain = True
    c.pretextEnvSeed = 977
    return c
