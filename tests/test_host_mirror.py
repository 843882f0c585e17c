"""Host mirror of the reference's plugin surface: CPU checks (construction, state_dict layout,
seeded initial weights vs the reference) and GPU checks (module forward/backward through autograd,
cached_sound rule, reward wrapper vs the reference's golden run, trainer end to end)."""
import os
import pickle
import types
from importlib import import_module

import numpy as np
import pytest
import torch

from oracle import model as omodel
from oracle import synth

PKG = "voicecontrolledrobot-var_b200"
DEV = "cuda:0"


class Cfg:
    pass


def kuka_cfg():
    c = Cfg()
    c.name = "ArmConfig"
    c.img_dim = (3, 96, 96); c.sound_dim = (1, 100, 40); c.representationDim = 3
    c.taskNum = 4; c.envFolder = os.path.join("pybullet", "arms"); c.tripletMargin = 1.0
    c.soundSource = {"dataset": ["GoogleCommand"]}
    c.RLRewardSoundSound = False; c.realTimeVec = False; c.RLTrain = True
    return c


def ithor_cfg():
    c = kuka_cfg()
    c.name = "AI2ThorConfig"; c.sound_dim = (1, 600, 40); c.envFolder = "ai2thor"
    return c


def _net(kind, cfg):
    mod = import_module(f"{PKG}.models.pretext." + ("arm_pretext_model" if kind == "kuka" else "ai2thor_pretext_model"))
    return mod.VARPretextNet(cfg)


def test_seeded_construction_matches_reference_initial_weights(vb, golden):
    g = golden("init")
    for kind, cfg, seed in (("kuka", kuka_cfg(), 453), ("ithor", ithor_cfg(), 977)):
        torch.manual_seed(seed)
        m = _net(kind, cfg)
        sd = m.state_dict()
        assert list(sd.keys()) == list(omodel.param_shapes(kind).keys())
        for k, v in sd.items():
            ref = g[f"{kind}.{k}"]
            assert tuple(v.shape) == tuple(omodel.param_shapes(kind)[k])
            assert abs(v.double().sum().item() - ref[0]) < 1e-9 and v.flatten()[0].item() == ref[2], k


def test_module_refuses_cpu_tensors(vb):
    m = _net("kuka", kuka_cfg())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 96, 96), None, None)
    d = m(None, None, None)  # all-None call is legal in the reference (returns the cache)
    assert d["image_feat"] is None and d["sound_feat_positive"] is None


def _reference_imgcnn():
    """The layer list of models/RL/ai2thor_RL_model.py:15-27 (plain torch, fp32, CPU)."""
    nn = torch.nn
    return nn.Sequential(
        nn.Conv2d(3, 32, 3, stride=1, padding=1), nn.ReLU(), nn.Conv2d(32, 32, 3, stride=1, padding=1), nn.ReLU(),
        nn.MaxPool2d(2, stride=2), nn.Conv2d(32, 64, 3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(2, stride=2),
        nn.Conv2d(64, 64, 3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(2, stride=2),
        nn.Conv2d(64, 128, 3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(2, stride=2),
        nn.Conv2d(128, 128, 3, stride=2, padding=1), nn.ReLU(), nn.Flatten())


def test_policy_imgcnn_state_dict_layout(vb):
    """the policy net's image CNN mirror keeps the reference's keys and shapes."""
    mod = import_module(f"{PKG}.models.RL.ai2thor_RL_model")
    ref = _reference_imgcnn()
    sd = {"imgCNN." + k: v for k, v in ref.state_dict().items()}
    sd["cnnMlp.0.weight"] = torch.zeros(512, 1152)  # other policy tensors are ignored
    m = mod.ai2thorImgCNN.from_policy_state_dict(sd)
    got = m.imgCNN.state_dict()
    assert list(got) == list(ref.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.equal(got[k], v)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 96, 96))


@pytest.mark.gpu
@pytest.mark.parametrize("N,dtype", [(3, torch.uint8), (40, torch.float32), (300, torch.uint8)])
def test_policy_imgcnn_forward_vs_torch(vb, N, dtype):
    """ai2thorNet_VAR.imgCNN (models/RL/ai2thor_RL_model.py:15-27) served by the VAR image-branch kernels: flattened
    [N, 1152] features against plain fp32 PyTorch (tf32 MMAs: 2e-3 of the feature scale, as for image_feat_raw)."""
    mod = import_module(f"{PKG}.models.RL.ai2thor_RL_model")
    torch.manual_seed(7)
    ref = _reference_imgcnn()
    m = mod.ai2thorImgCNN.from_policy_state_dict({"imgCNN." + k: v for k, v in ref.state_dict().items()})
    img_u8 = torch.from_numpy(synth.make_images(5, N))
    x = img_u8.float() / 255
    with torch.no_grad():
        want = ref(x)
    got = m((img_u8 if dtype == torch.uint8 else x).to(DEV))
    assert got.shape == (N, 1152) and got.dtype == torch.float32
    err = (got.cpu() - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-3, err
    # weights changed in place (a PPO update of the policy): the engine copy follows
    with torch.no_grad():
        for p in ref.parameters():
            p.mul_(0.5)
        m.imgCNN.load_state_dict(ref.state_dict())
        want2 = ref(x)
    got2 = m(img_u8.to(DEV))
    assert (got2.cpu() - want2).abs().max().item() / want2.abs().max().item() < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("kind,B", [("kuka", 6), ("ithor", 2)])
def test_module_forward_backward_through_autograd(vb, kind, B):
    cfg = kuka_cfg() if kind == "kuka" else ithor_cfg()
    m = _net(kind, cfg).to(DEV)
    sd = omodel.init_state_dict(kind, 31)
    m.load_state_dict(sd)
    m.train()
    images, sp, sn = synth.model_case(kind, B, 77)
    d = m(torch.from_numpy(images).to(DEV), torch.from_numpy(sp).to(DEV), torch.from_numpy(sn).to(DEV))
    assert set(d) == {"image_feat", "sound_feat_positive", "sound_feat_negative", "image_BCE", "sound_BCE",
                      "image_feat_raw", "pos_sound_raw"}
    crit = torch.nn.TripletMarginLoss(margin=1.0, p=2)            # the reference's own loss + autograd
    loss = crit(d["image_feat"], d["sound_feat_positive"], d["sound_feat_negative"])
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-6)
    opt.zero_grad()
    loss.backward()
    o = omodel.OracleVAR(kind, {k: v.clone().requires_grad_(True) for k, v in sd.items()})
    dr = o(torch.from_numpy(images), torch.from_numpy(sp), torch.from_numpy(sn))
    loss_ref = omodel.triplet_margin_loss(dr["image_feat"], dr["sound_feat_positive"], dr["sound_feat_negative"])
    loss_ref.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 1e-3 * abs(float(loss_ref.detach())) + 1e-6
    for k in ("image_feat", "sound_feat_positive", "sound_feat_negative"):
        assert np.abs(d[k].detach().cpu().numpy() - dr[k].detach().numpy()).max() < 1e-3
    for k in ("image_feat_raw", "pos_sound_raw"):
        a, b = d[k].detach().cpu().numpy(), dr[k].detach().numpy()
        assert a.shape == b.shape and np.abs(a - b).max() <= 2e-3 * np.abs(b).max()
    # Direction only: at B<=6 and random init the hinge gradient is ill-conditioned in the embedding
    # error and marginal ReLU / max-pool decisions move whole terms (see tests/test_gpu_parity.py,
    # where the exact backward arithmetic is pinned with a fixed upstream gradient).
    ga, gb = [], []
    for k, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, k
        a, b = p.grad.cpu().numpy().reshape(-1).astype(np.float64), o.sd[k].grad.numpy().reshape(-1).astype(np.float64)
        ga.append(a); gb.append(b)
        if "Triplet" in k or k.startswith("rnn."):
            assert a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300) > 0.97, k
    ga, gb = np.concatenate(ga), np.concatenate(gb)
    assert ga @ gb / (np.linalg.norm(ga) * np.linalg.norm(gb)) > 0.9
    opt.step()                                                    # parameters change -> engine repacks
    with torch.no_grad():
        d2 = m(torch.from_numpy(images).to(DEV), None, None)
    assert not torch.equal(d2["image_feat"], d["image_feat"].detach())
    # cached_sound rule (pretext_base.py:29-32): all-inf positive reuses the last positive embedding
    m.eval()
    with torch.no_grad():
        d3 = m(torch.from_numpy(images).to(DEV), torch.from_numpy(sp).to(DEV), None)
        inf = torch.full_like(torch.from_numpy(sp), float("inf")).to(DEV)
        d4 = m(torch.from_numpy(images).to(DEV), inf, None)
    assert torch.equal(d4["sound_feat_positive"], d3["sound_feat_positive"]) and d4["pos_sound_raw"] is None
    assert d3["sound_feat_negative"] is None
    # state_dict round trip through a legacy-format checkpoint (VAR/pretext_VAR.py:79)
    import io
    buf = io.BytesIO()
    torch.save(m.state_dict(), buf, _use_new_zipfile_serialization=False)
    buf.seek(0)
    m2 = _net(kind, cfg)
    m2.load_state_dict(torch.load(buf))
    with torch.no_grad():
        d5 = m2.to(DEV)(torch.from_numpy(images).to(DEV), None, None)
    assert torch.equal(d5["image_feat"], d3["image_feat"])


@pytest.mark.gpu
def test_reward_wrapper_matches_reference_golden(vb, golden):
    """Same stub env / seeds as oracle/make_golden.py::gold_reward, driven through the mirror."""
    g = golden("reward")
    vpn = import_module(f"{PKG}.Envs.vec_env.vec_pretext_normalize")
    cfg = kuka_cfg()
    N, steps = int(g["N"]), int(g["steps"])
    m = _net("kuka", cfg)
    m.load_state_dict(omodel.init_state_dict(omodel.KUKA, 11))
    m.to(DEV).eval()
    rng = np.random.default_rng(5)
    obs_seq, rew_seq, done_seq = [], [], []
    for t in range(steps + 1):
        obs_seq.append({"image": rng.integers(0, 256, (N, 3, 96, 96)).astype(np.uint8),
                        "goal_sound": (rng.standard_normal((N, 1, 100, 40)) * 5).astype(np.float32),
                        "robot_pose": rng.standard_normal((N, 4)).astype(np.float32)})
        rew_seq.append(rng.standard_normal(N))
        done_seq.append(rng.random(N) < 0.3)

    class Venv:
        num_envs = N
        observation_space = types.SimpleNamespace(shape=(1,))
        action_space = None
        t = 0

        def reset(self):
            return obs_seq[0]

        def step_wait(self):
            self.t += 1
            return obs_seq[self.t], rew_seq[self.t].copy(), done_seq[self.t].copy(), ({},) * N

    w = vpn.VecPretextNormalize(Venv(), ob=False, ret=True, gamma=0.99, config=cfg,
                                pretextObj=types.SimpleNamespace(pretextModel=m))
    o0 = w.reset()
    assert np.abs(o0["image_feat"] - g["reset_image_feat"]).max() < 1e-3
    assert np.abs(o0["goal_sound_feat"] - g["reset_goal_sound_feat"]).max() < 1e-3
    for t in range(steps):
        o, r, d, _ = w.step_wait()
        assert np.abs(o["image_feat"] - g[f"image_feat{t}"]).max() < 1e-3
        assert np.abs(o["goal_sound_feat"] - g[f"goal_sound_feat{t}"]).max() < 1e-3
        assert np.abs(w.origStepReward - g[f"orig{t}"]).max() < 3e-3
        assert np.abs(r - g[f"rew{t}"]).max() < 3e-3
        assert o["image"].dtype == np.float64 and np.allclose(o["image"], obs_seq[t + 1]["image"] / 255.)
        assert set(o) == {"robot_pose", "goal_sound_feat", "image", "image_feat"}
    assert abs(float(w.ret_rms.var) - float(g["ret_var"])) < 1e-2 * float(g["ret_var"])
    # iTHOR convention: goal sound sent as all-inf after the first step -> cached embedding is reused
    obs_inf = dict(obs_seq[1]); obs_inf["goal_sound"] = np.full((N, 1, 100, 40), np.inf, np.float32)
    f_img, f_goal, _ = w.getEmbeddings(obs_inf)
    assert np.abs(f_goal - g[f"goal_sound_feat{steps - 1}"]).max() < 1e-3


@pytest.mark.gpu
def test_reward_wrapper_device_path_matches_reference_golden(vb, golden):
    """step_wait_device(): reward query + float64 return normalisation on the device, same golden run."""
    g = golden("reward")
    vpn = import_module(f"{PKG}.Envs.vec_env.vec_pretext_normalize")
    cfg = kuka_cfg()
    N, steps = int(g["N"]), int(g["steps"])
    m = _net("kuka", cfg)
    m.load_state_dict(omodel.init_state_dict(omodel.KUKA, 11))
    m.to(DEV).eval()
    rng = np.random.default_rng(5)
    seq = []
    for t in range(steps + 1):
        seq.append(({"image": rng.integers(0, 256, (N, 3, 96, 96)).astype(np.uint8),
                     "goal_sound": (rng.standard_normal((N, 1, 100, 40)) * 5).astype(np.float32),
                     "robot_pose": rng.standard_normal((N, 4)).astype(np.float32)},
                    rng.standard_normal(N), rng.random(N) < 0.3))

    class Venv:
        num_envs = N
        observation_space = types.SimpleNamespace(shape=(1,))
        action_space = None
        t = 0

        def reset(self):
            return seq[0][0]

        def step_wait(self):
            self.t += 1
            return seq[self.t][0], seq[self.t][1].copy(), seq[self.t][2].copy(), ({},) * N

    w = vpn.VecPretextNormalize(Venv(), ob=False, ret=True, gamma=0.99, config=cfg,
                                pretextObj=types.SimpleNamespace(pretextModel=m))
    w.reset()
    for t in range(steps):
        o, r, d, _ = w.step_wait_device()
        assert r.is_cuda and r.shape == (N, 1) and all(v.is_cuda for v in o.values())
        assert np.abs(r[:, 0].cpu().numpy() - g[f"rew{t}"]).max() < 3e-3
        assert np.abs(o["image_feat"].cpu().numpy() - g[f"image_feat{t}"]).max() < 1e-3
        w.sync_device_stats()
        assert np.abs(w.origStepReward - g[f"orig{t}"]).max() < 3e-3
    assert abs(float(w.ret_rms.var) - float(g["ret_var"])) < 1e-2 * float(g["ret_var"])
    assert abs(float(w.ret_rms.mean) - float(g["ret_mean"])) < 1e-2 * max(1.0, abs(float(g["ret_mean"])))


@pytest.mark.gpu
@pytest.mark.parametrize("device_path", [False, True])
def test_ithor_reward_wrapper_matches_reference_golden(vb, golden, device_path):
    """processAI2Thor + the iTHOR net + the all-inf cached-goal schedule (Envs/ai2thor/RL_env_VAR.py:509-510)
    through step_wait (host protocol) and step_wait_device, against the reference wrapper's own run
    (oracle/make_golden.py::gold_reward_ithor)."""
    from conftest import ithor_config
    g = golden("reward_ithor")
    vpn = import_module(f"{PKG}.Envs.vec_env.vec_pretext_normalize")
    cfg = ithor_config()
    N, steps = int(g["N"]), int(g["steps"])
    m = _net("ithor", cfg)
    m.load_state_dict(omodel.init_state_dict(omodel.ITHOR, 13))
    m.to(DEV).eval()
    obs_seq, rew_seq, done_seq = synth.ithor_reward_case(N, steps)
    assert np.isinf(obs_seq[1]["goal_sound"]).all() and not np.isinf(obs_seq[3]["goal_sound"]).any()

    class Venv:
        num_envs = N
        observation_space = types.SimpleNamespace(shape=(1,))
        action_space = None
        t = 0

        def reset(self):
            return obs_seq[0]

        def step_wait(self):
            self.t += 1
            return obs_seq[self.t], rew_seq[self.t].copy(), done_seq[self.t].copy(), ({},) * N

    w = vpn.VecPretextNormalize(Venv(), ob=False, ret=True, gamma=0.99, config=cfg,
                                pretextObj=types.SimpleNamespace(pretextModel=m))
    o0 = w.reset()
    assert set(o0) == {"occupancy", "goal_sound_feat", "image", "image_feat"}
    assert np.abs(o0["image_feat"] - g["reset_image_feat"]).max() < 1e-3
    assert np.abs(o0["goal_sound_feat"] - g["reset_goal_sound_feat"]).max() < 1e-3
    assert abs(float(o0["occupancy"].sum()) - float(g["reset_occupancy_sum"])) < 1e-6 * float(g["reset_occupancy_sum"])
    for t in range(steps):
        if device_path:
            o, r, d, _ = w.step_wait_device()
            w.sync_device_stats()
            o = {k: v.cpu().numpy() for k, v in o.items()}
            r = r[:, 0].cpu().numpy()
        else:
            o, r, d, _ = w.step_wait()
        assert np.abs(r - g[f"rew{t}"]).max() < 3e-3, t
        assert np.abs(w.origStepReward - g[f"orig{t}"]).max() < 3e-3, t
        assert np.abs(o["image_feat"] - g[f"image_feat{t}"]).max() < 1e-3, t
        assert np.abs(o["goal_sound_feat"] - g[f"goal_sound_feat{t}"]).max() < 1e-3, t   # cached on the inf steps
        assert o["occupancy"].max() <= 1.0 and o["image"].max() <= 1.0
    assert abs(float(w.ret_rms.var) - float(g["ret_var"])) < 1e-2 * float(g["ret_var"])


@pytest.mark.gpu
def test_rl_var_surface(vb, tmp_path):
    """VAR/RL_VAR.py:8-76: RL_VAR(config).run() loads the VAR weights, builds the (sharded) envs through
    the env factory hook, and testRL rolls a policy reading `eval_envs.venv.origStepReward` after every
    batched reward query; results land in test_<policy>.csv."""
    rlv = import_module(f"{PKG}.VAR.RL_VAR")
    vpn = import_module(f"{PKG}.Envs.vec_env.vec_pretext_normalize")
    assert rlv.shard_envs(16, 1, 4) == (4, 8) and rlv.shard_envs(10, 2, 3) == (6, 10)
    cfg = kuka_cfg()
    cfg.pretextModel = import_module(f"{PKG}.models.pretext.arm_pretext_model").VARPretextNet
    cfg.pretextModelLoadDir = str(tmp_path / "var.pt")
    torch.save(omodel.init_state_dict(omodel.KUKA, 3), cfg.pretextModelLoadDir, _use_new_zipfile_serialization=False)
    cfg.RLManualControl = False; cfg.RLManualControlLoaded = False; cfg.RLTrain = False
    cfg.RLEnvName = "arms-RL-v2"; cfg.RLEnvSeed = 1; cfg.RLGamma = 0.99; cfg.RLDeterministic = True
    cfg.render = False; cfg.success_threshold = 1
    cfg.skillInfos = [{"path": str(tmp_path / "policy.pt")}]
    rng = np.random.default_rng(2)

    class BaseEnv:
        size_per_class = [1, 1, 1, 1]
        size_per_class_cumsum = np.cumsum([1, 1, 1, 1])
        episodeCounter = 0

    base = BaseEnv()

    class Venv:  # one env, 3 steps per episode
        num_envs = 1
        observation_space = types.SimpleNamespace(shape=(1,))
        action_space = None
        unwrapped = types.SimpleNamespace(envs=[base])
        t = 0

        def _obs(self):
            return {"image": rng.integers(0, 256, (1, 3, 96, 96)).astype(np.uint8),
                    "goal_sound": (rng.standard_normal((1, 1, 100, 40)) * 5).astype(np.float32),
                    "robot_pose": np.zeros((1, 4), np.float32)}

        def reset(self):
            return self._obs()

        def step_async(self, a):
            pass

        def step_wait(self):
            self.t += 1
            done = self.t % 3 == 0
            if done:
                base.episodeCounter += 1
            return self._obs(), np.array([0.25]), np.array([done]), ({"goal_area_count": self.t % 2},)

        def close(self):
            pass

    made = {}

    def make_vec_envs(**kw):
        made.update(kw)
        w = vpn.VecPretextNormalize(Venv(), ob=False, ret=True, gamma=kw["gamma"], config=kw["config"],
                                    pretextObj=kw["pretextObj"])
        return types.SimpleNamespace(venv=w, reset=w.reset, step=w.step, close=w.close, render=lambda: None)

    class Policy:
        recurrent_hidden_state_size = 4

        def act(self, obs, h, masks, deterministic=False):
            assert set(obs) == {"robot_pose", "goal_sound_feat", "image", "image_feat"}
            return None, torch.zeros(1, 1), None, h

    rl = rlv.RL_VAR(cfg, make_vec_envs=make_vec_envs, load_policy=lambda envs: [Policy()])
    rl.run()
    assert made["num_processes"] == 1 and made["pretextObj"] is rl.pretextObj and rl.pretextObj.pretextModel is not None
    import pandas as pd
    df = pd.read_csv(tmp_path / "test_policy.csv")
    assert list(df.columns) == ["objIdx", "goal area count", "rewards", "results"] and len(df) == 4
    assert df["objIdx"].tolist() == [0, 1, 2, 3] and np.isfinite(df["rewards"]).all()
    # per-episode reward = sum over 3 steps of (img . goal dot in [-1, 1]) + 0.25 env reward
    assert (df["rewards"].abs() <= 3 * 1.25 + 1e-6).all()


def _write_dataset(root, cfg, n_items=48, clips_per_class=5, media_only_records=False):
    from scipy.io import wavfile
    words = ["up", "down", "left", "right"]
    media = os.path.join(root, "commonMedia")
    for c, wd in enumerate(words):
        if media_only_records:
            break
        d = os.path.join(media, "GoogleCommand", "train", wd)
        os.makedirs(d)
        for i, clip in enumerate(synth.make_clips(100 + c, clips_per_class, 16000)):
            wavfile.write(os.path.join(d, f"{i}.wav"), 16000, clip)
    data = os.path.join(root, "data", "train")
    os.makedirs(data)
    gts = synth.make_labels(3, n_items)
    imgs = synth.make_images(4, n_items)
    for f in range(2):
        items = [{"image": imgs[i], "ground_truth": int(gts[i])} for i in range(f, n_items, 2)]
        with open(os.path.join(data, f"data_{f}.pickle"), "wb") as fh:
            pickle.dump(items, fh)
    if not media_only_records:
        cfg.commonMediaPath = media
        cfg.soundSource = {"dataset": ["GoogleCommand"], "train_test": "train", "items": {"GoogleCommand": words},
                           "size": {"GoogleCommand": [1000] * 4}, "max_sound_dur": {"GoogleCommand": 3.0}}
    cfg.pretextDataDir = [os.path.join(root, "data")]
    cfg.pretextDataFileLoadNum = ["all"]
    cfg.pretextModelSaveDir = os.path.join(root, "model")
    cfg.pretextModelLoadDir = os.path.join(root, "model", "2.pt")
    cfg.pretextModelSaveInterval = 2
    cfg.pretextTrainBatchSize = 16
    cfg.pretextDataNumWorkers = 4
    cfg.pretextLR = 1e-3; cfg.pretextAdamL2 = 1e-6; cfg.pretextLRStep = "step"
    cfg.pretextLRDecayEpoch = [2]; cfg.pretextLRDecayGamma = 0.2; cfg.pretextEpoch = 3
    cfg.pretextTrain = True; cfg.pretextCollection = False; cfg.pretextModelFineTune = False
    cfg.pretextEnvSeed = 453; cfg.plotRepresentation = -1
    return gts


ITHOR_CLIPS = {("none", "lights", "activate"): 4, ("none", "lights", "deactivate"): 3, ("none", "music", "activate"): 5,
               ("none", "music", "deactivate"): 2, ("none", "lamp", "activate"): 3, ("none", "lamp", "deactivate"): 4}


def _write_dataset_ithor(root, cfg, n_items=48):
    """Fluent-Speech-Commands layout read by loadFSCData_ai2thor (Envs/audioLoader.py:61-98): a csv with
    path / transcription / action / object / location columns + the wavs it names."""
    import pandas as pd
    from scipy.io import wavfile
    media = os.path.join(root, "commonMedia")
    os.makedirs(os.path.join(media, "FSC", "data"))
    os.makedirs(os.path.join(media, "FSC", "wavs"))
    rows = []
    for li, ((loc, obj, act), n) in enumerate(ITHOR_CLIPS.items()):
        for i, clip in enumerate(synth.make_clips(200 + li, n, (12000, 30000))):
            rel = os.path.join("wavs", f"{obj}_{act}_{i}.wav")
            wavfile.write(os.path.join(media, "FSC", rel), 16000, clip)
            rows.append({"path": rel, "transcription": f"{act} the {obj}", "action": act, "object": obj, "location": loc})
    rows.append({"path": "wavs/none.wav", "transcription": "x", "action": "bring", "object": "shoes", "location": "none"})
    pd.DataFrame(rows).to_csv(os.path.join(media, "FSC", "data", "train_data.csv"))
    gts = _write_dataset(root, cfg, n_items, media_only_records=True)
    cfg.commonMediaPath = media
    cfg.pretextEnvSeed = 977
    return gts


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["kuka", "ithor"])
def test_trainer_end_to_end(vb, tmp_path, kind):
    """VAR_Pretext(config).run(): wavs + pickled records on disk -> device sampler -> MFCC -> fused
    steps -> legacy checkpoints + progress.csv, and the loader's first epoch == the oracle's stream.
    iTHOR: config.name == 'AI2ThorConfig' (task list, synonym draws, python_speech_features-flavoured
    MFCC: dataset.py:17-53, audioLoader.py:203-237)."""
    from conftest import ITHOR_ALL_TASKS, ITHOR_OBJ_ACT, ITHOR_SYNONYM, ithor_config
    from oracle import mfcc as omfcc
    from oracle import sampler as osampler
    ds = import_module(f"{PKG}.dataset")
    if kind == "kuka":
        cfg = kuka_cfg()
        cfg.pretextModel = import_module(f"{PKG}.models.pretext.arm_pretext_model").VARPretextNet
        gts = _write_dataset(str(tmp_path), cfg)
        F = 100
    else:
        cfg = ithor_config()
        cfg.pretextModel = import_module(f"{PKG}.models.pretext.ai2thor_pretext_model").VARPretextNet
        gts = _write_dataset_ithor(str(tmp_path), cfg)
        F = 600
    cfg.pretextDataset = ds.VARDataset
    trainer = import_module(f"{PKG}.VAR.pretext_VAR").VAR_Pretext(cfg)
    # loader contract: reference tuple shapes / dtypes, index stream == oracle continuing torch's generator
    torch.manual_seed(cfg.pretextEnvSeed)
    torch.rand(11)  # the global generator is not fresh when the loader is built
    ogen = osampler.TorchCPUGenerator.from_torch_state(torch.get_rng_state().numpy())
    gen_loader, final = ds.loadEnvData(cfg.pretextDataDir, cfg, 16, True, 4, False, cfg.pretextDataFileLoadNum,
                                       dtype=ds.VARDataset)
    assert len(final) == 48 and len(gen_loader) == 3
    record_gt = [int(p["ground_truth"]) for d in final.datasets for p in d.ground_truth_pair]
    batches = osampler.epoch_batches(ogen, 48, 16)
    if kind == "ithor":
        words = gen_loader.audio.words
        assert {(l, o, a): len(words[l][o][a]) for l in words for o in words[l] for a in words[l][o]} == ITHOR_CLIPS
        n_loc, n_obj, lists = osampler.ithor_task_tables(ITHOR_ALL_TASKS, ITHOR_SYNONYM, ITHOR_OBJ_ACT, words)
        tsizes = [[[ITHOR_CLIPS[k] for k in col] for col in row] for row in lists]
    checked = 0
    for (image, sp, sn, gt), batch in zip(gen_loader, batches):
        assert image.shape == (16, 3, 96, 96) and image.dtype == torch.float32 and float(image.max()) <= 1.0
        assert sp.shape == (16, 1, F, 40) and sn.shape == (16, 1, F, 40) and gt.dtype == torch.int64
        assert gt.cpu().tolist() == [record_gt[i] for i in batch]
        for j, i in enumerate(batch):
            if kind == "kuka":
                sn_id, pos, neg = osampler.sample_triplet_kuka(ogen, record_gt[i], 4, {k: [5] for k in range(4)})
            else:
                sn_id, pos, neg = osampler.sample_triplet_ithor(ogen, record_gt[i], 4, n_loc, n_obj, tsizes)
            assert (float(sp[j].abs().max()) == 0.0) == (pos is None)
            assert (float(sn[j].abs().max()) == 0.0) == (neg is None)
            if kind == "ithor" and pos is not None and checked < 4:
                # the features are those of exactly the drawn clip, in the python_speech_features flavour
                t, li, oi, clip = pos
                l, o, a = lists[t][li][oi]
                ref = omfcc.process_sound_feat(omfcc.mfcc_psf(words[l][o][a][clip]), (1, 600, 40))[0]
                got = sp[j, 0].cpu().numpy()
                scale = np.maximum(np.abs(ref).max(axis=-1, keepdims=True), 1.0)
                assert (np.abs(got - ref) <= 1e-4 * np.abs(ref) + 1e-4 * scale).all()
                checked += 1
    losses = trainer.run() or None
    files = sorted(os.listdir(cfg.pretextModelSaveDir))
    assert "progress.csv" in files and "1.pt" in files and "2.pt" in files
    import pandas as pd
    prog = pd.read_csv(os.path.join(cfg.pretextModelSaveDir, "progress.csv"))
    assert len(prog) == 3 and np.isfinite(prog["avg_loss"]).all() and prog["avg_loss"].iloc[-1] < prog["avg_loss"].iloc[0]
    sd = torch.load(os.path.join(cfg.pretextModelSaveDir, "2.pt"))
    assert list(sd.keys()) == list(omodel.param_shapes(kind).keys())
    cfg.plotNumBatch = 1
    fp = trainer.project2representation_with_ground_truth(gen_loader)   # pretext.py:147-203
    assert fp['img'].shape == (32, 4) and fp['sound'].shape == (32, 4)
    assert np.allclose(np.linalg.norm(fp['img'][:, :3], axis=1), 1.0, atol=1e-5)
    trainer.pretextModel = None
    trainer.loadPretextModel()                                   # pretext.py:102-111
    with torch.no_grad():
        out = trainer.pretextModel(torch.zeros(2, 3, 96, 96, device=DEV), None, None)
    assert out["image_feat"].shape == (2, 3)
