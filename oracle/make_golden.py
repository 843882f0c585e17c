"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, read-only) on seeded synthetic inputs.  Run in the build container:

    python oracle/make_golden.py

The GPU box has no /root/reference; tests read only the committed .npz files.
Import-time stubs are supplied for packages the reference imports but this image
lacks (gym, matplotlib, python_speech_features); none of them is on the hot path
exercised here.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import model as omodel  # noqa: E402
from oracle import synth  # noqa: E402


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Space:
        def __init__(self, *a, **k):
            self.shape = k.get("shape", ())

    spaces = mod("gym.spaces", Box=_Space, Dict=dict, Discrete=_Space)
    mod("gym", spaces=spaces, Env=object, Wrapper=object)
    mod("gym.envs"); mod("gym.envs.registration", register=lambda **k: None)
    mod("matplotlib", use=lambda *a, **k: None); mod("matplotlib.pyplot")
    mod("python_speech_features", mfcc=None)
    mod("ai2thor"); mod("ai2thor.controller", Controller=object); mod("ai2thor.platform", CloudRendering=object)
    sys.path.insert(0, REF)


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


def gold_mfcc():
    from Envs.audioLoader import audioLoader
    out = {}
    for tag, cfg, n_samples, ds in (("kuka_1s", kuka_cfg(), 16000, "GoogleCommand"),
                                    ("kuka_fsc", kuka_cfg(), (16000, 64000), "FSC"),
                                    ("kuka_4s_nsynth", kuka_cfg(), 64000, "NSynth"),
                                    ("ithor_1s", ithor_cfg(), 16000, "FSC"),
                                    ("ithor_short", ithor_cfg(), 4000, "FSC")):
        al = audioLoader(cfg)
        al.fs = 16000
        clips = synth.make_clips(4321, 3, n_samples)
        feats = [np.asarray(al.get_mfcc(c, al.param_dict[ds], "torchaudio"), dtype=np.float64) for c in clips]
        out[tag] = np.stack(feats).astype(np.float32)
        out[tag + "_lens"] = np.array([len(c) for c in clips])
    np.savez_compressed(os.path.join(GOLD, "mfcc.npz"), **out)
    print("mfcc:", {k: v.shape for k, v in out.items()})


def gold_model(net, B, seed):
    if net == omodel.KUKA:
        from models.pretext.arm_pretext_model import VARPretextNet
        cfg = kuka_cfg()
    else:
        from models.pretext.ai2thor_pretext_model import VARPretextNet
        cfg = ithor_cfg()
        torch.Tensor.cuda = lambda self, *a, **k: self  # constructor calls .cuda() (line 46)
    m = VARPretextNet(cfg)
    sd = omodel.init_state_dict(net, seed)
    assert list(m.state_dict().keys()) == list(sd.keys()), (list(m.state_dict().keys()), list(sd.keys()))
    m.load_state_dict(sd)
    m.train()
    images, sp, sn = synth.model_case(net, B, seed)
    d = m(torch.from_numpy(images), torch.from_numpy(sp), torch.from_numpy(sn))
    crit = torch.nn.TripletMarginLoss(margin=1.0, p=2)
    loss = crit(d["image_feat"], d["sound_feat_positive"], d["sound_feat_negative"])
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-6)
    opt.zero_grad()
    loss.backward()
    out = {k: d[k].detach().numpy() for k in
           ("image_feat", "sound_feat_positive", "sound_feat_negative", "image_feat_raw", "pos_sound_raw")}
    out["loss"] = np.array(loss.item(), dtype=np.float32)
    for k, p in m.named_parameters():
        g = p.grad.detach().numpy()
        out["gsum." + k] = np.array([g.sum(dtype=np.float64), np.abs(g).sum(dtype=np.float64),
                                     np.abs(g).max()])
        if g.size <= 8192:
            out["grad." + k] = g
        else:
            out["grad." + k] = g.reshape(-1)[:: max(1, g.size // 4096)][:4096].copy()
    opt.step()
    for k, p in m.named_parameters():
        w = p.detach().numpy()
        out["w1." + k] = w.reshape(-1)[:: max(1, w.size // 1024)][:1024].copy()
    # cached_sound rule: all-inf positive reuses the cache (pretext_base.py:29-32)
    m.eval()
    with torch.no_grad():
        d2 = m(torch.from_numpy(images), torch.full_like(torch.from_numpy(sp), float("inf")), None)
    out["cached_sound_feat"] = d2["sound_feat_positive"].detach().numpy()
    out["eval_image_feat"] = d2["image_feat"].numpy()
    np.savez_compressed(os.path.join(GOLD, f"model_{net}.npz"), B=B, seed=seed, **out)
    print("model", net, "loss", float(loss))


def gold_sampler():
    """Drive the reference VARDataset through DataLoader(num_workers=0) with a recording
    audio stub: the recorded (intent, dataset, clip) draws are the golden index stream."""
    import pickle
    import tempfile
    from dataset import VARDataset
    import torch.utils.data as D
    cfg = kuka_cfg()
    sizes = {0: [7, 3], 1: [5], 2: [11, 2, 4], 3: [6]}
    rec = []

    class AudioStub:
        def genSoundFeat(self, intentIdx, featType, rand_fn, mfcc_from="torchaudio", trans_fn=None):
            if intentIdx > cfg.taskNum - 1:
                intentIdx = cfg.taskNum - 1
            s = sizes[intentIdx]
            ds = int(rand_fn(0, len(s), size=()))
            clip = int(rand_fn(0, s[ds], size=()))
            rec.append((intentIdx, ds, clip))
            return np.zeros(cfg.sound_dim), None

    gts = synth.make_labels(99, 23)
    items = [{"image": np.zeros((3, 2, 2), np.uint8), "ground_truth": int(g)} for g in gts]
    with tempfile.NamedTemporaryFile(suffix=".pickle", delete=False) as f:
        pickle.dump(items, f)
        path = f.name
    ds = VARDataset(path, cfg, audio=AudioStub())
    torch.manual_seed(453)  # pretextEnvSeed (config.py:55; pretext.py:294)
    dl = D.DataLoader(ds, batch_size=5, shuffle=True, num_workers=0, drop_last=False)
    epochs = []
    for ep in range(2):
        rec.clear()
        gt_stream = []
        for image, sp, sn, gt in dl:
            gt_stream += gt.tolist()
        epochs.append((list(gt_stream), [list(r) for r in rec]))
    os.unlink(path)
    np.savez_compressed(os.path.join(GOLD, "sampler_kuka.npz"), gts=gts, seed=453, batch=5,
                        sizes=np.array([str(sizes)]),
                        ep0_gt=np.array(epochs[0][0]), ep0_draws=np.array(epochs[0][1]),
                        ep1_gt=np.array(epochs[1][0]), ep1_draws=np.array(epochs[1][1]))
    print("sampler draws per epoch:", len(epochs[0][1]), len(epochs[1][1]))


ITHOR_LIST_SIZES = {("none", "lights", "activate"): 7, ("none", "lights", "deactivate"): 3,
                    ("none", "music", "activate"): 5, ("none", "music", "deactivate"): 11,
                    ("none", "lamp", "activate"): 4, ("none", "lamp", "deactivate"): 6}


def gold_sampler_ithor():
    """The reference VARDataset under the REAL AI2ThorConfig + EnvConfig (Envs/ai2thor/config.py,
    env_config.py), with the reference audioLoader's own getAudioFromTask / genSoundFeatFromTask
    doing the draws; only get_mfcc is replaced (python_speech_features is absent) by a recorder that
    identifies the chosen clip.  Also records the generator state at loader construction so the
    `continue torch's global generator` path is pinned."""
    import pickle
    import tempfile
    import torch.utils.data as D
    from Envs.ai2thor.config import AI2ThorConfig
    from Envs.ai2thor.env_config import EnvConfig
    from Envs.audioLoader import audioLoader
    from dataset import VARDataset
    cfg = AI2ThorConfig()
    cfg.get_env_config(EnvConfig)
    assert cfg.taskNum == 4 and cfg.name == "AI2ThorConfig"
    keys = list(ITHOR_LIST_SIZES)
    rec = []

    class RecAudio(audioLoader):
        def get_mfcc(self, audioSamples, param, mfcc_from):
            assert mfcc_from is None  # the iTHOR path never asks for torchaudio (audioLoader.py:203,234-236)
            rec.append((int(audioSamples[0]), int(audioSamples[1])))
            return np.zeros(cfg.sound_dim)

    al = RecAudio(cfg)
    al.fs = 16000
    al.transcription = {}
    for li, (loc, obj, act) in enumerate(keys):
        n = ITHOR_LIST_SIZES[(loc, obj, act)]
        al.words.setdefault(loc, {}).setdefault(obj, {})[act] = [np.array([li, j], np.int16) for j in range(n)]
        al.transcription.setdefault(loc, {}).setdefault(obj, {})[act] = ["t"] * n
    gts = synth.make_labels(77, 29)
    items = [{"image": np.zeros((3, 2, 2), np.uint8), "ground_truth": int(g)} for g in gts]
    with tempfile.NamedTemporaryFile(suffix=".pickle", delete=False) as f:
        pickle.dump(items, f)
        path = f.name
    ds = VARDataset(path, cfg, audio=al)
    torch.manual_seed(cfg.pretextEnvSeed)  # 977 (Envs/ai2thor/config.py:60; pretext.py:294)
    torch.rand(37)  # the generator is NOT fresh when training starts (model init consumed it)
    state0 = torch.get_rng_state().numpy().copy()
    dl = D.DataLoader(ds, batch_size=6, shuffle=True, num_workers=0, drop_last=False)
    out = {"gts": gts, "seed": cfg.pretextEnvSeed, "batch": 6, "rng_state": state0,
           "list_keys": np.array(["/".join(k) for k in keys]),
           "list_sizes": np.array([ITHOR_LIST_SIZES[k] for k in keys])}
    for ep in range(2):
        rec.clear()
        gt_stream = []
        for image, sp, sn, gt in dl:
            gt_stream += gt.tolist()
        out[f"ep{ep}_gt"] = np.array(gt_stream)
        out[f"ep{ep}_draws"] = np.array(rec)
    os.unlink(path)
    np.savez_compressed(os.path.join(GOLD, "sampler_ithor.npz"), **out)
    print("ithor sampler draws per epoch:", len(out["ep0_draws"]), len(out["ep1_draws"]))


def gold_reward():
    from Envs.vec_env.vec_pretext_normalize import VecPretextNormalize
    from models.pretext.arm_pretext_model import VARPretextNet
    cfg = kuka_cfg()
    N, steps = 6, 5
    m = VARPretextNet(cfg)
    m.load_state_dict(omodel.init_state_dict(omodel.KUKA, 11))
    m.eval()
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

    pre = types.SimpleNamespace(pretextModel=m)
    w = VecPretextNormalize(Venv(), ob=False, ret=True, gamma=0.99, config=cfg, pretextObj=pre)
    w.device = torch.device("cpu")
    o0 = w.reset()
    out = {"reset_image_feat": o0["image_feat"], "reset_goal_sound_feat": o0["goal_sound_feat"]}
    for t in range(steps):
        o, r, d, _ = w.step_wait()
        out[f"rew{t}"] = r
        out[f"orig{t}"] = w.origStepReward
        out[f"image_feat{t}"] = o["image_feat"]
        out[f"goal_sound_feat{t}"] = o["goal_sound_feat"]
    out["ret_var"] = np.array(w.ret_rms.var)
    out["ret_mean"] = np.array(w.ret_rms.mean)
    np.savez_compressed(os.path.join(GOLD, "reward.npz"), N=N, steps=steps, **out)
    print("reward ok, var", w.ret_rms.var)


def gold_reward_ithor():
    """The reference VecPretextNormalize with the iTHOR VARPretextNet (processAI2Thor, the all-inf
    cached-goal schedule, return normalisation) on the CPU."""
    from Envs.vec_env.vec_pretext_normalize import VecPretextNormalize
    from models.pretext.ai2thor_pretext_model import VARPretextNet
    torch.Tensor.cuda = lambda self, *a, **k: self  # constructor calls .cuda() (ai2thor_pretext_model.py:46)
    cfg = ithor_cfg()
    N, steps = 5, 7
    m = VARPretextNet(cfg)
    m.load_state_dict(omodel.init_state_dict(omodel.ITHOR, 13))
    m.eval()
    obs_seq, rew_seq, done_seq = synth.ithor_reward_case(N, steps)

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

    w = VecPretextNormalize(Venv(), ob=False, ret=True, gamma=0.99, config=cfg,
                            pretextObj=types.SimpleNamespace(pretextModel=m))
    w.device = torch.device("cpu")
    o0 = w.reset()
    out = {"reset_image_feat": o0["image_feat"], "reset_goal_sound_feat": o0["goal_sound_feat"],
           "reset_occupancy_sum": np.array(o0["occupancy"].sum())}
    for t in range(steps):
        o, r, d, _ = w.step_wait()
        out[f"rew{t}"] = r
        out[f"orig{t}"] = w.origStepReward
        out[f"image_feat{t}"] = o["image_feat"]
        out[f"goal_sound_feat{t}"] = o["goal_sound_feat"]
        assert set(o) == {"occupancy", "goal_sound_feat", "image", "image_feat"}
    out["ret_var"] = np.array(w.ret_rms.var)
    out["ret_mean"] = np.array(w.ret_rms.mean)
    np.savez_compressed(os.path.join(GOLD, "reward_ithor.npz"), N=N, steps=steps, **out)
    print("ithor reward ok, var", w.ret_rms.var)


def gold_adam():
    rng = np.random.default_rng(3)
    p0 = rng.standard_normal(257).astype(np.float32)
    grads = rng.standard_normal((4, 257)).astype(np.float32)
    p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([p], lr=1e-4, weight_decay=1e-6)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[2, 3], gamma=0.2)
    traj, lrs = [], []
    for g in grads:
        opt.zero_grad()
        p.grad = torch.from_numpy(g.copy())
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
        traj.append(p.detach().numpy().copy())
    np.savez_compressed(os.path.join(GOLD, "adam.npz"), p0=p0, grads=grads, traj=np.stack(traj),
                        lrs=np.array(lrs))
    print("adam ok")


def gold_init():
    """Seeded construction of the reference modules: per-tensor checksums pin that the host
    mirror reproduces the reference's initial weights (same layer construction order and RNG use)."""
    out = {}
    from models.pretext.arm_pretext_model import VARPretextNet as KukaNet
    torch.manual_seed(453)
    m = KukaNet(kuka_cfg())
    for k, v in m.state_dict().items():
        out["kuka." + k] = np.array([v.double().sum().item(), v.double().abs().sum().item(), v.flatten()[0].item()])
    from models.pretext.ai2thor_pretext_model import VARPretextNet as ThorNet
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.manual_seed(977)
    m = ThorNet(ithor_cfg())
    for k, v in m.state_dict().items():
        out["ithor." + k] = np.array([v.double().sum().item(), v.double().abs().sum().item(), v.flatten()[0].item()])
    np.savez_compressed(os.path.join(GOLD, "init.npz"), **out)
    print("init ok", len(out))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    install_stubs()
    torch.set_num_threads(8)
    if len(sys.argv) > 1:  # regenerate selected fixtures only: python oracle/make_golden.py gold_init ...
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    gold_mfcc()
    gold_sampler()
    gold_sampler_ithor()
    gold_adam()
    gold_reward()
    gold_reward_ithor()
    gold_model(omodel.KUKA, 4, 7)
    gold_model(omodel.ITHOR, 2, 9)
    gold_init()
