"""CPU: the oracle restatement against the golden vectors produced by the unmodified
reference (oracle/make_golden.py).  This pins the oracle; the GPU tests then compare the
CUDA path with the same vectors and with the oracle on fresh inputs."""
import ast

import numpy as np
import torch

from conftest import rel_to_max
from oracle import mfcc as omfcc
from oracle import model as omodel
from oracle import optim as ooptim
from oracle import reward as oreward
from oracle import sampler as osampler
from oracle import synth

MFCC_CASES = {  # tag -> (n_samples, (n_fft, win, hop), F)
    "kuka_1s": (16000, (512, 400, 160), 100),
    "kuka_fsc": ((16000, 64000), (512, 400, 160), 100),
    "kuka_4s_nsynth": (64000, (1024, 800, 640), 100),
    "ithor_1s": (16000, (512, 400, 160), 600),
    "ithor_short": (4000, (512, 400, 160), 600),
}


def test_mfcc_oracle_matches_reference(golden):
    g = golden("mfcc")
    for tag, (ns, (nfft, win, hop), F) in MFCC_CASES.items():
        clips = synth.make_clips(4321, 3, ns)
        assert [len(c) for c in clips] == g[tag + "_lens"].tolist()
        for i, c in enumerate(clips):
            feat = omfcc.process_sound_feat(omfcc.mfcc_torchaudio(c, 16000, nfft, win, hop), (1, F, 40))
            ref = g[tag][i]
            assert feat.shape == ref.shape
            # two fp32 FFT implementations (numpy pocketfft vs torch): allow fp32 round-off
            assert np.allclose(feat, ref, rtol=1e-4, atol=2e-3), (tag, i, np.abs(feat - ref).max())


def test_model_oracle_matches_reference(golden):
    for net, B, seed in ((omodel.KUKA, 4, 7), (omodel.ITHOR, 2, 9)):
        g = golden("model_" + net)
        sd = omodel.init_state_dict(net, seed)
        images, sp, sn = synth.model_case(net, B, seed)
        o = omodel.OracleVAR(net, sd)
        with torch.no_grad():
            d = o(torch.from_numpy(images), torch.from_numpy(sp), torch.from_numpy(sn))
        for k in ("image_feat", "sound_feat_positive", "sound_feat_negative", "image_feat_raw", "pos_sound_raw"):
            assert rel_to_max(d[k].numpy(), g[k]) < 2e-5, (net, k)
        loss = omodel.triplet_margin_loss(d["image_feat"], d["sound_feat_positive"], d["sound_feat_negative"])
        assert abs(float(loss) - float(g["loss"])) < 1e-5
        with torch.no_grad():
            d2 = o(torch.from_numpy(images), torch.full((B, 1, sp.shape[2], 40), float("inf")), None)
        assert rel_to_max(d2["sound_feat_positive"].numpy(), g["cached_sound_feat"]) < 2e-5


def test_sampler_oracle_matches_reference(golden):
    g = golden("sampler_kuka")
    sizes = ast.literal_eval(str(g["sizes"][0]))
    gts = g["gts"].tolist()
    gen = osampler.TorchCPUGenerator(int(g["seed"]))
    for ep in range(2):
        draws, order = [], []
        for batch in osampler.epoch_batches(gen, len(gts), int(g["batch"])):
            for idx in batch:
                order.append(gts[idx])
                _, pos, neg = osampler.sample_triplet_kuka(gen, gts[idx], 4, sizes)
                draws += [list(x) for x in (pos, neg) if x is not None]
        assert order == g[f"ep{ep}_gt"].tolist()
        assert draws == g[f"ep{ep}_draws"].tolist()


def test_adam_oracle_matches_torch(golden):
    g = golden("adam")
    p = g["p0"].copy()
    st = ooptim.AdamState(p.size)
    for i, grad in enumerate(g["grads"]):
        ooptim.adam_step(p, grad, st, float(g["lrs"][i]))
        assert np.allclose(p, g["traj"][i], rtol=0, atol=2e-7)
    assert np.allclose([ooptim.multistep_lr(1e-4, e, [2, 3], 0.2) for e in range(4)], g["lrs"])


def test_reward_oracle_matches_reference(golden):
    g = golden("reward")
    N, steps = int(g["N"]), int(g["steps"])
    rng = np.random.default_rng(5)
    norm = oreward.ReturnNormalizer(N, gamma=0.99)
    for t in range(steps + 1):
        rng.integers(0, 256, (N, 3, 96, 96)); rng.standard_normal((N, 1, 100, 40)); rng.standard_normal((N, 4))
        env_rew = rng.standard_normal(N)
        done = rng.random(N) < 0.3
        if t == 0:
            continue
        rew, _, _ = oreward.calc_reward(env_rew, g[f"image_feat{t-1}"], g[f"goal_sound_feat{t-1}"])
        out, orig = norm.step(rew, done)
        assert np.allclose(orig, g[f"orig{t-1}"], atol=1e-6)
        assert np.allclose(out, g[f"rew{t-1}"], atol=1e-6)
    assert abs(norm.rms.var - float(g["ret_var"])) < 1e-9


from conftest import ITHOR_ALL_TASKS, ITHOR_OBJ_ACT, ITHOR_SYNONYM  # noqa: E402


def ithor_golden_tables(g):
    keys = [tuple(k.split("/")) for k in g["list_keys"].tolist()]
    sizes = dict(zip(keys, g["list_sizes"].tolist()))
    words = {}
    for (loc, obj, act), n in sizes.items():
        words.setdefault(loc, {}).setdefault(obj, {})[act] = [None] * n
    n_loc, n_obj, lists = osampler.ithor_task_tables(ITHOR_ALL_TASKS, ITHOR_SYNONYM, ITHOR_OBJ_ACT, words)
    tsizes = [[[sizes[k] for k in col] for col in row] for row in lists]
    return keys, n_loc, n_obj, lists, tsizes


def test_ithor_sampler_oracle_matches_reference(golden):
    """dataset.py:17-53 + audioLoader.py:203-237 driven through the REAL AI2ThorConfig/EnvConfig by
    oracle/make_golden.py::gold_sampler_ithor; the stream continues a non-fresh torch generator."""
    g = golden("sampler_ithor")
    keys, n_loc, n_obj, lists, tsizes = ithor_golden_tables(g)
    gts = g["gts"].tolist()
    gen = osampler.TorchCPUGenerator.from_torch_state(g["rng_state"])
    for ep in range(2):
        gt_stream, draws = [], []
        for batch in osampler.epoch_batches(gen, len(gts), int(g["batch"])):
            for idx in batch:
                gt_stream.append(gts[idx])
                _, pos, neg = osampler.sample_triplet_ithor(gen, gts[idx], 4, n_loc, n_obj, tsizes)
                for d in (pos, neg):
                    if d is not None:
                        t, li, oi, clip = d
                        draws.append([keys.index(lists[t][li][oi]), clip])
        assert gt_stream == g[f"ep{ep}_gt"].tolist()
        assert draws == g[f"ep{ep}_draws"].tolist()


def test_torch_state_words_continue_global_generator():
    torch.manual_seed(5)
    torch.rand(1000)  # crosses a twist boundary
    gen = osampler.TorchCPUGenerator.from_torch_state(torch.get_rng_state().numpy())
    want = [int(torch.randint(0, 1000003, size=())) for _ in range(700)]
    assert [gen.randint(0, 1000003) for _ in range(700)] == want
    torch.manual_seed(6)  # fresh seed: left_ == 1, twist before the first draw
    gen = osampler.TorchCPUGenerator.from_torch_state(torch.get_rng_state().numpy())
    assert gen.randint(0, 97) == int(torch.randint(0, 97, size=()))
