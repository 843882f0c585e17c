"""ORACLE (test infrastructure only -- never imported by the product path).

Plain fp32 PyTorch-on-CPU restatement of the two VAR encoders and of
`PretextNetBase.VAR_forward`, written functionally over a reference-layout
`state_dict` so it shares no code with the product modules.

Follows
  models/pretext/arm_pretext_model.py:9-59      (Kuka net)
  models/pretext/ai2thor_pretext_model.py:5-64  (iTHOR net)
  models/pretext/pretext_base.py:10-42          (VAR_forward, cached_sound rule)
PINNED: oracle/make_golden.py runs the imported reference modules on the same
seeded weights/inputs and stores their outputs in tests/golden/model_*.npz;
tests/test_oracle.py checks this file against them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

KUKA = "kuka"
ITHOR = "ithor"

# (name, shape) in reference state_dict order
KUKA_PARAMS = [
    ("imgBranch.0", (32, 3, 3, 3)), ("imgBranch.2", (32, 32, 3, 3)), ("imgBranch.4", (64, 32, 3, 3)),
    ("imgBranch.6", (64, 64, 3, 3)), ("imgBranch.8", (64, 64, 3, 3)),
    ("soundCNN.0", (32, 1, 5, 40)), ("soundCNN.2", (32, 32, 3, 1)), ("soundCNN.4", (32, 32, 3, 1)),
    ("soundCNN.6", (32, 32, 3, 1)),
    ("imgTriplet.0", (128, 576)), ("imgTriplet.2", (3, 128)),
    ("soundTriplet.0", (128, 160)), ("soundTriplet.2", (3, 128)),
]
ITHOR_PARAMS = [
    ("imgBranch.0", (32, 3, 3, 3)), ("imgBranch.2", (32, 32, 3, 3)), ("imgBranch.5", (64, 32, 3, 3)),
    ("imgBranch.8", (64, 64, 3, 3)), ("imgBranch.11", (128, 64, 3, 3)), ("imgBranch.14", (128, 128, 3, 3)),
    ("cnn.0", (64, 1, 11, 11)), ("cnn.2", (64, 64, 11, 5)), ("cnn.4", (64, 64, 7, 3)),
    ("imgTriplet.0", (128, 1152)), ("imgTriplet.2", (3, 128)),
    ("soundTriplet.0", (128, 1024)), ("soundTriplet.2", (64, 128)), ("soundTriplet.4", (3, 64)),
]
ITHOR_RNN = [
    ("rnn.weight_ih_l0", (1536, 448)), ("rnn.weight_hh_l0", (1536, 512)),
    ("rnn.bias_ih_l0", (1536,)), ("rnn.bias_hh_l0", (1536,)),
    ("rnn.weight_ih_l0_reverse", (1536, 448)), ("rnn.weight_hh_l0_reverse", (1536, 512)),
    ("rnn.bias_ih_l0_reverse", (1536,)), ("rnn.bias_hh_l0_reverse", (1536,)),
]


def param_shapes(net):
    """Ordered {key: shape} exactly as the reference module's state_dict()."""
    out = {}
    if net == KUKA:
        for name, shp in KUKA_PARAMS:
            out[name + ".weight"] = shp
            out[name + ".bias"] = (shp[0],)
        # reference registration order: imgBranch, soundCNN, imgTriplet, soundTriplet
        return out
    # iTHOR registration order (ai2thor_pretext_model.py:41-61): imgBranch, rnn, cnn, imgTriplet, soundTriplet
    for name, shp in ITHOR_PARAMS[:6]:
        out[name + ".weight"] = shp
        out[name + ".bias"] = (shp[0],)
    for name, shp in ITHOR_RNN:
        out[name] = shp
    for name, shp in ITHOR_PARAMS[6:]:
        out[name + ".weight"] = shp
        out[name + ".bias"] = (shp[0],)
    return out


def init_state_dict(net, seed):
    """Deterministic synthetic weights (numpy RNG; torch-default-like fan-in uniform scale)."""
    rng = np.random.default_rng(seed)
    sd = {}
    shapes = param_shapes(net)
    for k, shp in shapes.items():
        if k.startswith("rnn."):
            bound = 1.0 / math.sqrt(512)
        elif k.endswith(".weight"):
            bound = 1.0 / math.sqrt(int(np.prod(shp[1:])))
        else:
            wshape = shapes[k[:-5] + ".weight"]
            bound = 1.0 / math.sqrt(int(np.prod(wshape[1:])))
        sd[k] = torch.from_numpy(rng.uniform(-bound, bound, size=shp).astype(np.float32))
    return sd


def _img_branch(net, sd, image):
    x = image[:, :3, :, :]
    if net == KUKA:
        for i in (0, 2, 4, 6, 8):
            x = F.relu(F.conv2d(x, sd[f"imgBranch.{i}.weight"], sd[f"imgBranch.{i}.bias"], stride=2, padding=1))
        return x.reshape(x.size(0), -1)
    x = F.relu(F.conv2d(x, sd["imgBranch.0.weight"], sd["imgBranch.0.bias"], padding=1))
    x = F.relu(F.conv2d(x, sd["imgBranch.2.weight"], sd["imgBranch.2.bias"], padding=1))
    x = F.max_pool2d(x, 2, 2)
    x = F.relu(F.conv2d(x, sd["imgBranch.5.weight"], sd["imgBranch.5.bias"], padding=1))
    x = F.max_pool2d(x, 2, 2)
    x = F.relu(F.conv2d(x, sd["imgBranch.8.weight"], sd["imgBranch.8.bias"], padding=1))
    x = F.max_pool2d(x, 2, 2)
    x = F.relu(F.conv2d(x, sd["imgBranch.11.weight"], sd["imgBranch.11.bias"], padding=1))
    x = F.max_pool2d(x, 2, 2)
    x = F.relu(F.conv2d(x, sd["imgBranch.14.weight"], sd["imgBranch.14.bias"], stride=2, padding=1))
    return x.flatten(1)


def gru_bidir_final(x, sd):
    """torch.nn.GRU(448, 512, batch_first, bidirectional) final hidden states, written out.
    Gate order r, z, n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn))."""
    B, T, _ = x.shape
    outs = []
    for suffix, order in (("", range(T)), ("_reverse", range(T - 1, -1, -1))):
        wih, whh = sd["rnn.weight_ih_l0" + suffix], sd["rnn.weight_hh_l0" + suffix]
        bih, bhh = sd["rnn.bias_ih_l0" + suffix], sd["rnn.bias_hh_l0" + suffix]
        h = torch.zeros(B, 512, dtype=x.dtype)
        for t in order:
            gi = x[:, t] @ wih.t() + bih
            gh = h @ whh.t() + bhh
            ir, iz, inn = gi.chunk(3, 1)
            hr, hz, hn = gh.chunk(3, 1)
            r = torch.sigmoid(ir + hr)
            z = torch.sigmoid(iz + hz)
            n = torch.tanh(inn + r * hn)
            h = (1 - z) * n + z * h
        outs.append(h)
    return torch.cat(outs, dim=1)


def _sound_branch(net, sd, sound):
    if net == KUKA:
        x = F.relu(F.conv2d(sound, sd["soundCNN.0.weight"], sd["soundCNN.0.bias"], stride=(2, 1)))
        for i in (2, 4, 6):
            x = F.relu(F.conv2d(x, sd[f"soundCNN.{i}.weight"], sd[f"soundCNN.{i}.bias"], stride=(2, 1)))
        return x.reshape(x.size(0), -1)
    x = F.relu(F.conv2d(sound, sd["cnn.0.weight"], sd["cnn.0.bias"], stride=2, padding=5))
    x = F.relu(F.conv2d(x, sd["cnn.2.weight"], sd["cnn.2.bias"], stride=2, padding=5))
    x = F.relu(F.conv2d(x, sd["cnn.4.weight"], sd["cnn.4.bias"], stride=2, padding=1))
    x = torch.reshape(torch.transpose(x, 1, 2), (-1, 73, 64 * 7))
    return gru_bidir_final(x, sd)


def _head(sd, prefix, raw, nlayers):
    x = raw
    idx = 0
    for li in range(nlayers):
        x = F.linear(x, sd[f"{prefix}.{idx}.weight"], sd[f"{prefix}.{idx}.bias"])
        if li < nlayers - 1:
            x = F.relu(x)
        idx += 2
    return x


def l2_normalize(x, eps=1e-12):
    """F.normalize(x, p=2, dim=1): x / max(||x||, eps)."""
    n = x.pow(2).sum(dim=1, keepdim=True).sqrt().clamp_min(eps)
    return x / n


class OracleVAR:
    """Stateful wrapper reproducing pretext_base.py:10-42 including `cached_sound`."""

    def __init__(self, net, state_dict):
        self.net = net
        self.sd = state_dict
        self.cached_sound = None

    def sound(self, s):
        raw = _sound_branch(self.net, self.sd, s)
        feat = l2_normalize(_head(self.sd, "soundTriplet", raw, 2 if self.net == KUKA else 3))
        return raw, feat

    def __call__(self, image, sound_positive, sound_negative):
        image_feat = image_feat_raw = sound_feat_negative = pos_sound_raw = None
        if image is not None:
            image_feat_raw = _img_branch(self.net, self.sd, image)
            image_feat = l2_normalize(_head(self.sd, "imgTriplet", image_feat_raw, 2))
        if sound_positive is not None and (not torch.isinf(sound_positive).all()):
            pos_sound_raw, feat = self.sound(sound_positive)
            self.cached_sound = feat
        sound_feat_positive = self.cached_sound
        if sound_negative is not None:
            _, sound_feat_negative = self.sound(sound_negative)
        return {"image_feat": image_feat, "sound_feat_positive": sound_feat_positive,
                "sound_feat_negative": sound_feat_negative, "image_BCE": None, "sound_BCE": None,
                "image_feat_raw": image_feat_raw, "pos_sound_raw": pos_sound_raw}


def triplet_margin_loss(a, p, n, margin=1.0, eps=1e-6):
    """torch.nn.TripletMarginLoss(margin, p=2) written out (VAR/pretext_VAR.py:38,64):
    mean(max(0, ||a-p+eps|| - ||a-n+eps|| + margin)), eps added element-wise."""
    dp = (a - p + eps).pow(2).sum(1).sqrt()
    dn = (a - n + eps).pow(2).sum(1).sqrt()
    return torch.clamp(dp - dn + margin, min=0).mean()
