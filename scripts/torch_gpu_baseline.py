"""Same-box comparison bar (SURVEY §8d "time the reference on 1 B200"): the reference's training
step written the reference's way -- eager PyTorch modules (cuDNN convs, cuDNN GRU), torchaudio
MFCC on the GPU, nn.TripletMarginLoss, torch.optim.Adam -- on the workload shapes bench.py uses.
Nothing of this repo's CUDA path runs here (only the layer spec tables are imported), and the
result is NOT a bench.py line: it is written to gpurun_out/torch_gpu_baseline.json and quoted
in DESIGN.md §5 next to the numbers of the hand-written path.

usage: python scripts/torch_gpu_baseline.py [ithor|kuka] [batch] [steps]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchaudio

import importlib
pkg = importlib.import_module("voicecontrolledrobot-var_b200.models.pretext._layers")
thor = importlib.import_module("voicecontrolledrobot-var_b200.models.pretext.ai2thor_pretext_model")
arm = importlib.import_module("voicecontrolledrobot-var_b200.models.pretext.arm_pretext_model")


class ThorNet(nn.Module):
    def __init__(self, D=3):
        super().__init__()
        self.imgBranch = pkg.conv_stack(thor.IMG_SPEC)
        self.rnn = nn.GRU(thor.GRU_IN, thor.GRU_HIDDEN, batch_first=True, bidirectional=True)
        self.cnn = pkg.conv_stack(thor.SND_SPEC, flatten=False)
        self.imgTriplet = pkg.mlp_head([1152, 128, D])
        self.soundTriplet = pkg.mlp_head([1024, 128, 64, D])

    def sound(self, s):
        x = self.cnn(s)                                  # [B, 64, 73, 7]
        x = x.transpose(1, 2).reshape(x.shape[0], thor.GRU_STEPS, -1)
        _, h = self.rnn(x)
        return F.normalize(self.soundTriplet(torch.cat([h[0], h[1]], 1)), p=2, dim=1)

    def forward(self, img, sp, sn):
        return F.normalize(self.imgTriplet(self.imgBranch(img)), p=2, dim=1), self.sound(sp), self.sound(sn)


class KukaNet(nn.Module):
    def __init__(self, D=3):
        super().__init__()
        self.imgBranch = pkg.conv_stack(arm.IMG_SPEC)
        self.soundCNN = pkg.conv_stack(arm.SND_SPEC)
        self.imgTriplet = pkg.mlp_head([576, 128, D])
        self.soundTriplet = pkg.mlp_head([160, 128, D])

    def forward(self, img, sp, sn):
        snd = lambda s: F.normalize(self.soundTriplet(self.soundCNN(s)), p=2, dim=1)
        return F.normalize(self.imgTriplet(self.imgBranch(img)), p=2, dim=1), snd(sp), snd(sn)


def run(net, B, steps, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    dev = "cuda:0"
    torch.manual_seed(0)
    model = (ThorNet() if net == "ithor" else KukaNet()).to(dev)
    Fr = 600 if net == "ithor" else 100
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-6)
    crit = nn.TripletMarginLoss(margin=1.0, p=2)
    mfcc = torchaudio.transforms.MFCC(sample_rate=16000, n_mfcc=40, log_mels=True, melkwargs=dict(
        n_fft=512, win_length=400, hop_length=160, n_mels=40, f_min=0, f_max=None,
        window_fn=torch.hamming_window)).to(dev)
    img_u8 = torch.randint(0, 256, (B, 3, 96, 96), dtype=torch.uint8, device=dev)
    wav = (torch.randn(2 * B, 16000, device=dev) * 8000).clamp(-32767, 32767).to(torch.int16)

    def step():
        feats = mfcc(wav.float() / 32768.0).transpose(1, 2)           # [2B, 101, 40]
        if Fr > feats.shape[1]:
            feats = F.pad(feats, (0, 0, 0, Fr - feats.shape[1]))
        else:
            feats = feats[:, :Fr]
        snd = feats[:, None].contiguous()
        img = img_u8.float() / 255.0
        opt.zero_grad()
        a, p, n = model(img, snd[:B], snd[B:])
        loss = crit(a, p, n)
        loss.backward()
        opt.step()
        return loss

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"net": net, "batch": B, "tf32": tf32, "ms_per_step": ms, "triplets_per_s": B / (ms * 1e-3)}


if __name__ == "__main__":
    net = sys.argv[1] if len(sys.argv) > 1 else "ithor"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (256 if net == "ithor" else 64)
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    out = {"what": "eager PyTorch (cuDNN conv + cuDNN GRU + torchaudio MFCC + Adam) port of the reference step, "
                   "same shapes as bench.py; inputs resident on the device",
           "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "runs": []}
    for tf32 in (True, False):
        t0 = time.time()
        r = run(net, B, steps, tf32)
        r["wall_s"] = time.time() - t0
        out["runs"].append(r)
        print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/torch_gpu_baseline_{net}_b{B}.json", "w") as f:
        json.dump(out, f, indent=1)
