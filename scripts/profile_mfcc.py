"""MFCC front-end alone: B clips resident in HBM, CUDA-event timing, GB/s of algorithmic bytes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import var_b200 as vb
from importlib import import_module
al = import_module("voicecontrolledrobot-var_b200.Envs.audioLoader")
cases = [("1s F=100", 2048, 16000, (512, 400, 160), 100, 0), ("1s F=600", 512, 16000, (512, 400, 160), 600, 0),
         ("4s F=100", 2048, 64000, (1024, 800, 640), 100, 0), ("1s F=600 psf", 512, 16000, (512, 400, 160), 600, 1)]
if len(sys.argv) > 1:
    cases = [cases[int(sys.argv[1])]]
rng = np.random.default_rng(0)
for name, B, S, (nfft, win, hop), F, flav in cases:
    wav = torch.from_numpy(rng.integers(-20000, 20000, B * S).astype(np.int16)).cuda()
    off = (torch.arange(B, device="cuda", dtype=torch.int64) * S).contiguous()
    ln = torch.full((B,), S, dtype=torch.int32, device="cuda")
    out = torch.empty(B, F, 40, device="cuda")
    for _ in range(3):
        al.mfcc_device(wav, off, ln, 16000, nfft, win, hop, F, flavour=flav, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    iters = 20
    e0.record()
    for _ in range(iters):
        al.mfcc_device(wav, off, ln, 16000, nfft, win, hop, F, flavour=flav, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = B * (S * 2 + F * 160)
    frames = B * min(F, 1 + S // hop)
    print(f"{name:13s} B={B:5d}  {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s  {frames/ms/1e3:8.1f} Mframes/s")
