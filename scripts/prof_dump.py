"""Per-launch CUDA-event dump of one training step (grouped by layer) -> gpurun_out/prof_dump.csv"""
import sys, os, ctypes, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import var_b200 as vb
from oracle import model as omodel
net = sys.argv[1] if len(sys.argv) > 1 else "ithor"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
F = 600 if net == "ithor" else 100
eng = vb.VarEngine(vb.ITHOR if net == "ithor" else vb.KUKA, F, 3, "cuda:0")
eng.load_state_dict(omodel.init_state_dict(net, 0))
if len(sys.argv) > 3 and sys.argv[3] == "serial":
    eng.set_overlap(False)
img = torch.randint(0, 256, (B, 3, 96, 96), dtype=torch.uint8, device="cuda")
snd = torch.randn(2 * B, F, 40, device="cuda") * 4
for _ in range(2):
    eng.zero_grad(); eng.triplet_step(img, snd)
torch.cuda.synchronize()
lib = vb._lib.lib
lib.var_prof_dump.argtypes = [ctypes.c_char_p]
lib.var_prof_begin()
eng.zero_grad(); eng.triplet_step(img, snd); eng.adam_step(1e-4)
os.makedirs("gpurun_out", exist_ok=True)
lib.var_prof_dump(b"gpurun_out/prof_dump.csv")
agg = collections.OrderedDict()
for line in open("gpurun_out/prof_dump.csv").read().splitlines()[1:]:
    tag, ms, fl, note = line.split(",", 3)
    k = (vb._lib.PROF_TAGS[int(tag)], note)
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += float(ms); a[2] += float(fl)
tot = sum(a[1] for a in agg.values())
print(f"total kernel ms {tot:.3f}")
for (tag, note), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  n={n:3d} {fl/ms/1e9 if ms else 0:7.1f} TF/s  {tag:12s} {note}")
