"""Per-launch CUDA-event times of one batched reward query (iTHOR net, cached goal embedding) at small N."""
import sys, os, ctypes, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import var_b200 as vb
from oracle import model as omodel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng = vb.VarEngine(vb.ITHOR, 600, 3, "cuda:0")
eng.load_state_dict(omodel.init_state_dict("ithor", 0))
img = torch.randint(0, 256, (N, 3, 96, 96), dtype=torch.uint8, device="cuda")
cached = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda"), dim=1)
er = torch.zeros(N, device="cuda")
for _ in range(3):
    eng.reward(img, goal_feat_cached=cached, env_reward=er)
torch.cuda.synchronize()
lib = vb._lib.lib
lib.var_prof_dump.argtypes = [ctypes.c_char_p]
lib.var_prof_begin()
eng.reward(img, goal_feat_cached=cached, env_reward=er)
os.makedirs("gpurun_out", exist_ok=True)
lib.var_prof_dump(b"gpurun_out/prof_reward.csv")
tot = 0.0
for line in open("gpurun_out/prof_reward.csv").read().splitlines()[1:]:
    tag, ms, fl, note = line.split(",", 3)
    tot += float(ms)
    print(f"{float(ms)*1e3:8.1f} us  {vb._lib.PROF_TAGS[int(tag)]:12s} {note}")
print(f"sum of kernels {tot*1e3:.1f} us (events around each launch: includes launch gaps)")
g = eng.reward_graph(N, torch.uint8, fresh_goal=False)
g.images.copy_(img); g.goal_feat_cached.copy_(cached)
for _ in range(5):
    g.launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(50):
    g.launch()
e1.record(); torch.cuda.synchronize()
print(f"captured graph: {e0.elapsed_time(e1)/50*1e3:.1f} us per query at N={N}")
