"""Forward / dgrad / wgrad of one conv shape through the C ABI (default: thor.snd.conv2), a warm-up
pass plus one measured pass -- small enough for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import var_b200 as vb
lib = vb._lib.lib
N, H, W, Cin, Cout, R, S, sh, sw, ph, pw = [int(x) for x in (sys.argv[1:12] if len(sys.argv) > 11 else
                                                         "128 300 20 64 64 11 5 2 2 5 5".split())]
P, Q = (H + 2 * ph - R) // sh + 1, (W + 2 * pw - S) // sw + 1
K = R * S * Cin
kpad = (K + 31) // 32 * 32
dev = "cuda:0"
x = torch.randn(N, H, W, Cin, device=dev)
w = torch.randn(Cout, kpad, device=dev) * 0.02
b = torch.zeros(Cout, device=dev)
y = torch.empty(N, P, Q, Cout, device=dev)
dy = torch.randn(N, P, Q, Cout, device=dev)
dx = torch.empty_like(x)
dw = torch.zeros(Cout, kpad, device=dev)
db = torch.zeros(Cout, device=dev)
ev = [torch.cuda.Event(True) for _ in range(4)]
for it in range(2):
    ev[0].record()
    assert lib.var_conv2d_fwd(x.data_ptr(), 0, None, 1.0, N, H, W, Cin, Cout, R, S, sh, sw, ph, pw, w.data_ptr(),
                              b.data_ptr(), y.data_ptr(), 1, 1, None) == 0
    ev[1].record()
    assert lib.var_conv2d_dgrad(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), x.data_ptr(), N, H, W, Cin, Cout, R, S,
                                sh, sw, ph, pw, 1, None) == 0
    ev[2].record()
    assert lib.var_conv2d_wgrad(x.data_ptr(), 0, None, 1.0, dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, H, W, Cin,
                                Cout, R, S, sh, sw, ph, pw, None) == 0
    ev[3].record()
    torch.cuda.synchronize()
fl = 2.0 * N * P * Q * Cout * K
for name, a, c in (("fwd", 0, 1), ("dgrad", 1, 2), ("wgrad+colsum", 2, 3)):
    ms = ev[a].elapsed_time(ev[c])
    print(f"{name:14s} {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TF/s")
