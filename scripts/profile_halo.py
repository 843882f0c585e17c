"""Sweeps the launch shape of the im2col-free (halo) f16 conv kernels on the iTHOR sound conv2 / conv3 shapes.
usage: python scripts/profile_halo.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import var_b200 as vb

lib = vb._lib.lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
DEV = "cuda:0"


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


VARIANTS = [
    {},                                   # shipped: im2col-free fwd (3 CTAs/SM) + dgrad (2 CTAs/SM), im2col wgrad
    {"VAR_HALO": "0"},                    # im2col kernels for all three passes
    {"VAR_HALO_WGRAD": "1"},              # im2col-free weight gradient (slower)
    {"VAR_HALO_CPS": "2"},
    {"VAR_HALO_DG_CPS": "3"},
    {"VAR_HALO_CPS": "1", "VAR_HALO_SLOTS": "5", "VAR_HALO_STAGES": "8"},   # one MMA stream per SM
]
for (H, W, Cin, Cout, R, S, sh, sw, ph, pw) in [(300, 20, 64, 64, 11, 5, 2, 2, 5, 5), (150, 13, 64, 64, 7, 3, 2, 2, 1, 1)]:
    P, Q = (H + 2 * ph - R) // sh + 1, (W + 2 * pw - S) // sw + 1
    K = R * S * Cin
    flop = 2.0 * N * P * Q * Cout * K
    xh = torch.randn(N, H, W, Cin, device=DEV).half()
    wh = (torch.randn(Cout, K, device=DEV) * 0.02).half()
    dyh = torch.randn(N, P, Q, Cout, device=DEV).half()
    yh = torch.empty(N, P, Q, Cout, device=DEV, dtype=torch.half)
    dx = torch.empty(N, H, W, Cin, device=DEV)
    b = torch.zeros(Cout, device=DEV)
    geo = (N, H, W, Cin, Cout, R, S, sh, sw, ph, pw)
    print(f"== conv {H}x{W} {R}x{S} N={N}: {flop/1e9:.1f} GFLOP (forward convention)")
    for env in VARIANTS:
        for k in list(os.environ):
            if k.startswith("VAR_HALO_"):
                os.environ.pop(k)
        os.environ.update(env)
        rc = lib.var_conv2d_fwd_h16(xh.data_ptr(), *geo, wh.data_ptr(), b.data_ptr(), yh.data_ptr(), 1, 1, 0, None)
        tf = timeit(lambda: lib.var_conv2d_fwd_h16(xh.data_ptr(), *geo, wh.data_ptr(), b.data_ptr(), yh.data_ptr(), 1, 1, 0, None)) if rc == 0 else float("nan")
        rc2 = lib.var_conv2d_dgrad_h16(dyh.data_ptr(), wh.data_ptr(), dx.data_ptr(), 0, xh.data_ptr(), 1, None, *geo, 1, None)
        td = timeit(lambda: lib.var_conv2d_dgrad_h16(dyh.data_ptr(), wh.data_ptr(), dx.data_ptr(), 0, xh.data_ptr(), 1, None, *geo, 1, None)) if rc2 == 0 else float("nan")
        dw = torch.zeros(Cout, K, device=DEV); db = torch.zeros(Cout, device=DEV)
        rc3 = lib.var_conv2d_wgrad_h16(xh.data_ptr(), dyh.data_ptr(), dw.data_ptr(), db.data_ptr(), None, *geo, None)
        tw = timeit(lambda: lib.var_conv2d_wgrad_h16(xh.data_ptr(), dyh.data_ptr(), dw.data_ptr(), db.data_ptr(), None, *geo, None)) if rc3 == 0 else float("nan")
        print(f"wgrad {tw:.3f} ms {flop/tw/1e9:5.0f} TF/s rc {rc3} | ", end="")
        print(f"fwd {tf:.3f} ms {flop/tf/1e9:5.0f} TF/s | dgrad {td:.3f} ms {flop/td/1e9:5.0f} TF/s | rc {rc} {rc2} | {env}")
