"""Times the 16-bit conv ops (and their tf32 counterparts) on the iTHOR sound conv shapes at training size.
usage: python scripts/profile_h16.py [N]"""
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


for (H, W, Cin, Cout, R, S, sh, sw, ph, pw) in [(300, 20, 64, 64, 11, 5, 2, 2, 5, 5), (150, 13, 64, 64, 7, 3, 2, 2, 1, 1)]:
    P, Q = (H + 2 * ph - R) // sh + 1, (W + 2 * pw - S) // sw + 1
    K = R * S * Cin
    flop = 2.0 * N * P * Q * Cout * K
    x = torch.randn(N, H, W, Cin, device=DEV)
    xh = x.half()
    w32 = torch.randn(Cout, K, device=DEV) * 0.02
    wh = w32.half()
    dy = torch.randn(N, P, Q, Cout, device=DEV)
    dyh = dy.half()
    y = torch.empty(N, P, Q, Cout, device=DEV)
    yh = torch.empty(N, P, Q, Cout, device=DEV, dtype=torch.half)
    dx = torch.empty(N, H, W, Cin, device=DEV)
    dw = torch.zeros(Cout, K, device=DEV)
    db = torch.zeros(Cout, device=DEV)
    b = torch.zeros(Cout, device=DEV)
    geo = (N, H, W, Cin, Cout, R, S, sh, sw, ph, pw)
    print(f"== conv {H}x{W} Cin{Cin} Cout{Cout} {R}x{S} N={N}: {flop/1e9:.1f} GFLOP per pass")
    t = timeit(lambda: lib.var_conv2d_fwd(x.data_ptr(), 0, None, 1.0, *geo, w32.data_ptr(), b.data_ptr(), y.data_ptr(), 1, 1, None))
    print(f"fwd   tf32 {t:.3f} ms {flop/t/1e9:.0f} TF/s")
    t = timeit(lambda: lib.var_conv2d_fwd_h16(xh.data_ptr(), *geo, wh.data_ptr(), b.data_ptr(), yh.data_ptr(), 1, 1, 0, None))
    print(f"fwd   f16  {t:.3f} ms {flop/t/1e9:.0f} TF/s")
    t = timeit(lambda: lib.var_conv2d_dgrad(dy.data_ptr(), w32.data_ptr(), dx.data_ptr(), x.data_ptr(), *geo, 1, None))
    print(f"dgrad tf32 {t:.3f} ms {flop/t/1e9:.0f} TF/s")
    t = timeit(lambda: lib.var_conv2d_dgrad_h16(dyh.data_ptr(), wh.data_ptr(), dx.data_ptr(), 0, xh.data_ptr(), 1, None, *geo, 1, None))
    print(f"dgrad f16  {t:.3f} ms {flop/t/1e9:.0f} TF/s")
    t = timeit(lambda: lib.var_conv2d_wgrad(x.data_ptr(), 0, None, 1.0, dy.data_ptr(), dw.data_ptr(), db.data_ptr(), *geo, None))
    print(f"wgrad tf32 {t:.3f} ms {flop/t/1e9:.0f} TF/s (+ colsum)")
    for pb, st in ((32, 8), (64, 4), (64, 6), (128, 3), (128, 4), (256, 2)):
        os.environ["VAR_WGRAD16_PB"], os.environ["VAR_WGRAD16_STAGES"] = str(pb), str(st)
        rc = lib.var_conv2d_wgrad_h16(xh.data_ptr(), dyh.data_ptr(), dw.data_ptr(), db.data_ptr(), None, *geo, None)
        if rc != 0:
            print(f"wgrad f16 pb={pb} stages={st}: rc {rc} {vb._lib.last_error()}")
            continue
        t = timeit(lambda: lib.var_conv2d_wgrad_h16(xh.data_ptr(), dyh.data_ptr(), dw.data_ptr(), db.data_ptr(), None, *geo, None))
        print(f"wgrad f16  pb={pb:3d} stages={st} {t:.3f} ms {flop/t/1e9:.0f} TF/s (bias grad fused)")
    os.environ.pop("VAR_WGRAD16_PB"); os.environ.pop("VAR_WGRAD16_STAGES")
