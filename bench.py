#!/usr/bin/env python
"""VAR hot-path benchmark: triplets/s of the full training step (+ reward queries/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ithor_b256|kuka_b64|kuka_dp8192|mfcc_4s]
    python bench.py --impl reference ...      # the CPU arm: oracle port on the host cores

One step = sample B triplets (device mt19937) -> MFCC of the 2B selected clips -> both encoders
forward -> fused head/normalise/triplet-margin loss + gradient -> full backward -> [NCCL
all-reduce of the flat gradient buffer] -> fused Adam.  Weights are random-init, data synthetic
(oracle/synth.py); everything the step reads is resident in HBM for `value`, and copied from
pinned host memory every step for `e2e`.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (net, per-GPU batch (None = global/N), global batch, clip samples, stft, F, scaling)
    "ithor_b256": dict(net="ithor", batch=256, clip=16000, stft=(512, 400, 160), F=600, scaling="weak",
                       desc="BASELINE configs[1]: iTHOR VAR pretext training, GoogleCommand-shaped 1 s 16 kHz clips "
                            "(512/400/160) padded to 600 frames, batch 256 per GPU"),
    "kuka_b64": dict(net="kuka", batch=64, clip=16000, stft=(512, 400, 160), F=100, scaling="weak",
                     desc="BASELINE configs[0]: Kuka VAR pretext training, 1 s clips, batch 64 per GPU"),
    "kuka_dp8192": dict(net="kuka", batch=None, global_batch=8192, clip=16000, stft=(512, 400, 160), F=100,
                        scaling="strong", desc="BASELINE configs[4]: Kuka data-parallel training, global batch 8192"),
    "mfcc_4s": dict(net="kuka", batch=1024, clip=64000, stft=(1024, 800, 640), F=100, scaling="weak",
                    desc="BASELINE configs[3]: NSynth-shaped 4 s clips (1024/800/640), batch 1024 per GPU"),
}
TASK_NUM, CLIPS_PER_CLASS, IMAGE_POOL, N_ITEMS = 4, 1000, 4096, 65536
FLOP_PER_TRIPLET = {"kuka": 75.5e6, "ithor": 9.95e9}       # SURVEY.md section 8(d), fwd + bwd
FLOP_PER_QUERY = {"kuka": 23.4e6, "ithor": 337.4e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ithor_b256", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reward", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------- data
def synth_data(wl, seed=4321):
    """Seeded synthetic pools.  Clip synthesis costs ~1 ms per second of
    audio on the host, so the pool is built from 64 distinct clips per class rolled to 1000
    variants (distinct sample offsets keep every clip unique for the MFCC)."""
    from oracle import synth
    rng = np.random.default_rng(seed)
    base = [synth.make_clips(seed + c, 64, wl["clip"]) for c in range(TASK_NUM)]
    words = {}
    for c in range(TASK_NUM):
        clips = []
        for i in range(CLIPS_PER_CLASS):
            b = base[c][i % 64]
            clips.append(np.roll(b, 37 * (i // 64)) if i >= 64 else b)
        words[c] = {"GoogleCommand": clips}
    images = rng.integers(0, 256, size=(IMAGE_POOL, 3, 96, 96), dtype=np.uint8)
    gts = synth.make_labels(seed + 99, N_ITEMS)
    return words, images, gts


class Cfg:
    pass


def make_config(wl):
    c = Cfg()
    c.name = "ArmConfig"
    c.img_dim = (3, 96, 96)
    c.sound_dim = (1, wl["F"], 40)
    c.representationDim = 3
    c.taskNum = TASK_NUM
    c.tripletMargin = 1.0
    c.envFolder = os.path.join("pybullet", "arms")
    c.soundSource = {"dataset": ["GoogleCommand"]}
    c.pretextAdamL2 = 1e-6
    c.pretextLR = 1e-4
    c.RLRewardSoundSound = False
    return c


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out = self.proc.communicate()[0]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def measure_tf32_peak(dev):
    """cuBLAS TF32 GEMM 8192^3, measured like MEASURED_PEAKS.json's bf16 figure (best of 5): the
    tensor roofline denominator for the tf32 conv kernels (no TF32 peak is recorded there)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
    torch.matmul(a, b)
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return best


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    import var_b200 as vb
    from importlib import import_module
    al = import_module("voicecontrolledrobot-var_b200.Envs.audioLoader")
    ds = import_module("voicecontrolledrobot-var_b200.dataset")
    lib = vb._lib.lib
    from oracle import model as omodel

    wl = WORKLOADS[args.workload]
    net = wl["net"]
    global_b = wl["batch"] * world if wl["batch"] else wl["global_batch"]
    local_b = global_b // world
    cfg = make_config(wl)
    words, images, gts = synth_data(wl)
    audio = al.audioLoader(cfg)
    audio.fs = 16000
    audio.words = words
    arena = audio.build_arena(dev)
    # N_ITEMS triplet records index an IMAGE_POOL-image pool (sampled with replacement)
    sampler = ds.DeviceTripletSampler(TASK_NUM, arena.dataset_sizes, gts, None, seed=977, device=dev,
                                      clip_off=arena.clip_off, clip_len=arena.clip_len)
    pool = torch.from_numpy(images).to(dev)
    eng = vb.VarEngine(vb.KUKA if net == "kuka" else vb.ITHOR, wl["F"], 3, dev)
    eng.load_state_dict(omodel.init_state_dict(net, 0))
    n_fft, win, hop = wl["stft"]
    loss_acc = torch.zeros((), device=dev)

    perm_holder = {"perm": None, "pos": 0}

    def next_items():
        if perm_holder["perm"] is None or perm_holder["pos"] + global_b > N_ITEMS:
            perm_holder["perm"] = sampler.begin_epoch()
            perm_holder["pos"] = 0
        s = perm_holder["pos"]
        perm_holder["pos"] += global_b
        return perm_holder["perm"][s:s + global_b]

    def step_resident(comm=True):
        rec = sampler.sample(next_items())
        lo, hi = local_b * rank, local_b * (rank + 1)
        off = torch.cat([rec["off"][lo:hi], rec["off"][global_b + lo:global_b + hi]])
        ln = torch.cat([rec["len"][lo:hi], rec["len"][global_b + lo:global_b + hi]])
        sounds = al.mfcc_device(arena.wav, off, ln, 16000, n_fft, win, hop, wl["F"])
        img = pool[(rec["item"][lo:hi] % IMAGE_POOL).long()]
        eng.zero_grad()
        eng.triplet_step(img, sounds, margin=1.0, loss_denominator=global_b, loss_out=loss_acc)
        if world > 1 and comm:
            dist.all_reduce(eng.grads)
        eng.adam_step(1e-4, weight_decay=1e-6)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = lib.var_launch_count()
    total_ms = timed(step_resident, args.steps)
    launches = lib.var_launch_count() - l0
    clk = clocks.stop()
    ms_per_step = total_ms / args.steps
    value = global_b / (ms_per_step * 1e-3)

    # ---- end to end: host buffers in, loss out, every step ------------------------------
    clip_len = wl["clip"]
    h_img = torch.empty(local_b, 3, 96, 96, dtype=torch.uint8).pin_memory()
    h_wav = torch.empty(2 * local_b, clip_len, dtype=torch.int16).pin_memory()
    rng = np.random.default_rng(rank)
    h_img.copy_(torch.from_numpy(images[rng.integers(0, IMAGE_POOL, local_b)]))
    flat = [c for k in range(TASK_NUM) for c in words[k]["GoogleCommand"]]
    for i in range(2 * local_b):
        h_wav[i].copy_(torch.from_numpy(flat[int(rng.integers(0, len(flat)))]))
    d_off = (torch.arange(2 * local_b, device=dev, dtype=torch.int64) * clip_len).contiguous()
    d_len = torch.full((2 * local_b,), clip_len, dtype=torch.int32, device=dev)
    h_loss = torch.zeros(1).pin_memory()

    def step_e2e():
        img = h_img.to(dev, non_blocking=True)
        wav = h_wav.to(dev, non_blocking=True)
        sounds = al.mfcc_device(wav, d_off, d_len, 16000, n_fft, win, hop, wl["F"])
        eng.zero_grad()
        loss = eng.triplet_step(img, sounds, margin=1.0, loss_denominator=global_b)
        if world > 1:
            dist.all_reduce(eng.grads)
        eng.adam_step(1e-4, weight_decay=1e-6)
        h_loss.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the user reads the loss every step

    for _ in range(3):
        step_e2e()
    e2e_steps = max(3, args.steps // 2)
    e2e_ms = timed(step_e2e, e2e_steps) / e2e_steps
    e2e = {"value": global_b / (e2e_ms * 1e-3), "unit": "triplets/s",
           "h2d_bytes_per_step": int(h_img.numel() + h_wav.numel() * 2) * world, "d2h_bytes_per_step": 4 * world,
           "ms_per_step": e2e_ms}

    out = None
    if rank == 0:
        # ---- per-kernel profile (CUDA events on the launching stream), a few extra steps ------
        # branches serialised for this pass: with the two-stream overlap on, the events around a
        # small kernel also count the time it waits for SMs held by the other branch
        eng.set_overlap(False)
        step_resident(comm=False)
        lib.var_prof_begin()
        prof_steps = 3
        for _ in range(prof_steps):
            step_resident(comm=False)  # rank 0 only: no collective inside the profiled steps
        prof = vb._lib.prof_end()
        eng.set_overlap(True)
        step_kernel_ms = sum(v[0] for v in prof.values()) / prof_steps
        hbm, bf16_burst, bf16_sus, src = measured_peaks()
        tf32_peak = measure_tf32_peak(dev)
        kernels = {}
        for tag, (ms, fl, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            kernels[tag] = {"ms_per_step": round(ms / prof_steps, 4), "launches_per_step": cnt / prof_steps,
                            "share": round(ms / prof_steps / step_kernel_ms, 4)}
            if fl > 0:
                kernels[tag]["tflops"] = round(fl / (ms * 1e-3) / 1e12, 2)
        top = max(prof.items(), key=lambda kv: kv[1][0])
        ttag, (tms, tfl, tcnt) = top
        clips_per_step = 2 * local_b
        mfcc_bytes = clips_per_step * (wl["clip"] * 2 + wl["F"] * 160)  # int16 in + [F,40] f32 out
        if tfl > 0:
            roof = {"kernel": ttag, "bound": "tensor", "achieved": tfl / (tms * 1e-3) / 1e12, "peak": tf32_peak,
                    "unit": "TFLOP/s", "peak_source": "cuBLAS tf32 8192^3 measured in this run "
                    f"(MEASURED_PEAKS bf16 burst {bf16_burst} TF/s, {src})", "traffic": None,
                    "flops_per_launch": tfl / tcnt, "avg_launch_ms": tms / tcnt}
        else:
            nbytes = mfcc_bytes * prof_steps if ttag == "mfcc" else None
            roof = {"kernel": ttag, "bound": "hbm", "achieved": (nbytes / (tms * 1e-3) / 1e9) if nbytes else None,
                    "peak": hbm, "unit": "GB/s", "peak_source": src, "traffic": None, "avg_launch_ms": tms / tcnt}
        roof["frac"] = (roof["achieved"] / roof["peak"]) if roof["achieved"] else None
        m = prof.get("mfcc")
        mfcc_roof = None
        if m:
            gbs = mfcc_bytes * prof_steps / (m[0] * 1e-3) / 1e9
            mfcc_roof = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                         "bytes_per_clip": wl["clip"] * 2 + wl["F"] * 160, "avg_launch_ms": m[0] / m[2],
                         "peak_source": src}
        # ---- reward queries/s (BASELINE configs[2]), device timed + end to end ---------------
        reward = None
        if not args.no_reward:
            reward = bench_reward(vb, eng, net, wl, dev, pool, al, arena)
        cpu = None if args.no_cpu_baseline else cpu_baseline(wl, net, sample_steps=1)
        if reward is not None and not args.no_cpu_baseline:
            reward["cpu_baseline"] = cpu_reward_baseline(wl, net)
        out = {
            "metric": "VAR train triplets/sec", "value": value, "unit": "triplets/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "tf32 tensor-core MMA, fp32 accumulate/state",
            "data": "synthetic (seeded uint8 frames + int16 16 kHz clips), random-init weights",
            "config": {"workload": args.workload, "description": wl["desc"], "net": net, "global_batch": global_b,
                       "per_gpu_batch": local_b, "clip_samples": wl["clip"], "stft": list(wl["stft"]),
                       "frames": wl["F"], "parallelism": f"dp{world}",
                       "l2": "no flush: one step streams >1 GB of activations (>> 126 MB L2) and draws fresh images/clips"},
            "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
            "clocks": clk, "roofline": roof, "mfcc_roofline": mfcc_roof, "kernels": kernels,
            "kernels_note": "per-family CUDA-event times of 3 extra steps with the image/sound stream overlap OFF "
                            f"(serial kernel sum {step_kernel_ms:.2f} ms/step vs {ms_per_step:.2f} ms/step measured "
                            "with overlap ON)",
            "model_tflops": value * FLOP_PER_TRIPLET[net] / 1e12, "tf32_peak_tflops": tf32_peak,
            "cpu_baseline": cpu, "reward": reward,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def bench_reward(vb, eng, net, wl, dev, pool, al, arena):
    """Batched reward query for N vectorised envs: device-timed and end-to-end (uint8 frames +
    goal MFCC from pinned host memory in, [N] rewards + embeddings out)."""
    res = {}
    F = wl["F"]
    for N in (16, 128, 1024):
        img = pool[:N].contiguous()
        snd = torch.randn(N, F, 40, device=dev) * 4
        env_r = torch.zeros(N, device=dev)
        cached = eng.reward(img, goal_sounds=snd, env_reward=env_r)[1]
        use_cache = net == "ithor"   # iTHOR re-sends inf (cached goal embedding); Kuka re-encodes every step

        def q():
            if use_cache:
                return eng.reward(img, goal_feat_cached=cached, env_reward=env_r)
            return eng.reward(img, goal_sounds=snd, env_reward=env_r)
        for _ in range(5):
            q()
        torch.cuda.synchronize()
        iters = 30
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            q()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        h_img = pool[:N].cpu().pin_memory()
        h_snd = snd.cpu().pin_memory()

        def q_e2e():
            i = h_img.to(dev, non_blocking=True)
            if use_cache:
                r = eng.reward(i, goal_feat_cached=cached, env_reward=env_r)
            else:
                r = eng.reward(i, goal_sounds=h_snd.to(dev, non_blocking=True), env_reward=env_r)
            return torch.cat([r[0], r[1], r[2][:, None], r[3][:, None]], 1).cpu()
        for _ in range(3):
            q_e2e()
        t0 = time.perf_counter()
        for _ in range(iters):
            q_e2e()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / iters
        res[str(N)] = {"queries_per_s": N / (ms * 1e-3), "ms": ms, "e2e_queries_per_s": N / (e2e_ms * 1e-3),
                       "e2e_ms": e2e_ms, "goal_sound": "cached" if use_cache else "re-encoded"}
    return res


# ----------------------------------------------------------------------------- CPU arm
def cpu_step_fn(wl, net, B):
    """The oracle port of one training step on the host cores: numpy MFCC of 2B clips, fp32
    torch-CPU encoders forward/backward, triplet loss, Adam."""
    from oracle import mfcc as omfcc, model as omodel, synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.clone().requires_grad_(True) for k, v in omodel.init_state_dict(net, 0).items()}
    o = omodel.OracleVAR(net, sd)
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4, weight_decay=1e-6)
    clips = synth.make_clips(1, 2 * B, wl["clip"])
    images = torch.from_numpy(synth.make_images(2, B).astype(np.float32) / np.float32(255))
    n_fft, win, hop = wl["stft"]

    def step():
        feats = np.stack([omfcc.process_sound_feat(omfcc.mfcc_torchaudio(c, 16000, n_fft, win, hop),
                                                   (1, wl["F"], 40)) for c in clips]).astype(np.float32)
        s = torch.from_numpy(feats)
        opt.zero_grad()
        d = o(images, s[:B], s[B:])
        loss = omodel.triplet_margin_loss(d["image_feat"], d["sound_feat_positive"], d["sound_feat_negative"])
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def cpu_reward_baseline(wl, net, N=16, iters=3):
    """Oracle port of getEmbeddings + calcReward (vec_pretext_normalize.py:82-101) for N envs on the host:
    uint8 frames -> /255 -> image branch (+ goal-sound branch for Kuka, cached for iTHOR) -> dot + envReward.
    Timed with one torch thread (RL.py:75 pins the reference to one) and with all host threads."""
    from oracle import model as omodel, reward as oreward, synth
    sd = omodel.init_state_dict(net, 0)
    o = omodel.OracleVAR(net, sd)
    img_u8 = synth.make_images(3, N)
    snd = torch.randn(N, 1, wl["F"], 40) * 4
    inf = torch.full((N, 1, wl["F"], 40), float("inf"))
    env_r = np.zeros(N, np.float32)
    out = {}

    def query(first):
        with torch.no_grad():
            image = torch.from_numpy(img_u8.astype(np.float64) / 255.0).float()
            d = o(image, snd if (first or net == "kuka") else inf, None)
            return oreward.calc_reward(env_r, d["image_feat"].numpy(), d["sound_feat_positive"].numpy())[0]
    for threads in (1, os.cpu_count() or 1):
        torch.set_num_threads(threads)
        query(True)
        t0 = time.perf_counter()
        for _ in range(iters):
            query(False)
        dt = (time.perf_counter() - t0) / iters
        out[f"threads_{threads}"] = {"queries_per_s": N / dt, "ms": dt * 1e3}
    torch.set_num_threads(os.cpu_count() or 1)
    return {"kind": "port", "n_envs": N, "sample": f"{iters} queries of {N} envs, oracle port on the host", **out}


def cpu_baseline(wl, net, sample_steps=1):
    B = 16 if net == "ithor" else 64
    step = cpu_step_fn(wl, net, B)
    step()
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        step()
    dt = (time.perf_counter() - t0) / sample_steps
    return {"value": B / dt, "unit": "triplets/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{sample_steps} step(s) of batch {B} (same clip shape / net as the workload), oracle port: "
                      f"numpy MFCC + fp32 torch-CPU fwd/bwd + Adam, {torch.get_num_threads()} threads",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    net = wl["net"]
    B = 16 if net == "ithor" else 64
    step = cpu_step_fn(wl, net, B)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    v = B / dt
    sample = (f"each step = batch {B} of the workload's shape on the host cores (oracle port of the reference CPU "
              f"path: numpy MFCC + fp32 torch-CPU fwd/bwd + Adam); {steps} timed steps")
    print(json.dumps({
        "impl": "reference", "metric": "VAR train triplets/sec", "value": v, "unit": "triplets/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": args.workload, "description": wl["desc"], "net": net,
                                        "sample_batch": B},
        "cpu_baseline": {"value": v, "unit": "triplets/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
