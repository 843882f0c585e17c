#!/usr/bin/env python
"""VAR hot-path benchmark: triplets/s of the full training step (+ reward queries/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ithor_b256|kuka_b64|kuka_dp8192|mfcc_4s]
    python bench.py --impl reference ...      # the CPU arm: oracle port on the host cores

Both timed numbers go through the repo's public API, the reference's own call chain
(VAR/pretext_VAR.py:19-26, :55-70):

    loadEnvData(config.pretextDataDir, config, ...)  ->  VAR_Pretext.train_epoch(engine, batches, lr)

on a synthetic triplet dataset written to disk in the reference's formats (pickled
{'image', 'ground_truth'} records; iTHOR: Fluent-Speech-Commands csv + wavs under
config.name == 'AI2ThorConfig', Kuka: GoogleCommand folders under 'ArmConfig').  One step = draw B
triplets (device mt19937, bit-exact index stream) -> MFCC of the 2B drawn clips -> both encoders
forward -> fused head / normalise / triplet-margin loss + gradient -> full backward -> [NCCL
all-reduce of the flat gradient buffer] -> fused Adam.

  value : frames + clips resident in HBM (config.pretextDataResident = True)
  e2e   : frames + clips in pinned host memory, gathered and uploaded EVERY step
          (pretextDataResident = False), and the step's loss read back to the host every step

Weights are random-init, data synthetic (oracle/synth.py generators, used for data only).  Prints ONE
JSON line (rank 0).
"""
import argparse
import json
import os
import pickle
import shutil
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "voicecontrolledrobot-var_b200"

WORKLOADS = {
    # name: net, per-GPU batch (None = global/N), clip samples, stft, F, scaling, dataset size
    "ithor_b256": dict(net="ithor", batch=256, clip=16000, stft=(512, 400, 160), F=600, scaling="weak", items=8192,
                       clips_per_list=1000,
                       desc="BASELINE configs[1]: iTHOR VAR pretext training (config.name == 'AI2ThorConfig': task-table "
                            "sampling, python_speech_features-flavoured MFCC), GoogleCommand-shaped 1 s 16 kHz clips "
                            "(512/400/160) padded to 600 frames, batch 256 per GPU"),
    "kuka_b64": dict(net="kuka", batch=64, clip=16000, stft=(512, 400, 160), F=100, scaling="weak", items=8192,
                     clips_per_list=1000,
                     desc="BASELINE configs[0]: Kuka VAR pretext training, 1 s clips, batch 64 per GPU"),
    "kuka_dp8192": dict(net="kuka", batch=None, global_batch=8192, clip=16000, stft=(512, 400, 160), F=100,
                        scaling="strong", items=32768, clips_per_list=1000,
                        desc="BASELINE configs[4]: Kuka data-parallel training, global batch 8192"),
    "mfcc_4s": dict(net="kuka", batch=1024, clip=64000, stft=(1024, 800, 640), F=100, scaling="weak", items=8192,
                    clips_per_list=250, dataset="NSynth",
                    desc="BASELINE configs[3]: NSynth-shaped 4 s clips (1024/800/640), batch 1024 per GPU"),
}
TASK_NUM = 4
FLOP_PER_TRIPLET = {"kuka": 75.5e6, "ithor": 9.95e9}       # SURVEY.md section 8(d), fwd + bwd
FLOP_PER_QUERY = {"kuka": 23.4e6, "ithor": 337.4e6}
# the reference's iTHOR tables (Envs/ai2thor/env_config.py:18-45, Envs/ai2thor/config.py:121-133)
ITHOR_ALL_TASKS = {"livingRoom": {"FloorLamp": ["ToggleObjectOn", "ToggleObjectOff"],
                                  "Television": ["ToggleObjectOn", "ToggleObjectOff"]}}
ITHOR_SYNONYM = {"livingRoom": ["none"], "FloorLamp": ["lights", "lamp"], "Television": ["music"],
                 "ToggleObjectOn": ["increase", "activate"], "ToggleObjectOff": ["decrease", "deactivate"]}
ITHOR_OBJ_ACT = {"lights": ["activate", "deactivate"], "music": ["activate", "deactivate"],
                 "lamp": ["activate", "deactivate"]}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ithor_b256", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reward", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--data-dir", default=os.environ.get("VAR_BENCH_DATA", "/tmp/var_b200_bench"))
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------- data on disk
class Cfg:
    pass


def make_config(name, wl, root):
    """The attributes of ArmConfig / AI2ThorConfig (+ EnvConfig) that the triplet path reads."""
    c = Cfg()
    net = wl["net"]
    c.img_dim = (3, 96, 96)
    c.sound_dim = (1, wl["F"], 40)
    c.representationDim = 3
    c.tripletMargin = 1.0
    c.pretextAdamL2, c.pretextLR, c.pretextLRStep = 1e-6, 1e-4, "step"
    c.pretextLRDecayEpoch, c.pretextLRDecayGamma = [20, 30], 0.2
    c.pretextDataDir = [os.path.join(root, "data")]
    c.pretextDataFileLoadNum = ["all"]
    c.pretextDataNumWorkers = 8
    c.pretextModelSaveDir = os.path.join(root, "model")
    c.pretextTrain, c.pretextCollection, c.pretextModelFineTune = True, False, False
    c.RLRewardSoundSound, c.realTimeVec, c.RLTrain = False, False, True
    c.commonMediaPath = os.path.join(root, "commonMedia")
    mod = __import__("importlib").import_module
    if net == "ithor":
        c.name = "AI2ThorConfig"
        c.envFolder = "ai2thor"
        c.allTasks, c.synonym = ITHOR_ALL_TASKS, ITHOR_SYNONYM
        c.taskNum = TASK_NUM
        c.soundSource = {"dataset": "FSC", "train_test": "train", "FSC_max_sound_dur": 6., "size": wl["clips_per_list"],
                         "FSC_obj_act": ITHOR_OBJ_ACT, "FSC_locations": ["none"], "FSC_csv": "train_data.csv"}
        c.pretextEnvSeed = 977
        c.pretextModel = mod(f"{PKG}.models.pretext.ai2thor_pretext_model").VARPretextNet
    else:
        ds = wl.get("dataset", "GoogleCommand")
        words = ["up", "down", "left", "right"]
        c.name = "ArmConfig"
        c.envFolder = os.path.join("pybullet", "arms")
        c.taskNum = TASK_NUM
        c.soundSource = {"dataset": [ds], "train_test": "train", "items": {ds: words},
                         "size": {ds: [wl["clips_per_list"]] * 4}, "max_sound_dur": {ds: 6.0}}
        c.pretextEnvSeed = 453
        c.pretextModel = mod(f"{PKG}.models.pretext.arm_pretext_model").VARPretextNet
    c.pretextDataset = mod(f"{PKG}.dataset").VARDataset
    return c


def write_dataset(root, wl, name, seed=4321):
    """Seeded synthetic dataset in the reference's on-disk formats.  Clip synthesis
    costs ~1 ms per second of audio on the host, so every clip list holds 64 distinct clips rolled to
    `clips_per_list` variants (distinct sample offsets keep every clip unique for the MFCC)."""
    from scipy.io import wavfile
    from oracle import synth
    marker = os.path.join(root, ".done")
    if os.path.exists(marker):
        return
    shutil.rmtree(root, ignore_errors=True)
    media = os.path.join(root, "commonMedia")
    n_clips = wl["clips_per_list"]

    def variants(list_seed):
        base = synth.make_clips(list_seed, 64, wl["clip"])
        return [np.roll(base[i % 64], 37 * (i // 64)) if i >= 64 else base[i] for i in range(n_clips)]

    if wl["net"] == "ithor":
        import pandas as pd
        os.makedirs(os.path.join(media, "FSC", "data"))
        os.makedirs(os.path.join(media, "FSC", "wavs"))
        rows = []
        for li, obj in enumerate(ITHOR_OBJ_ACT):
            for ai, act in enumerate(ITHOR_OBJ_ACT[obj]):
                for i, clip in enumerate(variants(seed + 10 * li + ai)):
                    rel = os.path.join("wavs", f"{obj}_{act}_{i}.wav")
                    wavfile.write(os.path.join(media, "FSC", rel), 16000, clip)
                    rows.append({"path": rel, "transcription": f"{act} the {obj}", "action": act, "object": obj,
                                 "location": "none"})
        pd.DataFrame(rows).to_csv(os.path.join(media, "FSC", "data", "train_data.csv"))
    else:
        ds = wl.get("dataset", "GoogleCommand")
        for c, wd in enumerate(["up", "down", "left", "right"]):
            d = os.path.join(media, ds, "train", wd)
            os.makedirs(d)
            for i, clip in enumerate(variants(seed + c)):
                wavfile.write(os.path.join(d, f"{i:04d}.wav"), 16000, clip)
    data = os.path.join(root, "data", "train")
    os.makedirs(data)
    rng = np.random.default_rng(seed)
    gts = synth.make_labels(seed + 99, wl["items"])
    per_file = 1024
    for f in range(0, wl["items"], per_file):
        imgs = rng.integers(0, 256, size=(min(per_file, wl["items"] - f), 3, 96, 96), dtype=np.uint8)
        items = [{"image": imgs[i], "ground_truth": int(gts[f + i])} for i in range(len(imgs))]
        with open(os.path.join(data, f"data_{f // per_file:03d}.pickle"), "wb") as fh:
            pickle.dump(items, fh, protocol=4)
    open(marker, "w").write("ok")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        if self.idx is None:
            return
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


def ncu_traffic(workload, family):
    """DRAM bytes per launch of a kernel family from the committed `ncu --set full` capture of the same
    workload (profiles/*_traffic.json, written by scripts/summarize_ncu.py); None when not captured."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            d = json.load(open(path))
        except (OSError, ValueError):
            continue
        fam = d.get(workload, {}).get(family)
        if fam:
            return {"dram_bytes_per_launch": fam["dram_bytes_per_launch"], "algorithmic_bytes_per_launch":
                    fam.get("algorithmic_bytes_per_launch"), "source": os.path.basename(path)}
    return None


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    from importlib import import_module
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))
    import var_b200 as vb
    ds = import_module(f"{PKG}.dataset")
    tr = import_module(f"{PKG}.VAR.pretext_VAR")
    lib = vb._lib.lib
    from oracle import model as omodel  # initial weights only (seeded, reference layout)

    wl = WORKLOADS[args.workload]
    net = wl["net"]
    global_b = wl["batch"] * world if wl["batch"] else wl["global_batch"]
    local_b = global_b // world
    root = os.path.join(args.data_dir, args.workload)
    if rank == 0:
        write_dataset(root, wl, args.workload)
    if world > 1:
        dist.barrier()
    cfg = make_config(args.workload, wl, root)
    cfg.pretextTrainBatchSize = global_b
    trainer = tr.VAR_Pretext(cfg)
    torch.manual_seed(cfg.pretextEnvSeed)
    torch.cuda.manual_seed_all(cfg.pretextEnvSeed)
    trainer.pretextModel = cfg.pretextModel(cfg).to(dev)
    trainer.pretextModel.load_state_dict(omodel.init_state_dict(net, 0))
    trainer.pretextModel.train()
    eng = trainer.pretextModel._get_engine(dev)
    eng.reset_optimizer()
    lr = cfg.pretextLR

    def build_loader(resident):
        cfg.pretextDataResident = resident
        loader, _ = ds.loadEnvData(data_dir=cfg.pretextDataDir, config=cfg, batch_size=global_b, shuffle=True,
                                   num_workers=cfg.pretextDataNumWorkers, drop_last=True,
                                   loadNum=cfg.pretextDataFileLoadNum, dtype=cfg.pretextDataset)
        loader.rank, loader.world_size = rank, world
        return loader

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- value: dataset resident in HBM ----------------------------------------------------
    loader = build_loader(True)
    batches = loader.stream()
    warm = max(args.warmup, 3)
    # clocks are sampled on rank 0's GPU only, and the sampler is started BEFORE the warm-up steps: the
    # start-up of nvidia-smi (NVML initialisation takes the driver lock) must not fall into the timed region
    clocks = ClockSampler(local if rank == 0 else None)
    clocks.start()
    trainer.train_epoch(eng, batches, lr, world, rank, max_steps=warm)
    l0 = lib.var_launch_count()
    total_ms = timed(lambda: trainer.train_epoch(eng, batches, lr, world, rank, max_steps=args.steps))
    launches = lib.var_launch_count() - l0
    clk = clocks.stop()
    ms_per_step = total_ms / args.steps
    value = global_b / (ms_per_step * 1e-3)

    # ---- end to end: pinned host dataset, gather + H2D every step, loss D2H every step ------
    # (diagnostics only, never set for a reported line: VAR_E2E_RESIDENT=1 keeps the dataset resident, VAR_E2E_NOSYNC=1
    # drops the per-step synchronisation -- they separate the cost of the host gather / upload from that of the read-back)
    sloader = build_loader(os.environ.get("VAR_E2E_RESIDENT") == "1")
    sbatches = sloader.stream()
    h_loss = torch.zeros(1).pin_memory()
    e2e_nosync = os.environ.get("VAR_E2E_NOSYNC") == "1"

    def read_loss(loss):
        h_loss.copy_(loss.reshape(1), non_blocking=True)
        if not e2e_nosync:
            torch.cuda.current_stream().synchronize()  # the user reads the loss every step

    trainer.train_epoch(eng, sbatches, lr, world, rank, max_steps=3, on_step=read_loss)
    e2e_steps = max(3, args.steps // 2)
    b0 = getattr(sloader, "h2d_bytes", 0)
    e2e_ms = timed(lambda: trainer.train_epoch(eng, sbatches, lr, world, rank, max_steps=e2e_steps,
                                               on_step=read_loss)) / e2e_steps
    e2e = {"value": global_b / (e2e_ms * 1e-3), "unit": "triplets/s",
           "h2d_bytes_per_step": int((getattr(sloader, "h2d_bytes", 0) - b0) / e2e_steps) * world, "d2h_bytes_per_step": 4 * world,
           "ms_per_step": e2e_ms,
           "api": "loadEnvData(config.pretextDataResident=False) -> VAR_Pretext.train_epoch(on_step=read loss): per step "
                  "the drawn uint8 frames and int16 clips are gathered from pinned host memory and uploaded, the loss is "
                  "copied back and the stream synchronised"}
    # the same loop with the loss read ONE STEP BEHIND (async D2H into a pinned ring, wait for the previous step's copy):
    # every step's loss still reaches the host inside the timed region, but the launch of step i+1 is not held back by
    # the read-back of step i -- what a training script that logs the loss would do; reported beside the strict figure
    ring = [(torch.zeros(1).pin_memory(), torch.cuda.Event()) for _ in range(2)]
    pend, cnt = [], [0]

    def read_loss_lagged(loss):
        buf, ev = ring[cnt[0] & 1]
        cnt[0] += 1
        buf.copy_(loss.reshape(1), non_blocking=True)
        ev.record()
        pend.append(ev)
        if len(pend) > 1:
            pend.pop(0).synchronize()

    def lagged_epoch():
        trainer.train_epoch(eng, sbatches, lr, world, rank, max_steps=e2e_steps, on_step=read_loss_lagged)
        while pend:
            pend.pop(0).synchronize()

    lag_ms = timed(lagged_epoch) / e2e_steps
    if getattr(sloader, "producer_stats", None):
        ps = sloader.producer_stats
        e2e["producer_ms_per_batch"] = {k: round(1e3 * v / max(1, ps["batches"]), 3) for k, v in ps.items() if k != "batches"}
    e2e["pipelined_read"] = {"ms_per_step": lag_ms, "value": global_b / (lag_ms * 1e-3),
                             "note": "same loop, loss of step i read after step i+1 was launched (every loss still read in the timed region)"}
    sbatches.close()

    # ---- reward queries/s (BASELINE configs[2]): envs sharded over the ranks, no collective ---
    reward = None
    if not args.no_reward:
        reward = bench_reward(vb, trainer, cfg, net, wl, dev, rank, world, dist)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None

    # ---- rank 0 only, outside any process group: per-kernel profile, peaks, baselines ---------
    # branches serialised for this pass: with the two-stream overlap on, the events around a small
    # kernel also count the time it waits for SMs held by the other branch
    eng.set_overlap(False)
    step_graph_was = getattr(eng, "use_step_graph", False)
    eng.use_step_graph = False  # per-launch events need the individual launches, not a graph replay
    trainer.train_epoch(eng, batches, lr, 1, 0, max_steps=1)
    lib.var_prof_begin()
    prof_steps = 3
    trainer.train_epoch(eng, batches, lr, 1, 0, max_steps=prof_steps)  # rank 0's slice, no collective
    torch.cuda.synchronize()
    prof = vb._lib.prof_end()
    eng.set_overlap(True)
    eng.use_step_graph = step_graph_was
    step_kernel_ms = sum(v[0] for v in prof.values()) / prof_steps
    hbm, bf16_burst, bf16_sus, src = measured_peaks()
    tf32_peak = measure_tf32_peak(dev)
    h16_flags = lib.var_h16_flags()
    # tensor peak a family is compared with: the kind::f16 kernels against MEASURED_PEAKS.json's bf16 figure (the
    # sustained one: they are timed inside a long step), the tf32 kernels against cuBLAS tf32 measured in this run
    f16_fams = {"gemm_fwd16", "gemm_dgrad16", "wgrad16"} | ({"gru_step", "gru_bwd"} if (h16_flags & 2 and net == "ithor") else set())
    clips_per_step = 2 * local_b
    bytes_per_clip = wl["clip"] * 2 + wl["F"] * 160  # int16 in + [F, 40] f32 out
    mfcc_bytes = clips_per_step * bytes_per_clip
    kernels, rooflines = {}, {}
    for tag, (ms, fl, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        kernels[tag] = {"ms_per_step": round(ms / prof_steps, 4), "launches_per_step": cnt / prof_steps,
                        "share": round(ms / prof_steps / step_kernel_ms, 4)}
        traffic = ncu_traffic(args.workload, tag)
        if fl > 0:
            tfs = fl / (ms * 1e-3) / 1e12
            kernels[tag]["tflops"] = round(tfs, 2)
            peak, psrc = ((bf16_sus, f"MEASURED_PEAKS.json bf16 sustained ({src}); f16 operands, kind::f16") if tag in f16_fams
                          else (tf32_peak, "cuBLAS tf32 8192^3 measured in this run (no tf32 entry in MEASURED_PEAKS.json)"))
            rooflines[tag] = {"kernel": tag, "bound": "tensor", "achieved": tfs, "peak": peak, "unit": "TFLOP/s",
                              "frac": tfs / peak, "peak_source": psrc, "flops_per_launch": fl / cnt,
                              "avg_launch_ms": ms / cnt, "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                              "traffic_detail": traffic}
            if tag in ("gru_step", "gru_bwd"):
                rooflines[tag]["note"] = ("sequential recurrence: one persistent launch runs 73 (72) dependent time steps, each bounded by a "
                                          "cross-CTA release/acquire round trip + the gate maths (>= 9 us per step measured, DESIGN section 7), "
                                          "not by the tensor pipe")
    ttag = max(prof.items(), key=lambda kv: kv[1][0])[0]
    if ttag in rooflines:
        roof = rooflines[ttag]
    else:
        tms, _, tcnt = prof[ttag]
        nbytes = mfcc_bytes * prof_steps if ttag == "mfcc" else None
        tr = ncu_traffic(args.workload, ttag)
        roof = {"kernel": ttag, "bound": "hbm", "achieved": (nbytes / (tms * 1e-3) / 1e9) if nbytes else None,
                "peak": hbm, "unit": "GB/s", "peak_source": src, "traffic": tr["dram_bytes_per_launch"] if tr else None,
                "traffic_detail": tr, "avg_launch_ms": tms / tcnt}
        roof["frac"] = (roof["achieved"] / roof["peak"]) if roof["achieved"] else None
    # the MFCC front-end timed ALONE (inside a step it runs on the loader's side stream under the previous step's
    # kernels, so its in-step events measure contention, not the kernel)
    mfcc_roof = mfcc_alone(ds, loader, wl, clips_per_step, bytes_per_clip, hbm, src, args.workload)
    cpu = torch_gpu = None
    if world == 1:
        if not args.no_torch_baseline and net in ("ithor", "kuka") and args.workload in ("ithor_b256", "kuka_b64"):
            torch_gpu = torch_gpu_baseline(net, local_b)
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(wl, net, 16 if net == "ithor" else 64, 1)
            if reward is not None:
                reward["cpu_baseline"] = cpu_reward_baseline(wl, net)
    out = {
        "metric": "VAR train triplets/sec", "value": value, "unit": "triplets/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": ("f16 (sound convs, GRU recurrences) + tf32 tensor-core MMA operands, fp32 accumulate / state / master weights"
                  if net == "ithor" and (h16_flags & 1) else "tf32 tensor-core MMA, fp32 accumulate/state"),
        "data": "synthetic (seeded uint8 frames + int16 16 kHz clips written to disk in the reference's formats), "
                "random-init weights",
        "config": {"workload": args.workload, "description": wl["desc"], "net": net, "config_name": cfg.name,
                   "global_batch": global_b, "per_gpu_batch": local_b, "clip_samples": wl["clip"],
                   "stft": list(wl["stft"]), "frames": wl["F"], "dataset_items": wl["items"],
                   "mfcc_flavour": "python_speech_features" if net == "ithor" else "torchaudio",
                   "api": "loadEnvData -> VAR_Pretext.train_epoch", "parallelism": f"dp{world}",
                   "l2": "no flush: one step streams >1 GB of activations (>> 126 MB L2) and draws fresh images/clips"},
        "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
        "clocks": clk, "roofline": roof, "rooflines": rooflines, "mfcc_roofline": mfcc_roof, "kernels": kernels,
        "kernels_note": "per-family CUDA-event times of 3 extra steps with the image/sound stream overlap OFF "
                        f"(serial kernel sum {step_kernel_ms:.2f} ms/step vs {ms_per_step:.2f} ms/step measured "
                        "with overlap ON); at N > 1 taken on rank 0 after the process group is destroyed",
        "model_tflops": value * FLOP_PER_TRIPLET[net] / 1e12, "tf32_peak_tflops": tf32_peak,
        "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu, "reward": reward,
    }
    return json.dumps(out)


def mfcc_alone(ds, loader, wl, clips, bytes_per_clip, hbm, src, workload):
    """The fused MFCC kernel on one step's worth of drawn clips, nothing else running: CUDA events over 20 launches."""
    al = __import__("importlib").import_module(f"{PKG}.Envs.audioLoader")
    arena, dev = loader.arena, loader.device
    n = int(arena.clip_off.numel())
    idx = torch.randint(0, n, (clips,), device=dev)
    off, ln = arena.clip_off[idx].contiguous(), arena.clip_len[idx].contiguous()
    n_fft, win, hop = loader.audio.stft_params(loader.param)
    out = torch.empty(clips, wl["F"], 40, device=dev)
    run = lambda: al.mfcc_device(arena.wav, off, ln, loader.audio.fs, n_fft, win, hop, wl["F"], flavour=loader.flavour, out=out)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = clips * bytes_per_clip / (ms * 1e-3) / 1e9
    tr = ncu_traffic(workload, "mfcc")
    return {"bound": "hbm (target) / fp32 issue (actual, DESIGN.md section 7)", "achieved": gbs, "peak": hbm, "unit": "GB/s",
            "frac": gbs / hbm, "bytes_per_clip": bytes_per_clip, "clips": clips, "avg_launch_ms": ms, "peak_source": src,
            "timed": "alone, 20 launches, CUDA events", "traffic": tr["dram_bytes_per_launch"] if tr else None}


def bench_reward(vb, trainer, cfg, net, wl, dev, rank, world, dist):
    """Batched reward query (Envs/vec_env/vec_pretext_normalize.py:82-101) for N vectorised envs sharded by
    env index over the ranks (rank r answers envs [r*N/G, (r+1)*N/G), no collective).  `queries_per_s`:
    captured-graph query with resident inputs, CUDA events, max over ranks.  `e2e_*`: the wrapper's own
    step_wait() over a stub venv that hands out numpy observations (H2D, query, D2H, host return
    normalisation), wall clock, max over ranks."""
    import types
    from importlib import import_module
    vpn = import_module(f"{PKG}.Envs.vec_env.vec_pretext_normalize")
    shard_envs = import_module(f"{PKG}.VAR.RL_VAR").shard_envs
    model = trainer.pretextModel
    model.eval()
    eng = model._get_engine(dev)
    F = wl["F"]
    fresh_every = 50 if net == "ithor" else 1   # iTHOR sends inf after an episode's first step (RLEnvMaxSteps 50)
    extra = "occupancy" if net == "ithor" else "robot_pose"
    res = {"sharding": f"envs split by index over {world} rank(s), no collective",
           "e2e": "VecPretextNormalize.step_wait(): numpy observations -> H2D -> captured-graph query -> one D2H -> the reference's "
                  "host post-processing (float64 image / 255, return normalisation); e2e_device_protocol: step_wait_device()",
           "goal_sound": "cached except every 50th step" if net == "ithor" else "re-encoded every step"}

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for N in (16, 128, 1024):
        lo, hi = shard_envs(N, rank, world)
        n = hi - lo
        if n == 0:
            continue
        rng = np.random.default_rng(100 + rank)
        g_f = eng.reward_graph(n, torch.uint8, fresh_goal=True)
        g_c = eng.reward_graph(n, torch.uint8, fresh_goal=False)
        for g in (g_f, g_c):
            g.images.copy_(torch.from_numpy(rng.integers(0, 256, (n, 3, 96, 96), dtype=np.uint8)))
        g_f.goal_sounds.copy_(torch.randn(n, F, 40) * 4)
        g_f.launch()
        g_c.goal_feat_cached.copy_(g_f.goal_feat)
        iters = 100
        sched = [(g_f if (i % fresh_every == 0) else g_c) for i in range(iters)]
        for g in sched[:5]:
            g.launch()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for g in sched:
            g.launch()
        e1.record(); torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1) / iters)
        # end to end through the wrapper
        obs_real = {"image": rng.integers(0, 256, (n, 3, 96, 96)).astype(np.uint8),
                    "goal_sound": (rng.standard_normal((n, 1, F, 40)) * 4).astype(np.float32),
                    extra: np.zeros((n, 1, 9, 9) if net == "ithor" else (n, 4), np.float32)}
        obs_inf = dict(obs_real, goal_sound=np.full((n, 1, F, 40), np.inf, np.float32))

        class Venv:
            num_envs = n
            observation_space = types.SimpleNamespace(shape=(1,))
            action_space = None
            t = 0

            def reset(self):
                return obs_real

            def step_wait(self):
                self.t += 1
                o = obs_real if self.t % fresh_every == 0 else obs_inf
                return o, np.zeros(n), np.zeros(n, bool), ({},) * n

        w = vpn.VecPretextNormalize(Venv(), ob=False, ret=True, gamma=0.99, config=cfg,
                                    pretextObj=types.SimpleNamespace(pretextModel=model))
        w.reset()
        for _ in range(3):
            w.step_wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            w.step_wait()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / iters)
        # the device-resident protocol (step_wait_device: numpy observations in, reward / embeddings / policy inputs
        # stay on the device, return normalisation on the device; one D2H of the [N] rewards to close the step)
        wd = vpn.VecPretextNormalize(Venv(), ob=False, ret=True, gamma=0.99, config=cfg,
                                     pretextObj=types.SimpleNamespace(pretextModel=model))
        wd.reset()
        for _ in range(3):
            wd.step_wait_device()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            _, r_dev, _, _ = wd.step_wait_device()
            r_dev.cpu()
        dev_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / iters)
        res[str(N)] = {"queries_per_s": N / (ms * 1e-3), "ms": ms, "envs_per_rank": n,
                       "e2e_queries_per_s": N / (e2e_ms * 1e-3), "e2e_ms": e2e_ms,
                       "e2e_device_protocol_queries_per_s": N / (dev_ms * 1e-3), "e2e_device_protocol_ms": dev_ms,
                       "tflops": N / (ms * 1e-3) * FLOP_PER_QUERY[net] / 1e12}
    model.train()
    return res


def torch_gpu_baseline(net, B):
    """The same training step as eager PyTorch on this GPU (cuDNN convs + cuDNN GRU + torchaudio MFCC +
    torch.optim.Adam; scripts/torch_gpu_baseline.py -- none of this repo's kernels): the same-box bar
    SURVEY 8(d) / BASELINE.md name, with TF32 allowed (cuDNN's default) and in strict fp32."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("torch_gpu_baseline", os.path.join(ROOT, "scripts", "torch_gpu_baseline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    runs = {}
    try:
        for tf32 in (True, False):
            r = mod.run(net, B, 10, tf32)
            runs["tf32_allowed" if tf32 else "fp32"] = {"ms_per_step": r["ms_per_step"], "triplets_per_s": r["triplets_per_s"]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
        torch.cuda.empty_cache()
    return {"what": "eager PyTorch port of the reference step (cuDNN conv + cuDNN GRU + torchaudio MFCC + Adam), inputs "
                    "resident on the device, same batch / shapes, 10 timed steps after 5 warm-ups", "batch": B,
            "torch": torch.__version__, **runs}


# ----------------------------------------------------------------------------- CPU arm
def cpu_step_fn(wl, net, B):
    """The oracle port of one training step on the host cores: numpy MFCC of 2B clips (the flavour the
    workload's config selects), fp32 torch-CPU encoders forward/backward, triplet loss, Adam."""
    from oracle import mfcc as omfcc, model as omodel, synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.clone().requires_grad_(True) for k, v in omodel.init_state_dict(net, 0).items()}
    o = omodel.OracleVAR(net, sd)
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4, weight_decay=1e-6)
    clips = synth.make_clips(1, 2 * B, wl["clip"])
    images = torch.from_numpy(synth.make_images(2, B).astype(np.float32) / np.float32(255))
    n_fft, win, hop = wl["stft"]

    def feat(c):
        if net == "ithor":  # getAudioFromTask leaves mfcc_from=None -> python_speech_features (audioLoader.py:159-161)
            return omfcc.mfcc_psf(c, 16000, win / 16000.0, hop / 16000.0, 40, 40, n_fft)
        return omfcc.mfcc_torchaudio(c, 16000, n_fft, win, hop)

    def step():
        feats = np.stack([omfcc.process_sound_feat(feat(c), (1, wl["F"], 40)) for c in clips]).astype(np.float32)
        s = torch.from_numpy(feats)
        opt.zero_grad()
        chunk = 64 if net == "ithor" else B
        total = 0.0
        for lo in range(0, B, chunk):  # gradient accumulation in chunks bounds the host memory at batch 256
            d = o(images[lo:lo + chunk], s[:B][lo:lo + chunk], s[B:][lo:lo + chunk])
            hinge = omodel.triplet_margin_loss(d["image_feat"], d["sound_feat_positive"], d["sound_feat_negative"])
            part = hinge * (min(chunk, B - lo) / B)
            part.backward()
            total += float(part.detach())
        opt.step()
        return total
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


def cpu_baseline(wl, net, B, sample_steps=1):
    step = cpu_step_fn(wl, net, B)
    step()
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        step()
    dt = (time.perf_counter() - t0) / sample_steps
    return {"value": B / dt, "unit": "triplets/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{sample_steps} step(s) of batch {B} (same clip shape / net / MFCC flavour as the workload), oracle "
                      f"port: numpy MFCC + fp32 torch-CPU fwd/bwd + Adam, {torch.get_num_threads()} threads",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    """The reference arm: the reference's own CPU path (its oracle port: the reference is pure Python and
    does not travel to the GPU box) on the workload's FULL per-GPU batch, all host threads."""
    rank, world, _ = dist_env()
    if rank != 0:
        return None
    wl = WORKLOADS[args.workload]
    net = wl["net"]
    B = wl["batch"] if wl["batch"] else min(wl["global_batch"], 1024)
    step = cpu_step_fn(wl, net, B)
    warm = max(1, min(args.warmup, 5))
    for _ in range(warm):
        step()
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    v = B / dt
    sample = (f"each step = the workload's batch ({B} triplets: {2 * B} clips through the numpy MFCC, fp32 torch-CPU "
              f"fwd/bwd, Adam) on {os.cpu_count()} host threads; {steps} timed steps after {warm} warm-up")
    return json.dumps({
        "impl": "reference", "metric": "VAR train triplets/sec", "value": v, "unit": "triplets/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": args.workload, "description": wl["desc"], "net": net,
                                        "global_batch": B, "per_gpu_batch": B},
        "cpu_baseline": {"value": v, "unit": "triplets/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


if __name__ == "__main__":
    import contextlib
    a = parse()
    _stdout = sys.stdout
    # the reference-shaped API prints progress ("Sound Loaded", class histogram ...): keep stdout for the ONE JSON line
    with contextlib.redirect_stdout(sys.stderr):
        line = run_reference(a) if a.impl == "reference" else run_b200(a)
    if line is not None:
        _stdout.write(line + "\n")
        _stdout.flush()
