"""ORACLE-side deterministic synthetic inputs.  Shared by the
golden generator, the tests and bench.py so that every party sees identical data."""
import numpy as np


def make_clip(rng, n_samples, fs=16000):
    """int16 mono clip: band-limited noise burst + 2-4 sinusoids with leading/trailing
    silence so the log(.+1e-6) floor is exercised."""
    t = np.arange(n_samples) / fs
    amp = rng.uniform(0.05, 0.9)
    x = np.zeros(n_samples)
    for _ in range(int(rng.integers(2, 5))):
        f = rng.uniform(80.0, 6000.0)
        x += rng.uniform(0.2, 1.0) * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
    noise = rng.standard_normal(n_samples)
    k = int(rng.integers(2, 16))
    noise = np.convolve(noise, np.ones(k) / k, mode="same")
    x += 0.5 * noise / (np.abs(noise).max() + 1e-9)
    x /= np.abs(x).max() + 1e-9
    lead = int(rng.integers(0, n_samples // 5))
    trail = int(rng.integers(0, n_samples // 5))
    x[:lead] = 0.0
    if trail:
        x[-trail:] = 0.0
    return np.round(x * amp * 32767.0).astype(np.int16)


def make_clips(seed, count, n_samples):
    rng = np.random.default_rng(seed)
    if np.isscalar(n_samples):
        return [make_clip(rng, int(n_samples)) for _ in range(count)]
    lo, hi = n_samples
    return [make_clip(rng, int(rng.integers(lo, hi + 1))) for _ in range(count)]


def make_images(seed, count):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(count, 3, 96, 96), dtype=np.uint8)


def make_labels(seed, count, task_num=4, mix=(50, 50, 50, 50, 100)):
    rng = np.random.default_rng(seed)
    p = np.asarray(mix, dtype=np.float64)
    return rng.choice(task_num + 1, size=count, p=p / p.sum()).astype(np.int64)


def model_case(net, B, seed):
    """Seeded inputs of the model golden vectors (oracle/make_golden.py::gold_model):
    images f32 in [0,1] from uint8, two MFCC-like sound batches [B,1,F,40] (iTHOR: frames
    >= 101 zero like processSoundFeat's padding)."""
    rng = np.random.default_rng(seed)
    images_u8 = make_images(seed, B)
    images = images_u8.astype(np.float32) / np.float32(255.0)
    F = 100 if net == "kuka" else 600

    def snd():
        s = (rng.standard_normal((B, 1, F, 40)) * np.array([20.0] + [4.0] * 39)).astype(np.float32)
        if net == "ithor":
            s[:, :, 101:, :] = 0.0
        return s
    sp = snd()
    sn = snd()
    return images, sp, sn


def ithor_reward_case(N, steps, seed=6, goal_every=3):
    """Seeded observation stream of the iTHOR reward golden: uint8 frames, occupancy maps, env rewards,
    dones; the goal sound is real on the reset observation and on every `goal_every`-th step and
    all-inf otherwise (Envs/ai2thor/RL_env_VAR.py:509-510 sends inf after an episode's first step)."""
    rng = np.random.default_rng(seed)
    obs_seq, rew_seq, done_seq = [], [], []
    for t in range(steps + 1):
        snd = np.zeros((N, 1, 600, 40), np.float32)
        if t % goal_every == 0:
            snd[:, :, :101] = (rng.standard_normal((N, 1, 101, 40)) * 5).astype(np.float32)
        else:
            snd[:] = np.inf
        obs_seq.append({"image": rng.integers(0, 256, (N, 3, 96, 96)).astype(np.uint8), "goal_sound": snd,
                        "occupancy": rng.integers(0, 256, (N, 1, 9, 9)).astype(np.uint8)})
        rew_seq.append(rng.standard_normal(N))
        done_seq.append(rng.random(N) < 0.3)
    return obs_seq, rew_seq, done_seq
