"""CPU, world_size 2 over gloo: the host-side data-parallel logic of VAR_Pretext.train_epoch --
identical global index stream on every rank, rank slices that tile the batch, per-rank loss /
gradient scaling by the GLOBAL batch, one all-reduce(sum) of the flat gradient buffer, identical
weights afterwards.  The device engine is replaced by a tiny numpy stand-in with the same
interface (the kernels themselves need a GPU and are covered by the -m gpu tests)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeEngine:
    """Linear 'model' y = w . x with hinge-free loss sum(x @ w) / denom: gradient is data dependent,
    so wrong slicing or scaling shows up in the weights."""

    def __init__(self, n):
        self.params = torch.arange(n, dtype=torch.float32) / n
        self.grads = torch.zeros(n)
        self.steps = 0

    def zero_grad(self):
        self.grads.zero_()

    def triplet_step(self, img, snd, margin=1.0, loss_denominator=None):
        x = img.float().reshape(img.shape[0], -1)[:, : self.params.numel()]
        self.grads += x.sum(0) / loss_denominator
        return (x @ self.params).sum() / loss_denominator

    def adam_step(self, lr, weight_decay=0.0):
        self.params -= lr * self.grads
        self.steps += 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sampler as osampler
    # every rank draws the same global stream (device sampler semantics, restated by the oracle)
    gen = osampler.TorchCPUGenerator(977)
    n_items, B = 64, 16
    batches = osampler.epoch_batches(gen, n_items, B)
    data = torch.arange(n_items * 8, dtype=torch.float32).reshape(n_items, 8) % 7
    eng = FakeEngine(8)
    losses = []
    for batch in batches:
        lo, hi = (len(batch) * rank) // world, (len(batch) * (rank + 1)) // world
        idx = torch.tensor(batch[lo:hi])
        eng.zero_grad()
        loss = eng.triplet_step(data[idx], None, loss_denominator=len(batch))
        dist.all_reduce(eng.grads)
        dist.all_reduce(loss)
        eng.adam_step(0.01)
        losses.append(float(loss))
    torch.save({"params": eng.params, "losses": losses, "batches": batches}, out + f".{rank}")
    dist.destroy_process_group()


def test_dp_slices_sum_to_single_process_run(tmp_path):
    world = 2
    out = str(tmp_path / "res")
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r = [torch.load(out + f".{i}") for i in range(world)]
    assert r[0]["batches"] == r[1]["batches"]                      # identical index stream
    assert torch.equal(r[0]["params"], r[1]["params"])             # replicas stay in lock step
    # single-process reference over the same stream
    sys.path.insert(0, ROOT)
    data = torch.arange(64 * 8, dtype=torch.float32).reshape(64, 8) % 7
    eng = FakeEngine(8)
    for batch, l in zip(r[0]["batches"], r[0]["losses"]):
        eng.zero_grad()
        loss = eng.triplet_step(data[torch.tensor(batch)], None, loss_denominator=len(batch))
        eng.adam_step(0.01)
        assert abs(float(loss) - l) < 1e-4 * max(1.0, abs(l))
    assert torch.allclose(eng.params, r[0]["params"], atol=1e-5)


def test_rank_slices_tile_every_batch_size():
    for B in (1, 7, 64, 255, 8192):
        for world in (1, 2, 3, 4, 8):
            cuts = [((B * r) // world, (B * (r + 1)) // world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))


def _trainer_worker(rank, world, port, out):
    """The REAL VAR_Pretext.train_epoch (host-tuple path) over gloo with a ragged tail batch that leaves
    rank 0 without a triplet: every rank must still reach both all-reduces and the optimiser step."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    tr = import_module("voicecontrolledrobot-var_b200.VAR.pretext_VAR")
    t = object.__new__(tr.VAR_Pretext)  # the constructor insists on a CUDA device; train_epoch's host logic does not
    t.device = torch.device("cpu")
    cfg = type("C", (), {})()
    cfg.sound_dim = (1, 2, 40); cfg.tripletMargin = 1.0; cfg.pretextAdamL2 = 0.0
    t.config = cfg
    n, bs = 17, 8   # 17 % 8 == 1 < world
    images = (torch.arange(n * 8, dtype=torch.float32).reshape(n, 8) % 5)
    batches = [(images[s:s + bs], torch.zeros(min(bs, n - s), 1, 2, 40), torch.zeros(min(bs, n - s), 1, 2, 40), None)
               for s in range(0, n, bs)]
    eng = FakeEngine(8)
    seen = []
    losses = t.train_epoch(eng, batches, 0.01, world, rank, on_step=lambda l: seen.append(float(l)))
    assert len(losses) == 3 and seen == [float(l) for l in losses]
    assert len(t.train_epoch(FakeEngine(8), batches, 0.01, world, rank, max_steps=2)) == 2
    torch.save({"params": eng.params, "losses": [float(l) for l in losses], "steps": eng.steps}, out + f".{rank}")
    dist.destroy_process_group()


def test_train_epoch_ragged_tail_smaller_than_world(tmp_path):
    world = 2
    out = str(tmp_path / "res")
    port = 31000 + os.getpid() % 2000
    mp.spawn(_trainer_worker, args=(world, port, out), nprocs=world, join=True)
    r = [torch.load(out + f".{i}") for i in range(world)]
    assert r[0]["steps"] == r[1]["steps"] == 3
    assert torch.equal(r[0]["params"], r[1]["params"]) and r[0]["losses"] == r[1]["losses"]
    images = (torch.arange(17 * 8, dtype=torch.float32).reshape(17, 8) % 5)
    eng = FakeEngine(8)
    for s in range(0, 17, 8):
        eng.zero_grad()
        eng.triplet_step(images[s:s + 8], None, loss_denominator=min(8, 17 - s))
        eng.adam_step(0.01)
    assert torch.allclose(eng.params, r[0]["params"], atol=1e-5)


class GraphedFakeEngine(FakeEngine):
    """FakeEngine + the graphed-step entry point of VarEngine (zero_grad + triplet_step in one call)."""

    def __init__(self, n):
        super().__init__(n)
        self.use_step_graph = True
        self.graphed_calls = 0
        self.eager_calls = 0

    def triplet_step(self, img, snd, margin=1.0, loss_denominator=None):
        self.eager_calls += 1
        return super().triplet_step(img, snd, margin, loss_denominator)

    def triplet_step_graphed(self, img, snd, margin=1.0, loss_denominator=None):
        self.graphed_calls += 1
        self.zero_grad()
        return FakeEngine.triplet_step(self, img, snd, margin, loss_denominator)


def test_train_epoch_routes_single_gpu_steps_through_the_graphed_step():
    """One process (world == 1), batches in loader slots: every non-empty batch goes through
    engine.triplet_step_graphed (which owns the zero_grad), with the same weights as the individually launched steps;
    engines without it, a switched-off flag and reference-style host tuples keep the eager sequence."""
    sys.path.insert(0, ROOT)
    from importlib import import_module
    tr = import_module("voicecontrolledrobot-var_b200.VAR.pretext_VAR")
    t = object.__new__(tr.VAR_Pretext)
    t.device = torch.device("cpu")
    cfg = type("C", (), {})()
    cfg.sound_dim = (1, 2, 40); cfg.tripletMargin = 1.0; cfg.pretextAdamL2 = 0.0
    t.config = cfg
    n, bs = 20, 8
    images = (torch.arange(n * 8, dtype=torch.float32).reshape(n, 8) % 5)
    batches = [(images[s:s + bs], torch.zeros(min(bs, n - s), 1, 2, 40), torch.zeros(min(bs, n - s), 1, 2, 40), None)
               for s in range(0, n, bs)]
    def stream():  # what DeviceTripletLoader.stream() yields: (images, sounds, gt, global batch, record) in loader slots
        for image, sp, sn, _ in batches:
            yield image, torch.cat([sp, sn]), None, image.shape[0], None

    g, e = GraphedFakeEngine(8), FakeEngine(8)
    lg = t.train_epoch(g, stream(), 0.01)
    le = t.train_epoch(e, batches, 0.01)
    assert g.graphed_calls == 3 and g.eager_calls == 0 and g.steps == 3
    assert torch.equal(g.params, e.params) and [float(a) for a in lg] == [float(b) for b in le]
    off = GraphedFakeEngine(8)
    off.use_step_graph = False
    t.train_epoch(off, stream(), 0.01)
    assert off.graphed_calls == 0 and off.eager_calls == 3 and torch.equal(off.params, e.params)
    host = GraphedFakeEngine(8)  # reference-style host tuples: fresh tensors per batch, nothing to replay
    t.train_epoch(host, batches, 0.01)
    assert host.graphed_calls == 0 and host.eager_calls == 3 and torch.equal(host.params, e.params)
