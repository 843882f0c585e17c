"""GPU, 2 ranks over NCCL (skipped with fewer than 2 GPUs): the data-parallel step -- rank slices,
1/B_global scaling inside the fused tail kernel, one all-reduce of the flat gradient buffer,
identical Adam -- reproduces the single-GPU step on the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out, net="kuka", B=24):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", device_id=dev)
    import var_b200 as vb
    from importlib import import_module
    from oracle import model as omodel, synth
    images, sp, sn = synth.model_case(net, B, 5)
    kind, F = (vb.KUKA, 100) if net == "kuka" else (vb.ITHOR, 600)
    eng = vb.VarEngine(kind, F, 3, dev)
    eng.load_state_dict(omodel.init_state_dict(net, 3))
    lo, hi = (B * rank) // world, (B * (rank + 1)) // world
    img = torch.from_numpy(images[lo:hi]).to(dev)
    snd = torch.from_numpy(np.concatenate([sp[lo:hi], sn[lo:hi]])[:, 0]).to(dev).contiguous()
    for _ in range(2):  # twice: the bucket event of the second step must not reuse the first step's record
        eng.zero_grad()
        loss = eng.triplet_step(img, snd, margin=1.0, loss_denominator=B)
        eng.allreduce_grads()  # iTHOR: rnn.* range reduced on a side stream from the mid-backward event
        dist.all_reduce(loss)
    eng.adam_step(1e-4, weight_decay=1e-6)
    res = {"loss": float(loss), "grads": eng.grads.cpu(), "params": eng.params.cpu(), "bucketed": bool(eng._bucket)}
    # reward queries sharded by env index: no collective, gathered here only to compare
    shard_envs = import_module("voicecontrolledrobot-var_b200.VAR.RL_VAR").shard_envs
    N = 16
    gen = torch.Generator().manual_seed(7)
    imgs = torch.randint(0, 256, (N, 3, 96, 96), dtype=torch.uint8, generator=gen)
    snds = torch.randn(N, F, 40, generator=gen) * 4
    elo, ehi = shard_envs(N)
    part = eng.reward(imgs[elo:ehi].to(dev), goal_sounds=snds[elo:ehi].to(dev).contiguous())
    res["reward_part"] = [t.cpu() for t in part]
    res["reward_range"] = (elo, ehi)
    if rank == 0:  # the single-GPU step on the whole batch
        res["reward_full"] = [t.cpu() for t in eng.reward(imgs.to(dev), goal_sounds=snds.to(dev).contiguous())]
        ref = vb.VarEngine(kind, F, 3, dev)
        ref.load_state_dict(omodel.init_state_dict(net, 3))
        ref.zero_grad()
        l1 = ref.triplet_step(torch.from_numpy(images).to(dev),
                              torch.from_numpy(np.concatenate([sp, sn])[:, 0]).to(dev).contiguous(), margin=1.0)
        ref.adam_step(1e-4, weight_decay=1e-6)
        res.update(ref_loss=float(l1), ref_grads=ref.grads.cpu(), ref_params=ref.params.cpu())
    torch.save(res, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("net,B", [("kuka", 24), ("ithor", 6)])
def test_two_rank_step_equals_single_gpu_step(tmp_path, net, B):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the same rule runs on one device in test_gpu_sizes.py)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 1000, out, net, B), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(r0["params"], r1["params"]) and torch.equal(r0["grads"], r1["grads"])
    assert r0["bucketed"] == (net == "ithor")
    for i in range(4):  # gathered per-rank reward queries == the single-GPU query, bit for bit
        gathered = torch.cat([r0["reward_part"][i], r1["reward_part"][i]])
        assert r0["reward_range"] == (0, 8) and r1["reward_range"] == (8, 16)
        assert torch.equal(gathered, r0["reward_full"][i]), i
    assert abs(r0["loss"] - r0["ref_loss"]) < 1e-5 * max(1.0, abs(r0["ref_loss"]))
    g, gr = r0["grads"].numpy(), r0["ref_grads"].numpy()
    assert np.abs(g - gr).max() <= 1e-4 * np.abs(gr).max()  # fp32 summation order only
    assert np.abs(r0["params"].numpy() - r0["ref_params"].numpy()).max() < 2.1e-4
