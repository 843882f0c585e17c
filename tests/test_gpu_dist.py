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


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", device_id=dev)
    import var_b200 as vb
    from oracle import model as omodel, synth
    B = 24
    images, sp, sn = synth.model_case("kuka", B, 5)
    eng = vb.VarEngine(vb.KUKA, 100, 3, dev)
    eng.load_state_dict(omodel.init_state_dict("kuka", 3))
    lo, hi = (B * rank) // world, (B * (rank + 1)) // world
    img = torch.from_numpy(images[lo:hi]).to(dev)
    snd = torch.from_numpy(np.concatenate([sp[lo:hi], sn[lo:hi]])[:, 0]).to(dev).contiguous()
    eng.zero_grad()
    loss = eng.triplet_step(img, snd, margin=1.0, loss_denominator=B)
    dist.all_reduce(eng.grads)
    dist.all_reduce(loss)
    eng.adam_step(1e-4, weight_decay=1e-6)
    res = {"loss": float(loss), "grads": eng.grads.cpu(), "params": eng.params.cpu()}
    if rank == 0:  # the single-GPU step on the whole batch
        ref = vb.VarEngine(vb.KUKA, 100, 3, dev)
        ref.load_state_dict(omodel.init_state_dict("kuka", 3))
        ref.zero_grad()
        l1 = ref.triplet_step(torch.from_numpy(images).to(dev),
                              torch.from_numpy(np.concatenate([sp, sn])[:, 0]).to(dev).contiguous(), margin=1.0)
        ref.adam_step(1e-4, weight_decay=1e-6)
        res.update(ref_loss=float(l1), ref_grads=ref.grads.cpu(), ref_params=ref.params.cpu())
    torch.save(res, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_gpu_step(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 1000, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(r0["params"], r1["params"]) and torch.equal(r0["grads"], r1["grads"])
    assert abs(r0["loss"] - r0["ref_loss"]) < 1e-5 * max(1.0, abs(r0["ref_loss"]))
    g, gr = r0["grads"].numpy(), r0["ref_grads"].numpy()
    assert np.abs(g - gr).max() <= 1e-4 * np.abs(gr).max()  # fp32 summation order only
    assert np.abs(r0["params"].numpy() - r0["ref_params"].numpy()).max() < 2.1e-4
