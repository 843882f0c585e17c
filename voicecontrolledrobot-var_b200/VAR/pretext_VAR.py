"""Host mirror of VAR/pretext_VAR.py::VAR_Pretext.

`trainRepresentation(epoch, lr, start_ep=0, plot=False)` keeps the reference's side effects
(`<ep>.pt` legacy-format checkpoints, `progress.csv`, per-epoch MultiStepLR) but each iteration
is ONE fused device step instead of the reference's zero_grad / forward / TripletMarginLoss /
backward / Adam sequence (VAR/pretext_VAR.py:55-70):

    var_sampler_batch -> var_mfcc_fwd -> var_net_triplet_step -> [NCCL all-reduce] -> var_adam_step

Under `torch.distributed` every rank draws the same global index stream, takes its slice of
the batch, and the flat gradient buffer is summed over ranks before an identical Adam step
(weights stay replicated).  The loss is read back once per epoch, not once per step."""
import os

import numpy as np
import torch
import torch.distributed as dist

from ..dataset import DeviceTripletLoader, loadEnvData
from ..engine import multistep_lr
from ..pretext import Pretext


class VAR_Pretext(Pretext):
    def __init__(self, config=None):
        if config is None:
            from cfg import main_config  # inside the reference tree, as VAR/pretext_VAR.py:14
            config = main_config()
        super().__init__(config)

    def _lr(self, base_lr, ep):
        if self.config.pretextLRStep == "step":
            return multistep_lr(base_lr, ep, self.config.pretextLRDecayEpoch, self.config.pretextLRDecayGamma)
        return base_lr

    def train_epoch(self, eng, data_generator, lr, world=1, rank=0, max_steps=None, on_step=None):
        """One pass over the generator (at most `max_steps` steps); returns the per-step device loss
        scalars.  `data_generator`: a DeviceTripletLoader, an iterator of its raw batches
        (`loader.stream()`), or any iterable of the reference's host tuples (image, sound_positive,
        sound_negative, gt).  `on_step(loss)` runs after each step's launches (e.g. to read the loss)."""
        cfg = self.config
        losses = []
        slot_batches = True  # batches live in the loader's preallocated slots (stable buffers)
        if isinstance(data_generator, DeviceTripletLoader):
            data_generator.rank, data_generator.world_size = rank, world
            batches = ((img, snd, gB) for img, snd, _, gB, _ in data_generator.raw_batches())
        elif getattr(data_generator, "gi_code", None) is not None and \
                data_generator.gi_code.co_name in ("stream", "raw_batches", "_resident_batches", "_streaming_batches"):
            batches = ((img, snd, gB) for img, snd, _, gB, _ in data_generator)
        else:
            def host_batches():
                for image, sp, sn, _ in data_generator:
                    b = image.shape[0]
                    lo, hi = (b * rank) // world, (b * (rank + 1)) // world
                    if hi == lo:
                        yield None, None, b
                        continue
                    F = cfg.sound_dim[1]
                    snd = torch.cat([sp[lo:hi].reshape(-1, F, 40), sn[lo:hi].reshape(-1, F, 40)]).float()
                    yield (image[lo:hi].to(self.device).contiguous(), snd.to(self.device).contiguous(), b)
            batches = host_batches()
            slot_batches = False  # fresh tensors every batch: a captured graph would be replayed once
        # one GPU: the step's ~80 launches are replayed as one CUDA graph per batch slot (engine.triplet_step_graphed);
        # with a process group the mid-backward bucket event of the gradient all-reduce keeps the eager launches
        graphed = (world == 1 and slot_batches and getattr(eng, "use_step_graph", False)
                   and hasattr(eng, "triplet_step_graphed"))
        for img, snd, global_b in batches:
            stepped = not (img is None or img.shape[0] == 0)
            if not stepped:
                # ragged tail batch smaller than the world size: this rank has no triplet, but it must
                # still take part in both collectives and in the (identical) Adam step
                eng.zero_grad()
                loss = torch.zeros((), dtype=torch.float32, device=self.device)
            elif graphed:
                loss = eng.triplet_step_graphed(img, snd, margin=cfg.tripletMargin, loss_denominator=global_b)
            else:
                eng.zero_grad()
                loss = eng.triplet_step(img, snd, margin=cfg.tripletMargin, loss_denominator=global_b)
            if world > 1:
                if hasattr(eng, "allreduce_grads"):
                    eng.allreduce_grads(stepped)   # bucketed: the recurrent-layer range overlaps the backward pass
                else:
                    dist.all_reduce(eng.grads)
                dist.all_reduce(loss)
            eng.adam_step(lr, weight_decay=cfg.pretextAdamL2)
            losses.append(loss)
            if on_step is not None:
                on_step(loss)
            if max_steps is not None and len(losses) >= max_steps:
                break
        return losses

    def trainRepresentation(self, epoch, lr, start_ep=0, plot=False):
        print('Begin representation training')
        cfg = self.config
        data_generator, ds = loadEnvData(data_dir=cfg.pretextDataDir, config=cfg,
                                         batch_size=cfg.pretextTrainBatchSize, shuffle=True,
                                         num_workers=cfg.pretextDataNumWorkers, drop_last=False,
                                         loadNum=cfg.pretextDataFileLoadNum, dtype=cfg.pretextDataset)
        os.makedirs(cfg.pretextModelSaveDir, exist_ok=True)
        self.pretextModel.train()
        eng = self.pretextModel._get_engine(self.device)
        eng.reset_optimizer()  # the reference builds a fresh optim.Adam per call (VAR/pretext_VAR.py:33)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        loss_list = []
        for ep in range(epoch):
            losses = self.train_epoch(eng, data_generator, self._lr(lr, ep), world, rank)
            if (ep + 1) % cfg.pretextModelSaveInterval == 0 or ep + 1 == epoch:
                self.pretextModel.sync_from_engine()
                if rank == 0:
                    fname = os.path.join(cfg.pretextModelSaveDir, str(start_ep + ep) + '.pt')
                    torch.save(self.pretextModel.state_dict(), fname, _use_new_zipfile_serialization=False)
                    print('Model saved to ' + fname)
            avg_loss = float(torch.stack(losses).mean()) if losses else float('nan')
            loss_list.append(avg_loss)
            print('average loss', avg_loss)
        if epoch > 0:
            self.pretextModel.sync_from_engine()
        if cfg.pretextTrain and rank == 0:
            import pandas as pd
            save_path = os.path.join(cfg.pretextModelSaveDir, 'progress.csv')
            pd.DataFrame({'avg_loss': loss_list}).to_csv(save_path, mode='w', header=True, index=False)
            print('results saved to', save_path)
        print('Pretext Training Complete')
        self.pretextModel.eval()
        return np.asarray(loss_list)
