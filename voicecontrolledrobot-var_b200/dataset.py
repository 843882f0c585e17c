"""Host mirror of dataset.py (VARDataset, VARFineTuneDataset, loadEnvData) with the triplet
assembly on the GPU.

The reference builds every triplet on the CPU inside `Dataset.__getitem__` (integer draws on
torch's global generator, then a per-clip MFCC).  Here the pickled triplet records are
uploaded once (uint8 images, labels), every clip lives in a device arena, and one batch is
assembled by two launches: `var_sampler_batch` (bit-exact index draws) and `var_mfcc_fwd`
(features of the 2B selected clips).  `loadEnvData` keeps the reference signature and returns
`(generator, final_dataset)`; the generator yields the reference's
`(image, sound_positive, sound_negative, gt)` tuples, already on the device.

Both sampling schemes are covered by the device sampler: pybullet/Kuka (intent -> dataset ->
clip, audioLoader.py:166-177, torchaudio-flavoured MFCC) and iTHOR (task list of dataset.py:17-29,
location-synonym / object-synonym / clip draws of audioLoader.py:203-237, and -- because
getAudioFromTask leaves mfcc_from=None -- the python_speech_features-flavoured MFCC).
"""
import glob
import os
import pickle
import queue
import threading

import numpy as np
import torch
from torch.utils.data.dataset import ConcatDataset, Dataset

from ._lib import check, lib, ptr, stream_ptr
from .Envs.audioLoader import audioLoader, mfcc_device


def torch_generator_words(rng_state):
    """torch.get_rng_state() of the CPU generator -> (uint32 words[624], read position).  The blob is
    the legacy THGeneratorState at::CPUGeneratorImpl serialises: u64 seed, i32 left, i32 seeded,
    u64 next, u64 state[624], ...; `left <= 1` means the next draw twists first (position 624)."""
    b = rng_state.numpy().tobytes()
    left = int(np.frombuffer(b, dtype=np.int32, count=1, offset=8)[0])
    nxt = int(np.frombuffer(b, dtype=np.uint64, count=1, offset=16)[0])
    words = np.ascontiguousarray(np.frombuffer(b, dtype=np.uint64, count=624, offset=24).astype(np.uint32))
    return words, (624 if left <= 1 else nxt)


class DeviceTripletSampler:
    """Device-resident mt19937 + the tables of one triplet dataset (see var_sampler_* in
    include/var_b200.h).  Runs the same generator algorithm and the same draw order as the
    reference's `for batch in DataLoader(shuffle=True, num_workers=0)` loop; `adopt_torch_state()`
    makes it continue torch's global CPU generator from its current state (what the reference
    draws from), `seed(s)` starts a fresh `torch.manual_seed(s)` stream.

    `dataset_sizes` (pybullet): clips per (intent, dataset).  `task_tables` (iTHOR): a TaskClipArena
    (per-task synonym counts and the [task, loc synonym, obj synonym] clip lists)."""

    def __init__(self, task_num, dataset_sizes, gt, stored_sn=None, seed=0, device=None, clip_off=None,
                 clip_len=None, task_tables=None):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceTripletSampler needs CUDA; there is no CPU fallback")
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.task_num = int(task_num)
        i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32), device=self.device)
        self.nds2 = None
        if task_tables is not None:
            t = task_tables
            if t.task_num != self.task_num:
                raise ValueError(f"config.taskNum = {self.task_num} but allTasks lists {t.task_num} tasks")
            self.max_ds, self.max_ds2, self.n_clips = t.max_loc * t.max_obj, t.max_obj, t.n_clips
            self.nds, self.nds2 = i32(t.n_loc), i32(t.n_obj)
            self.nclips, self.clip_base = i32(t.nclips.reshape(self.task_num, -1)), i32(t.clip_base.reshape(self.task_num, -1))
            cur = t.n_clips
        else:
            self.max_ds = max(1, max(len(s) for s in dataset_sizes))
            nds = [len(s) for s in dataset_sizes]
            if min(nds) < 1 or any(c < 1 for s in dataset_sizes for c in s):
                raise ValueError("every intent needs at least one dataset with at least one clip")
            nclips = np.zeros((self.task_num, self.max_ds), np.int32)
            base = np.zeros((self.task_num, self.max_ds), np.int32)
            cur = 0
            for i, s in enumerate(dataset_sizes):
                for j, c in enumerate(s):
                    nclips[i, j] = c
                    base[i, j] = cur
                    cur += c
            self.n_clips = cur
            self.nds, self.nclips, self.clip_base = i32(nds), i32(nclips), i32(base)
        self.clip_off = clip_off if clip_off is not None else torch.zeros(cur, dtype=torch.int64, device=self.device)
        self.clip_len = clip_len if clip_len is not None else torch.zeros(cur, dtype=torch.int32, device=self.device)
        self.gt = i32(gt)
        self.n_items = int(self.gt.numel())
        self.stored_sn = i32(stored_sn) if stored_sn is not None else None
        self.state = torch.zeros(625, dtype=torch.int32, device=self.device)
        if seed is None:
            self.adopt_torch_state()
        else:
            self.seed(seed)
        self.perm = torch.empty(self.n_items, dtype=torch.int32, device=self.device)

    def seed(self, seed):
        check(lib.var_sampler_seed(ptr(self.state), int(seed) & 0xFFFFFFFFFFFFFFFF, stream_ptr()), "var_sampler_seed")

    def adopt_torch_state(self, rng_state=None):
        """Continue torch's global CPU generator (or a given torch.get_rng_state() blob).  Under
        torch.distributed rank 0's state is broadcast, so every rank draws the same global batches."""
        import torch.distributed as dist
        if rng_state is None:
            rng_state = torch.get_rng_state()
        rng_state = torch.as_tensor(rng_state, dtype=torch.uint8).clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            buf = rng_state.to(self.device) if dist.get_backend() == "nccl" else rng_state
            dist.broadcast(buf, src=0)
            rng_state = buf.cpu()
        words, pos = torch_generator_words(rng_state)
        check(lib.var_sampler_set_state(ptr(self.state), words.ctypes.data, int(pos), stream_ptr()),
              "var_sampler_set_state")

    def begin_epoch(self):
        """DataLoader iterator creation + RandomSampler permutation -> int32 [n_items] on device."""
        check(lib.var_sampler_epoch(ptr(self.state), self.n_items, ptr(self.perm), stream_ptr()), "var_sampler_epoch")
        return self.perm

    def alloc_outputs(self, B):
        """Output buffers of one `sample()` call for batches of up to B items (reusable across calls)."""
        dev = self.device
        return {"item": torch.empty(B, dtype=torch.int32, device=dev),
                "gt": torch.empty(B, dtype=torch.int32, device=dev),
                "sn": torch.empty(B, dtype=torch.int32, device=dev),
                "rec": torch.empty(B, 6, dtype=torch.int32, device=dev),
                "off": torch.empty(2 * B, dtype=torch.int64, device=dev),
                "len": torch.empty(2 * B, dtype=torch.int32, device=dev),
                "_scratch": torch.empty(B, dtype=torch.int32, device=dev)}

    def sample(self, items, buffers=None):
        """items: int32 device tensor of dataset indices (a slice of the epoch permutation).  `buffers`: the
        result of alloc_outputs() to write into (views of the first B entries are returned)."""
        B = int(items.numel())
        if buffers is None:
            buffers = self.alloc_outputs(B)
        out = {"item": buffers["item"][:B], "gt": buffers["gt"][:B], "sn": buffers["sn"][:B], "rec": buffers["rec"][:B],
               "off": buffers["off"][:2 * B], "len": buffers["len"][:2 * B]}
        scratch = buffers["_scratch"][:B]
        items = items.contiguous()
        tail = (ptr(self.clip_off), ptr(self.clip_len), ptr(scratch), ptr(out["item"]), ptr(out["gt"]), ptr(out["sn"]),
                ptr(out["rec"]), ptr(out["off"]), ptr(out["len"]), stream_ptr())
        if self.nds2 is None:
            check(lib.var_sampler_batch(ptr(self.state), B, self.task_num, ptr(items), ptr(self.gt), ptr(self.stored_sn),
                                        ptr(self.nds), ptr(self.nclips), ptr(self.clip_base), self.max_ds, *tail),
                  "var_sampler_batch")
        else:
            check(lib.var_sampler_batch_tasks(ptr(self.state), B, self.task_num, ptr(items), ptr(self.gt),
                                              ptr(self.stored_sn), ptr(self.nds), ptr(self.nds2), ptr(self.nclips),
                                              ptr(self.clip_base), self.max_ds // self.max_ds2, self.max_ds2, *tail),
                  "var_sampler_batch_tasks")
        out["_keep"] = (scratch, items)
        return out


class VARDataset(Dataset):
    """Same constructor / fields as dataset.py:10-32: `ground_truth_pair` is the pickled list of
    {'image': u8[3,96,96], 'ground_truth': int[, 'sound_negative_id': int]} records."""

    def __init__(self, picklePath, config, **kwargs):
        self.filePath = picklePath
        self.config = config
        with open(self.filePath, 'rb') as f:
            self.ground_truth_pair = pickle.load(f)
        self.audio = kwargs['audio']
        if config.name == 'AI2ThorConfig':
            from .Envs.ai2thor.RL_env_VAR import Task
            self.Task = Task
            # task list, dataset.py:21-29
            self.tl = [Task(loc=loc, obj=obj, act=act) for loc in config.allTasks for obj in config.allTasks[loc]
                       for act in config.allTasks[loc][obj]]

    def __len__(self):
        return len(self.ground_truth_pair)

    def _draw_sn(self, item, gt):
        if 'sound_negative_id' in item:
            return int(item['sound_negative_id'])
        sn_id = torch.randint(low=0, high=self.config.taskNum, size=()).item()  # dataset.py:76
        return self.config.taskNum if gt == sn_id else sn_id

    def getImgSoundPair(self, gt, sn_id):
        """dataset.py:34-62 (both configs); features come from the GPU MFCC kernel."""
        T = self.config.taskNum
        zeros = lambda: np.zeros(shape=self.config.sound_dim)
        if self.config.name == 'AI2ThorConfig':
            feat = lambda i: self.audio.getAudioFromTask(torch, self.tl[i], self.Task)[0]
        else:
            feat = lambda i: self.audio.genSoundFeat(intentIdx=i, featType='MFCC', rand_fn=torch.randint)[0]
        if gt == T:
            return zeros(), feat(sn_id)
        pos = feat(gt)
        return pos, (zeros() if sn_id == T else feat(sn_id))

    def __getitem__(self, index):
        item = self.ground_truth_pair[index]
        image = (torch.from_numpy(item['image']) / 255.).float()
        gt = int(item['ground_truth'])
        if 'sound_negative' not in item:
            sp, sn = self.getImgSoundPair(gt, self._draw_sn(item, gt))
        else:
            sp, sn = item['sound_positive'], item['sound_negative']
        return image, sp, sn, gt


class VARFineTuneDataset(VARDataset):
    """dataset.py:94-133: the image-sound association is drawn once, at construction."""

    def __init__(self, picklePath, config, **kwargs):
        VARDataset.__init__(self, picklePath, config, **kwargs)
        for item in self.ground_truth_pair:
            if 'sound_negative' in item:
                continue
            gt = int(item['ground_truth'])
            item['sound_positive'], item['sound_negative'] = self.getImgSoundPair(gt, self._draw_sn(item, gt))

    def __getitem__(self, index):
        item = self.ground_truth_pair[index]
        image = (torch.from_numpy(item['image']) / 255.).float()
        return image, item['sound_positive'], item['sound_negative'], int(item['ground_truth'])


class _Slot:
    """One in-flight batch of the streaming loader: pinned host staging + device staging."""

    def __init__(self, bs, max_clip, device):
        self.h_img = torch.empty(bs, 3, 96, 96, dtype=torch.uint8).pin_memory()
        self.h_wav = torch.empty(2 * bs * (max_clip + 1), dtype=torch.int16).pin_memory()
        self.h_meta = torch.empty(2, 2 * bs, dtype=torch.int64).pin_memory()
        self.h_items = torch.empty(bs, dtype=torch.int64).pin_memory()
        self.d_img = torch.empty(bs, 3, 96, 96, dtype=torch.uint8, device=device)
        self.d_wav = torch.empty(2 * bs * (max_clip + 1), dtype=torch.int16, device=device)
        self.d_meta = torch.empty(2, 2 * bs, dtype=torch.int64, device=device)
        self.d_len32 = torch.empty(2 * bs, dtype=torch.int32, device=device)
        self.d_snd = None     # [2 * bs, F, 40] MFCC features of this slot (allocated on first use)
        self.free = threading.Event()
        self.free.set()
        self.consumed = None  # CUDA event: the compute stream is done reading this slot
        self.buf = None       # sampler output buffers of this slot (allocated on first use)


class DeviceTripletLoader:
    """Iterable replacing `DataLoader(ConcatDataset, batch_size, shuffle=True, num_workers=0)` for
    VARDataset records; batches are assembled by var_sampler_batch[_tasks] + var_mfcc_fwd.  Yields the
    reference tuples; `raw_batches()` yields the uint8 / [2B, F, 40] form the fused trainer consumes
    (rank-sliced under data parallelism).  Two residency modes:

    * resident=True (default): uint8 frames, labels and the int16 clip arena live in HBM (a B200 holds
      180 GB: 6 M frames); the sampler and the MFCC of batch k+1 run on a side stream under step k.
    * resident=False: frames and clips stay in pinned host memory (datasets beyond HBM).  A prefetch
      thread -- the counterpart of the reference's DataLoader workers -- draws batch k+1's indices on a
      side stream, gathers its frames / clips into pinned staging, uploads them and runs their MFCC on that
      stream while the GPU runs step k.

    seed=None (default) continues torch's global CPU generator from its state at construction, as the
    reference's DataLoader does; an int starts a fresh `torch.manual_seed(seed)` stream."""

    def __init__(self, images_u8, gt, stored_sn, audio, arena, config, batch_size, shuffle=True, drop_last=False,
                 seed=None, device=None, rank=0, world_size=1, resident=True):
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.config, self.audio, self.arena = config, audio, arena
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.rank, self.world_size = rank, world_size
        self.resident = bool(resident)
        images_u8 = torch.as_tensor(images_u8, dtype=torch.uint8).contiguous()
        self.n_items = images_u8.shape[0]
        if self.resident:
            self.images = images_u8.to(self.device)
        else:
            self.images_host = images_u8.pin_memory()
            self._clip_off_host = arena.clip_off.cpu().numpy()
            self.h2d_bytes = 0  # bytes uploaded by the batches yielded so far
        # one STFT parameter set per loader: the reference picks it per drawn dataset
        # (audioLoader.py:183); mixing NSynth/UrbanSound (1024-point) with the 512-point sets in
        # ONE intent table would need two plans per batch and is rejected here.
        names = sorted({n for row in arena.dataset_names for n in row})
        params = {audio.param_dict[n] for n in names}
        if len(params) != 1:
            raise NotImplementedError(f"datasets {names} mix STFT parameter sets")
        self.param = params.pop()
        ithor = config.name == 'AI2ThorConfig'
        # getAudioFromTask -> genSoundFeatFromTask(mfcc_from=None) -> python_speech_features (audioLoader.py:159-161,
        # 203, 234-236); genSoundFeat defaults to torchaudio (audioLoader.py:187)
        self.flavour = 1 if ithor else 0
        self.sampler = DeviceTripletSampler(config.taskNum, None if ithor else arena.dataset_sizes, gt, stored_sn, seed,
                                            self.device, arena.clip_off, arena.clip_len,
                                            task_tables=arena if ithor else None)
        # the sampler state is only ever touched on this stream (seeded above on the current one)
        # streaming mode: the producer thread WAITS for the sampler's indices, so the loader work (1-CTA sampler, uploads,
        # MFCC) runs at high priority instead of queueing behind the step's persistent kernels; resident mode is fully
        # asynchronous and keeps the default priority
        self._ls = torch.cuda.Stream(self.device, priority=0 if resident else -1)
        self._ls.wait_stream(torch.cuda.current_stream(self.device))
        self._slots = None

    def __len__(self):
        n = self.n_items // self.batch_size
        return n if (self.drop_last or self.n_items % self.batch_size == 0) else n + 1

    # ------------------------------------------------------------------ batch assembly
    def _epoch_perm(self):
        with torch.cuda.stream(self._ls):
            if self.shuffle:
                return self.sampler.begin_epoch().clone()
            return torch.arange(self.n_items, dtype=torch.int32, device=self.device)

    def _batch_starts(self):
        n, bs = self.n_items, self.batch_size
        return [s for s in range(0, n, bs) if not (self.drop_last and n - s < bs)]

    def _draw(self, perm, s, buffers=None):
        """Sampler launch for the batch starting at permutation position s (on the loader stream)."""
        items = perm[s:s + self.batch_size]
        B = int(items.numel())
        rec = self.sampler.sample(items, buffers)  # every rank draws the GLOBAL batch: identical streams
        lo, hi = (B * self.rank) // self.world_size, (B * (self.rank + 1)) // self.world_size
        return rec, B, lo, hi

    def _resident_batches(self):
        """Batch k+1's sampler + gather + MFCC launches are issued on the loader stream before batch k is handed
        out, so they run under step k instead of on its critical path.  Every batch writes into one of three
        preallocated slots (no allocation on the loader stream: a cross-stream free would stall the caching
        allocator until the consumer's work has drained)."""
        cs = torch.cuda.current_stream(self.device)
        n_fft, win, hop = self.audio.stft_params(self.param)
        F = self.config.sound_dim[1]
        wav = self.arena.wav  # first use uploads the arena on the compute stream
        bs = self.batch_size
        bmax = (bs + self.world_size - 1) // self.world_size + 1
        if getattr(self, "_rslots", None) is None or self._rslots[0]["img"].shape[0] < bmax:
            dev = self.device
            self._rslots = [dict(buf=self.sampler.alloc_outputs(bs), img=torch.empty(bmax, 3, 96, 96, dtype=torch.uint8, device=dev),
                                 snd=torch.empty(2 * bmax, F, 40, dtype=torch.float32, device=dev),
                                 idx=torch.empty(bmax, dtype=torch.int64, device=dev),
                                 off=torch.empty(2 * bmax, dtype=torch.int64, device=dev),
                                 ln=torch.empty(2 * bmax, dtype=torch.int32, device=dev), consumed=None) for _ in range(3)]
        self._ls.wait_stream(cs)
        perm = self._epoch_perm()

        def issue(k, s):
            slot = self._rslots[k % 3]
            with torch.cuda.stream(self._ls):
                if slot["consumed"] is not None:
                    self._ls.wait_event(slot["consumed"])  # the slot's previous batch has been read on cs
                rec, B, lo, hi = self._draw(perm, s, slot["buf"])
                b = hi - lo
                if b == 0:
                    out = (None, None, rec["gt"][lo:hi], B, rec)
                else:
                    torch.cat([rec["off"][lo:hi], rec["off"][B + lo:B + hi]], out=slot["off"][:2 * b])
                    torch.cat([rec["len"][lo:hi], rec["len"][B + lo:B + hi]], out=slot["ln"][:2 * b])
                    slot["idx"][:b].copy_(rec["item"][lo:hi])
                    torch.index_select(self.images, 0, slot["idx"][:b], out=slot["img"][:b])
                    mfcc_device(wav, slot["off"][:2 * b], slot["ln"][:2 * b], self.audio.fs, n_fft, win, hop, F,
                                flavour=self.flavour, out=slot["snd"][:2 * b])
                    out = (slot["img"][:b], slot["snd"][:2 * b], rec["gt"][lo:hi], B, rec)
                ev = torch.cuda.Event()
                ev.record(self._ls)
            return out, ev, slot

        starts = self._batch_starts()
        nxt = issue(0, starts[0]) if starts else None
        for k in range(len(starts)):
            out, ev, slot = nxt
            nxt = issue(k + 1, starts[k + 1]) if k + 1 < len(starts) else None
            cs.wait_event(ev)
            yield out
            # the consumer has launched its step on cs: the slot may be refilled once that work is done
            slot["consumed"] = torch.cuda.Event()
            slot["consumed"].record(cs)

    def _producer(self, q, stop, perm, starts, max_clip, mfcc_args):
        def put(x):
            while not stop.is_set():
                try:
                    q.put(x, timeout=0.05)
                    return True
                except queue.Full:
                    pass
            return False
        try:
            torch.cuda.set_device(self.device)
            wav_host = self.arena.wav_host
            import time as _t
            st_ = self.producer_stats = {"batches": 0, "wait_slot": 0.0, "draw_launch": 0.0, "draw_wait": 0.0, "gather": 0.0,
                                         "upload": 0.0, "put": 0.0}
            n_fft, win, hop, F = mfcc_args
            nth = max(1, min(4, (os.cpu_count() or 4) // max(1, self.world_size)))  # ranks share the host cores

            def launch_draw(k, s):
                """Stage 1 of batch k (asynchronous): indices drawn on the loader stream, index / clip tables copied to
                the slot's pinned buffers.  On a busy GPU the single-CTA sampler waits milliseconds for an SM, so it is
                issued one batch ahead and its latency hides behind the host gather of the previous batch."""
                t0 = _t.perf_counter()
                slot = self._slots[k % len(self._slots)]
                while not slot.free.wait(0.05):
                    if stop.is_set():
                        return None
                slot.free.clear()
                t1 = _t.perf_counter(); st_["wait_slot"] += t1 - t0
                with torch.cuda.stream(self._ls):
                    if slot.consumed is not None:
                        self._ls.wait_event(slot.consumed)   # the previous tenant of this slot was read on cs
                    if slot.buf is None:
                        slot.buf = self.sampler.alloc_outputs(self.batch_size)
                    rec, B, lo, hi = self._draw(perm, s, slot.buf)
                    b = hi - lo
                    if b:
                        meta = torch.stack([torch.cat([rec["off"][lo:hi], rec["off"][B + lo:B + hi]]),
                                            torch.cat([rec["len"][lo:hi], rec["len"][B + lo:B + hi]]).long()])
                        slot.h_meta[:, :2 * b].copy_(meta, non_blocking=True)
                        slot.h_items[:b].copy_(rec["item"][lo:hi], non_blocking=True)
                    drawn = torch.cuda.Event()
                    drawn.record(self._ls)
                st_["draw_launch"] += _t.perf_counter() - t1
                return slot, rec, B, lo, hi, b, drawn

            def finish(p):
                """Stage 2: wait for the indices, gather frames / clips into pinned staging with the library's memcpy
                threads (ctypes releases the GIL: the main thread keeps launching kernels meanwhile), upload, MFCC."""
                slot, rec, B, lo, hi, b, drawn = p
                t1 = _t.perf_counter()
                drawn.synchronize()
                gt = rec["gt"][lo:hi]
                t2 = _t.perf_counter(); st_["draw_wait"] += t2 - t1
                if not b:
                    return put((slot, 0, B, gt, rec, None, 0))
                check(lib.var_host_gather_rows(self.images_host.data_ptr(), 3 * 96 * 96, slot.h_items.data_ptr(), b,
                                               slot.h_img.data_ptr(), nth), "var_host_gather_rows")
                new_off = torch.empty(2 * b, dtype=torch.int64)
                cur = int(lib.var_host_gather_clips(wav_host.data_ptr(), slot.h_meta[0].data_ptr(), slot.h_meta[1].data_ptr(),
                                                    2 * b, slot.h_wav.data_ptr(), new_off.data_ptr(), nth))
                check(cur, "var_host_gather_clips")
                slot.h_meta[0, :2 * b] = new_off
                t3 = _t.perf_counter(); st_["gather"] += t3 - t2
                with torch.cuda.stream(self._ls):
                    slot.d_img[:b].copy_(slot.h_img[:b], non_blocking=True)
                    slot.d_wav[:cur].copy_(slot.h_wav[:cur], non_blocking=True)
                    slot.d_meta[:, :2 * b].copy_(slot.h_meta[:, :2 * b], non_blocking=True)
                    # the MFCC of this batch also runs here, ahead of the consumer (into the slot's own buffer: no
                    # allocation on the loader stream)
                    if slot.d_snd is None or slot.d_snd.shape[1] != F:
                        slot.d_snd = torch.empty(2 * self.batch_size, F, 40, dtype=torch.float32, device=self.device)
                    slot.d_len32[:2 * b].copy_(slot.d_meta[1, :2 * b])
                    mfcc_device(slot.d_wav, slot.d_meta[0, :2 * b], slot.d_len32[:2 * b], self.audio.fs, n_fft, win, hop, F,
                                flavour=self.flavour, out=slot.d_snd[:2 * b])
                    ev = torch.cuda.Event()
                    ev.record(self._ls)
                t4 = _t.perf_counter(); st_["upload"] += t4 - t3
                ok = put((slot, b, B, gt, rec, ev, b * 3 * 96 * 96 + cur * 2 + 2 * 2 * b * 8))
                st_["put"] += _t.perf_counter() - t4; st_["batches"] += 1
                return ok

            pending = None
            for k, s in enumerate(starts):
                nxt = launch_draw(k, s)
                if nxt is None:
                    return
                if pending is not None and not finish(pending):
                    return
                pending = nxt
            if pending is not None and not finish(pending):
                return
            put(None)
        except BaseException as e:  # noqa: BLE001 -- re-raised in the consumer
            put(e)

    def _streaming_batches(self):
        cs = torch.cuda.current_stream(self.device)
        n_fft, win, hop = self.audio.stft_params(self.param)
        F = self.config.sound_dim[1]
        max_clip = int(self.arena.clip_len.max()) if self.arena.clip_len.numel() else 0
        if self._slots is None or self._slots[0].h_wav.numel() < 2 * self.batch_size * (max_clip + 1):
            self._slots = [_Slot(self.batch_size, max_clip, self.device) for _ in range(3)]
        for sl in self._slots:
            sl.free.set()
        perm = self._epoch_perm()
        q, stop = queue.Queue(maxsize=2), threading.Event()
        th = threading.Thread(target=self._producer, args=(q, stop, perm, self._batch_starts(), max_clip, (n_fft, win, hop, F)),
                              daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                slot, b, B, gt, rec, ev, nbytes = item
                if not b:
                    yield None, None, gt, B, rec
                    slot.free.set()
                    continue
                cs.wait_event(ev)
                self.h2d_bytes += nbytes
                yield slot.d_img[:b], slot.d_snd[:2 * b], gt, B, rec
                # the consumer has launched its step on cs: mark the slot reusable once that work is done
                slot.consumed = torch.cuda.Event()
                slot.consumed.record(cs)
                slot.free.set()
        finally:
            stop.set()
            for sl in self._slots:
                sl.free.set()
            th.join(timeout=5.0)

    def raw_batches(self):
        """-> (uint8 images [b, 3, 96, 96], sounds [2b, F, 40], gt [b], global batch B, sampler record) with
        b = this rank's slice of the global batch (may be 0 on a ragged tail: images/sounds are then None)."""
        return self._resident_batches() if self.resident else self._streaming_batches()

    def stream(self):
        """Raw batches across epoch boundaries, endlessly (for step-counted training / benchmarking)."""
        while True:
            yield from self.raw_batches()

    def __iter__(self):
        for img_u8, sounds, gt, _, _ in self.raw_batches():
            if img_u8 is None:
                continue
            b = img_u8.shape[0]
            F = sounds.shape[1]
            yield ((img_u8.float() / 255.), sounds[:b].view(b, 1, F, 40), sounds[b:].view(b, 1, F, 40), gt.long())


def loadEnvData(data_dir, config, batch_size, shuffle, num_workers, drop_last, loadNum=None,
                dtype=VARDataset, train_test='train'):
    """dataset.py:136-168.  `num_workers` is accepted and ignored: batches are assembled on the
    device, in the draw order the reference has with num_workers=0."""
    audio = audioLoader(config=config)
    audio.loadData()
    all_datasets = []
    for i, dirs in enumerate(data_dir):
        assert os.path.exists(dirs)
        path = os.path.join(dirs, train_test)
        fileList = glob.glob(os.path.join(path, '*.pickle'))
        if not (loadNum is None or loadNum[i] == 'all') and len(fileList) > int(loadNum[i]):
            fileList = np.random.choice(fileList, size=int(loadNum[i]))
        for filePath in fileList:
            all_datasets.append(dtype(picklePath=str(filePath), config=config, audio=audio))
    final_dataset = ConcatDataset(all_datasets)
    records = [p for d in final_dataset.datasets for p in d.ground_truth_pair]
    num = [0] * (config.taskNum + 1)
    for p in records:
        num[int(p['ground_truth'])] += 1
    if dtype is VARFineTuneDataset or any('sound_negative' in p for p in records):
        # features are fixed per record (dataset.py:94-133): plain tensors, no sampling left
        generator = torch.utils.data.DataLoader(final_dataset, batch_size=batch_size, shuffle=shuffle, num_workers=0,
                                                pin_memory=True, drop_last=drop_last)
    else:
        images = np.stack([np.asarray(p['image'], dtype=np.uint8)[:3] for p in records])
        gt = [int(p['ground_truth']) for p in records]
        has_sn = ['sound_negative_id' in p for p in records]
        if any(has_sn) and not all(has_sn):
            raise ValueError("records mix stored and drawn sound_negative_id")
        stored = [int(p['sound_negative_id']) for p in records] if all(has_sn) else None
        # config.pretextDataResident (optional, default True): keep frames + clips in HBM, or stream them from
        # pinned host memory every step
        generator = DeviceTripletLoader(images, gt, stored, audio, audio.build_arena(), config, batch_size,
                                        shuffle=shuffle, drop_last=drop_last,
                                        resident=getattr(config, "pretextDataResident", True))
    print("The number of pairs for each object in the dataset is:", num)
    return generator, final_dataset
