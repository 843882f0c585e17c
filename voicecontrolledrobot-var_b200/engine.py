"""Host-side owner of one VAR encoder pair on one GPU.

`VarEngine` wraps a `var_net_*` object of libvar_b200.so: it owns the flat packed
parameter / tf32-operand / gradient / Adam buffers and the activation workspace (torch
tensors used purely as device memory), converts to and from the reference `state_dict`
layout (models/pretext/*.py), and exposes the four calls the hot path needs:
forward, backward, the fused triplet step (VAR/pretext_VAR.py:56-69) and the batched
reward query (Envs/vec_env/vec_pretext_normalize.py:82-101).
"""
import ctypes as C
import math
import os

import torch

from ._lib import check, lib, ptr, stream_ptr

KUKA, ITHOR = 0, 1


class VarEngine:
    def __init__(self, kind, sound_frames, rep_dim=3, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("VarEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.kind, self.F, self.D = kind, sound_frames, rep_dim
        h = C.c_void_p()
        check(lib.var_net_create(kind, sound_frames, rep_dim, C.byref(h)), "var_net_create")
        self._net = h
        self.nparams = int(lib.var_net_param_floats(h))
        z = lambda: torch.zeros(self.nparams, dtype=torch.float32, device=self.device)
        self.params, self.params_mma, self.grads = z(), z(), z()
        self.adam_m = self.adam_v = None
        self.adam_steps = 0
        check(lib.var_net_bind(h, ptr(self.params), ptr(self.params_mma), ptr(self.grads)), "var_net_bind")
        self.tensors = []  # (name, shape, offset, packed)
        name = C.create_string_buffer(128)
        nd, shp, off, pk = C.c_int(), (C.c_int * 4)(), C.c_int64(), C.c_int64()
        for i in range(lib.var_net_num_tensors(h)):
            check(lib.var_net_tensor_info(h, i, name, 128, C.byref(nd), shp, C.byref(off), C.byref(pk)),
                  "var_net_tensor_info")
            self.tensors.append((name.value.decode(), tuple(shp[:nd.value]), off.value, pk.value))
        self.index = {t[0]: i for i, t in enumerate(self.tensors)}
        ir, sr = C.c_int(), C.c_int()
        lib.var_net_raw_dims(h, C.byref(ir), C.byref(sr))
        self.img_raw_dim, self.snd_raw_dim = ir.value, sr.value
        self._ws = None
        self._ws_need = {}
        self._reward_graphs = {}
        self._bucket = None
        # graphed training step (triplet_step_graphed): VAR_STEP_GRAPH=0 keeps the ~80 individual launches per step
        self.use_step_graph = os.environ.get("VAR_STEP_GRAPH", "1") != "0"
        self._step_graphs = {}

    def __del__(self):
        try:
            if getattr(self, "_net", None):
                if getattr(self, "_bucket", None):
                    lib.var_net_set_bucket_event(self._net, 0, None)
                lib.var_net_destroy(self._net)
                self._net = None
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, sd):
        """reference-layout tensors (any device) -> packed master + tf32 copy."""
        missing = [t[0] for t in self.tensors if t[0] not in sd]
        if missing:
            raise KeyError(f"state_dict is missing {missing}")
        st = stream_ptr()
        keep = []
        for i, (name, shape, _, _) in enumerate(self.tensors):
            src = sd[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(src.shape) != shape:
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(src.shape)}")
            keep.append(src)
            check(lib.var_net_load_tensor(self._net, i, ptr(src), st), f"var_net_load_tensor({name})")
        torch.cuda.current_stream().synchronize()  # sources may be freed after return
        self.reset_optimizer()  # moments of the replaced weights are stale

    def _store(self, which):
        out = {}
        st = stream_ptr()
        for i, (name, shape, _, _) in enumerate(self.tensors):
            dst = torch.empty(shape, dtype=torch.float32, device=self.device)
            check(lib.var_net_store_tensor(self._net, i, which, ptr(dst), st), f"var_net_store_tensor({name})")
            out[name] = dst
        return out

    def state_dict(self):
        return self._store(0)

    def grad_dict(self):
        return self._store(1)

    def set_overlap(self, on):
        """Image / sound branch overlap on two streams (default on)."""
        check(lib.var_net_set_overlap(self._net, 1 if on else 0), "var_net_set_overlap")
        self._step_graphs.clear()  # captured with the other stream layout

    def zero_grad(self):
        self.grads.zero_()

    # ------------------------------------------------------------------- workspace
    def _workspace(self, n_img, n_snd, train):
        key = (n_img, n_snd, bool(train))
        need = self._ws_need.get(key)
        if need is None:
            need = int(lib.var_net_workspace_bytes(self._net, n_img, n_snd, 1 if train else 0))
            check(need, "var_net_workspace_bytes")
            self._ws_need[key] = need
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _image_kind(images):
        if images.dtype == torch.uint8:
            return 0
        if images.dtype == torch.float32:
            return 1
        raise TypeError(f"images must be uint8 or float32, got {images.dtype}")

    def _check_inputs(self, images, sounds):
        if images is not None:
            if not images.is_cuda or not images.is_contiguous() or tuple(images.shape[1:]) != (3, 96, 96):
                raise ValueError("images must be a contiguous CUDA tensor [N, 3, 96, 96]")
        if sounds is not None:
            if (not sounds.is_cuda or not sounds.is_contiguous() or sounds.dtype != torch.float32
                    or tuple(sounds.shape[-2:]) != (self.F, 40)):
                raise ValueError(f"sounds must be a contiguous CUDA float32 tensor [N, (1,) {self.F}, 40]")

    # --------------------------------------------------------------------- compute
    def forward(self, images, sounds, train=False, want_raw=True):
        """-> (img_feat [Ni, D], img_raw [Ni, raw], snd_feat [Ns, D], snd_raw [Ns, raw]); None where absent."""
        self._check_inputs(images, sounds)
        ni = 0 if images is None else images.shape[0]
        ns = 0 if sounds is None else sounds.shape[0]
        ws = self._workspace(ni, ns, train)
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        img_feat = e(ni, self.D) if ni else None
        snd_feat = e(ns, self.D) if ns else None
        img_raw = e(ni, self.img_raw_dim) if (ni and want_raw) else None
        snd_raw = e(ns, self.snd_raw_dim) if (ns and want_raw) else None
        check(lib.var_net_forward(self._net, ptr(images), self._image_kind(images) if ni else 0, ni, ptr(sounds),
                                  ns, ptr(ws), ws.numel(), 1 if train else 0, ptr(img_feat), ptr(img_raw),
                                  ptr(snd_feat), ptr(snd_raw), stream_ptr()), "var_net_forward")
        return img_feat, img_raw, snd_feat, snd_raw

    def backward(self, d_img_feat, d_snd_feat):
        """Backward of the last forward(train=True); accumulates into self.grads."""
        ws = self._ws
        if ws is None:
            raise RuntimeError("backward() before forward(train=True)")
        for d in (d_img_feat, d_snd_feat):
            if d is not None and (not d.is_cuda or not d.is_contiguous() or d.dtype != torch.float32):
                raise ValueError("gradients must be contiguous CUDA float32 tensors")
        check(lib.var_net_backward(self._net, ptr(d_img_feat), ptr(d_snd_feat), ptr(ws), ws.numel(), stream_ptr()),
              "var_net_backward")

    def triplet_step(self, images, sounds, margin=1.0, loss_denominator=None, loss_out=None, feats_out=None):
        """Forward + fused triplet loss + backward for B images and 2B sounds (positives, then
        negatives).  Gradients accumulate into self.grads; returns the device scalar that received
        sum(hinge) / loss_denominator."""
        self._check_inputs(images, sounds)
        B = images.shape[0]
        if sounds.shape[0] != 2 * B:
            raise ValueError("sounds must hold B positives followed by B negatives")
        ws = self._workspace(B, 2 * B, True)
        if loss_out is None:
            loss_out = torch.zeros((), dtype=torch.float32, device=self.device)
        denom = float(B if loss_denominator is None else loss_denominator)
        check(lib.var_net_triplet_step(self._net, ptr(images), self._image_kind(images), ptr(sounds), B,
                                       float(margin), denom, ptr(ws), ws.numel(), ptr(loss_out), ptr(feats_out),
                                       stream_ptr()), "var_net_triplet_step")
        return loss_out

    def triplet_step_graphed(self, images, sounds, margin=1.0, loss_denominator=None):
        """zero_grad() + triplet_step() of one batch as ONE CUDA-graph replay (VAR/pretext_VAR.py:56-68: the step is
        ~80 dependent kernel launches; when the trainer reads the loss every step the GPU otherwise idles while the
        host issues them).  A graph belongs to the buffers it was captured on -- the loaders hand out a ring of three
        preallocated batch slots, so three graphs serve a whole run.  The first step of a shape runs eagerly (that run
        sets kernel attributes and sizes the workspace, none of which may happen inside a capture) and its slot is
        captured right behind it; every other slot is captured the first time it is met and replayed from then on, so
        after three steps nothing is captured any more.  Returns a fresh device scalar with the loss.  Results are
        those of the eager step: the graph holds the same kernels on the same streams."""
        if not self.use_step_graph:
            self.zero_grad()
            return self.triplet_step(images, sounds, margin, loss_denominator)
        B = images.shape[0]
        denom = float(B if loss_denominator is None else loss_denominator)
        ws = self._workspace(B, 2 * B, True)
        shape_key = (tuple(images.shape), images.dtype, tuple(sounds.shape), float(margin), denom)
        key = (images.data_ptr(), sounds.data_ptr(), ws.data_ptr()) + shape_key
        ent = self._step_graphs.get(key)
        if ent is None:
            if len(self._step_graphs) >= 24:  # buffers keep changing: stay eager
                self.zero_grad()
                return self.triplet_step(images, sounds, margin, loss_denominator)
            if shape_key not in self._step_graphs:
                self._step_graphs[shape_key] = True
                self.zero_grad()
                loss = self.triplet_step(images, sounds, margin, loss_denominator)
                self._capture_step(key, images, sounds, margin, loss_denominator, ws)  # records only; runs nothing
                return loss
            ent = self._capture_step(key, images, sounds, margin, loss_denominator, ws)
            if ent is None:  # capture failed: eager from now on
                self.zero_grad()
                return self.triplet_step(images, sounds, margin, loss_denominator)
        ent["graph"].replay()
        if ent["counted"]:
            ent["counted"] = False  # the library counted these launches while they were recorded
        else:
            check(lib.var_launch_count_add(ent["launches"]), "var_launch_count_add")
        return ent["loss"].clone()

    def _capture_step(self, key, images, sounds, margin, loss_denominator, ws):
        self._check_inputs(images, sounds)
        loss = torch.zeros((), dtype=torch.float32, device=self.device)
        l0 = lib.var_launch_count()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self.zero_grad()
                loss.zero_()  # the tail kernel ADDS the batch's hinge sum into it
                self.triplet_step(images, sounds, margin, loss_denominator, loss_out=loss)
        except Exception as exc:  # a driver / torch build that cannot capture the step: same kernels, launched one by one
            import warnings
            warnings.warn(f"CUDA-graph capture of the training step failed ({exc!r}); using individual launches")
            self.use_step_graph = False
            self._step_graphs.clear()
            torch.cuda.synchronize(self.device)
            return None
        ent = dict(graph=g, loss=loss, launches=int(lib.var_launch_count() - l0), counted=True, keep=(images, sounds, ws))
        self._step_graphs[key] = ent
        return ent

    def reward(self, images, goal_sounds=None, goal_feat_cached=None, env_reward=None, out=None):
        """-> (img_feat [N, D], goal_feat [N, D], img_sound_dot [N], reward [N]); `out` = the four
        destination tensors (else they are allocated)."""
        self._check_inputs(images, goal_sounds)
        N = images.shape[0]
        ws = self._workspace(N, N if goal_sounds is not None else 0, False)
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        img_feat, goal_feat, dot, rew = out if out is not None else (e(N, self.D), e(N, self.D), e(N), e(N))
        check(lib.var_net_reward(self._net, ptr(images), self._image_kind(images), ptr(goal_sounds),
                                 ptr(goal_feat_cached), ptr(env_reward), N, ptr(ws), ws.numel(), ptr(img_feat),
                                 ptr(goal_feat), ptr(dot), ptr(rew), stream_ptr()), "var_net_reward")
        return img_feat, goal_feat, dot, rew

    # ------------------------------------------------------------------- data parallel
    def allreduce_grads(self, stepped=True):
        """Sum the flat gradient buffer over the ranks (the one collective of a training step).  For the
        iTHOR net the recurrent-layer gradients (rnn.*, 77 % of the bytes, one contiguous range) are final
        half way through the backward pass: the library records an event there, and that range is reduced
        on a side stream under the remaining conv backward kernels; the rest follows on the compute
        stream.  `stepped=False` (this rank ran no triplet_step since zero_grad) reduces everything in order."""
        import os
        import torch.distributed as dist
        if not stepped or os.environ.get("VAR_DP_BUCKET", "1") == "0" or not self._bucket_setup():
            dist.all_reduce(self.grads)
            return
        off, cnt, ev, comm = self._bucket
        comm.wait_event(ev)  # recorded by the backward pass of the last triplet_step
        with torch.cuda.stream(comm):
            work = dist.all_reduce(self.grads[off:off + cnt], async_op=True)
        if off > 0:
            dist.all_reduce(self.grads[:off])
        if off + cnt < self.nparams:
            dist.all_reduce(self.grads[off + cnt:])
        work.wait()  # the compute stream waits for the side-stream reduction

    def _bucket_setup(self):
        if self._bucket is None:
            off, cnt = C.c_int64(), C.c_int64()
            if lib.var_net_grad_bucket(self._net, 0, C.byref(off), C.byref(cnt)) != 0:
                self._bucket = False
            else:
                ev = torch.cuda.Event()
                ev.record()  # creates the underlying cudaEvent_t
                check(lib.var_net_set_bucket_event(self._net, 0, ev.cuda_event), "var_net_set_bucket_event")
                self._bucket = (off.value, cnt.value, ev, torch.cuda.Stream(self.device))
        return bool(self._bucket)

    def reset_optimizer(self):
        """Forget the Adam moments and the bias-correction step count (a fresh torch.optim.Adam)."""
        if self.adam_m is not None:
            self.adam_m.zero_()
            self.adam_v.zero_()
        self.adam_steps = 0

    def reward_graph(self, N, image_dtype=torch.uint8, fresh_goal=False):
        """Cached RewardGraph for (N, image dtype, goal sound re-encoded or cached).  Graphs pin the
        workspace they were captured with, so a later, larger workspace request drops them."""
        key = (int(N), image_dtype, bool(fresh_goal))
        g = self._reward_graphs.get(key)
        if g is None or g.ws_ptr != (self._ws.data_ptr() if self._ws is not None else 0):
            g = RewardGraph(self, N, image_dtype, fresh_goal)
            g.ws_ptr = self._ws.data_ptr()
            self._reward_graphs[key] = g
        return g

    def adam_step(self, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.params)
            self.adam_v = torch.zeros_like(self.params)
        self.adam_steps += 1
        check(lib.var_adam_step(ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v),
                                ptr(self.params_mma), self.nparams, float(lr), float(betas[0]), float(betas[1]),
                                float(eps), float(weight_decay), self.adam_steps, float(grad_scale), stream_ptr()),
              "var_adam_step")


class RewardGraph:
    """The batched reward query for a FIXED number of envs captured once as a CUDA graph
    (Envs/vec_env/vec_pretext_normalize.py:82-101 is issued every rollout step with the same N =
    config.RLNumEnvs, 8 by default): a query then costs one graph launch instead of ~12 dependent
    kernel launches, which is what bounds the latency at small N.  Inputs are written into the static
    buffers (`images`, `goal_sounds` / `goal_feat_cached`, `env_reward`), `launch()` replays, and `out`
    is ONE flat device buffer [img_feat | goal_feat | dot | reward] so the caller reads everything
    back with a single D2H copy."""

    def __init__(self, eng, N, image_dtype=torch.uint8, fresh_goal=False):
        dev, D = eng.device, eng.D
        self.eng, self.N, self.fresh = eng, int(N), bool(fresh_goal)
        self.images = torch.zeros(N, 3, 96, 96, dtype=image_dtype, device=dev)
        self.goal_sounds = torch.zeros(N, eng.F, 40, dtype=torch.float32, device=dev) if fresh_goal else None
        self.goal_feat_cached = None if fresh_goal else torch.zeros(N, D, dtype=torch.float32, device=dev)
        self.env_reward = torch.zeros(N, dtype=torch.float32, device=dev)
        self.out = torch.zeros(N * (2 * D + 2), dtype=torch.float32, device=dev)
        self.img_feat = self.out[:N * D].view(N, D)
        self.goal_feat = self.out[N * D:2 * N * D].view(N, D)
        self.dot = self.out[2 * N * D:2 * N * D + N]
        self.reward = self.out[2 * N * D + N:]
        self._run()  # eager once: sizes the workspace and sets the kernels' attributes outside the capture
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()

    def _run(self):
        self.eng.reward(self.images, goal_sounds=self.goal_sounds, goal_feat_cached=self.goal_feat_cached,
                        env_reward=self.env_reward, out=(self.img_feat, self.goal_feat, self.dot, self.reward))

    def launch(self):
        self.graph.replay()
        return self.img_feat, self.goal_feat, self.dot, self.reward


def multistep_lr(base_lr, epoch, milestones, gamma):
    """MultiStepLR as built by utils.get_scheduler (utils.py:42-46), stepped once per epoch
    (VAR/pretext_VAR.py:72-73): LR in effect during `epoch`."""
    return base_lr * math.pow(gamma, sum(1 for m in milestones if epoch >= m))
