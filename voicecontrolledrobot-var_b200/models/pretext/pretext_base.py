"""Host mirror of models/pretext/pretext_base.py::PretextNetBase.

`VAR_forward` keeps the reference contract (pretext_base.py:10-42): any of the three inputs
may be None, the positive sound is encoded only when it is present and not all-`inf`
(otherwise `self.cached_sound` is reused), and the result is the same 7-key dict.  The
arithmetic runs in libvar_b200.so through `VarEngine`; parameters stay ordinary
`nn.Parameter`s in the reference layout so `state_dict()` / optimisers / checkpoints are
interchangeable with the reference module.
"""
import torch
import torch.nn as nn

from ...engine import VarEngine


class _VarForward(torch.autograd.Function):
    """(image, sounds, *params) -> (image_feat, image_raw, sound_feat, sound_raw) with the backward
    pass of the CUDA engine wired into autograd (VAR/pretext_VAR.py:64-68 calls loss.backward())."""

    @staticmethod
    def forward(ctx, module, image, sounds, *params):
        eng = module._engine
        img_feat, img_raw, snd_feat, snd_raw = eng.forward(image, sounds, train=True)
        ctx.module = module
        ctx.has = (image is not None, sounds is not None)
        outs = tuple(o if o is not None else torch.empty(0, device=eng.device)
                     for o in (img_feat, img_raw, snd_feat, snd_raw))
        ctx.mark_non_differentiable(outs[1], outs[3])
        return outs

    @staticmethod
    def backward(ctx, d_img_feat, _d_img_raw, d_snd_feat, _d_snd_raw):
        module = ctx.module
        eng = module._engine
        eng.zero_grad()
        di = d_img_feat.contiguous().float() if (ctx.has[0] and d_img_feat is not None) else None
        ds = d_snd_feat.contiguous().float() if (ctx.has[1] and d_snd_feat is not None) else None
        eng.backward(di, ds)
        g = eng.grad_dict()
        grads = tuple(g[name].view_as(p) for name, p in module._ordered_params())
        return (None, None, None) + grads


class PretextNetBase(nn.Module):
    KIND = None  # set by subclasses (engine.KUKA / engine.ITHOR)

    def __init__(self):
        super().__init__()
        self._engine = None
        self._engine_key = None
        self.cached_sound = None

    # ----------------------------------------------------------------- engine plumbing
    def _ordered_params(self):
        return [(k, v) for k, v in self.named_parameters()]

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != device:
            self._engine = VarEngine(self.KIND, self.config.sound_dim[1], self.config.representationDim, device)
            self._engine_key = None
        key = tuple((p.data_ptr(), p._version) for _, p in self._ordered_params())
        if key != self._engine_key:  # parameters changed (optimizer step, load_state_dict, .to())
            self._engine.load_state_dict({k: v for k, v in self._ordered_params()})
            self._engine_key = key
        return self._engine

    def sync_from_engine(self):
        """Copy the engine's (trained) packed weights back into the nn.Parameters."""
        sd = self._engine.state_dict()
        with torch.no_grad():
            for k, p in self._ordered_params():
                p.copy_(sd[k].view_as(p))
        self._engine_key = tuple((p.data_ptr(), p._version) for _, p in self._ordered_params())

    @staticmethod
    def _prep_image(image):
        if image.dim() != 4 or image.shape[1] < 3:
            raise ValueError("image must be [B, >=3, 96, 96]")
        if image.shape[1] != 3:
            image = image[:, :3, :, :]  # pretext_base.py:22
        if image.dtype not in (torch.uint8, torch.float32):
            image = image.float()
        return image.contiguous()

    def VAR_forward(self, image, sound_positive, sound_negative, is_train=False):
        ref = next((t for t in (image, sound_positive, sound_negative) if t is not None), None)
        if ref is None:
            return {'image_feat': None, 'sound_feat_positive': self.cached_sound, 'sound_feat_negative': None,
                    'image_BCE': None, 'sound_BCE': None, 'image_feat_raw': None, 'pos_sound_raw': None}
        if not ref.is_cuda:
            raise RuntimeError("VARPretextNet (B200) runs on CUDA tensors only; there is no CPU fallback")
        eng = self._get_engine(ref.device)
        encode_pos = sound_positive is not None and (not torch.isinf(sound_positive).all())
        parts = []
        if encode_pos:
            parts.append(sound_positive)
        if sound_negative is not None:
            parts.append(sound_negative)
        sounds = None
        if parts:
            F = self.config.sound_dim[1]
            parts = [p.reshape(-1, F, 40).float() for p in parts]
            sounds = (parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)).contiguous()
        img = self._prep_image(image) if image is not None else None
        params = [p for _, p in self._ordered_params()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            img_feat, img_raw, snd_feat, snd_raw = _VarForward.apply(self, img, sounds, *params)
        else:
            img_feat, img_raw, snd_feat, snd_raw = eng.forward(img, sounds, train=False)
        image_feat = image_feat_raw = None
        if image is not None:
            image_feat, image_feat_raw = img_feat, img_raw
        pos_sound_raw = sound_feat_negative = None
        off = 0
        if encode_pos:
            n = sound_positive.shape[0]
            pos_sound_raw, self.cached_sound = snd_raw[:n], snd_feat[:n]
            off = n
        sound_feat_positive = self.cached_sound
        if sound_negative is not None:
            sound_feat_negative = snd_feat[off:off + sound_negative.shape[0]]
        return {'image_feat': image_feat, 'sound_feat_positive': sound_feat_positive,
                'sound_feat_negative': sound_feat_negative, 'image_BCE': None, 'sound_BCE': None,
                'image_feat_raw': image_feat_raw, 'pos_sound_raw': pos_sound_raw}
