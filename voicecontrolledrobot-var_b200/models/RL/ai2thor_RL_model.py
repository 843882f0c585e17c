"""Host mirror of the image CNN of the iTHOR policy net (models/RL/ai2thor_RL_model.py:15-27,
`ai2thorNet_VAR.imgCNN`), SURVEY.md section 8 row f4.

`imgCNN` has the topology of the VAR image branch (models/pretext/ai2thor_pretext_model.py:15-27): five
3x3 convs with ReLU, four 2x2 max-pools, a stride-2 3x3 conv, Flatten -> [N, 128 * 3 * 3].  PPO evaluates it on
every rollout step next to the VAR reward query, on the same [N, 3, 96, 96] observation, so it is served by
the same kernels: the weights are loaded into the `imgBranch.*` slots of an iTHOR `VarEngine` (all other
tensors zero) and the flattened feature row is the engine's raw image feature.  Inference only: the rest of
`ai2thorNet_VAR` (MLPs, GRU, PPO losses and their backward pass) is a consumer of this output and stays with
the reference (section 8, out of scope).

The torch layers are a parameter container in the reference layout (`imgCNN.<idx>.weight|bias`, same
Sequential indices), so `load_state_dict(policy.base.imgCNN.state_dict(), ...)` works unchanged."""
import torch
import torch.nn as nn

from ...engine import ITHOR, VarEngine
from ..pretext._layers import conv_stack
from ..pretext.ai2thor_pretext_model import IMG_SPEC


class ai2thorImgCNN(nn.Module):
    OUT_DIM = 128 * 3 * 3

    def __init__(self):
        super().__init__()
        self.imgCNN = conv_stack(IMG_SPEC)  # ai2thor_RL_model.py:15-27
        self._engine = None
        self._engine_key = None

    @classmethod
    def from_policy_state_dict(cls, sd, prefix="imgCNN."):
        """Build from any state dict that holds the policy net's `imgCNN.*` tensors (e.g. `actor_critic.base`)."""
        m = cls()
        m.imgCNN.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)})
        return m

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != device:
            self._engine = VarEngine(ITHOR, 600, 3, device)
            self._engine_key = None
        key = tuple((p.data_ptr(), p._version) for p in self.imgCNN.parameters())
        if key != self._engine_key:
            mine = {"imgBranch." + k: v for k, v in self.imgCNN.state_dict().items()}
            sd = {}
            for name, shape, _, _ in self._engine.tensors:
                sd[name] = mine[name] if name in mine else torch.zeros(shape)
            self._engine.load_state_dict(sd)
            self._engine_key = key
        return self._engine

    @torch.no_grad()
    def forward(self, image):
        """image: CUDA [N, 3, 96, 96], uint8 (scaled by 1/255 on the device, as processAI2Thor feeds the policy:
        Envs/vec_env/vec_pretext_normalize.py:135) or float32 already in [0, 1]  ->  [N, 1152] float32."""
        if not image.is_cuda:
            raise RuntimeError("ai2thorImgCNN (B200) runs on CUDA tensors only; there is no CPU fallback")
        if image.dtype not in (torch.uint8, torch.float32):
            image = image.float()
        eng = self._get_engine(image.device)
        _, raw, _, _ = eng.forward(image.contiguous(), None, train=False, want_raw=True)
        return raw
