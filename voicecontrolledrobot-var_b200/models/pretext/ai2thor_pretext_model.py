"""Host mirror of models/pretext/ai2thor_pretext_model.py (iTHOR VAR encoders with the
bidirectional GRU(448 -> 512) sound branch).

Same parameter tree and registration order as the reference (imgBranch, rnn, cnn, imgTriplet,
soundTriplet), rebuilt from layer specs; unlike the reference constructor (line 46) nothing here
needs CUDA at construction time.  The torch layers are parameter containers only: forward /
backward run in libvar_b200.so."""
import torch
import torch.nn as nn

from ...engine import ITHOR
from ._layers import conv_stack, mlp_head
from .pretext_base import PretextNetBase

_C3 = lambda cin, cout, stride=1: ("c", cin, cout, 3, stride, 1)
IMG_SPEC = [_C3(3, 32), _C3(32, 32), ("p",), _C3(32, 64), ("p",), _C3(64, 64), ("p",), _C3(64, 128), ("p",),
            _C3(128, 128, 2)]
SND_SPEC = [("c", 1, 64, (11, 11), 2, 5), ("c", 64, 64, (11, 5), 2, 5), ("c", 64, 64, (7, 3), 2, 1)]
GRU_IN, GRU_HIDDEN, GRU_STEPS = 64 * 7, 512, 73


class VARPretextNet(PretextNetBase):
    KIND = ITHOR

    def __init__(self, config):
        super().__init__()
        self.config = config
        if tuple(config.img_dim) != (3, 96, 96) or tuple(config.sound_dim) != (1, 600, 40):
            raise ValueError("iTHOR VARPretextNet is built for img_dim (3,96,96), sound_dim (1,600,40)")
        self.zero_feat = torch.zeros((config.representationDim,))
        self.imgBranch = conv_stack(IMG_SPEC)
        self.rnn = nn.GRU(input_size=GRU_IN, hidden_size=GRU_HIDDEN, batch_first=True, bidirectional=True)
        self.cnn = conv_stack(SND_SPEC, flatten=False)
        self.imgTriplet = mlp_head([128 * 9, 128, config.representationDim])
        self.soundTriplet = mlp_head([2 * GRU_HIDDEN, 128, 64, config.representationDim])

    def forward(self, image, sound_positive, sound_negative, is_train=False):
        return self.VAR_forward(image, sound_positive, sound_negative, is_train)
