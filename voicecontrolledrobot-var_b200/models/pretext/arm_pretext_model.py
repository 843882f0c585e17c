"""Host mirror of models/pretext/arm_pretext_model.py (Kuka VAR encoders).

The module tree is rebuilt from layer specs (`_layers.py`) with the reference's names, shapes and
construction order, and the constructor consumes the global RNG exactly where the reference's two
shape probes do (`get_layer_output_shape` at arm_pretext_model.py:44 and the sound-branch probe at
:50), so `torch.manual_seed(s); VARPretextNet(config)` yields the reference's initial weights and
checkpoints load both ways.  The torch layers are parameter containers only: forward / backward
run in libvar_b200.so."""
import torch

from ...engine import KUKA
from ._layers import conv_stack, mlp_head
from .pretext_base import PretextNetBase

# image: five 3x3 stride-2 convs 96 -> 3;  sound: a 5x40 conv over the MFCC, then three 3x1 convs
IMG_SPEC = [("c", cin, cout, 3, 2, 1) for cin, cout in ((3, 32), (32, 32), (32, 64), (64, 64), (64, 64))]
SND_SPEC = [("c", 1, 32, (5, 40), (2, 1), 0)] + [("c", 32, 32, (3, 1), (2, 1), 0)] * 3
IMG_FLAT, SND_FLAT, HIDDEN = 64 * 3 * 3, 32 * 5, 128


class VARPretextNet(PretextNetBase):
    KIND = KUKA

    def __init__(self, config):
        super().__init__()
        self.config = config
        if tuple(config.img_dim) != (3, 96, 96) or tuple(config.sound_dim) != (1, 100, 40):
            raise ValueError("Kuka VARPretextNet is built for img_dim (3,96,96), sound_dim (1,100,40)")
        self.imgBranch = conv_stack(IMG_SPEC)
        self.soundCNN = conv_stack(SND_SPEC)
        torch.rand((1, *config.img_dim))       # RNG parity: image-branch shape probe
        self.imgCNN_outputShape = torch.Size([1, IMG_FLAT])
        self.imgTriplet = mlp_head([IMG_FLAT, HIDDEN, config.representationDim])
        torch.rand(*config.sound_dim)          # RNG parity: sound-branch shape probe
        self.soundBranch_outputShape = torch.Size([1, SND_FLAT])
        self.soundTriplet = mlp_head([SND_FLAT, HIDDEN, config.representationDim])

    def forward(self, image, sound_positive, sound_negative, is_train=False):
        return self.VAR_forward(image, sound_positive, sound_negative, is_train=False)
