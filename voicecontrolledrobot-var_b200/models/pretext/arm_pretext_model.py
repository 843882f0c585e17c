"""Host mirror of models/pretext/arm_pretext_model.py (Kuka VAR encoders).

Parameter containers reproduce the reference module tree (names, shapes, construction order
and the two `torch.rand` shape probes of arm_pretext_model.py:44,50), so a seeded
construction yields the reference's initial weights and checkpoints load both ways.  The
nn layers are never called: forward/backward run in libvar_b200.so."""
import torch
import torch.nn as nn

from ...engine import KUKA
from .pretext_base import PretextNetBase


class Flatten(nn.Module):  # utils.py:9-11 (parameter-free placeholder, keeps Sequential indices)
    def forward(self, x):
        return x.view(x.size(0), -1)


def buildCNN(nn_module, config=None):
    chans = [3, 32, 32, 64, 64, 64]
    mods = []
    for i in range(5):
        mods += [nn.Conv2d(chans[i], chans[i + 1], 3, stride=2, padding=1), nn.ReLU()]
    nn_module.imgBranch = nn.Sequential(*mods, Flatten())


def buildSoundBranch(nn_module, config=None):
    mods = [nn.Conv2d(1, 32, (5, 40), stride=(2, 1)), nn.ReLU()]
    for _ in range(3):
        mods += [nn.Conv2d(32, 32, (3, 1), stride=(2, 1)), nn.ReLU()]
    nn_module.soundCNN = nn.Sequential(*mods, Flatten())


class VARPretextNet(PretextNetBase):
    KIND = KUKA

    def __init__(self, config):
        super().__init__()
        self.config = config
        if tuple(config.img_dim) != (3, 96, 96) or tuple(config.sound_dim) != (1, 100, 40):
            raise ValueError("Kuka VARPretextNet is built for img_dim (3,96,96), sound_dim (1,100,40)")
        buildCNN(self, config)
        buildSoundBranch(self, config)
        torch.rand((1, *config.img_dim))  # RNG parity with get_layer_output_shape (arm_pretext_model.py:44)
        self.imgCNN_outputShape = torch.Size([1, 576])
        self.imgTriplet = nn.Sequential(nn.Linear(576, 128), nn.ReLU(), nn.Linear(128, config.representationDim))
        torch.rand(*config.sound_dim)  # RNG parity with the sound-branch probe (arm_pretext_model.py:50)
        self.soundBranch_outputShape = torch.Size([1, 160])
        self.soundTriplet = nn.Sequential(nn.Linear(160, 128), nn.ReLU(), nn.Linear(128, config.representationDim))

    def forward(self, image, sound_positive, sound_negative, is_train=False):
        return self.VAR_forward(image, sound_positive, sound_negative, is_train=False)
