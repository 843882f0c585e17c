"""Host mirror of models/pretext/ai2thor_pretext_model.py (iTHOR VAR encoders, incl. the
bidirectional GRU(448 -> 512)).  Same parameter tree and registration order as the
reference (imgBranch, rnn, cnn, imgTriplet, soundTriplet); unlike the reference constructor
(line 46) nothing here requires CUDA at construction time.  The nn layers are parameter
containers only: forward/backward run in libvar_b200.so."""
import torch
import torch.nn as nn

from ...engine import ITHOR
from .pretext_base import PretextNetBase


def buildSoundBranch(nn_module, config=None):
    nn_module.rnn = nn.GRU(input_size=64 * 7, hidden_size=512, batch_first=True, bidirectional=True)
    nn_module.cnn = nn.Sequential(
        nn.Conv2d(1, 64, (11, 11), stride=(2, 2), padding=(5, 5)), nn.ReLU(),
        nn.Conv2d(64, 64, (11, 5), stride=(2, 2), padding=(5, 5)), nn.ReLU(),
        nn.Conv2d(64, 64, (7, 3), stride=(2, 2), padding=(1, 1)), nn.ReLU())


def buildCNN(nn_module, config=None):
    nn_module.imgBranch = nn.Sequential(
        nn.Conv2d(3, 32, 3, stride=1, padding=1), nn.ReLU(),
        nn.Conv2d(32, 32, 3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(2, stride=2),
        nn.Conv2d(32, 64, 3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(2, stride=2),
        nn.Conv2d(64, 64, 3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(2, stride=2),
        nn.Conv2d(64, 128, 3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(2, stride=2),
        nn.Conv2d(128, 128, 3, stride=2, padding=1), nn.ReLU(), nn.Flatten())


class VARPretextNet(PretextNetBase):
    KIND = ITHOR

    def __init__(self, config):
        super().__init__()
        self.config = config
        if tuple(config.img_dim) != (3, 96, 96) or tuple(config.sound_dim) != (1, 600, 40):
            raise ValueError("iTHOR VARPretextNet is built for img_dim (3,96,96), sound_dim (1,600,40)")
        self.zero_feat = torch.zeros((config.representationDim,))
        buildCNN(self, config)
        buildSoundBranch(self, config)
        self.imgTriplet = nn.Sequential(nn.Linear(128 * 9, 128), nn.ReLU(),
                                        nn.Linear(128, config.representationDim))
        self.soundTriplet = nn.Sequential(nn.Linear(2 * 512, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(),
                                          nn.Linear(64, config.representationDim))

    def forward(self, image, sound_positive, sound_negative, is_train=False):
        return self.VAR_forward(image, sound_positive, sound_negative, is_train)
