"""Host mirror of pretext.py::Pretext for the hot path: model construction, checkpoint load and
the `run()` -> `trainRepresentation()` hand-off (pretext.py:23-28, 102-111, 292-325).

Data collection (pretext.py:31-100) and the matplotlib / t-SNE plotting (pretext.py:147-290)
drive the pybullet / Unity simulators and a GUI; they are callers of the path, not part of it,
and stay with the reference."""
import os

import torch


class Pretext(object):
    def __init__(self, config):
        self.config = config
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 VAR path needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        print("Using device:", self.device)
        self.pretextModel = None

    def loadPretextModel(self):
        """pretext.py:102-111."""
        weight_path = self.config.pretextModelLoadDir
        if self.pretextModel is None:
            self.pretextModel = self.config.pretextModel(self.config)
        self.pretextModel.load_state_dict(torch.load(weight_path, map_location="cpu"))
        self.pretextModel.to(self.device).eval()
        print('Load weights for pretextModel from', weight_path)

    def collectPretextData(self, fileName=None):
        raise NotImplementedError("data collection runs the simulators: use the reference's pretext.py:31-100")

    def project2representation_with_ground_truth(self, data_generator, project_for='plot', req_grad=False):
        """pretext.py:147-203 without the image dump / cv2 side effects: every batch of the generator
        goes through the encoders as one batched device call; returns {'img', 'sound'} arrays of
        [n, representationDim + 1] (embedding, ground-truth label) for plotting / medoid code."""
        import numpy as np
        if project_for != 'plot':
            raise NotImplementedError
        feat_point = {'img': [], 'sound': [], 'lastBatchNum': -1}
        limit = getattr(self.config, "plotNumBatch", None)
        with torch.set_grad_enabled(req_grad):
            for n, data in enumerate(data_generator):
                if limit is not None and n > limit:
                    break
                img, sp, gt = data[0], data[1], data[3]
                features = self.pretextModel(img.to(self.device), sp.float().to(self.device), None)
                g = gt.cpu().numpy()[:, None]
                feat_point['img'].append(np.concatenate([features['image_feat'].detach().cpu().numpy(), g], axis=1))
                feat_point['sound'].append(np.concatenate([features['sound_feat_positive'].detach().cpu().numpy(), g],
                                                          axis=1))
        feat_point['img'] = np.concatenate(feat_point['img'], axis=0)
        feat_point['sound'] = np.concatenate(feat_point['sound'], axis=0)
        return feat_point

    def plotRepresentation(self, data_generator):
        raise NotImplementedError("plotting stays with the reference (pretext.py:205-264)")

    def run(self):
        torch.manual_seed(self.config.pretextEnvSeed)
        torch.cuda.manual_seed_all(self.config.pretextEnvSeed)
        if getattr(self.config, "pretextCollection", False):
            # pretext.py:297-304 runs the simulators here; that stage stays with the reference, so the
            # triplet files under config.pretextDataDir must already exist
            import warnings
            warnings.warn("config.pretextCollection is set: data collection drives the pybullet / Unity simulators "
                          "and is not part of this package -- using the triplet files already on disk")
        if self.config.pretextTrain:
            self.pretextModel = self.config.pretextModel(self.config).to(self.device)
            if self.config.pretextModelFineTune:
                self.loadPretextModel()
            os.makedirs(self.config.pretextModelSaveDir, exist_ok=True)
            self.trainRepresentation(epoch=self.config.pretextEpoch, lr=self.config.pretextLR, start_ep=0, plot=False)
        elif not getattr(self.config, "pretextCollection", False):
            self.loadPretextModel()

    def trainRepresentation(self, epoch, lr, start_ep=0, plot=False):
        raise NotImplementedError("Please Implement this method")
