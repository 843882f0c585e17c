"""Parameter-container builders for the VAR encoders.

The product never calls these torch layers: they exist so that `state_dict()` keys, shapes,
registration order and seeded initial values equal the reference modules'
(models/pretext/arm_pretext_model.py, ai2thor_pretext_model.py).  A branch is described by a
compact spec -- the same table the C++ layer plan in csrc/net.cu is written from -- and expanded
into an `nn.Sequential` whose indices match the reference (activation / pooling / flatten modules
occupy the same slots)."""
import torch.nn as nn


def conv_stack(spec, flatten=True):
    """spec: list of ("c", cin, cout, kernel, stride, padding) | ("p",) for MaxPool2d(2, 2).
    Every conv is followed by a ReLU slot, as in the reference."""
    mods = []
    for item in spec:
        if item[0] == "c":
            _, cin, cout, k, s, p = item
            mods += [nn.Conv2d(cin, cout, k, stride=s, padding=p), nn.ReLU()]
        elif item[0] == "p":
            mods.append(nn.MaxPool2d(2, stride=2))
        else:
            raise ValueError(item)
    if flatten:
        mods.append(nn.Flatten())
    return nn.Sequential(*mods)


def mlp_head(dims):
    """Linear(d0, d1) ReLU Linear(d1, d2) ReLU ... Linear(d_{n-1}, d_n) -- no ReLU after the last."""
    mods = []
    for i in range(len(dims) - 1):
        mods.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            mods.append(nn.ReLU())
    return nn.Sequential(*mods)
