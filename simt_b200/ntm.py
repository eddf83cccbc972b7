"""The noise-transition "T layer" and the convex-hull weight layer.

Same constructor signatures, parameter names (``NTM``, ``weight``) and forward maths as
model/deeplab_multi.py:244-263 (``sig_NTM``) and :265-286 (``sig_W``) of the reference.
These stay PyTorch on purpose (a few hundred elements; autograd
carries dT -> dNTM); what changes is that buffers are registered, so ``.cuda()`` / ``.to()``
moves the module instead of ``forward`` hard-calling ``.cuda()``, and the class prior is
found next to the package instead of at ``../ClassDist`` relative to the cwd.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEFAULT_CLASS_DIST = os.path.join(_DATA, "ClassDist_bapa.npy")


class sig_NTM(nn.Module):
    def __init__(self, num_classes, open_classes=0, init=None, class_dist=None):
        super().__init__()
        ck = num_classes + open_classes
        self.NTM = nn.Parameter(torch.ones(ck, num_classes))
        nn.init.kaiming_normal_(self.NTM, mode="fan_out", nonlinearity="relu")
        prior = torch.cat([torch.eye(num_classes, num_classes), torch.zeros(open_classes, num_classes)], 0)
        if class_dist is None:
            class_dist = DEFAULT_CLASS_DIST
        if isinstance(class_dist, (str, os.PathLike)):
            class_dist = np.load(class_dist)
        dist = torch.from_numpy(np.tile(np.asarray(class_dist, dtype=np.float64), (ck, 1))).to(torch.float32)
        self.register_buffer("Identity_prior", prior, persistent=False)
        self.register_buffer("Class_dist", dist, persistent=False)

    def forward(self):
        T = torch.sigmoid(self.NTM)
        T = T.mul(self.Class_dist.detach()) + self.Identity_prior.detach()
        return F.normalize(T, p=1, dim=1)


class sig_W(nn.Module):
    def __init__(self, num_classes, open_classes=0):
        super().__init__()
        self.classes = num_classes + open_classes
        init = 1.0 / (self.classes - 1.0)
        self.weight = nn.Parameter(init * torch.ones(self.classes, self.classes))
        self.register_buffer("identity", torch.zeros(self.classes, self.classes) - torch.eye(self.classes),
                             persistent=False)

    def forward(self):
        with torch.no_grad():
            self.weight.fill_diagonal_(-10000.0)
        w = torch.softmax(self.weight, dim=1)
        return self.identity.detach() + w
