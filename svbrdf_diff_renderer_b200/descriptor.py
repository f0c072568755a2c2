"""Drop-in for the reference's ``src/descriptor.py`` (class ``VGGLoss``), the second consumer of ``Microfacet.eval``.

Mirrors ``/root/reference/src/descriptor.py:7-79``: VGG19 feature network in eval mode with max-pooling replaced by
average pooling, forward hooks on layers 1, 3, 13, 22 (r11, r12, r32, r42), features flattened, weighted and
concatenated, MSE against the features of the loaded target images; inputs are normalised per channel first.

Differences, all additive:
* the network can be injected (``net=``) or built without pretrained weights (``pretrained=None``): the reference
  downloads ``vgg19(weights='DEFAULT')``, which needs network access;
* ``normalize`` is one broadcast expression instead of a Python loop over images — the same two float32 operations per
  element (``sub`` then ``div``), so the result is bit-identical;
* ``forward_normalized`` takes an already normalised batch: ``Microfacet.eval_normalized`` produces it inside the
  render kernel together with the L2 image loss.
The convolutions themselves are torch/cuDNN library code (out of scope of the hand-written path).
"""

from __future__ import annotations

import numpy as np
import torch as th

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.255)          # sic: descriptor.py:69 (the usual ImageNet value is 0.225)
HOOK_LAYERS = (1, 3, 13, 22)         # descriptor.py:25: r11, r12, r32, r42


class VGGLoss(th.nn.Module):
    def __init__(self, device, weights=np.array([1, 1, 1, 1]) / 4, net=None, pretrained="DEFAULT"):
        super().__init__()
        self.criterion = th.nn.MSELoss().to(device)
        if net is None:
            from torchvision.models import vgg19
            net = vgg19(weights=pretrained).features
        self.net = net.to(device)
        self.net.eval()
        for p in self.net.parameters():
            p.requires_grad_(False)
        for i, x in enumerate(self.net):                      # descriptor.py:16-19
            if isinstance(x, th.nn.MaxPool2d):
                self.net[i] = th.nn.AvgPool2d(kernel_size=2)
        self.outputs = []

        def hook(module, input, output):
            self.outputs.append(output)

        for i in HOOK_LAYERS:
            self.net[i].register_forward_hook(hook)
        self.weights = weights
        self.mean, self.std = MEAN, STD
        self._mean_t = th.tensor(MEAN, dtype=th.float32, device=device).view(1, 3, 1, 1)
        self._std_t = th.tensor(STD, dtype=th.float32, device=device).view(1, 3, 1, 1)

    def compute_feature_vector(self, x, is_gram=False):
        self.outputs = []
        self.net(x)
        result = []
        for i, feature in enumerate(self.outputs):
            if is_gram:
                n, f, s1, s2 = feature.shape
                s = s1 * s2
                feature = feature.view((n * f, s))
                result.append((th.mm(feature, feature.t()) / s).flatten() * self.weights[i])
            else:
                result.append(feature.flatten() * self.weights[i])
        return th.cat(result)

    def normalize(self, im):
        return (im - self._mean_t) / self._std_t              # descriptor.py:65-75, without the per-image loop

    def load(self, im):
        with th.no_grad():
            self.im_feature = self.compute_feature_vector(self.normalize(im))

    def forward(self, x):
        return self.forward_normalized(self.normalize(x))

    def forward_normalized(self, x_norm):
        return self.criterion(self.compute_feature_vector(x_norm), self.im_feature)
