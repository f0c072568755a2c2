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
        # The reference registers forward hooks on the four layers and runs the WHOLE network (descriptor.py:25-37,41), i.e.
        # the 6 convolutions past relu4_2 (conv4_3 ... conv5_4) whose outputs nobody reads.  Here the layers are walked
        # explicitly up to the last tapped one: same four feature maps, 28 % fewer multiply-adds per call.
        self.taps = tuple(HOOK_LAYERS)
        self.trunk = self.net[:max(self.taps) + 1]
        self.weights = weights
        self.mean, self.std = MEAN, STD
        self._mean_t = th.tensor(MEAN, dtype=th.float32, device=device).view(1, 3, 1, 1)
        self._std_t = th.tensor(STD, dtype=th.float32, device=device).view(1, 3, 1, 1)

    def feature_maps(self, x):
        """Activations of the tapped layers (r11, r12, r32, r42), in network order."""
        maps = []
        for index, layer in enumerate(self.trunk):
            x = layer(x)
            if index in self.taps:
                maps.append(x)
        return maps

    def compute_feature_vector(self, x, is_gram=False):
        """One flat descriptor per batch (descriptor.py:39-60): the weighted tapped activations, or — ``is_gram`` — the
        weighted Gram matrices of the (image x channel) rows of each tap, concatenated."""
        parts = []
        for weight, fmap in zip(self.weights, self.feature_maps(x)):
            if is_gram:
                rows = fmap.reshape(fmap.shape[0] * fmap.shape[1], -1)
                fmap = rows @ rows.t() / rows.shape[1]
            parts.append(fmap.reshape(-1) * weight)
        return th.cat(parts)

    def normalize(self, im):
        return (im - self._mean_t) / self._std_t              # descriptor.py:65-75, without the per-image loop

    def load(self, im):
        with th.no_grad():
            self.im_feature = self.compute_feature_vector(self.normalize(im))

    def forward(self, x):
        return self.forward_normalized(self.normalize(x))

    def forward_normalized(self, x_norm):
        return self.criterion(self.compute_feature_vector(x_norm), self.im_feature)
