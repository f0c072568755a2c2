"""ORACLE (test infrastructure, not product code): CPU restatement of the combined loss of the reference's second
consumer of ``Microfacet.eval``:

    loss = MSELoss(rendered, targets) + 0.1 * VGGLoss(rendered)        /root/reference/src/materialgan.py:141-147
    VGGLoss                                                            /root/reference/src/descriptor.py:7-79

``VGGLoss`` restated with torch CPU ops, op for op: per-image ``torchvision.transforms.Normalize`` loop
(descriptor.py:65-75: ``sub_(mean).div_(std)`` with mean (0.485, 0.456, 0.406), std (0.229, 0.224, 0.255)), VGG19
``features`` in eval mode with MaxPool2d -> AvgPool2d(2) (descriptor.py:16-19), hooks on layers 1, 3, 13, 22
(descriptor.py:25), flattened features times weights[i] concatenated (descriptor.py:39-58), MSE against the features
of the target images (descriptor.py:60-79).

The pretrained VGG19 weights are a download (descriptor.py:13) that is not available offline: the oracle, the
reference run that generated tests/golden/features_*.npz (oracle/make_golden_features.py, which patches only the
``vgg19(weights=...)`` constructor call) and the GPU tests all build ``vgg19(weights=None)`` after
``torch.manual_seed(seed)`` — same architecture and arithmetic, seeded random weights.
"""

from __future__ import annotations

import numpy as np
import torch as th

from oracle import torch_port as tp

MEAN = [0.485, 0.456, 0.406]
STD = [0.229, 0.224, 0.255]
LAYERS = (1, 3, 13, 22)


def seeded_vgg_features(seed: int):
    from torchvision.models import vgg19
    th.manual_seed(seed)
    net = vgg19(weights=None).features
    net.eval()
    for i, x in enumerate(net):
        if isinstance(x, th.nn.MaxPool2d):
            net[i] = th.nn.AvgPool2d(kernel_size=2)
    for p in net.parameters():
        p.requires_grad_(False)
    return net


def normalize(im):
    out = im.clone()
    mean = th.tensor(MEAN, dtype=im.dtype).view(3, 1, 1)
    std = th.tensor(STD, dtype=im.dtype).view(3, 1, 1)
    for i in range(im.shape[0]):
        out[i] = (im[i] - mean) / std
    return out


def feature_vector(net, x, weights=np.array([1, 1, 1, 1]) / 4):
    feats, k = [], 0
    for i, layer in enumerate(net):
        x = layer(x)
        if i in LAYERS:
            feats.append(x.flatten() * weights[k])
            k += 1
    return th.cat(feats)


def vgg_loss(net, rendered, target_feature):
    return th.nn.functional.mse_loss(feature_vector(net, normalize(rendered)), target_feature)


def combined_loss_and_grad(scene, tex, targets, net, feature_weight=0.1):
    """(image loss, weighted feature loss, d(total)/d tex) for ``tex`` [1,9,R,R] already in [-1,1]."""
    with th.no_grad():
        tfeat = feature_vector(net, normalize(targets))
    t = tex.clone().requires_grad_(True)
    img = tp.shade(scene, t)
    l2 = tp.l2_loss(img, targets)
    lf = vgg_loss(net, img, tfeat) * feature_weight
    (l2 + lf).backward()
    return float(l2), float(lf), t.grad.detach()
