"""Drop-in for the reference's ``src/optimization.py`` (abstract base ``Optim``).

Mirrors ``/root/reference/src/optimization.py:10-42``: holds the device, the L2 loss,
the renderer handle; ``gradient`` makes leaf tensors; ``save_loss`` plots the curve
(matplotlib is optional here — without it the curve is written as text next to the
requested image path, so the per-100-epoch dump cadence still leaves a record).
"""

from __future__ import annotations

import numpy as np
import torch as th


class Optim:
    def __init__(self, device, renderer_obj):
        self.device = device
        self.eps = 1e-4
        self.loss_l2 = th.nn.MSELoss().to(device)
        self.renderer_obj = renderer_obj

    def gradient(self, parameters):
        """Turn tensors into autograd leaves (optimization.py:17-23)."""
        if isinstance(parameters, list):
            for i, p in enumerate(parameters):
                parameters[i] = p.detach().requires_grad_(True)
            return parameters
        return parameters.detach().requires_grad_(True)

    def load_targets(self, targets):
        raise NotImplementedError("Should be implemented in derived class!")

    def compute_image_loss(self, predicts):
        return self.loss_l2(predicts, self.targets)

    def optim(self, epochs, lr, svbrdf_obj, optim_light):
        raise NotImplementedError("Should be implemented in derived class!")

    def save_loss(self, losses, labels, save_dir, N):
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            with open(str(save_dir) + ".txt", "w") as f:
                for curve, label in zip(losses, labels):
                    f.write(label + " (log1p): " + " ".join(f"{v:.6g}" for v in np.log1p(np.asarray(curve))) + "\n")
            return
        plt.figure(figsize=(8, 4))
        for curve, label in zip(losses, labels):
            plt.plot(np.log1p(curve), label=label)
        plt.xlim(0, N)
        plt.legend()
        plt.title("log(1+loss)")
        plt.savefig(save_dir)
        plt.close()
