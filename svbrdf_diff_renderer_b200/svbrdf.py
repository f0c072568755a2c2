"""Drop-in for the reference's ``src/svbrdf.py``: ``SvbrdfOptim`` (the per-pixel optimiser)
and ``SvbrdfIO`` (JSON config + PNG ingest/egress).

``SvbrdfOptim.optim`` keeps the signature and side effects of
``/root/reference/src/svbrdf.py:44-83`` — Adam(lr, betas=(0.9, 0.999)) on the unclamped
``[1,9,R,R]`` parameter, loss = MSE(render(clamp(textures)), targets), dumps at epoch 1, every
100th and the last — but the loop body (svbrdf.py:60-71: clamp -> eval -> MSE -> zero_grad ->
backward -> Adam.step) is ONE fused kernel launch per epoch (``svbrdf_l2_adam_run``): the
rendered image, the autograd graph and the gradient tensor are never materialised, and the
per-epoch loss stays on the device until a dump point (the reference syncs five times per
epoch for its progress bar, svbrdf.py:64-66).
"""

from __future__ import annotations

import ctypes
import json
from datetime import datetime

import numpy as np
import torch as th

from . import _native as nv
from . import maps
from .imageio import imread, imread_raw, imwrite, imwrite_u8, img9to1, tex4to1
from .microfacet import _log
from .optimization import Optim


class SvbrdfOptim(Optim):
    def __init__(self, device, renderer_obj):
        super().__init__(device, renderer_obj)
        self.res = renderer_obj.res
        self.losses = []

    # ---- initialisers (svbrdf.py:20-39) ----------------------------------------------------------
    def init_from_tex(self, textures):
        self.textures = self.gradient(textures.to(device=self.device, dtype=th.float32).contiguous())

    def init_from_const(self, dif=0.5, spe=0.04, rgh=0.2):
        shape = lambda c: (1, c, self.res, self.res)  # noqa: E731
        full = lambda c, v: th.full(shape(c), v * 2 - 1, dtype=th.float32, device=self.device)  # noqa: E731
        normal = th.zeros(shape(2), dtype=th.float32, device=self.device)
        self.textures = self.gradient(th.cat((full(3, dif), normal, full(1, rgh), full(3, spe)), 1))

    def init_from_randn(self, dif=0.5, spe=0.04, rgh=0.2):
        rn = lambda c: th.randn(1, c, self.res, self.res, device=self.device)  # noqa: E731
        normal = (rn(2) / 4).clamp(-1, 1)
        diffuse = (rn(3) / 8 + dif).clamp(0, 1) * 2 - 1
        specular = (rn(3) / 32 + spe).clamp(0, 1) * 2 - 1
        roughness = (rn(1) / 16 + rgh).clamp(0, 1) * 2 - 1
        self.textures = self.gradient(th.cat((diffuse, normal, roughness, specular), 1))

    def load_targets(self, targets):
        """``[N,3,R,R]`` float32 (what ``load_images_th`` returns), uint8 (the PNG bytes) or float16."""
        if targets.dtype not in (th.float32, th.uint8, th.float16):
            targets = targets.float()
        self.targets = targets.to(self.device).contiguous()

    def compute_image_loss(self, predicts):
        tg = self.targets if self.targets.dtype == th.float32 else (self.targets.float() / 255 if self.targets.dtype == th.uint8 else self.targets.float())
        return self.loss_l2(predicts, tg)

    # ---- the loop (svbrdf.py:44-83) --------------------------------------------------------------
    def optim(self, epochs, lr, svbrdf_obj, optim_light, fused=True, progress=True, read_back=True):
        """Run ``epochs`` Adam iterations.  ``fused=False`` takes the mode-B route instead
        (native render fwd/bwd under autograd + torch.optim.Adam), the path any non-L2 loss uses.
        ``read_back=False`` (no dumps, no progress bar): nothing is synchronised — the loss curve is returned as a device
        tensor and ``self.losses`` is filled lazily, so a batch driver can enqueue many materials back to back."""
        dump = svbrdf_obj is not None and hasattr(svbrdf_obj, "optimize_dir")
        tmp_dir = None
        if dump:
            stamp = str(datetime.now()).replace(" ", "-").replace(":", "-").replace(".", "-")
            tmp_dir = svbrdf_obj.optimize_dir / "tmp" / stamp
            tmp_dir.mkdir(parents=True, exist_ok=True)

        r = self.renderer_obj
        pw = None
        if optim_light:
            src = svbrdf_obj.cl[2] if svbrdf_obj is not None else r._pow
            pw = src.detach().to(device=self.device, dtype=th.float32).clone().contiguous()
            if svbrdf_obj is not None:
                svbrdf_obj.cl[2] = pw
            r.update_light(pw)
        if not fused:
            return self._optim_autograd(epochs, lr, svbrdf_obj, pw, tmp_dir)

        tex = nv.dev_f32(self.textures.data, "textures")
        tgt = self.targets
        if tuple(tgt.shape) != (r.n_of_imgs, 3, self.res, self.res):
            raise RuntimeError(f"targets must be [{r.n_of_imgs},3,{self.res},{self.res}], got {tuple(tgt.shape)}")
        dtype_code = nv.target_dtype_code(tgt)
        m, v = th.zeros_like(tex), th.zeros_like(tex)
        pow_state = th.zeros(6, dtype=th.float32, device=self.device) if optim_light else None
        curve = th.zeros(max(epochs, 1), dtype=th.float32, device=self.device)
        ws = r._workspace()
        light = r._pow if pw is None else pw
        geom = r._geom(nv.dev_f32(light, "light_pow"))
        L = nv.lib()

        self.losses = []
        bar = None
        if progress:
            import tqdm
            bar = tqdm.tqdm(total=epochs)
        done = 0
        while done < epochs:
            # next dump point of the reference cadence (svbrdf.py:73-83): epoch 1, multiples of 100, the last epoch.  With
            # nothing to dump and no progress bar nobody consumes the intermediate state: ONE call, one loss read-back.
            if dump or bar is not None:
                stop = 1 if done == 0 else min((done // 100 + 1) * 100, epochs)
            else:
                stop = epochs
            adam = nv.Adam(float(lr), 0.9, 0.999, 1e-8, done + 1)
            nv.check(L.svbrdf_l2_adam_run(ctypes.byref(geom), nv.ptr(tex), nv.ptr(m), nv.ptr(v), nv.ptr(tgt), dtype_code,
                                          ctypes.byref(adam), stop - done, ctypes.c_void_p(curve.data_ptr() + 4 * done),
                                          nv.ptr(pow_state), nv.ptr(ws), nv.stream_ptr(self.device)), "svbrdf_l2_adam_run")
            if not read_back and not dump and bar is None:
                self.losses = curve[:epochs]             # device tensor; the caller reads it when it needs it
                return self.losses
            seg = curve[done:stop].tolist()          # one device->host sync per segment
            self.losses.extend(seg)
            if bar is not None:
                bar.update(stop - done)
                bar.set_postfix({"Loss": seg[-1], "Light": [int(x) for x in light.tolist()]})
            done = stop
            if dump:
                self._dump(svbrdf_obj, tmp_dir, done, epochs)
        if bar is not None:
            bar.close()
        return self.losses

    def _optim_autograd(self, epochs, lr, svbrdf_obj, pw, tmp_dir):
        params = [self.textures]
        if pw is not None:
            pw.requires_grad_(True)
            params.append(pw)
        self.optimizer = th.optim.Adam(params, lr=lr, betas=(0.9, 0.999))
        losses = []
        for epoch in range(epochs):
            if pw is not None:
                self.renderer_obj.update_light(pw)
            loss = self.compute_image_loss(self.renderer_obj.eval(self.textures.clamp(-1, 1)))
            losses.append(loss.detach())
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
            if tmp_dir is not None and ((epoch + 1) % 100 == 0 or epoch == 0 or epoch == epochs - 1):
                self.losses = th.stack(losses).tolist()
                self._dump(svbrdf_obj, tmp_dir, epoch + 1, epochs)
        self.losses = th.stack(losses).tolist() if losses else []
        if pw is not None:
            pw.requires_grad_(False)
        return self.losses

    def optim_with_features(self, epochs, lr, feature_loss, feature_weight=0.1, optim_light=False):
        """Per-pixel Adam on the combined loss of the reference's second workflow (materialgan.py:141-147):
        ``MSELoss(rendered, targets) + feature_weight * VGGLoss(rendered)`` — BASELINE configs[1] "L2 + descriptor loss".
        The render, the descriptor's input normalisation and the L2 term are one native kernel each way
        (``Microfacet.eval_normalized``); the feature network is torch/cuDNN.  ``feature_loss`` is a
        ``descriptor.VGGLoss`` whose ``load(targets)`` has been called.  Returns ``(image losses, feature losses)``."""
        r = self.renderer_obj
        params = [self.textures]
        pw = None
        if optim_light:
            pw = r._pow.detach().clone().requires_grad_(True)
            params.append(pw)
        self.optimizer = th.optim.Adam(params, lr=lr, betas=(0.9, 0.999))
        li, lf = [], []
        for _ in range(epochs):
            if pw is not None:
                r.update_light(pw)
            norm, l2 = r.eval_normalized(self.textures.clamp(-1, 1), feature_loss.mean, feature_loss.std, self.targets)
            feat = feature_loss.forward_normalized(norm) * feature_weight
            li.append(l2.detach())
            lf.append(feat.detach())
            self.optimizer.zero_grad()
            (l2 + feat).backward()
            self.optimizer.step()
        if pw is not None:
            pw.requires_grad_(False)
        self.losses = th.stack(li).tolist() if li else []
        self.losses_feature = th.stack(lf).tolist() if lf else []
        return self.losses, self.losses_feature

    def _dump(self, svbrdf_obj, tmp_dir, epoch, epochs):
        """svbrdf.py:74-83: loss curve, the four maps and the N re-renders."""
        this_dir = tmp_dir / f"{epoch}"
        this_dir.mkdir(parents=True, exist_ok=True)
        self.save_loss([self.losses], ["image loss"], tmp_dir / "loss.jpg", epochs)
        with th.no_grad():
            maps = self.textures.detach().clamp(-1, 1)
            svbrdf_obj.save_textures_th(maps, this_dir)
            svbrdf_obj.save_images_th(self.renderer_obj.eval(maps), this_dir)


class SvbrdfIO:
    """JSON capture config + PNG I/O (svbrdf.py:86-221)."""

    _PATH_KEYS = ("reference_dir", "target_dir", "optimize_dir", "rerender_dir")

    def __init__(self, json_dir, device):
        self.device = device
        if not json_dir.exists():
            raise FileNotFoundError(f"[ERROR:SvbrdfIO:init] {json_dir} is not exists")
        with open(json_dir, "r") as f:
            data = json.load(f)
        for key in self._PATH_KEYS:
            if key in data:
                setattr(self, key, json_dir.parent / data[key])
        for key in ("im_size", "camera_pos", "light_pos", "light_pow"):
            if key in data:
                setattr(self, key, data[key])
        if "idx" in data:
            self.idx = data["idx"]
            self.n_of_imgs = len(self.idx)
        if "light_pow" in data:
            self.load_calibration_th()
        _log("[DONE:SvbrdfIO] Initial object")

    def np_to_th(self, arr):
        return th.from_numpy(np.ascontiguousarray(arr)).to(self.device)

    def th_to_np(self, arr):
        return arr.detach().cpu().numpy()

    def reconstruct_normal(self, texture):
        xy = texture[:, 0:2, :, :].clamp(-1, 1)
        z = (1 - (xy * xy).sum(1, keepdim=True).clamp(0, 1)).sqrt()
        n = th.cat((xy, z), 1)
        return n / n.norm(2.0, 1, keepdim=True)

    def load_calibration_th(self):
        cam = np.array(self.camera_pos, "float32")[self.idx, :]
        light = np.array(self.light_pos, "float32")[self.idx, :]
        power = np.array(self.light_pow, "float32")
        self.cl = [self.np_to_th(cam), self.np_to_th(light), self.np_to_th(power)]
        _log("[DONE:SvbrdfIO] Load parameters")

    def load_textures_th(self, textures_dir, res, on_device=None):
        """svbrdf.py:150-166.  On a CUDA device the decoded PNG bytes are uploaded as they are (10 B per texel) and
        resized (cv2-exact Lanczos-4) and decoded by the native kernels of ``maps.py``; ``on_device=False`` forces the
        host path (cv2 + numpy, the reference's own behaviour).  Both give the same maps bit for bit."""
        if not textures_dir.exists():
            raise FileNotFoundError(f"[ERROR:SvbrdfIO:load_textures_th] {textures_dir} is not exists")
        if on_device is None:
            on_device = th.device(self.device).type == "cuda"
        if on_device:
            arrays = {k: imread_raw(textures_dir / f"{k}.png") for k in ("nom", "dif", "spe", "rgh")}
            eight_bit = all(a.dtype == np.uint8 for a in arrays.values())
            # the device path stacks the four maps into one [10,h,w] byte array: same source size required (the reference
            # resizes each file on its own, so mixed-size map sets take the host path below, as they do there)
            same_size = len({a.shape[:2] for a in arrays.values()}) == 1
            if eight_bit and same_size and arrays["rgh"].ndim == 2 and all(arrays[k].ndim == 3 and arrays[k].shape[2] == 3 for k in ("nom", "dif", "spe")):
                planes = maps.png_arrays_to_planes(arrays, self.device)
                out = maps.decode_u8(maps.resize_lanczos4_u8(planes, res, res))
                _log("[DONE:SvbrdfIO] Load textures (numbers in range [-1,1])")
                return out
        normal = imread(textures_dir / "nom.png", "normal", (res, res))
        diffuse = imread(textures_dir / "dif.png", "srgb", (res, res))
        specular = imread(textures_dir / "spe.png", "srgb", (res, res))
        roughness = imread(textures_dir / "rgh.png", "rough", (res, res))
        chw = lambda a: self.np_to_th(a).permute(2, 0, 1).unsqueeze(0)  # noqa: E731
        out = th.cat((chw(diffuse * 2 - 1), chw(normal)[:, :2], self.np_to_th(roughness * 2 - 1)[None, None], chw(specular * 2 - 1)), 1)
        _log("[DONE:SvbrdfIO] Load textures (numbers in range [-1,1])")
        return out.contiguous()

    def save_textures_th(self, textures_th, textures_dir):
        """svbrdf.py:168-189.  CUDA maps are quantised by the native encoder (``maps.encode_u8``: 10 B per texel cross
        the bus instead of 36) — the PNG files are identical to the host path's."""
        textures_dir.mkdir(parents=True, exist_ok=True)
        if textures_th.is_cuda and textures_th.dtype == th.float32:
            for name, arr in maps.planes_to_png_arrays(maps.encode_u8(textures_th.contiguous(), clamp=False)).items():
                imwrite_u8(arr, textures_dir / f"{name}.png")
            tex4to1(textures_dir)
            _log("[DONE:SvbrdfIO] Save textures")
            return
        hwc = lambda t: self.th_to_np(t.squeeze(0).permute(1, 2, 0))  # noqa: E731
        imwrite(hwc(self.reconstruct_normal(textures_th[:, 3:5])), textures_dir / "nom.png", "normal")
        imwrite(hwc((textures_th[:, 0:3] + 1) / 2), textures_dir / "dif.png", "srgb")
        imwrite(hwc((textures_th[:, 6:9] + 1) / 2), textures_dir / "spe.png", "srgb")
        imwrite(self.th_to_np(((textures_th[:, 5] + 1) / 2).squeeze(0)), textures_dir / "rgh.png", "rough")
        tex4to1(textures_dir)
        _log("[DONE:SvbrdfIO] Save textures")

    def load_images_th(self, images_dir, res=256, as_uint8=False):
        """Targets as ``[N,3,res,res]``.  ``as_uint8=True`` keeps the decoded PNG bytes (the fused
        kernel divides by 255 itself, bit-identical to the host's ``/255``): a quarter of the PCIe upload and of the HBM
        traffic of the float32 stack.  ``as_uint8="auto"`` (what ``optim_perpixel`` passes): bytes when every file is an
        8-bit 3-channel PNG — what ``save_images_th`` and the capture pipeline write — float32 otherwise."""
        if not images_dir.exists():
            raise FileNotFoundError(f"[ERROR:SvbrdfIO:load_images_th] {images_dir} is not exists")
        files = [images_dir / f"{idx:02d}.png" for idx in self.idx]
        if as_uint8:
            raw = [imread_raw(fn, (res, res)) for fn in files]
            ok = all(im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3 for im in raw)
            if ok:
                _log("[DONE:SvbrdfIO] Load images (uint8)")
                return self.np_to_th(np.stack([np.ascontiguousarray(im[:, :, ::-1].transpose(2, 0, 1)) for im in raw], 0))
            if as_uint8 != "auto":
                bad = next(fn for fn, im in zip(files, raw) if not (im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3))
                raise ValueError(f"{bad}: as_uint8 needs an 8-bit 3-channel PNG")
        stack = [imread(fn, "srgb", (res, res)).transpose(2, 0, 1) for fn in files]
        _log("[DONE:SvbrdfIO] Load images")
        return self.np_to_th(np.stack(stack, 0))

    def save_images_th(self, images_th, images_dir):
        images_dir.mkdir(parents=True, exist_ok=True)
        if images_th.shape[0] != self.n_of_imgs:
            raise RuntimeError("[ERROR:SvbrdfIO:save_images_th]")
        host = self.th_to_np(images_th.permute(0, 2, 3, 1))
        for i, idx in enumerate(self.idx):
            imwrite(host[i], images_dir / f"{idx:02d}.png", "srgb")
        if self.n_of_imgs == 9:
            img9to1(images_dir)
        _log("[DONE:SvbrdfIO] Save images")
