"""GenerateImages under the reference's import path (reference utils/callback.py:33-165): the output writer of generate.py.

`save_raw_data` keeps the reference's on-disk format, which the metric scripts consume (scripts/metrics_eval.py:106-152):
    <save_dir>/sample/<token>/<cam_name>.jpg        generated views        <save_dir>/sample_gt/<token>/<cam_name>.jpg   ground truth
    <save_dir>/sample/<token>/bev.npz, sample_gt/<token>/bev.npz           the BEV segmentation (np.savez_compressed)
    <save_dir>/{gt,rec,gen}/<image_path> (+ .npz intrinsics next to gen/)  nuScenes layout, when the batch carries `image_paths`
B200-native part: the (B, cams, 3, H, W) fp32 tensors are converted to uint8 HWC on the GPU (bevgen_to_uint8_hwc), copied to pinned host
memory in one transfer per tensor, and encoded / written by a thread pool so the next batch's sampling overlaps the file I/O.
Not reproduced: the wandb / matplotlib figure panels (`viz/`, batched_camera_bev_grid, viz_bev need the absent image_utils package).
"""
import logging
import os
import random
import string
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch

log = logging.getLogger(__name__)

try:
    from pytorch_lightning import Callback as _Base
except Exception:  # pragma: no cover
    _Base = object


def _to_host_u8(x):
    """(..., 3, H, W) fp32 in [0,1] on the GPU -> uint8 numpy (..., H, W, 3) through pinned memory."""
    from bevgen_b200 import ops
    if not x.is_cuda:
        raise RuntimeError("GenerateImages expects the CUDA tensors produced by log_images / test_step (there is no CPU path)")
    lead = x.shape[:-3]
    u8 = ops.to_uint8_hwc(x.reshape(-1, *x.shape[-3:]).float().contiguous())
    host = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
    host.copy_(u8, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy().reshape(*lead, *u8.shape[1:])


def save_img(arr_hwc_u8, save_path: Path):
    from PIL import Image
    os.makedirs(save_path.parents[0], exist_ok=True)
    Image.fromarray(arr_hwc_u8).save(save_path)


class GenerateImages(_Base):
    def __init__(self, save_dir=None, figure_format=False, rand_str=False, workers=8, **kwargs):
        self.save_dir, self.figure_format, self.rand_str = save_dir, figure_format, rand_str
        self._pool = ThreadPoolExecutor(max_workers=workers)
        self._pending = []

    def flush(self):
        """Wait for every queued file (raises the first writer error)."""
        pending, self._pending = self._pending, []
        for f in pending:
            f.result()

    def _submit(self, fn, *a):
        self._pending.append(self._pool.submit(fn, *a))

    def save_raw_data(self, trainer, pl_module, outputs, batch, save_nuscenes_fmt: bool = True):
        if self.save_dir is None:
            if trainer is None or getattr(trainer, "log_dir", None) is None:
                raise ValueError("GenerateImages needs save_dir (or a trainer with log_dir)")
            from datetime import datetime
            root = Path(os.path.join(trainer.log_dir, "results", datetime.now().strftime("%Y_%m_%d-%H_%M")))
        else:
            root = Path(self.save_dir)
        gen, gt = _to_host_u8(outputs["gen"]), _to_host_u8(outputs["gt"])
        rec = _to_host_u8(outputs["rec"]) if "rec" in outputs else None
        seg = batch["segmentation"].to(dtype=torch.float).detach().cpu()
        nusc = save_nuscenes_fmt and "image_paths" in batch and str(getattr(pl_module.cfg, "dataset", "")).upper().endswith("NUSCENES")
        for b in range(gen.shape[0]):
            tok = batch["sample_token"][b]
            if self.rand_str:
                tok = tok + "_" + "".join(random.choices(string.ascii_uppercase + string.digits, k=5))
            for cam in range(gen.shape[1]):
                name = batch["cam_name"][cam][b]
                self._submit(save_img, gen[b, cam], root / "sample" / tok / f"{name}.jpg")
                self._submit(save_img, gt[b, cam], root / "sample_gt" / tok / f"{name}.jpg")
                if nusc:
                    rel = batch["image_paths"][cam][b]
                    self._submit(save_img, gt[b, cam], root / "gt" / rel)
                    if rec is not None:
                        self._submit(save_img, rec[b, cam], root / "rec" / rel)
                    self._submit(save_img, gen[b, cam], root / "gen" / rel)
                    self._submit(self._save_npz, (root / "gen" / rel).with_suffix(".npz"), batch["intrinsics"][b, cam].to(dtype=torch.float).detach().cpu(), False)
            for sub in ("sample", "sample_gt"):
                self._submit(self._save_npz, root / sub / tok / "bev.npz", seg[b], True)

    @staticmethod
    def _save_npz(path, tensor, compressed):
        os.makedirs(path.parents[0], exist_ok=True)
        (np.savez_compressed if compressed else np.savez)(path, tensor)

    def on_test_batch_end(self, trainer, pl_module, outputs, batch, batch_idx=0, dataloader_idx=0):
        self.save_raw_data(trainer, pl_module, outputs, batch)

    def on_test_end(self, trainer=None, pl_module=None):
        self.flush()
