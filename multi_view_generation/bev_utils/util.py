"""bev_utils/util.py subset: `Cameras`, `Dataset`, `denormalize_tensor` (reference util.py:20-39,97-118)."""
import torch

from bevgen_b200.geometry import Cameras, Dataset  # noqa: F401

IMG_MEAN = [0.4265, 0.4489, 0.4769]
IMG_STD = [0.2053, 0.2206, 0.2578]


def denormalize_tensor(x, keep_tensor=False):
    """x*std+mean per channel, clamp to [0,1]; (B,3,H,W) or (3,H,W).  CUDA tensors go through the C-ABI kernel."""
    squeeze = x.dim() == 3
    xb = x.unsqueeze(0) if squeeze else x
    if xb.is_cuda:
        from bevgen_b200 import ops
        xb = xb.float().contiguous()
        out = torch.empty_like(xb)
        ops.denormalize(xb, out, IMG_MEAN, IMG_STD)
    else:   # host tensors (e.g. ground-truth images for logging): trivial torch arithmetic, not a hot path
        m = torch.tensor(IMG_MEAN, dtype=xb.dtype).view(1, 3, 1, 1)
        s = torch.tensor(IMG_STD, dtype=xb.dtype).view(1, 3, 1, 1)
        out = torch.clamp(xb * s + m, 0, 1)
    out = out.squeeze(0) if squeeze else out
    if not keep_tensor:
        out = (out.detach().cpu().permute(*range(out.dim() - 3), out.dim() - 2, out.dim() - 1, out.dim() - 3).numpy() * 255).astype("uint8")
    return out
