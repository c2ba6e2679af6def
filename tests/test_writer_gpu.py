"""Output writer (GenerateImages.save_raw_data, reference utils/callback.py:72-132): on-disk layout and contents."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from bevgen_b200 import ops  # noqa: E402
from multi_view_generation.utils.callback import GenerateImages  # noqa: E402


def test_to_uint8_hwc_matches_torch():
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(5, 3, 20, 28, generator=g) * 1.2 - 0.1).cuda()            # includes values outside [0, 1]
    got = ops.to_uint8_hwc(x).cpu()
    want = (x.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).cpu()
    assert torch.equal(got, want)


def test_save_raw_data_layout(tmp_path):
    from PIL import Image
    B, cams, H, W = 2, 6, 64, 64
    g = torch.Generator().manual_seed(1)
    smooth = lambda: torch.nn.functional.interpolate(torch.rand(B * cams, 3, 8, 8, generator=g), size=(H, W), mode="bilinear").view(B, cams, 3, H, W).cuda()
    outputs = {"gen": smooth(), "gt": smooth(), "rec": smooth()}
    names = ["CAM_FRONT_LEFT", "CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_BACK_LEFT", "CAM_BACK", "CAM_BACK_RIGHT"]
    batch = {"sample_token": ["tokA", "tokB"], "cam_name": [[n] * B for n in names],
             "segmentation": (torch.rand(B, 32, 32, 7, generator=g) > 0.5).float(),
             "image_paths": [[f"samples/{n}/{t}.jpg" for t in ("a", "b")] for n in names],
             "intrinsics": torch.rand(B, cams, 3, 3, generator=g)}

    class Mod:
        class cfg:
            dataset = "Dataset.NUSCENES"

    cb = GenerateImages(save_dir=str(tmp_path))
    cb.on_test_batch_end(None, Mod(), outputs, batch, 0, 0)
    cb.on_test_end()
    for b, tok in enumerate(batch["sample_token"]):
        for c, n in enumerate(names):
            for sub, key in (("sample", "gen"), ("sample_gt", "gt")):
                img = np.asarray(Image.open(tmp_path / sub / tok / f"{n}.jpg")).astype(np.float32) / 255
                ref = outputs[key][b, c].permute(1, 2, 0).cpu().numpy()
                assert img.shape == (H, W, 3) and np.abs(img - ref).mean() < 0.02          # JPEG is lossy
            for sub, key in (("gt", "gt"), ("rec", "rec"), ("gen", "gen")):
                assert (tmp_path / sub / batch["image_paths"][c][b]).exists()
            k = np.load((tmp_path / "gen" / batch["image_paths"][c][b]).with_suffix(".npz"))["arr_0"]
            np.testing.assert_array_equal(k, batch["intrinsics"][b, c].numpy())
        for sub in ("sample", "sample_gt"):
            np.testing.assert_array_equal(np.load(tmp_path / sub / tok / "bev.npz")["arr_0"], batch["segmentation"][b].numpy())


def test_writer_rejects_cpu_tensors(tmp_path):
    cb = GenerateImages(save_dir=str(tmp_path))
    out = {"gen": torch.rand(1, 6, 3, 8, 8), "gt": torch.rand(1, 6, 3, 8, 8)}
    with pytest.raises(RuntimeError, match="no CPU path"):
        cb.save_raw_data(None, None, out, {"sample_token": ["t"], "cam_name": [["c"]] * 6, "segmentation": torch.zeros(1, 4, 4, 7)})
