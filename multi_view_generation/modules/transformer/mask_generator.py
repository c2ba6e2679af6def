"""Mask / camera-bias / layout generation under the reference's import path (reference modules/transformer/mask_generator.py).
The closed-form implementations live in bevgen_b200.geometry (numpy, no O(L^2) index gymnastics); these functions keep the reference's
names, argument meaning, return tuples and dtypes.  Pinned to goldens minted from the reference's own functions
(tests/golden/outward_*.npz, tests/test_geometry_cpu.py)."""
import torch

from bevgen_b200 import geometry
from multi_view_generation.modules.transformer.permuter import get_seq_pixel_mappings  # noqa: F401


def outward_pattern(cfg, return_camera_bias_matrix=False):
    """reference :131-214.  With return_camera_bias_matrix: the (L, L) float64 camera-bias prior (cond columns 1, image -> BEV bearing
    similarity in the cond columns of the image rows).  Otherwise (allowed_pattern (heads, L, L) float32, static_layout (nb, nb) int64,
    prob_layout (nb, nb) float32, prob_matrix (L, L) float64 with 0.5 on the cond columns)."""
    if return_camera_bias_matrix:
        return torch.from_numpy(cfg._full_prior()).clone()
    nc, ni = cfg.num_cond_tokens, cfg.num_img_tokens
    full = cfg._full_prior()
    static_l, prob_l, pfull = geometry.layout_components(cfg.sparse_block_size, ni, nc, cfg.num_pad_tokens, cfg.window_len,
                                                         cfg.forward_shuffle_idx.numpy(), cfg.causal_order, full[nc:nc + ni, nc:nc + ni])
    allowed = cfg.attention_mask[None].repeat(cfg.num_heads, 1, 1)
    return allowed, torch.from_numpy(static_l.astype("int64")), prob_l, torch.from_numpy(pfull)


def multi_outward_pattern(cfg):
    """reference :217-251: (layouts (heads, nb, nb) int64 drawn per head from outward_pattern's block prior, allowed_pattern)."""
    return cfg.get_mask()
