"""Mask / camera-bias / layout generation under the reference's import path (reference modules/transformer/mask_generator.py).
The closed-form implementations live in bevgen_b200.geometry; these wrappers keep the reference's function names."""
from multi_view_generation.modules.transformer.permuter import get_seq_pixel_mappings  # noqa: F401


def outward_pattern(cfg, return_camera_bias_matrix=False):
    if return_camera_bias_matrix:
        return cfg.prob_matrix.clone()
    raise NotImplementedError("use cfg.get_mask() (layouts, allowed) — the intermediate tuple of the reference is not exposed")


def multi_outward_pattern(cfg):
    return cfg.get_mask()
