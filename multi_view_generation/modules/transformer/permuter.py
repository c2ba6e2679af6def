"""Decode-order permuter under the reference's import path (reference modules/transformer/permuter.py:10-88)."""
import torch
import torch.nn as nn

from bevgen_b200 import geometry


class AbstractPermuter(nn.Module):
    def forward(self, x, reverse=False):
        raise NotImplementedError


class Identity(AbstractPermuter):
    def forward(self, x, reverse=False):
        return x


def get_seq_pixel_mappings(cfg):
    """(pixel_to_seq (cam,h,w) -> seq index, seq_to_pixel (seq,3) -> (cam,h,w)); permuter.py:26-30."""
    n = cfg.num_cams * cfg.cam_latent_h * cfg.cam_latent_w
    pixel_to_seq = torch.arange(n).reshape(cfg.num_cams, cfg.cam_latent_h, cfg.cam_latent_w)
    seq_to_pixel = torch.stack(torch.meshgrid(torch.arange(cfg.num_cams), torch.arange(cfg.cam_latent_h), torch.arange(cfg.cam_latent_w),
                                              indexing="ij"), -1).reshape(-1, 3)
    return pixel_to_seq, seq_to_pixel


class CustomPermuter(AbstractPermuter):
    def __init__(self, cfg):
        super().__init__()
        fwd = torch.from_numpy(geometry.decode_order(cfg.num_cams, cfg.cam_latent_h, cfg.cam_latent_w, cfg.dataset, cfg.causal_order))
        self.register_buffer("forward_shuffle_idx", fwd)
        self.register_buffer("backward_shuffle_idx", torch.argsort(fwd))

    def forward(self, x, reverse=False):
        return x[:, self.backward_shuffle_idx] if reverse else x[:, self.forward_shuffle_idx]
