"""GPTConfig — same field names / derived attributes as the reference dataclass
(multi_view_generation/modules/transformer/mingpt_sparse.py:26-113) so Hydra YAML nodes with
`_target_: ...mingpt_sparse.GPTConfig` keep instantiating.  Mask / camera-bias prior / decode order
come from `bevgen_b200.geometry` (closed forms), computed once on the host.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import geometry
from .geometry import Cameras, Dataset


@dataclass(unsafe_hash=True)
class GPTConfig:
    embd_pdrop: float
    resid_pdrop: float
    attn_pdrop: float
    num_layers: int
    num_heads: int
    num_embed: int
    hidden_size: int
    vocab_size: int
    cond_vocab_size: int
    num_cams: int
    window_len: int
    density: float
    sparse_block_size: int
    n_unmasked: int
    backend: str
    plot: bool
    cam_res: Tuple[int, int]
    cam_latent_res: Tuple[int, int]
    bev_latent_res: Tuple[int, int]
    camera_bias: bool
    bev_embed: bool
    image_embed: bool
    cam_names: Any
    cam_name_to_idx: Dict[str, Any] = field(init=False, compare=False)
    num_cond_tokens: int = field(init=False)
    num_cam_tokens: int = field(init=False)
    num_img_tokens: int = field(init=False)
    num_pad_tokens: int = field(init=False)
    gpt_block_size: int = field(init=False)
    cam_latent_h: int = field(init=False)
    cam_latent_w: int = field(init=False)
    attention_mask: Any = field(init=False, compare=False)
    causal_order: bool = False
    forward_shuffle_idx: Any = field(init=False, compare=False)
    backward_shuffle_idx: Any = field(init=False, compare=False)
    layout: Optional[Any] = field(init=False, compare=False, repr=False, default=None)
    only_front_cams: bool = field(init=False, compare=False, repr=False, default=False)
    dataset_name: str = field(init=False)
    output_dir: str = "output"
    prob_matrix = None
    legacy_prob_matrix: bool = True
    cam_intrinsics: Any = None
    cam_extrinsics: Any = None
    dataset: Any = Dataset.NUSCENES

    def __post_init__(self):
        if isinstance(self.dataset, Dataset):
            pass
        elif isinstance(self.dataset, int):
            self.dataset = Dataset(self.dataset)
        else:
            self.dataset = Dataset[self.dataset]
        self.dataset_name = self.dataset.name.lower()
        if not isinstance(self.cam_names, Cameras):
            self.cam_names = Cameras[self.cam_names]
        assert len(self.cam_names) == self.num_cams
        self.cam_res = tuple(self.cam_res)
        self.cam_latent_res = tuple(self.cam_latent_res)
        self.bev_latent_res = tuple(self.bev_latent_res)
        self.cam_name_to_idx = {k: v for v, k in enumerate(self.cam_names.value)}

        self.cam_latent_h, self.cam_latent_w = self.cam_latent_res
        self.num_cond_tokens = self.bev_latent_res[0] * self.bev_latent_res[1]
        self.num_cam_tokens = self.cam_latent_h * self.cam_latent_w
        self.num_img_tokens = self.num_cam_tokens * self.num_cams
        blk = self.sparse_block_size
        self.gpt_block_size = blk * int(np.ceil((self.num_img_tokens + self.num_cond_tokens) / blk))
        self.num_pad_tokens = self.gpt_block_size - (self.num_img_tokens + self.num_cond_tokens)
        if not self.legacy_prob_matrix and self.camera_bias:
            # needs pretrained/cam_data_<dataset>.pt which is not distributed (mask_generator.py:89-110)
            raise NotImplementedError("legacy_prob_matrix=False needs pretrained/cam_data_*.pt (not available)")

        fwd = geometry.decode_order(self.num_cams, self.cam_latent_h, self.cam_latent_w, self.dataset, self.causal_order)
        self.forward_shuffle_idx = torch.from_numpy(fwd)
        self.backward_shuffle_idx = torch.argsort(self.forward_shuffle_idx)
        mask = geometry.attention_mask(self.num_img_tokens, self.num_cond_tokens, self.num_pad_tokens,
                                       self.window_len, fwd, self.causal_order)
        self.attention_mask = torch.from_numpy(mask.astype(np.float32))
        self._prior = None
        if self.camera_bias:
            self.prob_matrix = torch.from_numpy(self._full_prior())

    # -- helpers --------------------------------------------------------------------------------
    def _full_prior(self):
        if self._prior is None:
            self._prior = geometry.camera_bias_prior(
                self.num_cams, self.cam_latent_h, self.cam_latent_w, self.bev_latent_res[0], self.bev_latent_res[1],
                self.num_pad_tokens, self.window_len, self.forward_shuffle_idx.numpy(), self.causal_order)
        return self._prior

    def get_mask(self):
        """(layouts (heads,nb,nb) int64, allowed (heads,L,L) float) like mask_generator.multi_outward_pattern."""
        full = self._full_prior()
        nc, ni = self.num_cond_tokens, self.num_img_tokens
        layouts = geometry.block_layouts(self.num_heads, self.sparse_block_size, self.density, ni, nc,
                                         self.num_pad_tokens, self.window_len, self.forward_shuffle_idx.numpy(),
                                         self.causal_order, full[nc:nc + ni, nc:nc + ni])
        allowed = self.attention_mask[None].repeat(self.num_heads, 1, 1)
        return layouts, allowed

    def forward_permuter(self, x):
        return x[:, self.forward_shuffle_idx]

    def backward_permuter(self, x):
        return x[:, self.backward_shuffle_idx]

    def layout_covers_mask(self, layouts) -> bool:
        """True when every allowed position lies in a present block for every head (density=1.0 case):
        then the block layout never removes anything and attention is governed by the mask alone."""
        blk = self.sparse_block_size
        nb = self.gpt_block_size // blk
        need = self.attention_mask.bool().reshape(nb, blk, nb, blk).any(3).any(1)
        return bool((layouts.bool() | ~need[None]).all())
