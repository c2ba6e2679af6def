"""Net2NetTransformer of the MaskGit variant under the reference's import path (reference
modules/stage2/cond_transformer_multi_view_muse.py:28-286).  Inference surface: sample / encode_to_z / encode_to_c / decode_to_img /
get_input / get_xc / log_images / test_step / forward with the reference's signatures; training steps and wandb panels are out of scope."""
import logging
import time
from typing import Optional

import torch

from multi_view_generation.bev_utils import util
from multi_view_generation.modules.stage2.cond_transformer_multi_view import Net2NetTransformer as _ARNet2Net
from multi_view_generation.modules.stage2.cond_transformer_multi_view import _Base
from multi_view_generation.modules.transformer.permuter import Identity
from multi_view_generation import utils

log = logging.getLogger(__name__)


class Net2NetTransformer(_Base):
    def __init__(self, maskgit, first_stage, cond_stage, cfg, permuter=None, ckpt_path=None, ignore_keys=[], unfrozen_keys=[],
                 first_stage_key="image", cond_stage_key="segmentation", downsample_cond_size=-1, pkeep=1.0, sos_token=0,
                 unconditional=False, skip_sampling: bool = False, bbox_ce_weight: float = 0.0, reset_random_mask: int = 0,
                 debug_viz: bool = False, partial_decoding: Optional[int] = None, bbox_warmup_steps: int = -1,
                 top_k: Optional[int] = None, warmup_steps: int = 500, lr_decay: bool = False, sample_iterations: int = 18, **kwargs):
        super().__init__()
        for k, v in kwargs.items():
            if k != "self":
                setattr(self, k, v)
        self.be_unconditional, self.sos_token = unconditional, sos_token
        self.first_stage_key, self.cond_stage_key = first_stage_key, cond_stage_key
        self.skip_sampling, self.bbox_ce_weight, self.reset_random_mask = skip_sampling, bbox_ce_weight, reset_random_mask
        self.debug_viz, self.partial_decoding, self.bbox_warmup_steps = debug_viz, partial_decoding, bbox_warmup_steps
        self.top_k, self.lr_decay, self.warmup_steps, self.sample_iterations = top_k, lr_decay, warmup_steps, sample_iterations
        self.first_stage_model = _ARNet2Net._freeze(first_stage)
        self.cond_stage_model = _ARNet2Net._freeze(cond_stage)
        self.cfg = cfg
        self.maskgit = maskgit
        self.permuter = Identity() if permuter is None else permuter
        if ckpt_path is not None:
            utils.init_from_ckpt(self, ckpt_path, ignore_keys=ignore_keys, unfrozen_keys=unfrozen_keys)
        self.downsample_cond_size, self.pkeep = downsample_cond_size, pkeep

    expand_all_images = _ARNet2Net.expand_all_images
    combine_all_images = _ARNet2Net.combine_all_images
    encode_to_z = _ARNet2Net.encode_to_z
    encode_to_c = _ARNet2Net.encode_to_c
    decode_to_img = _ARNet2Net.decode_to_img
    get_input = _ARNet2Net.get_input
    get_xc = _ARNet2Net.get_xc
    _memo_call = _ARNet2Net._memo_call

    def _dev(self):
        return self.maskgit.transformer.to_logits.weight.device

    @torch.no_grad()
    def sample(self, cond, batch, partial_decoding_idx=None):
        """-> LongTensor (b*cam, h, w) (reference :124-141): cameras in `partial_decoding_idx` keep their ground-truth tokens."""
        init_ids = None
        if partial_decoding_idx is not None:
            init_ids = torch.full((cond.shape[0], self.cfg.num_cams, self.cfg.num_cam_tokens), self.maskgit.mask_id, dtype=torch.long, device=cond.device)
            _, z_indices = self.encode_to_z(self.get_input(self.first_stage_key, batch).to(cond.device), batch)
            z_indices = self.expand_all_images(z_indices)
            init_ids[:, partial_decoding_idx, :] = z_indices[:, partial_decoding_idx]
            init_ids = init_ids.reshape(-1, self.cfg.num_cam_tokens)
        ids = self.maskgit.generate(init_ids=init_ids, cond_images=cond, fmap_size=(self.cfg.cam_latent_h, self.cfg.cam_latent_w), batch=batch,
                                    timesteps=self.sample_iterations)
        assert ids.max() < self.cfg.vocab_size
        return ids

    def test_step(self, batch, batch_idx=0):
        return self.log_images(batch, generate_only=True)

    def forward(self, batch):
        return self.log_images(batch, generate_only=True)

    @torch.no_grad()
    def log_images(self, batch, generate_only=False, **kwargs):
        """-> {'gen','rec','gt'} each (b, num_cams, 3, H, W) fp32 in [0,1] (reference :231-285)."""
        start = time.time()
        dev = self._dev()
        x, c = self.get_xc(batch)
        x, c = x.to(dev), c.to(dev)
        quant_z, z_indices = self.encode_to_z(x, batch)
        _, c_indices = self.encode_to_c(c, batch)
        zshape = quant_z.shape
        rec = util.denormalize_tensor(self.decode_to_img(z_indices, zshape), keep_tensor=True)
        pidx = None
        if self.partial_decoding:
            if self.partial_decoding == 2:
                pidx = torch.randint(self.cfg.num_cams, (torch.randint(1, self.cfg.num_cams, ()).item(),))
            elif self.partial_decoding == 3:
                pidx = torch.tensor([0]) if torch.rand(()).item() > 0.5 else torch.tensor([0, 2])
            else:
                pidx = torch.randint(self.cfg.num_cams, (1,))
        index_sample = self.sample(c_indices, batch, partial_decoding_idx=pidx)
        gen = util.denormalize_tensor(self.decode_to_img(index_sample.reshape(index_sample.shape[0], -1), zshape), keep_tensor=True)
        gt = util.denormalize_tensor(x, keep_tensor=True)
        log.info("Generating images took %.3f s", time.time() - start)
        return {"gen": self.expand_all_images(gen), "rec": self.expand_all_images(rec), "gt": self.expand_all_images(gt)}
