"""Net2NetTransformer under the reference's import path (reference modules/stage2/cond_transformer_multi_view.py:30-561).

Inference surface only: forward / sample / encode_to_z / encode_to_c / decode_to_img / get_input / get_xc / shared_step /
inference_step / log_images / test_step with the reference's signatures.  `sample` runs the KV-cache CUDA-graph sampler
(bevgen_b200.gpt_decode) instead of 1536 full forwards; training hooks, wandb visualisations and bbox-weighted loss are out
of scope.
"""
import logging
import time
from typing import Optional

import torch
import torch.nn.functional as F

from multi_view_generation import utils
from multi_view_generation.bev_utils import util
from multi_view_generation.modules.transformer.permuter import Identity

try:
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pragma: no cover
    _Base = torch.nn.Module

log = logging.getLogger(__name__)


def disabled_train(self, mode=True):
    return self


class Net2NetTransformer(_Base):
    def __init__(self, transformer, first_stage, cond_stage, permuter=None, ckpt_path=None, ignore_keys=[], unfrozen_keys=[],
                 first_stage_key="image", cond_stage_key="segmentation", downsample_cond_size=-1, pkeep=1.0, sos_token=0,
                 unconditional=False, skip_sampling: bool = False, bbox_ce_weight: float = 0.0, reset_random_mask: int = 0,
                 debug_viz: bool = False, partial_decoding: Optional[int] = None, bbox_weight_epoch: int = -1,
                 top_k: Optional[int] = None, warmup_steps: int = 500, lr_decay: bool = False, **kwargs):
        super().__init__()
        for k, v in kwargs.items():           # the reference setattr()s every unknown kwarg (:58-61)
            if k != "self":
                setattr(self, k, v)
        self.be_unconditional, self.sos_token = unconditional, sos_token
        self.first_stage_key, self.cond_stage_key = first_stage_key, cond_stage_key
        self.skip_sampling, self.bbox_ce_weight, self.reset_random_mask = skip_sampling, bbox_ce_weight, reset_random_mask
        self.debug_viz, self.partial_decoding, self.bbox_weight_epoch = debug_viz, partial_decoding, bbox_weight_epoch
        self.top_k, self.lr_decay, self.warmup_steps = top_k, lr_decay, warmup_steps
        self.first_stage_model = self._freeze(first_stage)
        self.cond_stage_model = self._freeze(cond_stage)
        self.transformer = transformer
        self.cfg = self.transformer.cfg
        self.permuter = Identity() if permuter is None else permuter
        if ckpt_path is not None:
            utils.init_from_ckpt(self, ckpt_path, ignore_keys=ignore_keys, unfrozen_keys=unfrozen_keys)
        self.downsample_cond_size, self.pkeep = downsample_cond_size, pkeep
        self.sample_seed = None               # set to an int for reproducible sampling; default draws from torch's RNG

    @staticmethod
    def _freeze(model):
        model = model.eval()
        model.train = disabled_train.__get__(model)
        return model

    def expand_all_images(self, arr):
        return arr.reshape(-1, self.cfg.num_cams, *arr.shape[1:])

    def combine_all_images(self, arr):
        return arr.reshape(-1, *arr.shape[2:])

    # ---------------------------------------------------------------- forward / sample
    @torch.no_grad()
    def forward(self, x, c, batch):
        _, z_indices = self.encode_to_z(x, batch)
        _, c_indices = self.encode_to_c(c, batch)
        z_indices = self.expand_all_images(z_indices)
        target = z_indices.reshape(z_indices.shape[0], -1).clone()
        # GPT.forward overwrites the last token of its argument with PAD (reference mingpt_sparse.py:328-329): hand it a copy, the
        # encoder's index tensor may be shared with log_images through the test_step memo
        logits = self.transformer(z_indices.clone(), c_indices, batch, sampling=False)
        return logits.contiguous(), target.contiguous()

    def top_k_logits(self, logits, k):
        v, _ = torch.topk(logits, k)
        out = logits.clone()
        out[out < v[..., [-1]]] = -float("Inf")
        return out

    @torch.no_grad()
    def inference_step(self, batch):
        if not hasattr(self, "z_indices"):
            x, c = self.get_xc(batch)
            _, self.c_indices = self.encode_to_c(c.to(self._dev()), batch)
            self.z_indices = torch.full((c.shape[0], self.cfg.num_cams, self.cfg.num_cam_tokens), self.cfg.vocab_size, dtype=torch.int64,
                                        device=self._dev())
        return self.transformer(self.z_indices, self.c_indices, batch, sampling=False)

    @torch.no_grad()
    def sample(self, x, c, batch, temperature=1.0, sample=False, top_k=None, callback=lambda k: None, partial_decoding_idx=None):
        """-> LongTensor (B, num_cams, cam_tokens).  Only x.shape[0] of `x` is used (as in the reference, :157)."""
        B = x.shape[0]
        if self.skip_sampling:
            return torch.zeros((B, self.cfg.num_cams, self.cfg.num_cam_tokens), dtype=torch.int64, device=c.device)
        assert not self.transformer.training
        forced = None
        if partial_decoding_idx is not None:
            # the listed cameras keep their ground-truth tokens (reference :161-165) and are skipped by the loop (:181-182); under the
            # causal mask this is exactly a forced token at their decode positions and a sampled one everywhere else
            _, z_indices = self.encode_to_z(self.get_input(self.first_stage_key, batch).to(c.device), batch)
            z_indices = self.expand_all_images(z_indices)                                   # (B, cams, tokens)
            given = torch.full_like(z_indices, -1)
            cams = [int(i) for i in partial_decoding_idx]
            given[:, cams, :] = z_indices[:, cams, :]
            fwd = torch.as_tensor(self.cfg.forward_shuffle_idx, device=c.device, dtype=torch.int64)
            forced = given.reshape(B, -1)[:, fwd].contiguous()                              # decode order
        seed = self.sample_seed if self.sample_seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        out = self.transformer.sampler(B).sample(c, batch, temperature=temperature, top_k=top_k, greedy=not sample, seed=seed,
                                                 forced_tokens=forced)
        assert out.max() < self.cfg.vocab_size
        return out

    # ---------------------------------------------------------------- stage-1 plumbing
    def _dev(self):
        return self.transformer.head.weight.device

    @torch.no_grad()
    def encode_to_z(self, x, batch):
        quant_z, _, info = self._memo_call("z", x, lambda: self.first_stage_model.encode(x, batch))
        indices = self.permuter(info[2].view(quant_z.shape[0], -1))
        return quant_z, indices

    @torch.no_grad()
    def encode_to_c(self, c, batch):
        if self.downsample_cond_size > -1:
            c = F.interpolate(c, size=(self.downsample_cond_size, self.downsample_cond_size))
        quant_c, _, (_, _, indices) = self._memo_call("c", c, lambda: self.cond_stage_model.encode(c, batch))
        return quant_c, indices.view(c.shape[0], -1)

    @torch.no_grad()
    def decode_to_img(self, index, zshape):
        index = self.permuter(index, reverse=True)
        bhwc = (zshape[0], zshape[2], zshape[3], zshape[1])
        return self.first_stage_model.decode_indices(index.reshape(-1), bhwc)      # get_codebook_entry + decode, NHWC throughout

    def get_input(self, key, batch):
        memo = getattr(self, "_step_memo", None)
        if memo is not None and memo.get(("in", key, id(batch[key]))) is not None:
            return memo[("in", key, id(batch[key]))]
        x = batch[key]
        if x.dtype == torch.double or x.dtype == torch.uint8:
            x = x.float()
        x = x.movedim(-1, -3)                                   # ... h w c -> ... c h w
        if key == "image":
            if len(x.shape) == 4:
                x = x[None, ...]
            x = self.combine_all_images(x)
        x = x.contiguous()
        if memo is not None:
            memo[("in", key, id(batch[key]))] = x
        return x

    def _memo_call(self, tag, x, fn):
        """Within one test_step the reference encodes the same images twice (shared_step and log_images); the results are identical, so
        the second call returns the first one's tensors.  Outside test_step (no memo) every call computes."""
        memo = getattr(self, "_step_memo", None)
        if memo is None:
            return fn()
        key = (tag, x.data_ptr(), tuple(x.shape))
        if key not in memo:
            memo[key] = fn()
        return memo[key]

    def batch_to_device(self, batch):
        """One asynchronous upload of the tensors the hot path reads (pinned host memory -> no staging copy); everything else in the
        dict (cam_name, sample_token, ...) is passed through untouched."""
        dev = self._dev()
        keys = (self.first_stage_key, self.cond_stage_key, "intrinsics_inv", "extrinsics_inv")
        return {k: (v.to(dev, non_blocking=True) if (k in keys and torch.is_tensor(v) and not v.is_cuda) else v) for k, v in batch.items()}

    def get_xc(self, batch, N=None):
        x, c = self.get_input(self.first_stage_key, batch), self.get_input(self.cond_stage_key, batch)
        if N is not None:
            x, c = x[:N], c[:N]
        return x, c

    @torch.no_grad()
    def shared_step(self, batch, batch_idx=0, inference=False):
        if self.bbox_ce_weight > 0:
            raise NotImplementedError("bbox-weighted cross entropy is a training-only path")
        x, c = self.get_xc(batch)
        logits, target = self(x.to(self._dev()), c.to(self._dev()), batch)
        return sum(F.cross_entropy(l.view(-1, l.shape[-1]), t.view(-1)) for l, t in zip(logits, target)) / len(logits)

    def test_step(self, batch, batch_idx=0):
        """generate.py's per-batch call (reference :378-384): teacher-forced loss, then log_images(generate_only=True).  The batch is
        uploaded once and the stage-1 encodes are shared between the two halves."""
        batch = self.batch_to_device(batch)
        self._step_memo = {}
        try:
            loss = self.shared_step(batch, batch_idx)
            if hasattr(self, "log") and _Base is not torch.nn.Module:
                self.log("test/loss", loss, prog_bar=True, on_step=False, on_epoch=True)
            self.last_test_loss = loss
            return self.log_images(batch, generate_only=True, top_k=self.top_k)
        finally:
            self._step_memo = None

    @torch.no_grad()
    def log_images(self, batch, temperature=None, top_k=None, callback=None, generate_only=False, **kwargs):
        """-> {'gen','rec','gt'} each (B, num_cams, 3, H, W) fp32 in [0,1] (reference :479-544, generate_only branch)."""
        start = time.time()
        dev = self._dev()
        x, c = self.get_xc(batch)
        x, c = x.to(dev), c.to(dev)
        quant_z, z_indices = self.encode_to_z(x, batch)
        _, c_indices = self.encode_to_c(c, batch)
        zshape = quant_z.shape
        rec = util.denormalize_tensor(self.decode_to_img(z_indices, zshape), keep_tensor=True)
        partial_decoding_idx = self._draw_partial_decoding_idx()
        index_sample = self.sample(self.expand_all_images(z_indices)[:, :0], c_indices, batch,
                                   temperature=temperature if temperature is not None else 1.0, sample=True,
                                   top_k=top_k if top_k is not None else 100,
                                   callback=callback if callback is not None else (lambda k: None),
                                   partial_decoding_idx=partial_decoding_idx)
        gen = util.denormalize_tensor(self.decode_to_img(self.combine_all_images(index_sample), zshape), keep_tensor=True)
        gt = util.denormalize_tensor(x, keep_tensor=True)
        gen = self.expand_all_images(gen)
        if partial_decoding_idx is not None:
            # reference :526-529: the cameras that kept their ground-truth tokens are framed in green (3 px, colour (0, 249, 0));
            # the text overlays drawn on the GT panels (:530-533, third-party image_utils) are visualisation only and not reproduced
            kept = gen[:, partial_decoding_idx]
            colour = torch.tensor([0.0, 249.0 / 255.0, 0.0], device=gen.device, dtype=gen.dtype).view(1, 1, 3, 1, 1)
            frame = torch.ones(kept.shape[-2:], dtype=torch.bool, device=gen.device)
            frame[3:-3, 3:-3] = False
            gen[:, partial_decoding_idx] = torch.where(frame, colour.expand_as(kept), kept)
        log.info("Generating images took %.3f s", time.time() - start)
        return {"gen": gen, "rec": self.expand_all_images(rec), "gt": self.expand_all_images(gt)}

    def _draw_partial_decoding_idx(self):
        """Which cameras keep their ground-truth tokens (reference log_images :503-515): mode 2 = one or two random cameras, mode 3 =
        [0] or [0, 2] with equal probability, mode 4 = [3, 0, 2], any other truthy value = one random camera; None when off."""
        if not self.partial_decoding:
            return None
        n = self.transformer.cfg.num_cams
        if self.partial_decoding == 2:
            return torch.randint(n, (int(torch.randint(1, 3, ()).item()),))
        if self.partial_decoding == 3:
            return torch.tensor([0]) if torch.rand(()).item() > 0.5 else torch.tensor([0, 2])
        if self.partial_decoding == 4:
            return torch.tensor([3, 0, 2])
        return torch.randint(n, (1,))
