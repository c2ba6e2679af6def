"""GPTConfig / GPT under the reference's import path (reference modules/transformer/mingpt_sparse.py:26-391).

`GPT` holds its parameters under the reference's state-dict key names (x_tok_emb, cond_tok_emb, x_pos_emb, cond_pos_emb,
blocks.{i}.{ln1,ln2,attention.{query,key,value},mlp.{0,2}}, ln_f, head, img_embed, cam_embed, bev_embed, bev_cam_pos_emb,
camera_bias_emb, blocks.{i}.attention.sparse_self_attention.master_layout) and runs `forward` on the bevgen_b200 engine.
"""
import logging

import torch
import torch.nn as nn

from bevgen_b200.engine_cache import EngineCacheMixin
from bevgen_b200.geometry_torch import bev_grid as _bev_grid
from bevgen_b200.geometry_torch import generate_grid  # noqa: F401
from bevgen_b200.geometry_torch import image_plane as _image_plane
from bevgen_b200.gpt_config import GPTConfig  # noqa: F401
from multi_view_generation.modules.transformer.sparse_self_attention import SparseSelfAttention

logger = logging.getLogger(__name__)


def get_bev_grid(cfg, offset=0):
    return _bev_grid(cfg.bev_latent_res[0], cfg.bev_latent_res[1], offset)


class CustomSparsityConfig:
    def __init__(self, num_heads, layout, block, different_layout_per_head=True):
        self.num_heads, self.block, self.different_layout_per_head, self.layout = num_heads, block, different_layout_per_head, layout

    def make_layout(self, seq_len):
        return self.layout


class CustomSparseSelfAttention(nn.Module):
    """q/k/v projections + block-sparse attention holder; NO output projection (reference :157-212)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        if cfg.hidden_size % cfg.num_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)" % (cfg.hidden_size, cfg.num_heads))
        self.num_attention_heads = cfg.num_heads
        self.attention_head_size = int(cfg.hidden_size / cfg.num_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(cfg.hidden_size, self.all_head_size)
        self.key = nn.Linear(cfg.hidden_size, self.all_head_size)
        self.value = nn.Linear(cfg.hidden_size, self.all_head_size)
        layout, _ = cfg.get_mask()          # drawn per layer, like the reference (identical for density = 1.0)
        self.sparse_self_attention = SparseSelfAttention(CustomSparsityConfig(cfg.num_heads, layout, cfg.sparse_block_size), attn_mask_mode="mul")


class Block(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.ln1 = nn.LayerNorm(cfg.num_embed)
        self.ln2 = nn.LayerNorm(cfg.num_embed)
        if cfg.backend != "deepspeed":
            raise ValueError("only backend='deepspeed' semantics are implemented (the reference's 'pytorch' branch inverts the mask, SURVEY App. B)")
        self.attention = CustomSparseSelfAttention(cfg)
        self.attention_mask = cfg.attention_mask
        self.mlp = nn.Sequential(nn.Linear(cfg.num_embed, 4 * cfg.num_embed), nn.GELU(), nn.Linear(4 * cfg.num_embed, cfg.num_embed),
                                 nn.Dropout(cfg.resid_pdrop))


class GPT(EngineCacheMixin, nn.Module):
    def __init__(self, cfg: GPTConfig, precision="f16f8", **kwargs):
        super().__init__()
        self.cfg = cfg
        self.precision = precision
        self.x_tok_emb = nn.Embedding(cfg.vocab_size + 1, cfg.num_embed)
        self.cond_tok_emb = nn.Embedding(cfg.cond_vocab_size, cfg.num_embed)
        self.x_pos_emb = nn.Parameter(torch.zeros(1, cfg.num_img_tokens, cfg.num_embed))
        self.cond_pos_emb = nn.Parameter(torch.zeros(1, cfg.num_cond_tokens, cfg.num_embed))
        self.drop = nn.Dropout(cfg.embd_pdrop)
        self.blocks = nn.Sequential(*[Block(cfg) for _ in range(cfg.num_layers)])
        self.ln_f = nn.LayerNorm(cfg.num_embed)
        self.head = nn.Linear(cfg.num_embed, cfg.vocab_size, bias=False)
        if cfg.image_embed:
            plane = _image_plane(cfg.cam_latent_h, cfg.cam_latent_w, cfg.cam_res).t().reshape(1, 1, 3, cfg.cam_latent_h, cfg.cam_latent_w)
            self.register_buffer("image_plane", plane.contiguous(), persistent=False)
            self.img_embed = nn.Conv2d(4, cfg.num_embed, 1, bias=False)
            self.cam_embed = nn.Conv2d(4, cfg.num_embed, 1, bias=False)
        if cfg.bev_embed:
            self.register_buffer("bev_grid", get_bev_grid(cfg))
            self.bev_embed = nn.Conv2d(2, cfg.num_embed, 1)
            self.bev_cam_pos_emb = nn.Parameter(torch.zeros(1, cfg.num_cams, cfg.num_cond_tokens, cfg.num_embed))
        if cfg.camera_bias:
            L = cfg.gpt_block_size
            self.camera_bias_emb = nn.Parameter(torch.zeros(1, L * (L + 1) // 2))
        self.apply(self._init_weights)
        self._engine, self._engine_key, self._samplers = None, None, {}
        logger.info("number of parameters: %e", sum(p.numel() for p in self.parameters()))

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(module, nn.Linear) and module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    # ---------------------------------------------------------------- engine cache
    def engine(self):
        from bevgen_b200.gpt_engine import GPTEngine
        p = self.head.weight
        if not p.is_cuda:
            raise RuntimeError("bevgen_b200 GPT runs on a CUDA device only (no CPU fallback): call .cuda() first")
        key = self._engine_cache_key(p.device, self.precision)       # parameters AND buffers (master_layout) by (data_ptr, _version)
        if self._engine is None or self._engine_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            # one layout per layer (each CustomSparseSelfAttention draws / loads its own master_layout, reference :143-154,177)
            layouts = (torch.stack([blk.attention.sparse_self_attention.master_layout for blk in self.blocks])
                       if self.cfg.density < 1.0 else None)
            self._engine = GPTEngine(sd, self.cfg, device=p.device, precision=self.precision, layouts=layouts)
            self._engine_key, self._samplers = key, {}
        return self._engine

    def sampler(self, batch_size):
        from bevgen_b200.gpt_decode import GPTSampler
        eng = self.engine()
        if batch_size not in self._samplers:
            self._samplers = {batch_size: GPTSampler(eng, batch_size)}     # one live KV cache at a time
        return self._samplers[batch_size]

    @torch.no_grad()
    def forward(self, cam_indices, bev_indices, batch, sampling, **kwargs):
        """-> logits (B, num_img_tokens, vocab) in (cam,h,w) order.  Like the reference, `cam_indices` is modified in place when
        sampling=False (last token <- PAD id, reference :328-329)."""
        eng = self.engine()
        if not sampling:
            cam_indices[:, -1, -1] = self.cfg.vocab_size
        logits = eng.forward(cam_indices, bev_indices, batch, sampling)
        if not torch.isfinite(logits).all():
            raise AssertionError("non-finite logits")
        return logits
