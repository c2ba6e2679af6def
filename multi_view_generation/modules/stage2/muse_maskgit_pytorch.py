"""MaskGit stage-2 variant under the reference's import path (reference modules/stage2/muse_maskgit_pytorch.py).

Inference surface: `MaskGitTransformerMultiView` / `TokenCritic` (.forward, .forward_with_cond_scale), `SelfCritic`, `MaskGit.generate`
with the reference's constructor arguments and state_dict key names (parameters live here as plain tensors, the arithmetic runs in
bevgen_b200.maskgit_engine on the CUDA library).  The training forward (random masking + cross-entropy / critic BCE, :629-728) is out of
scope like the rest of training.
"""
from typing import Callable, Iterable, Optional

import torch
from torch import nn

from bevgen_b200.engine_cache import EngineCacheMixin, fingerprint
from multi_view_generation.modules.transformer.mingpt_sparse import GPTConfig, get_bev_grid


def cosine_schedule(t):
    return torch.cos(t * torch.pi * 0.5)


class LayerNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))


def _feed_forward(dim, mult=4):
    inner = int(dim * mult * 2 / 3)
    return nn.Sequential(LayerNorm(dim), nn.Linear(dim, inner * 2, bias=False), nn.Identity(), LayerNorm(inner), nn.Linear(inner, dim, bias=False))


class Attention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, cross_attend=False, scale=8, cfg: Optional[GPTConfig] = None):
        super().__init__()
        inner = dim_head * heads
        self.scale, self.heads, self.cross_attend, self.cfg = scale, heads, cross_attend, cfg
        self.norm = LayerNorm(dim)
        self.null_kv = nn.Parameter(torch.randn(2, heads, 1, dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.q_scale = nn.Parameter(torch.ones(dim_head))
        self.k_scale = nn.Parameter(torch.ones(dim_head))
        self.to_out = nn.Linear(inner, dim, bias=False)


class TransformerBlocks(nn.Module):
    def __init__(self, *, dim, depth, dim_head=64, heads=8, ff_mult=4, cfg: Optional[GPTConfig] = None):
        super().__init__()
        self.layers = nn.ModuleList([nn.ModuleList([Attention(dim=dim, dim_head=dim_head, heads=heads, cfg=cfg),
                                                    Attention(dim=dim, dim_head=dim_head, heads=heads, cross_attend=True, cfg=cfg),
                                                    _feed_forward(dim=dim, mult=ff_mult)]) for _ in range(depth)])
        self.norm = LayerNorm(dim)


class TransformerMultiView(EngineCacheMixin, nn.Module):
    def __init__(self, *, num_tokens, dim, seq_len, dim_out=None, self_cond=False, add_mask_id=False, cfg: Optional[GPTConfig] = None,
                 precision="f16f8", **kwargs):
        super().__init__()
        if self_cond:
            raise NotImplementedError("self-conditioning is off in the reference's config (muse_stage_two_multi_view.yaml) and not built")
        self.cfg, self.dim, self.precision = cfg, dim, precision
        self.seq_len = seq_len[0] * seq_len[1] if isinstance(seq_len, Iterable) else seq_len
        self.mask_id = num_tokens if add_mask_id else None
        self.num_tokens = num_tokens
        self.token_emb = nn.Embedding(num_tokens + int(add_mask_id), dim)
        self.pos_emb = nn.Embedding(cfg.num_img_tokens, dim)
        self.cond_token_emb = nn.Embedding(cfg.cond_vocab_size, dim)
        self.cond_pos_emb = nn.Embedding(cfg.num_cond_tokens, dim)
        self._block_kw = dict(depth=kwargs["depth"], heads=kwargs.get("heads", 8), dim_head=kwargs.get("dim_head", 64), ff_mult=kwargs.get("ff_mult", 4))
        self.transformer_blocks = TransformerBlocks(dim=dim, cfg=cfg, **kwargs)
        self.norm = LayerNorm(dim)
        self.dim_out = num_tokens if dim_out is None else dim_out
        self.to_logits = nn.Linear(dim, self.dim_out, bias=False)
        self.self_cond = self_cond
        self.self_cond_to_init_embed = _feed_forward(dim)
        if cfg.image_embed:
            self.img_embed = nn.Conv2d(4, cfg.num_embed, 1, bias=False)
            self.cam_embed = nn.Conv2d(4, cfg.num_embed, 1, bias=False)
        if cfg.bev_embed:
            self.register_buffer("bev_grid", get_bev_grid(cfg))
            self.bev_embed = nn.Conv2d(2, cfg.num_embed, 1)
            self.bev_cam_pos_emb = nn.Parameter(torch.zeros(1, cfg.num_cams, cfg.num_cond_tokens, cfg.num_embed))
        if cfg.camera_bias:
            L = cfg.gpt_block_size
            self.camera_bias_emb = nn.Parameter(torch.zeros(1, L * (L + 1) // 2))
        self._engine, self._engine_key = None, None

    def engine(self, critic=None):
        from bevgen_b200.maskgit_engine import MaskGitEngine
        p = self.to_logits.weight
        if not p.is_cuda:
            raise RuntimeError("bevgen_b200 MaskGit runs on a CUDA device only (no CPU fallback): call .cuda() first")
        cw = None if critic is None else fingerprint(critic)
        key = (self._engine_cache_key(p.device, self.precision), cw)
        if self._engine is None or self._engine_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            cr = None if critic is None else {"weight": critic.weight.detach(), "bias": critic.bias.detach()}
            self._engine = MaskGitEngine(sd, self.cfg, device=p.device, precision=self.precision, critic=cr, **self._block_kw)
            self._engine_key = key
        return self._engine

    @torch.no_grad()
    def forward(self, x, return_embed=False, return_logits=False, labels=None, ignore_index=0, self_cond_embed=None, cond_drop_prob=0.,
                conditioning_token_ids: Optional[torch.Tensor] = None, batch=None, weights=None):
        if labels is not None or self.training:
            raise NotImplementedError("training forward / losses are out of scope; call .eval()")
        if self.dim_out != self.num_tokens:
            raise NotImplementedError("a separate TokenCritic network (dim_out=1) is not built; the reference's config uses self_token_critic")
        logits, embed = self.engine().forward(x, conditioning_token_ids, batch)
        return (logits, embed) if return_embed else logits

    def forward_with_cond_scale(self, *args, cond_scale=3., return_embed=False, **kwargs):
        """Reference :262-281.  In eval mode the conditioning dropout is inactive (:341), the null pass equals the conditional pass and
        null + (logits - null) * cond_scale == logits: one forward."""
        kwargs.pop("cond_drop_prob", None)
        return self.forward(*args, return_embed=return_embed, cond_drop_prob=0., **kwargs)


class MaskGitTransformerMultiView(TransformerMultiView):
    def __init__(self, *args, **kwargs):
        assert "add_mask_id" not in kwargs
        super().__init__(*args, add_mask_id=True, **kwargs)


class TokenCritic(TransformerMultiView):
    def __init__(self, *args, **kwargs):
        assert "dim_out" not in kwargs
        super().__init__(*args, dim_out=1, **kwargs)


class SelfCritic(nn.Module):
    def __init__(self, net):
        super().__init__()
        self.net = net
        self.to_pred = nn.Linear(net.dim, 1)

    @torch.no_grad()
    def forward_with_cond_scale(self, x, *args, conditioning_token_ids=None, batch=None, **kwargs):
        return self.net.engine(self.to_pred).critic_scores(x, conditioning_token_ids, batch)[..., None]

    forward = forward_with_cond_scale


class MaskGit(nn.Module):
    def __init__(self, image_size, transformer: MaskGitTransformerMultiView, noise_schedule: Callable = cosine_schedule,
                 token_critic: Optional[TokenCritic] = None, self_token_critic=False, cond_image_size=None, cond_drop_prob=0.5,
                 self_cond_prob=0.9, no_mask_token_prob=0., critic_loss_weight=1.):
        super().__init__()
        self.image_size = image_size[0] * image_size[1] if isinstance(image_size, Iterable) else image_size
        self.cond_image_size, self.cond_drop_prob = cond_image_size, cond_drop_prob
        self.transformer = transformer
        self.self_cond = transformer.self_cond
        self.mask_id = transformer.mask_id
        self.noise_schedule = noise_schedule
        assert not (self_token_critic and token_critic is not None)
        if token_critic is not None:
            raise NotImplementedError("a separate TokenCritic network is not built; use self_token_critic=True as the reference's config does")
        self.token_critic = SelfCritic(transformer) if self_token_critic else None
        self.critic_loss_weight, self.self_cond_prob, self.no_mask_token_prob = critic_loss_weight, self_cond_prob, no_mask_token_prob
        self.sample_seed = None          # int for reproducible sampling; default draws from torch's RNG

    @torch.no_grad()
    def generate(self, init_ids: Optional[torch.Tensor] = None, cond_images: Optional[torch.Tensor] = None, fmap_size=None, temperature=1.,
                 topk_filter_thres=0.9, can_remask_prev_masked=False, force_not_use_token_critic=False, timesteps=12, cond_scale=3,
                 critic_noise_scale=1, batch=None):
        """-> LongTensor (b*cam, h, w).  `cond_scale` is accepted and has no effect, as in the reference's eval mode (see engine doc)."""
        use_critic = self.token_critic is not None and not force_not_use_token_critic
        if not use_critic and can_remask_prev_masked:
            assert self.no_mask_token_prob > 0.
        eng = self.transformer.engine(self.token_critic.to_pred if use_critic else None)
        gen = None
        if self.sample_seed is not None:
            gen = torch.Generator(device=eng.dev).manual_seed(int(self.sample_seed))
        assert tuple(fmap_size) == (self.transformer.cfg.cam_latent_h, self.transformer.cfg.cam_latent_w)
        return eng.generate(cond_images, batch, timesteps=timesteps, temperature=temperature, topk_filter_thres=topk_filter_thres,
                            critic_noise_scale=critic_noise_scale, init_ids=init_ids, use_critic=use_critic, generator=gen)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("MaskGit training forward (random masking + losses) is out of scope")
