"""VQModel / VQSegmentationModel under the reference's import path (reference modules/stage1/vqgan.py:31-261).

Constructor kwargs, attribute names (`encoder`, `decoder`, `quantize`, `quant_conv`, `post_quant_conv`, `colorize`) and
`encode(x, batch) -> (quant, emb_loss, (None, None, indices))` / `decode(quant)` match the reference so Hydra configs and
checkpoints load unchanged.  All arithmetic runs in the bevgen_b200 kernels; training steps are out of scope.
"""
import torch
import torch.nn as nn

from bevgen_b200.engine_cache import EngineCacheMixin
from multi_view_generation import utils
from multi_view_generation.modules.stage1.model import Decoder, Encoder
from multi_view_generation.modules.stage1.quantize import VectorQuantizer2 as VectorQuantizer

try:  # Lightning is optional (absent in the build image); the module surface is identical either way
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pragma: no cover
    _Base = nn.Module


class VQModel(EngineCacheMixin, _Base):
    def __init__(self, ddconfig, lossconfig, n_embed, embed_dim, cam_res, cam_latent_res, cam_emd_dim, geometric_embedding=False,
                 ckpt_path=None, ignore_keys=[], image_key="image", colorize_nlabels=None, monitor=None, remap=None,
                 sane_index_shape=False, denormalize=True, legacy=True, precision="f16f8", **kwargs):
        super().__init__()
        self.image_key, self.denormalize = image_key, denormalize
        self.cam_res = tuple(cam_res)
        self.ddconfig = dict(ddconfig)
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.loss = lossconfig
        self.geometric_embedding = geometric_embedding
        self.n_embed, self.embed_dim = n_embed, embed_dim
        if geometric_embedding:      # reference :62-69: ray / camera-centre embeddings added to the encoder output
            if cam_emd_dim != ddconfig["z_channels"]:
                raise ValueError("geometric_embedding adds a cam_emd_dim-channel embedding to the z_channels-channel encoder output")
            fh, fw = cam_latent_res
            xs, ys = torch.linspace(0, 1, fw), torch.linspace(0, 1, fh)
            gx, gy = torch.meshgrid((xs, ys), indexing="xy")
            plane = torch.stack([gx * cam_res[1], gy * cam_res[0], torch.ones_like(gx)], 0)[None, None]          # 1 1 3 h w
            self.register_buffer("image_plane", plane.contiguous(), persistent=False)
            self.img_embed = nn.Conv2d(4, cam_emd_dim, 1, bias=False)
            self.cam_embed = nn.Conv2d(4, cam_emd_dim, 1, bias=False)
        self.quantize = VectorQuantizer(n_embed, embed_dim, beta=0.25, remap=remap, sane_index_shape=sane_index_shape, legacy=legacy)
        self.quant_conv = nn.Conv2d(ddconfig["z_channels"], embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.precision = precision
        self._engine, self._engine_key = None, None
        if ckpt_path is not None:
            utils.init_from_ckpt(self, ckpt_path, ignore_keys=ignore_keys, strict=False)
        if colorize_nlabels is not None:
            assert type(colorize_nlabels) == int
            self.register_buffer("colorize", torch.randn(3, colorize_nlabels, 1, 1))
        if monitor is not None:
            self.monitor = monitor

    # ---------------------------------------------------------------- engine cache
    def engine(self):
        from bevgen_b200.vqgan_engine import VQGANEngine
        p = self.quant_conv.weight
        if not p.is_cuda:
            raise RuntimeError("bevgen_b200 VQModel runs on a CUDA device only (no CPU fallback): call .cuda() first")
        key = self._engine_cache_key(p.device, self.precision)       # (data_ptr, _version) of every parameter and buffer
        if self._engine is None or self._engine_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items() if k != "colorize"}
            self._engine = VQGANEngine(sd, self.ddconfig, self.n_embed, self.embed_dim, device=p.device, precision=self.precision)
            self._engine_key = key
        return self._engine

    # ---------------------------------------------------------------- reference surface
    @torch.no_grad()
    def encode(self, x, batch=None):
        eng = self.engine()
        zq, idx, h = eng.encode(x, batch, self.cam_res) if self.geometric_embedding else eng.encode(x)
        mse = torch.mean((zq - h) ** 2)
        emb_loss = mse + self.quantize.beta * mse          # value of quantize.py:290-295 (legacy and non-legacy coincide at inference)
        if self.quantize.sane_index_shape:
            idx = idx.reshape(zq.shape[0], zq.shape[1], zq.shape[2])
        return eng.nhwc_to_nchw(zq), emb_loss, (None, None, idx)

    @torch.no_grad()
    def decode(self, quant):
        eng = self.engine()
        return eng.decode_nhwc(eng.nchw_to_nhwc(quant))

    @torch.no_grad()
    def decode_indices(self, indices, shape_bhwc):
        """get_codebook_entry + decode without the NHWC->NCHW->NHWC round trip (used by stage 2's decode_to_img)."""
        b, h, w, _ = shape_bhwc
        return self.engine().decode_indices(indices, b, h, w)

    def forward(self, input, batch=None):
        quant, diff, _ = self.encode(input, batch)
        return self.decode(quant), diff

    def get_input(self, batch, k):
        x = batch[k]
        if len(x.shape) == 3:
            x = x[..., None]
        if len(x.shape) == 5:
            x = x.flatten(0, 1)
        return x.permute(0, 3, 1, 2).to(memory_format=torch.contiguous_format).float()

    def get_last_layer(self):
        return self.decoder.conv_out.weight


class VQSegmentationModel(VQModel):
    def __init__(self, n_labels, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("colorize", torch.randn(3, n_labels, 1, 1))
