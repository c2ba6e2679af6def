"""Encoder / Decoder of the stage-1 VQGAN under the reference's import path and state-dict key names
(reference modules/stage1/model.py:34-192,342-537).  The nn.Conv2d / GroupNorm children only HOLD parameters
(so checkpoints load unchanged); `forward` runs the bevgen_b200 VQGAN engine (NHWC sm_100a kernels)."""
import torch
import torch.nn as nn


def nonlinearity(x):
    return x * torch.sigmoid(x)


def Normalize(in_channels):
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=0):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = Normalize(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            if conv_shortcut:
                raise NotImplementedError("conv_shortcut=True is never used by the shipped configs")
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)


class _EngineBacked(nn.Module):
    """Builds (and caches) a VQGANEngine from this module's own parameters; rebuilt after load_state_dict / .to()."""
    _prefix = ""

    def _init_engine_cache(self, ddconfig):
        self._dd = dict(ddconfig)
        self._engine = None
        self._engine_key = None
        self.precision = "f16f8"

    def _load_from_state_dict(self, *a, **k):
        self._engine = None
        return super()._load_from_state_dict(*a, **k)

    def engine(self):
        from bevgen_b200.vqgan_engine import VQGANEngine
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("bevgen_b200 VQGAN runs on a CUDA device only (no CPU fallback): call .cuda() first")
        key = (p.device, self.precision, tuple(q._version for q in self.parameters()))
        if self._engine is None or self._engine_key != key:
            sd = {self._prefix + k: v.detach() for k, v in self.state_dict().items()}
            self._engine = VQGANEngine(sd, self._dd, device=p.device, precision=self.precision)
            self._engine_key = key
        return self._engine


class Encoder(_EngineBacked):
    _prefix = "encoder."

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, **ignore_kwargs):
        super().__init__()
        if not resamp_with_conv:
            raise NotImplementedError("resamp_with_conv=False (avg-pool) is never used by the shipped configs")
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res //= 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, 1, 1)
        self._init_engine_cache(dict(ch=ch, ch_mult=list(ch_mult), num_res_blocks=num_res_blocks, in_channels=in_channels,
                                     out_ch=out_ch, z_channels=z_channels, resolution=resolution))

    @torch.no_grad()
    def forward(self, x):
        eng = self.engine()
        return eng.nhwc_to_nchw(eng.encoder(x))


class Decoder(_EngineBacked):
    _prefix = "decoder."

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, **ignorekwargs):
        super().__init__()
        if give_pre_end or not resamp_with_conv:
            raise NotImplementedError("give_pre_end / resamp_with_conv=False are never used by the shipped configs")
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)
        self._init_engine_cache(dict(ch=ch, ch_mult=list(ch_mult), num_res_blocks=num_res_blocks, in_channels=in_channels,
                                     out_ch=out_ch, z_channels=z_channels, resolution=resolution))

    @torch.no_grad()
    def forward(self, z):
        eng = self.engine()
        self.last_z_shape = z.shape
        return eng.decoder(eng.nchw_to_nhwc(z))
