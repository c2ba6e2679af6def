"""VectorQuantizer2 under the reference's import path (reference modules/stage1/quantize.py:213-329).
Inference semantics only: nearest code (bit-exact argmin of |z|^2+|e|^2-2z.e), gather, commitment loss value."""
import torch
import torch.nn as nn


class VectorQuantizer2(nn.Module):
    def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
        super().__init__()
        if remap is not None:
            raise NotImplementedError("index remapping is never used by the shipped configs")
        self.n_e, self.e_dim, self.beta, self.legacy = n_e, e_dim, beta, legacy
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)
        self.remap, self.re_embed = None, n_e
        self.sane_index_shape = sane_index_shape
        self._sq = None

    def _code_sqnorm(self):
        from bevgen_b200 import ops
        w = self.embedding.weight
        key = (w.device, w._version, w.data_ptr())
        if self._sq is None or self._sq[0] != key:
            out = torch.empty(self.n_e, device=w.device)
            ops.row_sqnorm(w.detach().float().contiguous(), out)
            self._sq = (key, out)
        return self._sq[1]

    @torch.no_grad()
    def forward_nhwc(self, z_nhwc):
        """z (N,H,W,e) fp32 CUDA -> (z_q NHWC, idx (N*H*W,), loss)"""
        from bevgen_b200 import ops
        n, h, w, e = z_nhwc.shape
        rows = n * h * w
        book = self.embedding.weight.detach().float().contiguous()
        idx = torch.empty(rows, dtype=torch.int64, device=z_nhwc.device)
        zq = torch.empty_like(z_nhwc)
        ws = torch.empty(rows, dtype=torch.float32, device=z_nhwc.device)
        ops.vq_nearest(z_nhwc.view(rows, e), book, self._code_sqnorm(), ws, idx, zq.view(rows, e))
        mse = torch.mean((zq - z_nhwc) ** 2)      # scalar diagnostic only (quantize.py:290-295); not on the hot path
        loss = mse + self.beta * mse
        return zq, idx, loss

    @torch.no_grad()
    def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
        assert temp is None or temp == 1.0, "Only for interface compatible with Gumbel"
        assert not rescale_logits and not return_logits, "Only for interface compatible with Gumbel"
        from bevgen_b200 import ops
        if not z.is_cuda:
            raise RuntimeError("bevgen_b200 VectorQuantizer2 runs on a CUDA device only (no CPU fallback)")
        n, c, h, w = z.shape
        z = z.float().contiguous()
        z_nhwc = torch.empty((n, h, w, c), dtype=torch.float32, device=z.device)
        ops.transpose_f32(z, z_nhwc, n, c, h * w)
        zq_nhwc, idx, loss = self.forward_nhwc(z_nhwc)
        zq = torch.empty_like(z)
        ops.transpose_f32(zq_nhwc, zq, n, h * w, c)
        if self.sane_index_shape:
            idx = idx.reshape(n, h, w)
        return zq, loss, (None, None, idx)

    @torch.no_grad()
    def get_codebook_entry(self, indices, shape):
        """indices (flat int64) -> (B,C,H,W) fp32 (shape given as (B,H,W,C)); quantize.py:314-329."""
        from bevgen_b200 import ops
        book = self.embedding.weight.detach().float().contiguous()
        idx = indices.reshape(-1).to(device=book.device, dtype=torch.int64).contiguous()
        out = torch.empty((idx.numel(), self.e_dim), dtype=torch.float32, device=book.device)
        ops.codebook_gather(book, idx, out)
        if shape is None:
            return out
        b, h, w, c = shape
        res = torch.empty((b, c, h, w), dtype=torch.float32, device=book.device)
        ops.transpose_f32(out, res, b, h * w, c)
        return res
