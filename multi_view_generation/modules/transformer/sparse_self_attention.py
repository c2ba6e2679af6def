"""`SparseSelfAttention` under the reference's import path (reference modules/transformer/sparse_self_attention.py:11-177).

Only the state that callers touch is kept: the per-head block layout buffer `master_layout` (logged by
Net2NetTransformer.on_test_start and present in checkpoints).  The arithmetic — QK^T + camera bias (added BEFORE the
1/sqrt(d_head) scale) + mask + softmax + PV — runs in the bevgen_b200 attention kernels.  At density = 1.0 the layout is RNG-free
and identical on every rank; at density < 1 it is drawn from the global RNG, and the reference's `dist.broadcast(master_layout)`
(:50-52) happens inside bevgen_b200.sharding.broadcast_module_weights, which ships integer buffers with the weights.
"""
import torch.nn as nn


class SparseSelfAttention(nn.Module):
    def __init__(self, sparsity_config=None, key_padding_mask_mode="add", attn_mask_mode="mul", max_seq_length=2048, layout=None):
        super().__init__()
        self.sparsity_config = sparsity_config
        if layout is None and sparsity_config is not None:
            layout = sparsity_config.make_layout(max_seq_length)
        self.register_buffer("master_layout", layout)
        self._need_layout_synchronization = False
        self.key_padding_mask_mode, self.attn_mask_mode = key_padding_mask_mode, attn_mask_mode

    def get_layout(self, L):
        block = self.master_layout.shape[-1] and (L // self.master_layout.shape[-1])
        if self.sparsity_config is not None:
            block = self.sparsity_config.block
        if L % block != 0:
            raise ValueError(f"Sequence Length, {L}, needs to be dividable by Block size {block}!")
        nb = L // block
        return self.master_layout[..., :nb, :nb].cpu()

    def forward(self, *a, **k):
        raise RuntimeError("SparseSelfAttention is executed inside bevgen_b200's fused attention path (GPT.forward); "
                           "there is no standalone eager implementation")
