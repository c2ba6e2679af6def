"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference from /root/reference.

Used exclusively by ``oracle/make_golden.py`` (in the build container, where /root/reference
is mounted) to mint the golden vectors under ``tests/golden/`` and to pin the CPU restatements
in ``oracle/``.  Nothing here may be imported by the product (``bevgen_b200``), and nothing here
can run on the GPU box (the reference tree does not travel).

Recipe follows SURVEY.md Appendix C: stage-1 files load by path as they only need
torch/numpy/einops; the stage-2 transformer files need ``sys.modules`` stubs for the absent
third-party packages (pyrootutils, deepspeed, matplotlib, nuscenes, ...), and DeepSpeed's Triton
block-sparse ops are replaced by a dense fp32 restatement of
``multi_view_generation/modules/transformer/sparse_self_attention.py:128-177``.
"""
import importlib.util
import sys
import types
from pathlib import Path
from unittest import mock

import torch

REF = Path("/root/reference")


def available() -> bool:
    return (REF / "multi_view_generation").is_dir()


def _load_by_path(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def stage1():
    """Returns (model_module, quantize_module) of the reference, loaded unmodified."""
    m = _load_by_path("ref_stage1_model", "multi_view_generation/modules/stage1/model.py")
    q = _load_by_path("ref_stage1_quantize", "multi_view_generation/modules/stage1/quantize.py")
    return m, q


def dense_sparse_self_attention_forward(self, query, key, value, rpe=None, key_padding_mask=None,
                                        attn_mask=None, add_mask=None):
    """Dense fp32 stand-in for the DeepSpeed sdd -> (+bias) -> softmax(scale, mul-mask) -> dsd chain.

    Order of operations follows sparse_self_attention.py:153 (QK^T), :155-163 (bias add, BEFORE
    scaling), :166-173 (softmax of scale*x with 'mul' mask: 0 -> -inf), :176 (PV); the per-head
    block layout (master_layout, :59-60) restricts the support.
    """
    b, h, L, d = query.shape
    blk = self.sparsity_config.block
    s = torch.matmul(query.float(), key.float().transpose(-1, -2))
    if add_mask is not None:
        s = s + add_mask.float()[:, None]
    s = s * (float(d) ** -0.5)
    layout = self.master_layout[..., : L // blk, : L // blk].to(torch.bool)
    dense = layout.repeat_interleave(blk, -2).repeat_interleave(blk, -1)  # (h, L, L)
    s = s.masked_fill(~dense[None], float("-inf"))
    if attn_mask is not None:
        s = s.masked_fill(attn_mask.squeeze()[None, None] == 0, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, value.float()).to(query.dtype)


_stage2_cache = None


def stage2():
    """Returns the reference's mingpt_sparse module (GPTConfig, GPT) with the dense stand-in patched in."""
    global _stage2_cache
    if _stage2_cache is not None:
        return _stage2_cache
    for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.ticker", "image_utils", "nuscenes",
              "nuscenes.map_expansion", "nuscenes.map_expansion.map_api", "nuscenes.nuscenes", "pyquaternion",
              "shapely", "shapely.geometry", "cv2", "PIL", "PIL.Image", "PIL.ImageOps", "torchvision",
              "torchvision.transforms", "torchvision.transforms.functional", "av2", "wandb", "seaborn"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = mock.MagicMock()
    pr = types.ModuleType("pyrootutils")
    pr.setup_root = lambda **k: REF
    sys.modules["pyrootutils"] = pr
    ds, dso, dsa = types.ModuleType("deepspeed"), types.ModuleType("deepspeed.ops"), types.ModuleType(
        "deepspeed.ops.sparse_attention")

    class SparsityConfig:
        def __init__(self, num_heads, block=16, different_layout_per_head=False):
            self.num_heads, self.block, self.different_layout_per_head = num_heads, block, different_layout_per_head

    dsa.SparsityConfig = SparsityConfig
    sys.modules["deepspeed"], sys.modules["deepspeed.ops"], sys.modules["deepspeed.ops.sparse_attention"] = ds, dso, dsa
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    # our own drop-in package has the same top-level name; make sure the REFERENCE one wins here
    for k in [k for k in sys.modules if k == "multi_view_generation" or k.startswith("multi_view_generation.")]:
        del sys.modules[k]
    from multi_view_generation.modules.transformer import mingpt_sparse
    from multi_view_generation.modules.transformer import sparse_self_attention as ssa
    ssa.SparseSelfAttention.forward = dense_sparse_self_attention_forward
    _stage2_cache = mingpt_sparse
    return mingpt_sparse


def muse():
    """Returns the reference's stage2/muse_maskgit_pytorch module (MaskGitTransformerMultiView, SelfCritic, MaskGit), unmodified.  The
    lucidrains `muse_maskgit_pytorch` package it imports two unused symbols from (:16-17) is absent offline and stubbed."""
    stage2()
    for n in ["muse_maskgit_pytorch", "muse_maskgit_pytorch.vqgan_vae", "muse_maskgit_pytorch.t5"]:
        if n not in sys.modules:
            sys.modules[n] = mock.MagicMock()
    try:
        import beartype  # noqa: F401
    except Exception:
        bt = types.ModuleType("beartype")
        bt.beartype = lambda f: f
        sys.modules["beartype"] = bt
    import torchvision.transforms as T
    if isinstance(sys.modules.get("torchvision"), mock.MagicMock):
        pass
    return _load_by_path("ref_muse_maskgit", "multi_view_generation/modules/stage2/muse_maskgit_pytorch.py")


_STUB_ROOTS = {"pytorch_lightning", "hydra", "omegaconf", "rich", "lpips", "kornia", "torchmetrics", "taming", "image_utils", "wandb", "cv2",
               "nuscenes", "pyquaternion", "shapely", "av2", "seaborn", "matplotlib", "PIL", "imageio", "lightning_utilities",
               "lightning_fabric", "deepspeed", "descartes", "skimage"}


def vqgan():
    """Returns the reference's modules/stage1/vqgan.py (the LightningModule VQModel itself, unmodified).  pytorch_lightning, hydra and the
    logging / plotting packages it pulls in through multi_view_generation.utils are absent offline: a meta-path finder answers every import
    below those roots with a MagicMock package, `pytorch_lightning.LightningModule` being a bare nn.Module."""
    import importlib.abc
    import importlib.machinery
    stage2()

    class _LM(torch.nn.Module):
        pass

    class _Loader(importlib.abc.Loader):
        def create_module(self, spec):
            m = mock.MagicMock(name=spec.name)
            m.__path__, m.__name__, m.__spec__ = [], spec.name, spec
            if spec.name == "pytorch_lightning":
                m.LightningModule = _LM
            return m

        def exec_module(self, module):
            pass

    class _RefStubFinder(importlib.abc.MetaPathFinder):
        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in _STUB_ROOTS:
                return importlib.machinery.ModuleSpec(name, _Loader(), is_package=True)
            return None
    for k in [k for k in sys.modules if k.split(".")[0] in _STUB_ROOTS]:
        del sys.modules[k]
    if not any(type(f).__name__ == "_RefStubFinder" for f in sys.meta_path):
        sys.meta_path.insert(0, _RefStubFinder())
    import multi_view_generation.utils  # noqa: F401  (first: the reference has a utils <-> vqgan import cycle that only resolves in this order)
    from multi_view_generation.modules.stage1 import vqgan as rv
    return rv


def release_stage2():
    """Drop the reference package from sys.modules / sys.path so our drop-in package can be imported."""
    global _stage2_cache
    _stage2_cache = None
    for k in [k for k in sys.modules if k == "multi_view_generation" or k.startswith("multi_view_generation.")]:
        del sys.modules[k]
    if str(REF) in sys.path:
        sys.path.remove(str(REF))
