"""Host-side geometry (decode order, mask, camera-bias prior, block layout) vs goldens minted from the
reference's GPTConfig.__post_init__ (mingpt_sparse.py:74-102).  Integer/bool artefacts: bit-exact."""
import numpy as np
import pytest
import torch

from bevgen_b200.gpt_config import GPTConfig
from tests.cases import CONFIG_CASES


@pytest.mark.parametrize("name", list(CONFIG_CASES))
def test_gptconfig_matches_reference(name, golden_dir):
    g = np.load(golden_dir / f"gptconfig_{name}.npz")
    cfg = GPTConfig(**CONFIG_CASES[name])
    L, nc, ni, npad = g["sizes"]
    assert (cfg.gpt_block_size, cfg.num_cond_tokens, cfg.num_img_tokens, cfg.num_pad_tokens) == (L, nc, ni, npad)
    assert np.array_equal(cfg.forward_shuffle_idx.numpy(), g["forward_shuffle_idx"])
    assert torch.equal(cfg.backward_shuffle_idx, torch.argsort(cfg.forward_shuffle_idx))
    mask = np.unpackbits(g["attention_mask_bits"])[: L * L].reshape(L, L).astype(bool)
    assert np.array_equal(cfg.attention_mask.numpy().astype(bool), mask)
    layouts, allowed = cfg.get_mask()
    shp = tuple(g["layout_shape"])
    ref_layout = np.unpackbits(g["layout_bits"])[: int(np.prod(shp))].reshape(shp).astype(bool)
    assert np.array_equal(layouts.numpy().astype(bool), ref_layout)
    assert allowed.shape == (cfg.num_heads, L, L)
    assert cfg.prob_matrix.dtype == torch.float64
    np.testing.assert_allclose(cfg.prob_matrix[g["prob_rows"]].numpy(), g["prob_values"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(torch.diagonal(cfg.prob_matrix).numpy(), g["prob_diag"], rtol=0, atol=1e-12)
    assert abs(cfg.prob_matrix.sum().item() - float(g["prob_sum"])) < 1e-6
    assert cfg.layout_covers_mask(layouts)        # density=1.0: the block layout never removes an allowed position


@pytest.mark.parametrize("name", ["nusc6_16x16", "nusc6_16x16_noncausal", "nusc6_7x9"])
def test_outward_pattern_matches_reference(name, golden_dir):
    """mask_generator.outward_pattern under the reference's import path: the intermediate tuple (allowed pattern, static block layout,
    block prior, padded prior) against the reference's own function (goldens minted by oracle/make_golden.py outward), dtypes included."""
    from multi_view_generation.modules.transformer import mask_generator as mg
    g = np.load(golden_dir / f"outward_{name}.npz")
    cfg = GPTConfig(**CONFIG_CASES[name])
    allowed, static_layout, prob_layout, prob_matrix = mg.outward_pattern(cfg)
    assert (str(allowed.dtype), str(static_layout.dtype), str(prob_layout.dtype), str(prob_matrix.dtype)) == (
        str(g["allowed_dtype"]), str(g["static_dtype"]), str(g["prob_layout_dtype"]), str(g["prob_dtype"]))
    assert tuple(allowed.shape) == tuple(g["allowed_shape"]) and tuple(static_layout.shape) == tuple(g["static_shape"])
    L = cfg.gpt_block_size
    ref_allowed = np.unpackbits(g["allowed_bits"])[: L * L].reshape(L, L).astype(bool)
    for h in (0, cfg.num_heads - 1):
        assert np.array_equal(allowed[h].numpy().astype(bool), ref_allowed)
    nb = static_layout.shape[0]
    assert np.array_equal(static_layout.numpy().astype(bool), np.unpackbits(g["static_bits"])[: nb * nb].reshape(nb, nb).astype(bool))
    np.testing.assert_allclose(prob_layout.numpy(), g["prob_layout"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(prob_matrix[g["prob_rows"]].numpy(), g["prob_values"], rtol=0, atol=1e-12)
    assert abs(prob_matrix.sum().item() - float(g["prob_sum"])) < 1e-6
    bias = mg.outward_pattern(cfg, return_camera_bias_matrix=True)
    assert bias.dtype == torch.float64 and abs(bias.sum().item() - float(g["bias_sum"])) < 1e-6
    layouts, allowed2 = mg.multi_outward_pattern(cfg)
    assert torch.equal(allowed2, allowed) and layouts.dtype == torch.int64


def test_mask_closed_form():
    """allowed(i,j) = (j < n_cond) or (i >= n_cond and j <= i) — the property the KV cache rests on."""
    cfg = GPTConfig(**CONFIG_CASES["nusc6_16x16"])
    L, nc = cfg.gpt_block_size, cfg.num_cond_tokens
    i, j = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    expect = (j < nc) | ((i >= nc) & (j <= i))
    assert np.array_equal(cfg.attention_mask.numpy().astype(bool), expect)
    assert abs(cfg.attention_mask.mean().item() - 0.5104) < 1e-3


def test_first_decode_entries():
    cfg = GPTConfig(**CONFIG_CASES["nusc6_16x16"])
    assert cfg.forward_shuffle_idx[:8].tolist() == [7, 263, 8, 264, 6, 262, 9, 265]     # SURVEY §8 a10 probe


def test_errors():
    kw = dict(CONFIG_CASES["nusc6_16x16"])
    with pytest.raises(AssertionError):
        GPTConfig(**{**kw, "num_cams": 5})
    with pytest.raises(NotImplementedError):
        GPTConfig(**{**kw, "legacy_prob_matrix": False})


@pytest.mark.parametrize("block", [16, 32, 64, 128])
def test_layout_bit_table_for_fused_attention(block):
    """ops.layout_to_tiles64: the 16-position bit table the fused attention kernel ANDs with its closed-form mask reproduces the DeepSpeed
    block layout (sparse_self_attention.py:59-60) element for element, for every supported sparse_block_size."""
    from bevgen_b200 import ops
    L, H = 1792, 3
    nb = -(-L // block)
    g = torch.Generator().manual_seed(block)
    lay = (torch.rand(H, nb, nb, generator=g) < 0.4).to(torch.uint8)
    table = ops.layout_to_tiles64(lay, block, L)
    assert table.shape == (H, L // 128, L // 128) and table.dtype == torch.int64
    i = torch.arange(L)
    want = lay[:, (i // block)[:, None], (i // block)[None, :]].bool()                           # [H][L][L]
    e = table[:, (i // 128)[:, None], (i // 128)[None, :]]                                       # [H][L][L] int64
    bit = (8 * ((i % 128) // 16))[:, None] + ((i % 128) // 16)[None, :]
    got = ((e >> bit) & 1).bool()
    assert torch.equal(got, want)
    assert torch.equal(table == 0, ~want.reshape(H, L // 128, 128, L // 128, 128).any(4).any(2))  # zero entry <=> tile skipped


def test_maskgit_module_state_dict_matches_the_reference():
    """Checkpoint compatibility of the MaskGit variant: `MaskGit(...).state_dict()` of the drop-in modules has exactly the keys and shapes
    of the reference's (minted from the unmodified muse_maskgit_pytorch.py with the same constructor arguments; includes the SelfCritic
    head, which the reference registers both as `token_critic.net.*` and `transformer.*`)."""
    import json
    from pathlib import Path
    from multi_view_generation.modules.stage2 import muse_maskgit_pytorch as m
    from tests.cases import GPT_SMALL
    want = json.load(open(Path(__file__).parent / "golden" / "maskgit_small_state_dict_keys.json"))
    cfg = GPTConfig(**GPT_SMALL)
    tr = m.MaskGitTransformerMultiView(num_tokens=cfg.vocab_size, dim=cfg.num_embed, seq_len=tuple(cfg.cam_latent_res), depth=2, dim_head=64,
                                       heads=cfg.num_heads, ff_mult=4, cfg=cfg)
    model = m.MaskGit(image_size=tuple(cfg.cam_latent_res), transformer=tr, self_token_critic=True)
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == want, (sorted(set(want) - set(got))[:5], sorted(set(got) - set(want))[:5])
