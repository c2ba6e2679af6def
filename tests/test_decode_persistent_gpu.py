"""The persistent KV-cache decode kernel (csrc/decode_persistent.cu) piece by piece: the pack kernel against the numpy statement of the
format, every phase of ONE decode step against fp32 torch arithmetic on the same cache contents (localises a bug to QKV / attention /
MLP / head), and the whole loop against the per-launch chain and the reference goldens."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from bevgen_b200 import _lib, decode_format  # noqa: E402
from bevgen_b200.gpt_config import GPTConfig  # noqa: E402
from bevgen_b200.gpt_decode import GPTSampler  # noqa: E402
from bevgen_b200.gpt_engine import GPTEngine  # noqa: E402
from oracle import synth  # noqa: E402
from tests.cases import GPT_CASES, GPT_SMALL, gpt_sizes  # noqa: E402


def test_pack_kernel_matches_the_host_format():
    lib = _lib.init()
    g = torch.Generator().manual_seed(0)
    for rows, d, nq in ((24, 64, 1), (40, 128, 4), (13, 64, 1), (128, 256, 1)):
        w = torch.randn(rows, nq * d, generator=g) * 0.02
        lo_mul = 2.0 ** (19 - int(np.floor(np.log2(float(w.abs().max())))))
        n = lib.bevgen_pack_decode_linear(None, rows, nq * d, d, nq, lo_mul, None, None)
        out = torch.zeros(int(n), dtype=torch.uint8, device="cuda")
        wd = w.cuda()
        assert lib.bevgen_pack_decode_linear(wd.data_ptr(), rows, nq * d, d, nq, lo_mul, out.data_ptr(), None) == n
        torch.cuda.synchronize()
        want = decode_format.pack_reference(w.numpy(), d, nq, lo_mul)
        assert np.array_equal(out.cpu().numpy(), want), (rows, d, nq)


def _one_layer_case(layers=1, B=3):
    kw = {**GPT_SMALL, "num_layers": layers}
    cfg = GPTConfig(**kw)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=4)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3")
    return cfg, sd, cam, bev, batch, eng, B


def _ws_views(smp, cfg, B, c_last):
    """The kernel's workspace layout (launch_decode_persistent): X, X1 [2][16][d], QKV [2][16][3d], LOGITS [16][vpad], ...  Everything that
    crosses CTAs inside a step exists twice; layer number c of a launch (counted over its steps) writes instance c & 1."""
    d, ws = cfg.num_embed, smp._pk["ws"]
    o, out = 0, {}
    for name, n in (("X", 16 * d), ("X1", 16 * d), ("QKV", 16 * 3 * d)):
        out[name] = ws[o + (c_last & 1) * n:o + (c_last & 1) * n + n]
        o += 2 * n
    vpad = (cfg.vocab_size + 7) // 8 * 8
    out["LOGITS"] = ws[o:o + 16 * vpad]
    return {k: v.view(16, -1) for k, v in out.items()}


@pytest.mark.parametrize("last_step", [1, 129, 257, 258, 260, 769])
def test_one_step_phase_by_phase(last_step):
    """One layer; the LAST decode step of a run of `last_step` steps: every intermediate the kernel leaves in its workspace against fp32
    torch on the GPU (steps 129 / 257 / 769: the newest key opens a new 128-key cache block)."""
    cfg, sd, cam, bev, batch, eng, B = _one_layer_case()
    d, H, nc = cfg.num_embed, cfg.num_heads, cfg.num_cond_tokens
    sdg = {k: v.cuda() for k, v in sd.items()}
    smp = GPTSampler(eng, B)
    assert smp.persistent
    forced = cam.reshape(B, -1)[:, cfg.forward_shuffle_idx]
    toks, trace = smp.sample(bev, batch, forced_tokens=forced, trace_logits=True, steps=last_step + 1)
    torch.cuda.synchronize()
    ws = _ws_views(smp, cfg, B, c_last=last_step - 1)       # one layer per step, steps 1 .. last_step in the launch
    # the row the kernel processed last: sequence position nc + last_step - 1 (decode-order token last_step - 1, forced)
    nc = nc + last_step - 1
    x0 = eng.embed(smp.cam_idx, bev.cuda(), batch, sampling=True, row0=nc, nrows=1)[:, 0]
    p = "blocks.0"
    y = F.layer_norm(x0, (d,), sdg[f"{p}.ln1.weight"], sdg[f"{p}.ln1.bias"])
    wqkv = torch.cat([sdg[f"{p}.attention.{n}.weight"] for n in ("query", "key", "value")])
    bqkv = torch.cat([sdg[f"{p}.attention.{n}.bias"] for n in ("query", "key", "value")])
    qkv = y @ wqkv.t() + bqkv
    e_qkv = (ws["QKV"][:B] - qkv).abs().max().item()
    assert e_qkv < 2e-5, f"QKV linear: {e_qkv}"
    # attention of row nc over the cached keys 0..nc (cache contents as the kernel saw them, fp16)
    n = nc + 1
    kc = smp.kc[0].float().permute(0, 1, 2, 4, 3).reshape(B, H, -1, 64)[:, :, :n]        # [B][H][L/128][64][128] -> [B][H][L][64]
    vc = smp.vc[0].float()[:, :, :n]
    q = qkv[:, :d].view(B, H, 64)
    knew, vnew = qkv[:, d:2 * d].view(B, H, 64), qkv[:, 2 * d:].view(B, H, 64)
    # appended key / value: fp16 of the kernel's own fp32 values (one fp16 ulp of slack for values rounding the other way)
    assert (kc[:, :, nc] - knew).abs().max().item() < 1e-3, "appended key"
    assert (vc[:, :, nc] - vnew).abs().max().item() < 1e-3, "appended value"
    s = torch.einsum("bhc,bhjc->bhj", q, kc)
    s = (s + eng.bias[nc, :n][None, None]) * (64 ** -0.5)
    att = torch.einsum("bhj,bhjc->bhc", torch.softmax(s, -1), vc).reshape(B, d)
    x1 = y + att
    e_x1 = (ws["X1"][:B] - x1).abs().max().item()
    assert e_x1 < 2e-5, f"attention + residual: {e_x1}"
    z = F.layer_norm(x1, (d,), sdg[f"{p}.ln2.weight"], sdg[f"{p}.ln2.bias"])
    h = F.gelu(z @ sdg[f"{p}.mlp.0.weight"].t() + sdg[f"{p}.mlp.0.bias"])
    x2 = x1 + h @ sdg[f"{p}.mlp.2.weight"].t() + sdg[f"{p}.mlp.2.bias"]
    e_h = (ws["X"][:B] - x2).abs().max().item()            # the residual stream after MLP1 + GELU + MLP2 (the kernel's X after the last layer)
    assert e_h < 3e-5, f"MLP1 + GELU + MLP2 + residual: {e_h}"
    logits = F.layer_norm(x2, (d,), sdg["ln_f.weight"], sdg["ln_f.bias"]) @ sdg["head.weight"].t()
    e_l = (ws["LOGITS"][:B, : cfg.vocab_size] - logits).abs().max().item()
    assert e_l < 3e-5, f"head logits: {e_l}"
    assert (trace[last_step] - logits).abs().max().item() < 3e-5
    assert torch.equal(toks.reshape(B, -1)[:, cfg.forward_shuffle_idx[: last_step + 1]].cpu(), forced[:, : last_step + 1])
    print(f"step {last_step}, phase by phase: qkv {e_qkv:.1e} x1 {e_x1:.1e} x2 {e_h:.1e} logits {e_l:.1e}")


@pytest.mark.parametrize("name", ["small", "wide2"])
def test_persistent_equals_launch_chain_and_reference_golden(name, golden_dir):
    """All 1536 steps through the persistent kernel: same logits as the per-launch chain (different weight format: 3-byte vs bf16 hi+lo)
    and as the reference's full forward (golden rows) at the 1e-3 bar."""
    g = np.load(golden_dir / f"gpt_{name}.npz")
    kw, B = GPT_CASES[name]
    cfg = GPTConfig(**kw)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=4)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3")
    forced = cam.reshape(B, -1)[:, cfg.forward_shuffle_idx]
    smp = GPTSampler(eng, B)
    assert smp.persistent
    toks, trace = smp.sample(bev, batch, forced_tokens=forced, trace_logits=True)
    torch.cuda.synchronize()
    assert torch.equal(toks.cpu(), cam)
    steps = cfg.backward_shuffle_idx[g["rows"]]
    err = np.abs(trace[steps].permute(1, 0, 2).cpu().numpy() - g["logits_s"]).max()
    old = GPTSampler(eng, B)
    old.persistent = False
    _, trace_old = old.sample(bev, batch, forced_tokens=forced, trace_logits=True)
    torch.cuda.synchronize()
    diff = (trace - trace_old).abs().max().item()
    err_old = np.abs(trace_old[steps].permute(1, 0, 2).cpu().numpy() - g["logits_s"]).max()
    per_step = (trace - trace_old).abs().amax(dim=(1, 2))
    bad = torch.nonzero(per_step > 1e-3).flatten()
    print(f"[{name}] persistent kernel: max logit err vs reference golden {err:.2e}; per-launch chain vs golden {err_old:.2e}; "
          f"persistent vs chain {diff:.2e} (first step over 1e-3: {int(bad[0]) if len(bad) else None}, count {len(bad)})")
    assert err < 1e-3 and err_old < 1e-3 and diff < 5e-4


def test_persistent_sampling_is_reproducible_and_respects_top_k():
    kw, B = GPT_CASES["small"]
    cfg = GPTConfig(**kw)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    _, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=4)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3")
    smp = GPTSampler(eng, B)
    a, tr = smp.sample(bev, batch, temperature=1.0, top_k=5, seed=11, steps=200, trace_logits=True)
    b = smp.sample(bev, batch, temperature=1.0, top_k=5, seed=11, steps=200)
    c = smp.sample(bev, batch, temperature=1.0, top_k=5, seed=12, steps=200)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and not torch.equal(a, c)
    drawn = a.reshape(B, -1)[:, cfg.forward_shuffle_idx[:200]]                     # [B][200]
    kth = tr[:200].topk(5, dim=-1).values[..., -1]                                  # [200][B]
    chosen = tr[:200].gather(2, drawn.t()[..., None])[..., 0]
    assert bool((chosen >= kth).all()), "a token outside the top-k set was drawn"
    assert int(a.max()) <= cfg.vocab_size                                           # un-decoded positions stay PAD (= vocab_size)


@pytest.mark.parametrize("d,heads,vocab,B,layers,steps", [(512, 8, 100, 5, 3, 300), (256, 4, 1000, 16, 2, 140), (1024, 16, 128, 1, 2, 200), (768, 12, 333, 7, 2, 260)])
def test_persistent_kernel_other_widths_batches_and_vocabularies(d, heads, vocab, B, layers, steps):
    """Shapes the goldens do not cover: 8 / 12 k-groups per row (d = 512 / 768: fewer consumer warps hold a k-group than exist), odd batch
    sizes incl. the full 16, vocabularies that are not a multiple of 8 (padded head units).  The launch chain (different weight format and
    kernels, pinned to the reference goldens elsewhere) is the checker; forced tokens keep both on the same path."""
    kw = {**GPT_SMALL, "hidden_size": d, "num_embed": d, "num_heads": heads, "vocab_size": vocab, "num_layers": layers}
    cfg = GPTConfig(**kw)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=9)
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=6)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3")
    forced = cam.reshape(B, -1)[:, cfg.forward_shuffle_idx]
    new = GPTSampler(eng, B)
    assert new.persistent
    toks, trace = new.sample(bev, batch, forced_tokens=forced, trace_logits=True, steps=steps)
    old = GPTSampler(eng, B)
    old.persistent = False
    toks_old, trace_old = old.sample(bev, batch, forced_tokens=forced, trace_logits=True, steps=steps)
    torch.cuda.synchronize()
    assert torch.equal(toks, toks_old)
    diff = (trace[:steps] - trace_old[:steps]).abs().max().item()
    assert diff < 5e-4, diff
    # free-running greedy decoding picks the same tokens on both paths
    g_new = GPTSampler(eng, B).sample(bev, batch, greedy=True, steps=40)
    g_old_s = GPTSampler(eng, B)
    g_old_s.persistent = False
    g_old = g_old_s.sample(bev, batch, greedy=True, steps=40)
    torch.cuda.synchronize()
    assert (g_new != g_old).float().mean().item() < 0.01
