"""Stage-2 measurements on one B200 (BASELINE configs[2] teacher-forced forward, configs[3] KV-cache sampling), full-size model."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from bevgen_b200 import ops  # noqa: E402
from bevgen_b200.gpt_config import GPTConfig  # noqa: E402
from bevgen_b200.gpt_decode import GPTSampler  # noqa: E402
from bevgen_b200.gpt_engine import GPTEngine  # noqa: E402
from oracle import synth  # noqa: E402

KW = dict(embd_pdrop=0., resid_pdrop=0., attn_pdrop=0., n_unmasked=0, num_cams=6, vocab_size=1024, cond_vocab_size=1024, hidden_size=1024,
          num_embed=1024, num_heads=16, num_layers=24, backend="deepspeed", sparse_block_size=16, window_len=32, cam_res=(256, 256),
          cam_latent_res=(16, 16), plot=False, causal_order=True, camera_bias=True, image_embed=True, bev_embed=True, bev_latent_res=(16, 16),
          density=1.0, cam_names="NUSCENES_CAMERAS", dataset="NUSCENES")


def sizes(cfg):
    return dict(num_embed=cfg.num_embed, gpt_block_size=cfg.gpt_block_size, num_img_tokens=cfg.num_img_tokens, num_cond_tokens=cfg.num_cond_tokens,
                num_cams=cfg.num_cams, vocab_size=cfg.vocab_size, cond_vocab_size=cfg.cond_vocab_size, num_layers=cfg.num_layers)


def timed(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run(precision, B=16, layers=24, fwd_iters=3, sample=True, kv_dtype=None, use_pdl=True):
    cfg = GPTConfig(**{**KW, "num_layers": layers})
    t0 = time.time()
    sd = synth.gpt_state_dict(sizes(cfg), seed=2)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision=precision)
    del sd
    cam, bev, batch = synth.stage2_inputs(B, seed=0)
    cam, bev = cam.cuda(), bev.cuda()
    batch = {k: v.cuda() for k, v in batch.items()}
    res = {"precision": precision, "B": B, "layers": layers, "build_s": time.time() - t0}
    eng.forward(cam, bev, batch, sampling=True)
    ms = timed(lambda: eng.forward(cam, bev, batch, sampling=True), fwd_iters)
    fl = eng.flops_per_sample() * B
    res["forward"] = {"ms": ms, "samples_per_s": B / ms * 1e3, "algorithmic_tflops": fl / ms / 1e9, "tflop_per_sample": fl / B / 1e12}
    # attention share: time the composed attention alone on one layer
    d = cfg.num_embed
    qkv = (torch.randn(B, cfg.gpt_block_size, 3 * d, device="cuda").bfloat16(), torch.randn(B, cfg.gpt_block_size, 3 * d, device="cuda").bfloat16() * 1e-2)
    if eng.npass == 1:
        qkv = (qkv[0], None)
    y = torch.randn(B, cfg.gpt_block_size, d, device="cuda")
    eng.attention(qkv, y, B, cfg.gpt_block_size)
    ams = timed(lambda: eng.attention(qkv, y, B, cfg.gpt_block_size), 3)
    afl = 4.0 * B * d * eng._allowed
    res["attention_layer"] = {"ms": ams, "allowed_tflops": afl / ams / 1e9, "dense_equiv_tflops": 4.0 * B * cfg.gpt_block_size ** 2 * d / ams / 1e9}
    if sample:
        sampler = GPTSampler(eng, B, kv_dtype=kv_dtype)
        sampler.use_pdl = use_pdl
        sampler.sample(bev, batch, temperature=1.0, top_k=100, seed=1, steps=64)         # warm-up + graph capture
        ops.Stats.reset()
        torch.cuda.synchronize()
        t0 = time.time()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        toks = sampler.sample(bev, batch, temperature=1.0, top_k=100, seed=1)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        byt = sampler.bytes_per_batch()
        # full generation (configs[3]): tokens -> codebook gather -> VQGAN decoder -> denormalise, 6 images per scene
        from bevgen_b200.vqgan_engine import VQGANEngine
        dd = synth.vqgan_ddconfig()
        vq = VQGANEngine(synth.vqgan_state_dict(dd, seed=1), dd, device="cuda:0", precision="bf16" if precision == "bf16" else "f16f8")
        idx = toks.reshape(-1).contiguous()
        dec = lambda: vq.decode_indices(idx, B * 6, 16, 16)
        dec()
        dms = timed(dec, 2)
        res["generate"] = {"sample_ms": ms, "vqgan_decode_ms": dms, "images_per_s": B * 6 / (ms + dms) * 1e3,
                           "note": "KV-cache sampling of B scenes + VQGAN decode of their 6 x 256x256 images (the generate.py hot path)"}
        res["sample"] = {"kv_dtype": str(kv_dtype or "default"), "pdl": use_pdl, "ms": ms, "wall_s": time.time() - t0, "images_per_s": B * 6 / ms * 1e3, "ms_per_token_step": ms / 1536,
                         "algorithmic_GB": byt / 1e9, "achieved_GBps": byt / ms / 1e6, "tokens_ok": bool(int(toks.max()) < 1024)}
    return res


if __name__ == "__main__":
    out = []
    for prec in sys.argv[1:] or ["f16f8", "fp32x3", "bf16"]:
        kv = None
        if prec.endswith("+kv16"):
            prec, kv = prec[:-5], torch.float16
        pdl = True
        if prec.endswith("-pdl"):
            prec, pdl = prec[:-4], False
        r = run(prec, kv_dtype=kv, use_pdl=pdl, fwd_iters=1 if (kv or not pdl) else 3)
        print(json.dumps(r), flush=True)
        out.append(r)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "stage2_perf.json").write_text(json.dumps(out, indent=1))
