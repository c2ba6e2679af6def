"""Which decode steps deviate from the teacher-forced forward (small configs): persistent kernel and per-launch chain."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_decode import GPTSampler
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth
from tests.cases import GPT_CASES, GPT_VARIANTS, gpt_sizes, gpt_variant_inputs

for name in sys.argv[1:] or ["small", "small_nusc14x25"]:
    if name in GPT_CASES:
        kw, B = GPT_CASES[name]
        cfg = GPTConfig(**kw)
        sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
        cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=4)
    else:
        cfg, sd, cam, bev, batch = gpt_variant_inputs(name, synth, GPTConfig)
        B = cam.shape[0]
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3")
    full = eng.forward(cam.cuda(), bev.cuda(), batch, sampling=True)
    want = full[:, cfg.forward_shuffle_idx.cuda()].permute(1, 0, 2)
    forced = cam.reshape(B, -1)[:, cfg.forward_shuffle_idx]
    for pers in (True, False):
        smp = GPTSampler(eng, B)
        smp.persistent = pers
        _, trace = smp.sample(bev, batch, forced_tokens=forced, trace_logits=True)
        torch.cuda.synchronize()
        per = (trace - want).abs().amax(dim=(1, 2))
        bad = torch.nonzero(per > 1e-3).flatten().tolist()
        print(f"[{name}] persistent={pers}: max err vs forward {per.max().item():.2e}; bad steps ({len(bad)}): {bad[:40]}; errs {[round(per[i].item(), 4) for i in bad[:12]]}", flush=True)
