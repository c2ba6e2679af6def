"""Stage-2 forward with block-sparse layouts of decreasing density (scripts/inference.py:170-175 is the reference's only
density sweep): python tools/layout_perf.py"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tools.stage2_perf import KW, sizes
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth

B = 16
cfg = GPTConfig(**KW)
sd = synth.gpt_state_dict(sizes(cfg), seed=2)
cam, bev, batch = synth.stage2_inputs(B, seed=0)
cam, bev = cam.cuda(), bev.cuda()
batch = {k: v.cuda() for k, v in batch.items()}
nb = cfg.gpt_block_size // cfg.sparse_block_size
out = []
for keep in (None, 0.5, 0.25, 0.1):
    layouts = None
    if keep is not None:      # random block layouts at 128-position granularity (whole key tiles drop out), diagonal + first key block kept
        g = torch.Generator().manual_seed(1)
        coarse = torch.rand(cfg.num_layers, cfg.num_heads, nb * cfg.sparse_block_size // 128, nb * cfg.sparse_block_size // 128, generator=g) < keep
        rep = 128 // cfg.sparse_block_size
        lay = coarse.repeat_interleave(rep, 2).repeat_interleave(rep, 3)
        lay |= torch.eye(nb, dtype=torch.bool)[None, None]
        lay[..., 0] = True
        layouts = lay.long()
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="f16f8", layouts=layouts)
    for _ in range(2):
        eng.forward(cam, bev, batch, sampling=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.forward(cam, bev, batch, sampling=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    dens = 1.0 if layouts is None else float((eng.layouts[0].float().mean()).item())
    out.append({"keep": keep, "layout_density": dens, "forward_ms": ms, "samples_per_s": B / ms * 1e3})
    print(out[-1], flush=True)
    del eng
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "layout_perf.json", "w"), indent=1)
