"""Experiment: split the batch over k concurrent CUDA streams (one GPTSampler + CUDA graph each) to overlap the latency-bound
weight GEMMs of one half with the bandwidth-bound cache attention of the other."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tools.stage2_perf import KW, sizes
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_decode import GPTSampler
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
B = 16
cfg = GPTConfig(**KW)
eng = GPTEngine(synth.gpt_state_dict(sizes(cfg), seed=2), cfg, device="cuda:0", precision=prec)
_, bev, batch = synth.stage2_inputs(B, seed=0)
bev = bev.cuda(); batch = {k: v.cuda() for k, v in batch.items()}

for nsplit in (1, 2, 4):
    bs = B // nsplit
    samplers = [GPTSampler(eng, bs) for _ in range(nsplit)]
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    parts = [(bev[i * bs:(i + 1) * bs].contiguous(), {k: v[i * bs:(i + 1) * bs].contiguous() for k, v in batch.items()}) for i in range(nsplit)]
    def run(steps=None):
        outs = [None] * nsplit
        torch.cuda.synchronize()
        # interleave graph replays across streams: issue step-by-step so both streams progress together
        for i, (s, st) in enumerate(zip(samplers, streams)):
            with torch.cuda.stream(st):
                outs[i] = s.sample(parts[i][0], parts[i][1], top_k=100, seed=1 + i, steps=steps)
        torch.cuda.synchronize()
        return outs
    run(steps=40)
    t0 = time.time()
    outs = run()
    dt = time.time() - t0
    print(f"{prec} nsplit={nsplit}: {dt:.3f} s for {B} scenes -> {B * 6 / dt:.1f} img/s", flush=True)
    del samplers
    torch.cuda.empty_cache()
