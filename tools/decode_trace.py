"""Needs a build with the probes: VARIANT_FLAGS=-DBEVGEN_DP_TRACE tools/build_variant.sh trace bevgen_b200/csrc/decode_persistent.cu, then
BEVGEN_B200_LIB=$PWD/bevgen_b200/variants/lib_trace.so python tools/decode_trace.py.
Clock trace of thread 0 of one CTA through one layer of the persistent decode kernel (BEVGEN_DP_DBG=64 [+ other flags]):
python tools/decode_trace.py [steps]   env: BEVGEN_DP_TRACE_CTA, BEVGEN_DP_TRACE_STEP.  Prints probe id, cycles since the previous probe."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["BEVGEN_DP_DBG"] = str(int(os.environ.get("BEVGEN_DP_DBG", "0")) | 64)
import torch
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_decode import GPTSampler
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth
from tests.cases import GPT_FULL, gpt_sizes

NAMES = {1: "layer start", 2: "qkv done", 3: "phase boundary done", 4: "attention done", 5: "phase boundary done", 6: "mlp1 done", 7: "phase boundary done", 8: "mlp2 (partials + finalise) done", 9: "phase boundary done",
         10: "lin: enter", 11: "lin: part_range", 12: "lin: act TMA issued", 13: "lin: stats + fragments loaded (tags valid)", 14: "lin: act landed", 15: "lin: fragments loaded", 16: "lin: barrier A",
         17: "lin: unit landed", 18: "lin: unit mma + sts", 19: "lin: stats finished", 20: "lin: barrier B", 21: "lin: epilogue",
         30: "mlp2: enter", 31: "mlp2: act quarter landed", 32: "mlp2: frags + barrier", 33: "mlp2: unit mma", 34: "mlp2: reduced + finalised", 35: "mlp2: barrier",
         50: "att: enter", 51: "att: prologue loads/stores", 52: "att: bias row landed", 53: "att: barrier", 54: "att: loop head", 55: "att: prev released", 56: "att: block landed",
         57: "att: block math", 58: "att: released", 59: "att: loop done", 60: "att: end barrier", 61: "att: merge + finish", 62: "att: proxy fence",
         90: "sync: enter", 91: "gs: barrier 1", 92: "gs: red.release", 93: "gs: poll done", 94: "sync: CTA barrier done"}


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    cfg = GPTConfig(**GPT_FULL)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="f16f8")
    _, bev, batch = synth.stage2_inputs(16, seed=0)
    bd = {k: v.cuda() for k, v in batch.items()}
    smp = GPTSampler(eng, 16)
    smp.profile_phases = True
    smp.sample(bev, bd, temperature=1.0, top_k=100, seed=3, steps=steps)
    torch.cuda.synchronize()
    tr = smp._pk["prof"][-32:].reshape(-1).cpu()
    n = int(tr[1023])
    ev = [(int(v) >> 48, int(v) & ((1 << 48) - 1)) for v in tr[:n].tolist()]
    t0 = ev[0][1]
    for i, (pid, c) in enumerate(ev):
        print(f"{c - t0:8d} (+{c - ev[i - 1][1] if i else 0:6d})  {pid:3d} {NAMES.get(pid, '')}")


if __name__ == "__main__":
    main()
