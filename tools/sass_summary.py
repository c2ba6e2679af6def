"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA paths (B200_PROFILING.md): python tools/sass_summary.py > profiles/rNN_sass_summary.txt
Reads bevgen_b200/libbevgen_b200.so with cuobjdump (no GPU needed)."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "bevgen_b200" / "libbevgen_b200.so"
PATTERNS = OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTCQMMA", r"\bUTCQMMA"), ("2CTA MMA", r"\bUTC[HQ]MMA\.2CTA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"),
                        ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA (mma.sync)", r"\bHMMA"), ("LDL/STL", r"\b(LDL|STL)")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            kernels[cur] = Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        kernels[cur]["instructions"] += 1 if re.search(r"^\s+/\*[0-9a-f]{4,}\*/", line) else 0
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                kernels[cur][name] += 1
    cols = ["instructions"] + list(PATTERNS)
    print(f"SASS mnemonic counts per kernel of {LIB.name} (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':70s} " + " ".join(f"{c:>10s}" for c in cols))
    tot = Counter()
    for k, c in kernels.items():
        print(f"{k[:70]:70s} " + " ".join(f"{c[n]:10d}" for n in cols))
        tot.update(c)
    print(f"{'TOTAL':70s} " + " ".join(f"{tot[n]:10d}" for n in cols))


if __name__ == "__main__":
    sys.exit(main())
