"""Summarise an .ncu-rep (read here, no GPU needed) into the JSON kept under profiles/: python tools/ncu_summary.py rep.ncu-rep out.json"""
import csv, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.max.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = {"Kernel Name": r[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = f"{r[i]} {units[i]}".strip()
    res.append(d)
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1))
