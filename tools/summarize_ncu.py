#!/usr/bin/env python
"""One line per captured launch of `ncu --set full` reports: duration, DRAM bytes and bandwidth, L2 and
tensor-pipe percentages, registers.  Usage: python tools/summarize_ncu.py a.ncu-rep [b.ncu-rep ...]

umma_pct is the tcgen05 tensor-pipe utilisation: sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off
.avg.pct_of_peak_sustained_elapsed = bf16 UTCHMMA operations executed / (8192 per SM per elapsed cycle).  (The generic
sm__pipe_tensor_cycles_active_realtime counter, tensor_rt_pct, does not count the tcgen05 path properly: it reads
17 % for a 1300 TFLOP/s launch.)  umma_tflops = the same counter's op count / duration: executed FLOPs including the
zero rows of partial M tiles, so it is >= the algorithmic figure."""
import csv
import re
import subprocess
import sys

KEYS = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_rt_pct": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "umma_pct": "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "umma_ops": "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum",
    "xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
}
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = name.replace("<unnamed>::", "").replace("ace::", "").replace("void ", "")
    m = re.match(r"([\w:]+)(<[^(]*>)?\(", name)
    if not m:
        return name[:50]
    targs = re.sub(r"\(int\)", "", m.group(2) or "")
    return (m.group(1) + targs)[:50]


for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(v) for k, v in KEYS.items() if v in hdr}
    print(f"# {rep}")
    for r in rows[2:]:
        g = lambda k: r[col[k]] if k in col else "?"
        dur_us = float(g("dur")) * SCALE.get(units[col["dur"]], 1.0)
        by = (float(g("rd")) * SCALE.get(units[col["rd"]], 1.0) + float(g("wr")) * SCALE.get(units[col["wr"]], 1.0))
        print(f"{short(r[hdr.index('Kernel Name')]):52s} grid={g('grid'):>6s} dur_us={dur_us:9.1f} dram_MB={by / 1e6:9.1f} "
              f"dram_GBs={by / dur_us / 1e3:7.1f} dram_pct={g('dram_pct')[:5]:>5s} lts_pct={g('lts_pct')[:5]:>5s} "
              f"umma_pct={g('umma_pct')[:5]:>5s} umma_tflops={(float(g('umma_ops')) / dur_us / 1e6) if g('umma_ops') != '?' else 0:7.1f} "
              f"tensor_rt_pct={g('tensor_rt_pct')[:5]:>5s} xu_pct={g('xu_pct')[:5]:>5s} issue_pct={g('issue_pct')[:5]:>5s} regs={g('regs')}")
