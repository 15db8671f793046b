#!/usr/bin/env python
"""Kernel-class shares of ONE C2 SONG from the ncu launch list of one DiT step + one VAE decode
(`profiles/r1_vN_launches.csv`): the step's launches weighted x27 (steps per song), the decode's x1 — next to the
shares bench.py measures itself with CUDA events around every launch of a song (`breakdown_ms_per_song` in its
JSON line).  The ncu per-launch times are cold-cache and serialised, so only the SHARES are comparable.
Usage: python tools/song_shares.py profiles/r1_v6_launches.csv gpurun_out/bench_c2.json [steps]"""
import csv
import json
import sys

VAE = ("res_unit_kernel", "EpiConv", "final_conv_kernel", "first_conv_kernel", "posterior_kernel")


def klass(name):
    if "attention" in name:
        return "attention"
    if "gemm_tc" in name or "res_unit" in name:
        return "gemm"
    if "final_conv" in name or "first_conv" in name:
        return "simt_conv"
    return "elementwise"


def main(path, bench_json, steps=27):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    tot = {}
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        us = float(r["Metric Value"]) / 1e3
        w = 1 if any(v in name for v in VAE) else steps
        tot[klass(name)] = tot.get(klass(name), 0.0) + w * us
    ncu_total = sum(tot.values())
    d = json.loads(open(bench_json).read().strip().splitlines()[-1])
    b = {k: v for k, v in d["breakdown_ms_per_song"].items() if k in ("gemm", "attention", "elementwise", "simt_conv")}
    b_total = sum(b.values())
    print(f"# one C2 song = {steps} DiT steps + 1 VAE decode; ncu launch list weighted accordingly vs bench.py's own")
    print("# per-launch CUDA-event profile of a song (graphs bypassed in both, so launch gaps are not in either)")
    print(f"{'class':12s} {'ncu ms':>9s} {'ncu share':>10s} {'bench ms':>9s} {'bench share':>12s}")
    for k in ("gemm", "attention", "elementwise", "simt_conv"):
        print(f"{k:12s} {tot.get(k, 0) / 1e3:9.2f} {100 * tot.get(k, 0) / ncu_total:9.1f}% {b[k]:9.2f} {100 * b[k] / b_total:11.1f}%")
    print(f"{'total':12s} {ncu_total / 1e3:9.2f} {'':>10s} {b_total:9.2f}   (song, graph + PDL, event-timed: {d['ms_per_step']:.1f} ms)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 27)
