"""bench.py's output contract (the driver parses ONE JSON line per run).

CPU: the FLOP model behind `dit_step_tensor_util` / `roofline` against an independent count, the workload table
against BASELINE.json's configs.  GPU: a short real run of the b200 arm as a subprocess — every key the contract
names is present and sane, the end-to-end number is a separate measurement with its byte counts, and the launches
were this repo's kernels."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_flop_model_matches_an_independent_count():
    Bc, S, E, L, D, I = 2, 750, 512, 24, 2048, 6144
    M = Bc * S
    gemm = 2 * M * D * (4096 + 2048 + 2048 + 2048 + 2 * I + I)          # qkv, self_o, cross_q, cross_o, gate_up, down
    # attention: sliding layers see at most 2*128+1 keys, full layers all S; cross-attention sees E keys
    self_full, self_win = 4 * S * S * 128 * 16 * Bc, 4 * S * min(S, 257) * 128 * 16 * Bc
    cross = 4 * S * E * 128 * 16 * Bc
    want = L * gemm + (L // 2) * (self_full + self_win) + L * cross
    got = bench.dit_flops(Bc, S, E)
    assert abs(got - want) / want < 0.02, (got, want)   # proj_in / proj_out and the K/V cache are the remainder
    assert 4.4e12 < got < 4.7e12                         # the 4.53 TFLOP per C2 step DESIGN.md quotes


def test_workloads_are_baseline_configs():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        configs = json.load(f)["configs"]
    wl = bench.WORKLOADS
    assert "10 s" in configs[0] and wl["c1"]["seconds"] == 10 and wl["c1"]["steps"] == 8
    assert "60 s" in configs[1] and "27 steps" in configs[1] and (wl["c2"]["seconds"], wl["c2"]["steps"]) == (60, 27)
    assert "240 s" in configs[2] and "60 steps" in configs[2] and (wl["c3"]["seconds"], wl["c3"]["steps"]) == (240, 60)
    assert "120 s" in configs[4] and wl["c5"]["seconds"] == 120 and wl["c5"].get("repaint")
    for w in wl.values():
        assert w["T"] == w["seconds"] * 25               # 25 latent frames per second


@pytest.mark.gpu
def test_b200_arm_prints_the_contract_line():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("CUDA device required")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-extra",
                        "--no-gpu-baseline", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["metric"] == "generated-audio-sec/wall-sec" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak" and d["dtype"] == "bf16"
    assert d["vs_baseline"] is None and "synthetic" in d["data"] and "60 s" in d["config"]["workload"]
    assert d["config"]["outputs_finite"] is True
    assert 100.0 < d["value"] < 2000.0 and abs(d["value"] - 60.0 * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"] + 1e-3
    e = d["e2e"]
    assert e["unit"] == "audio-s/s" and 100.0 < e["value"] < 2000.0 and e["value"] != d["value"]
    assert e["h2d_bytes_per_step"] > 1_000_000 and e["d2h_bytes_per_step"] == 2 * 1500 * 1920 * 4
    assert d["gpu_launches"] >= 2 * 27 * 195               # this repo's kernels, counted by the library
    rf = d["roofline"]
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and 0.0 < rf["frac"] < 1.0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["achieved"] < rf["peak"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(d["clocks"]["reasons"])
