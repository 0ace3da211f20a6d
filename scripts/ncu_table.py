"""ncu report of scripts/kernel_zoo.py (`ncu --set full --nvtx --nvtx-include "zoo/" -o X`) -> one row per kernel variant:
duration, tensor-pipe utilisation, issue / XU utilisation, DRAM bytes, achieved algorithmic TFLOP/s and GB/s against
MEASURED_PEAKS.json.  Runs where the report was copied to (needs `ncu` for reading only, no GPU).

  python scripts/ncu_table.py gpurun_out/r02_zoo.ncu-rep profiles/r02_ncu_kernel_table.csv
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
           "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def algorithmic(name):
    """(FLOPs, bytes) of one launch: 2*MAC per conv / linear, 4*N*L*d per attention head; every operand / result once."""
    m = re.search(r"self_attn N=(\d+) B=(\d+) h=(\d+)", name)
    if m:
        N, B, h = map(int, m.groups())
        return 4.0 * B * h * N * N * 64, 4 * B * N * h * 64 * 2
    m = re.search(r"(?:cross|ip)_attn N=(\d+) L=(\d+) B=(\d+) h=(\d+)", name)
    if m:
        N, Lk, B, h = map(int, m.groups())
        return 4.0 * B * h * N * Lk * 64, (2 * B * N * h * 64 + 2 * B * Lk * h * 64) * 2
    m = re.search(r"(\d+)x(\d+)x(\d+)", name)
    if m and re.search(r"ff|qkv|out-proj", name):
        M, N, K = map(int, m.groups())
        nout = N // 2 if "GEGLU" in name else N
        return 2.0 * M * N * K, (M * K + N * K + M * nout + (M * nout if "residual" in name else 0)) * 2
    m = re.search(r"conv3x3 (\d+)->(\d+) @(\d+)x\d+ B=(\d+)", name)
    if m:
        ci, co, H, B = map(int, m.groups())
        return 2.0 * B * H * H * co * 9 * ci, (B * H * H * ci + 9 * ci * co + B * H * H * co) * 2
    m = re.search(r"GroupNorm\+SiLU (\d+)ch @(\d+)x\d+ B=(\d+)", name)
    if m:
        c, H, B = map(int, m.groups())
        return 0.0, 2 * B * H * H * c * 2
    if "fuse_step" in name:
        HW = 128 * 128
        return 0.0, (4 + 4) * HW * 16 + 2 * HW * 4 + 2 * HW * 32 + HW * 32 + 6 * HW * 16
    return 0.0, 0.0


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    rng_col = next(i for i, h in enumerate(hdr) if "Push/Pop" in h)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    with open(out, "w", newline="") as f:
        f.write("# one launch per kernel variant at its SDXL shape (scripts/kernel_zoo.py), ncu --set full --clock-control none; "
                f"peaks from MEASURED_PEAKS.json: {peaks['bf16_tflops_sustained']} TFLOP/s sustained bf16, {peaks['hbm_gbs']} GB/s HBM copy; "
                "algorithmic work = 2*MAC / 4*N*L*d, operands and results once\n")
        w = csv.writer(f)
        w.writerow(["case", "kernel", "us", "tensor_pipe_pct_active", "tensor_pipe_pct_elapsed", "issue_slots_pct", "xu_pipe_pct",
                    "lts_pct", "dram_MB", "algorithmic_TFLOPs", "algorithmic_GBs", "bound", "frac_of_measured_peak", "grid",
                    "regs", "sm_ghz"])
        for r in rows[2:]:
            names = re.findall(r"<default domain>:([^:]+):none", r[rng_col])
            case = names[-1] if names else ""
            g = lambda m: float(r[ix[m]].replace(",", ""))  # noqa: E731
            us = g("gpu__time_duration.sum")
            fl, by = algorithmic(case)
            tf, gbs = (fl / us / 1e6 if fl else 0.0), by / us / 1e3
            tensor_bound = fl > 0 and ("self_attn" in case or "attn" not in case)
            frac = tf / peaks["bf16_tflops_sustained"] if tensor_bound else gbs / peaks["hbm_gbs"]
            w.writerow([case, r[ix["Kernel Name"]].split("(")[0], f"{us:.1f}",
                        f"{g(METRICS[1]):.1f}", f"{g(METRICS[2]):.1f}", f"{g(METRICS[5]):.1f}", f"{g(METRICS[6]):.1f}",
                        f"{g(METRICS[7]):.1f}", f"{g(METRICS[3]) + g(METRICS[4]):.1f}", f"{tf:.1f}", f"{gbs:.1f}",
                        "tensor" if tensor_bound else "hbm", f"{frac:.3f}", r[ix["launch__grid_size"]],
                        r[ix["launch__registers_per_thread"]], f"{g(METRICS[8]):.3f}"])
    print(open(out).read())


if __name__ == "__main__":
    main()
