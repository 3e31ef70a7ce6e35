#!/usr/bin/env python
"""Turns an ncu report into the short text summary committed under profiles/ (run where ncu is installed).
usage: profile_summary.py <report.ncu-rep> [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pat and not pat.search(r[ki]):
            continue
        print("kernel:", r[ki])
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = r[i]
                print(f"  {w:95s} {r[i]:>16s} {units[i]}")
        stalls = []
        for i, n in enumerate(hdr):
            if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i]), n[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        tot_s = sum(v for v, _ in stalls)
        if tot_s > 0:
            print("  warp-state samples (pc sampling):", ", ".join(f"{n} {100 * v / tot_s:.1f}%" for v, n in sorted(stalls, reverse=True) if v / tot_s >= 0.02))
        try:
            def tobytes(name):
                i = hdr.index(name)
                v = float(r[i])
                u = units[i].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            tot = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
            i = hdr.index("gpu__time_duration.sum")
            t = float(r[i]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(units[i].lower().replace("second", "s"), 1e-9)
            print(f"  dram bytes per launch (read+write): {tot:.0f}   -> {tot / t / 1e9:.1f} GB/s over {t * 1e6:.1f} us")
        except Exception as e:  # noqa
            print("  (no dram summary:", e, ")")
        print()


if __name__ == "__main__":
    main()
