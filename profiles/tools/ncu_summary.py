#!/usr/bin/env python
"""Condense an `ncu --set full --import-source on` report into the text summary committed under profiles/ (no GPU needed):

    python profiles/tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_full_summary.txt

Per kernel: duration, pipe utilisation, issue rate, DRAM bytes, warp-stall mix, and the opcode mix with the shared-memory
wavefronts per instruction (the bank-conflict multiplier that bounds both count kernels)."""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = [
    ("duration (us)", "gpu__time_duration.sum"),
    ("SM cycles elapsed (max)", "sm__cycles_elapsed.max"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("LSU instr pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("L1/shared data pipe (LSU wavefronts) % of peak", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("shared-memory wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("achieved occupancy % (warps active)", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("DRAM read (ncu unit)", "dram__bytes_read.sum"),
    ("DRAM write (ncu unit)", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("registers / thread", "launch__registers_per_thread"),
    ("dynamic smem / block (ncu unit)", "launch__shared_mem_per_block_dynamic"),
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    ki = hdr.index("Kernel Name")
    for r in rows:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        name = re.sub(r"^.*::", "", name)
        print("=" * 100)
        print(name)
        for label, key in RAW:
            if key in hdr:
                i = hdr.index(key)
                print(f"  {label:52s} {r[i]:>16s} {units[i]}")
        stalls = {h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(r[i]) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:8]
        print("  warps stalled per issue slot (top): " + ", ".join(f"{k} {v:.2f}" for k, v in top))
        src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + re.match(r"\w+", name).group(0)]))))
        h = [i for i, x in enumerate(src) if "Source" in x and "Instructions Executed" in x]
        if not h:
            continue
        sh = src[h[0]]
        ix = {n: i for i, n in enumerate(sh)}
        byop, wav, ideal, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0.0
        for x in src[h[0] + 1:(h[1] - 1 if len(h) > 1 else None)]:
            try:
                n = float(x[ix["Instructions Executed"]])
            except (ValueError, IndexError):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x[ix["Source"]])
            op = m.group(2) if m else "?"
            base = op.split(".")[0]
            if base in ("LDS", "STS", "ATOMS", "LDG", "STG", "RED", "ATOM", "ATOMG", "SHFL"):
                base = ".".join(op.split(".")[:3])
            byop[base] += n
            tot += n
            wav[base] += float(x[ix["L1 Wavefronts Shared"]] or 0)
            ideal[base] += float(x[ix["L1 Wavefronts Shared Ideal"]] or 0)
        print(f"  opcode mix (warp instructions, {tot:.0f} total):")
        for op, n in byop.most_common(18):
            extra = f"   shared wavefronts {wav[op]:.0f} = {wav[op] / n:.2f} per instruction (ideal {ideal[op] / n:.2f})" if wav[op] else ""
            print(f"    {op:22s} {n:14.0f} {100 * n / tot:5.1f} %{extra}")


if __name__ == "__main__":
    main()
