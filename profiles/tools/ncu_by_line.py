#!/usr/bin/env python
"""Attribute the instructions one kernel executed (ncu --set full --import-source on, built with -lineinfo)
to CUDA source lines, with no GPU:

    python profiles/tools/ncu_by_line.py gpurun_out/prof.ncu-rep bucketize_kernelILi3 findkmer_b200/csrc/fkb_bucket.cu [top_n]

It joins `ncu --page source --csv` (per-SASS-instruction counters, by address) with `nvdisasm -g` of the cubin
inside findkmer_b200/libfindkmer_b200.so (SASS offset -> source line)."""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]


def main():
    rep, mangled, src = sys.argv[1], sys.argv[2], Path(sys.argv[3])
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    kern_regex = re.sub(r"ILi\d+.*", "", mangled)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "findkmer_b200" / "libfindkmer_b200.so")], cwd=d, check=True,
                       stdout=subprocess.DEVNULL)
        sass = ""
        for cubin in Path(d).glob("*.cubin"):
            out = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
            if mangled in out:
                sass = out
                break
    off2line, infn, cur = {}, False, None
    for ln in sass.split("\n"):
        if ".text." in ln and (ln.startswith(".text.") or ".section" in ln):
            infn = mangled in ln
        if not infn:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:  # code inlined from a header keeps its own (file, line)
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
    csv_text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern_regex],
                              capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csv_text)))
    h = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r][0]
    hdr = rows[h]
    ei, ai = hdr.index("Instructions Executed"), hdr.index("Address")
    si = hdr.index("# Samples") if "# Samples" in hdr else None
    seen, data = set(), []
    for r in rows[h + 1:]:
        try:
            a, n = int(r[ai], 16), float(r[ei])
        except (ValueError, IndexError):
            continue
        if a in seen:
            continue
        seen.add(a)
        data.append((a, n, float(r[si]) if si is not None and r[si] else 0.0))
    base = min(a for a, _, _ in data)
    by_line, samples, tot, stot = collections.Counter(), collections.Counter(), 0.0, 0.0
    for a, n, smp in data:
        line = off2line.get(a - base)
        by_line[line] += n
        samples[line] += smp
        tot += n
        stot += smp
    texts = {}

    def source_line(key):
        if not key:
            return "(no line info)"
        name, line = key
        if name not in texts:
            cands = [ROOT / src] + list((ROOT / "findkmer_b200" / "csrc").glob(name))
            hit = [c for c in cands if c.name == name and c.exists()]
            texts[name] = hit[0].read_text().split("\n") if hit else []
        t = texts[name]
        return (t[line - 1].strip()[:100] if 0 < line <= len(t) else "?")
    print(f"kernel {mangled}: {tot:.0f} warp-level instructions executed, {stot:.0f} stall samples")
    print(f"{'instr%':>7} {'stall%':>7}  file:line  source")
    for line, n in by_line.most_common(top):
        where = f"{line[0]}:{line[1]}" if line else "-"
        print(f"{n / tot * 100:7.2f} {samples[line] / max(stot, 1) * 100:7.2f}  {where:>22}  {source_line(line)}")


if __name__ == "__main__":
    main()
