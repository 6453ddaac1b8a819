#!/usr/bin/env python
"""Print the SASS of one kernel of the built library with the CUDA source line of every instruction (no GPU needed):

    python profiles/tools/sass_fn.py bucketize_kernelILi3 [lib.so]  > /tmp/p1.sass

Used to count instructions per pipe in the hot loops before spending GPU time."""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]


def main():
    mangled = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else str(ROOT / "findkmer_b200" / "libfindkmer_b200.so")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
        for cubin in Path(d).glob("*.cubin"):
            out = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
            if mangled not in out:
                continue
            infn, cur = False, None
            for ln in out.split("\n"):
                if ln.startswith("\t.section") or ln.startswith("//-----"):
                    infn = mangled in ln if ".text." in ln else False
                    continue
                if not infn:
                    continue
                m = re.search(r'//## File "(.*?)", line (\d+)', ln)
                if m:
                    cur = f"{Path(m.group(1)).name}:{m.group(2)}"
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m:
                    print(f"{m.group(1)} {cur or '-':24s} {m.group(2)}")
                elif re.match(r"^\.L_x_\d+:", ln):
                    print(ln)
            return


if __name__ == "__main__":
    main()
