#!/usr/bin/env python
"""DRAM traffic of ONE step of bench.py from an ncu metrics pass (no GPU needed to run this):

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
    python profiles/tools/traffic_summary.py gpurun_out/traffic.csv > profiles/rNN_traffic_k11.json

Takes the kernels of the LAST step (from the last bucketize_kernel launch, which starts a step, to the end of the capture) and prints the JSON bench.py reads for `roofline.traffic`."""
import csv
import json
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[0]), {"kernel": r[4]})
        d[r[12]] = float(r[14].replace(",", ""))
        d["unit_" + r[12]] = r[13]
    ids = sorted(launches)
    names = [launches[i]["kernel"] for i in ids]
    # a step = [bucketize, count_buckets, count_direct_kernel (the edge slivers, side stream), (fold levels), finalize levels]: take the last one
    # (the edge slivers' count_direct_kernel runs on a side stream and may be listed after the bucketed kernels)
    starts = [i for i, n in zip(ids, names) if "bucketize" in n] or [i for i, n in zip(ids, names) if "count_direct_kernel" in n]
    if not starts:
        sys.exit("no count kernel launch in the capture")
    step = [i for i in ids if i >= starts[-1] and ("fkb" in launches[i]["kernel"] or "unnamed" in launches[i]["kernel"])]
    kernels, total = [], 0
    for i in step:
        d = launches[i]
        name = re.sub(r"\(.*", "", d["kernel"].replace("(anonymous namespace)", "").replace("<unnamed>", ""))
        name = name.replace("unnamed>::", "").replace("void ", "").replace("fkb::", "").replace("::", "").strip()
        us = d.get("gpu__time_duration.sum", 0.0)
        unit = d.get("unit_gpu__time_duration.sum", "ns")
        us = us / 1000.0 if unit in ("ns", "nsecond") else (us * 1000.0 if unit in ("ms", "msecond") else us)
        rd, wr = int(d.get("dram__bytes_read.sum", 0)), int(d.get("dram__bytes_write.sum", 0))
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            u = d.get("unit_" + key, "byte")
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            if key.endswith("read.sum"):
                rd = int(d.get(key, 0) * scale)
            else:
                wr = int(d.get(key, 0) * scale)
        total += rd + wr
        kernels.append({"kernel": name, "us": round(us, 1), "dram_read_bytes": rd, "dram_write_bytes": wr})
    print(json.dumps({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none; "
                                "python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline (last step); B200",
                      "dram_bytes_per_step": total, "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
