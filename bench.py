#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Gbases/s counted, k = 11 on a 3.1 Gbp synthetic genome, N B200s.

    python bench.py [--gpus N --steps K --warmup W]            # N > 1: launched by torch.distributed.run
    python bench.py --impl reference [...]                     # the reference's own CPU implementation

One "step" = one pass of the counting hot path over the whole workload (config 4 of BASELINE.json:
3.1e9 bases, 24 records, 60-column FASTA, k = 11):
    value : stripped stream already resident in HBM -> zero accumulators, count kernel(s), [NCCL reduce],
            finalize (N, base counts, node count); timed with CUDA events, max over ranks.
    e2e   : the same job through the host-buffer C-ABI call -- raw FASTA bytes in PINNED HOST memory ->
            H2D -> device strip -> kernels -> [reduce] -> finalize -> table + counts back on the host.
Strong scaling: the 3.1 Gbp are split into N contiguous shards (16-byte halo), one per GPU; per-GPU tables
are summed with one NCCL reduce (plus a max-reduce of the prefix flags and a sum of 12 scalars).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Gbases/s counted, k=11 on 3.1 Gbp"
UNIT = "Gbases/s"
N_BASES = 3_100_000_000
K = 11


# --------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# the reference arm: the reference's own CPU implementation on a bounded sample of the workload
# --------------------------------------------------------------------------------------------------
def reference_sample_run(sample_bases: int, k: int, workdir: str):
    """Run the untouched reference binary (oracle/_ref/findKmer) on a `sample_bases` file of the workload's
    shape; returns (count-phase seconds, total seconds, kind).  Falls back to the C port of the oracle only when
    the binary is absent."""
    from findkmer_b200 import synth
    from oracle import harness
    lay = synth.config4(n_bases=sample_bases)
    data = synth.render(lay)
    if harness.reference_available():
        # -z with a huge threshold: the count phase is identical, the CSV stays header-only (the launcher's own
        # runs use -z 1000 on the full genome, k6thru11fullANDupstream.sh:21-24)
        r = harness.run_reference(data, k, z=1000000, probe=False, time_phases=True, timeout=1800)
        if not r.ok or r.count_seconds is None:
            raise RuntimeError("reference run failed: " + r.stdout[-400:])
        return r.count_seconds, r.total_seconds, "reference"
    t0 = time.perf_counter()
    o = harness.oracle_count_fasta(data, k)
    dt = time.perf_counter() - t0
    assert o.rc == 0
    return dt, dt, "port"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = args.ref_sample_bases
    times = []
    kind = "reference"
    with tempfile.TemporaryDirectory() as d:
        for i in range(args.warmup + args.steps):
            cs, ts, kind = reference_sample_run(sample, args.k, d)
            if i >= args.warmup:
                times.append(cs)
    mean_s = sum(times) / len(times)
    value = sample / mean_s / 1e9
    cores = 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, extra={"sample": f"first-principles sample: {sample} bases of the same shape per step"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} bases (config-4 shape, 24 records, 60-col), k={args.k}, count phase only, "
                                   f"single-threaded reference on {os.cpu_count()} host cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, extra=None):
    cfg = {"workload": f"config 4: synthetic {args.bases / 1e9:.3g} Gbp genome, 24 records, 60-column FASTA, k={args.k} (4^{args.k}-entry uint32 table)",
           "k": args.k, "n_bases": args.bases, "records": 24, "line_width": 60, "seed": 20144,
           "l2": "inputs exceed L2 (3.1 GB stream per pass vs 126 MB L2); no flush needed",
           "sharding": f"{args.gpus} contiguous shard(s), 16-byte left halo, NCCL sum-reduce of per-GPU tables" if args.gpus > 1 else "1 shard"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bases", type=int, default=N_BASES, help="workload size (default: the metric's 3.1e9)")
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-bases", type=int, default=20_000_000, help="bounded sample for the cpu_baseline leg")
    ap.add_argument("--ref-sample-bases", type=int, default=4_000_000, help="per-step sample of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-5 (k = 8, dirty input) measurement")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from findkmer_b200 import synth
    from findkmer_b200.engine import KmerCounter
    from findkmer_b200.sharded import reduce_accumulators, shard_cuts
    from findkmer_b200._lib import FkbCounts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torch.distributed.run with {args.gpus} ranks", file=sys.stderr)
            return 2
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    counter = KmerCounter(local_rank)
    k = args.k
    raw_lay = synth.config4(n_bases=args.bases)
    str_lay = raw_lay.stripped()
    stream_total = str_lay.total_bytes
    raw_total = raw_lay.total_bytes

    # ---- device-resident shard of the stripped stream (value leg) ----
    cuts = shard_cuts(stream_total, world)
    b, e = cuts[rank], cuts[rank + 1]
    halo = 16 if rank > 0 else 0
    d_stream = counter.synth_fasta_device(str_lay, b - halo, e - b + halo)
    acc = counter.new_accumulators(k)
    torch.cuda.synchronize()
    lib, ctx = counter._lib, counter._ctx
    st = torch.cuda.current_stream(dev)
    d_counts = torch.zeros(12, dtype=torch.int64, device=dev)
    import ctypes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_acc():
        reduce_accumulators(acc.table, acc.flags, acc.partials, dst=0, fused=acc.buf)

    ev_pairs = []

    def value_step(record: bool):
        counter._check(lib.fkb_zero_device(ctx, k, acc.table.data_ptr(), acc.flags.data_ptr(), acc.partials.data_ptr(), st.cuda_stream))
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
        counter.count_stream_device(d_stream, k, acc, halo, d_stream.numel())
        if record:
            e1.record(st)
            ev_pairs.append((e0, e1))
        reduce_acc()
        if rank == 0:
            counter._check(lib.fkb_finalize_device(ctx, k, acc.table.data_ptr(), acc.flags.data_ptr(), acc.partials.data_ptr(),
                                                   stream_total, d_counts.data_ptr(), st.cuda_stream))

    for _ in range(args.warmup):
        value_step(False)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = counter.launches
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record(st)
    for _ in range(args.steps):
        value_step(True)
    t_end.record(st)
    barrier()
    launches = counter.launches - launches0
    clock_info = clocks.stop() if rank == 0 else None
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = sum(a.elapsed_time(b_) for a, b_ in ev_pairs) / len(ev_pairs)
    if world > 1:
        t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kernel_ms = float(t[0]), float(t[1])
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])
    ms_per_step = total_ms / args.steps
    value = args.bases / (ms_per_step * 1e-3) / 1e9

    # live per-kernel durations of the count path (CUDA events recorded by the library around its kernels, on the stream it
    # launches on), from a few extra steps outside the timed region: the dominant kernel's own roofline figure, and the shares
    # that the committed ncu launch list must agree with
    phase_ms = None
    if lib.fkb_set_option(ctx, b"phase_events", 1) == 0:
        acc_ms = [0.0, 0.0, 0.0]
        n_extra = 5
        for _ in range(n_extra):
            value_step(False)
            ms3 = (ctypes.c_double * 3)()
            if lib.fkb_phase_times(ctx, ctypes.byref(ms3)) == 0:
                acc_ms = [a + b_ for a, b_ in zip(acc_ms, ms3)]
        lib.fkb_set_option(ctx, b"phase_events", 0)
        phase_ms = [a / n_extra for a in acc_ms]
        if world > 1:
            t3 = torch.tensor(phase_ms, dtype=torch.float64, device=dev)
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            phase_ms = [float(x) for x in t3]

    # ---- parity of what was timed (outside the timed region) ----
    # (1) conservation over the record layout; (2) at the metric's own size, the table and scalars against the oracle's full-size
    # record (tests/golden/golden_fullsize.json: sha256 of the table of a record-parallel run of oracle/kmer_oracle.c);
    # (3) N > 1: rank 0 counts the WHOLE stream alone and the NCCL-reduced table, the flag-derived node count and every scalar
    # must be bit-identical to that single-GPU result.
    parity = {}
    expect = sum(max(0, raw_lay.record_bases(r) - k + 1) for r in range(raw_lay.n_records))
    if rank == 0:
        import hashlib
        c = FkbCounts.from_buffer_copy(d_counts.cpu().numpy().tobytes())
        assert c.n_kmers == expect and c.base_total == args.bases and not c.rollover, (c.n_kmers, expect, c.base_total)
        reduced_table = acc.table.clone()
        parity["conservation"] = "ok"
        gpath = ROOT / "tests" / "golden" / "golden_fullsize.json"
        if gpath.exists() and args.bases == N_BASES:
            g = json.loads(gpath.read_text())["records"].get(f"config4_k{k}")
            if g:
                sha = hashlib.sha256(reduced_table.cpu().numpy().astype("<u4").tobytes()).hexdigest()
                assert sha == g["table_sha256"], "count table differs from the oracle's full-size record"
                assert (c.n_kmers, c.base_total, list(c.base_count), c.node_count, c.unknown_chars) == \
                       (g["n_kmers"], g["base_total"], g["base_count"], g["node_count"], g["unknown_chars"])
                parity["oracle_fullsize"] = "ok: sha256(table) + 8 scalars == tests/golden/golden_fullsize.json config4_k%d" % k
        if world > 1:
            whole = counter.synth_fasta_device(str_lay)
            acc1 = counter.new_accumulators(k)
            counter.count_stream_device(whole, k, acc1)
            one = counter.finalize_device(acc1, stream_total, fetch_table=False)
            assert torch.equal(acc1.table, reduced_table), "NCCL-reduced table differs from the single-GPU table"
            assert (one.n_kmers, one.base_total, tuple(one.base_count), one.node_count, one.unknown_chars, one.runs_ge_k, one.valid_bases) == \
                   (c.n_kmers, c.base_total, tuple(c.base_count), c.node_count, c.unknown_chars, c.runs_ge_k, c.valid_bases)
            parity["parity_n"] = f"ok: {world}-rank reduced table, node count and scalars bit-identical to one GPU counting the whole stream"
            del whole, acc1
            torch.cuda.empty_cache()
        del reduced_table

    # ---- roofline of the dominant kernel (count): algorithmic bytes = 1 B/base of stream + the 4^k x 4 B table ----
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    shard_bytes = (e - b) + (4 ** k) * 4  # per launch (per rank): this rank's stream bytes + its table
    achieved = shard_bytes / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic of the same launch sequence, from the committed ncu capture (profiles/r01_traffic_k11.json:
    # dram__bytes_read.sum + dram__bytes_write.sum of every kernel of one step; it cannot be measured live)
    traffic, traffic_note, shares = None, None, None
    tpaths = sorted((ROOT / "profiles").glob("r*_traffic_k11.json"))  # the newest committed capture (named per round)
    tpath = tpaths[-1] if tpaths else None
    if tpath is not None and world == 1 and k == 11 and args.bases == N_BASES:
        tj = json.loads(tpath.read_text())
        traffic = tj["dram_bytes_per_step"]
        traffic_note = ("ncu capture " + tpath.name + f": {traffic / shard_bytes:.1f}x the algorithmic bytes BY DESIGN -- the routed 32-bit items "
                        "(one 16-mer per 6 bases) are written and read once (2 x 2/3 B/base); the stream itself is read exactly once and there is "
                        "no intermediate table in HBM (pass 2 folds in shared memory)")
        count_kernels = [kk for kk in tj["kernels"] if not kk["kernel"].startswith("finalize")]
        tot_us = sum(kk["us"] for kk in count_kernels) or 1.0
        shares = " + ".join(f"{kk['kernel']} ({100 * kk['us'] / tot_us:.0f} %)" for kk in sorted(count_kernels, key=lambda x: -x["us"]))
    dominant = None
    if phase_ms and phase_ms[0] > 0:
        p1 = (e - b) / (phase_ms[0] * 1e-3) / 1e9  # pass 1 reads this rank's stream bytes once: its algorithmic bytes
        p1_name = "bucketize16_kernel (pass 1: 16-mer items, one per 6 bases)" if k == 11 else "bucketize_kernel<%d> (pass 1)" % (13 - k + 1)
        dominant = {"kernel": p1_name, "ms": phase_ms[0], "achieved": p1, "frac": p1 / peak,
                    "algorithmic_bytes": e - b, "share_of_count_path": phase_ms[0] / kernel_ms,
                    "other_kernels_ms": {"count_buckets kernel (pass 2, incl. shared-memory fold)": phase_ms[1], "fold kernels": phase_ms[2]},
                    "how": "CUDA events recorded by the library around its kernels (option phase_events), mean of 5 steps after the timed region"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "dominant": dominant,
                "kernel": "count path of one step: " + (shares or "bucketize_kernel + count_buckets_kernel + edges (shares: profiles/)"),
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": shard_bytes, "peak_source": peak_src, "traffic_note": traffic_note,
                "binding_resource": "pass 1 (69 % of the count path): the ALU pipe (57 % busy on average, math_pipe_throttle the second stall "
                                    "reason; its ALU-pipe instructions alone are 0.75 ms) plus the synchronous flush of the staging rows (0.28 ms of "
                                    "store-path time); the load latency that bound it before is hidden by an L2 prefetch two iterations beyond the "
                                    "register pipeline (profiles/r02_prefetch.txt).  pass 2: shared-memory atomics (data pipe 76 % busy, 3.3 wavefronts "
                                    "per counter atomic; without any load latency it would take 0.525 instead of 0.569 ms).  DRAM ~47 % of the copy "
                                    "bandwidth: not HBM bound (profiles/r02_ceiling_model.md)"}

    # ---- e2e leg: raw FASTA bytes in pinned host memory -> counts on the host ----
    e2e = None
    if not args.no_e2e:
        rcuts = shard_cuts(raw_total, world)
        ra, rb = rcuts[rank], rcuts[rank + 1]
        look = min(ra, 1 << 16)
        d_raw = counter.synth_fasta_device(raw_lay, ra - look, rb - ra + look)
        h_raw = torch.empty(d_raw.numel(), dtype=torch.uint8, pin_memory=True)
        h_raw.copy_(d_raw)
        torch.cuda.synchronize()
        del d_raw, d_stream
        torch.cuda.empty_cache()
        h_table = torch.empty(4 ** k, dtype=torch.int32, pin_memory=True)
        h_counts = FkbCounts()
        h2d = d2h = 0

        def e2e_step():
            nonlocal h2d, d2h
            if world == 1:
                counter._check(lib.fkb_count_fasta_host(ctx, h_raw.data_ptr(), h_raw.numel(), k, h_table.data_ptr(), ctypes.byref(h_counts)))
                h2d = h_raw.numel() + 64  # pinned input: the raw file bytes cross PCIe once and are stripped on the GPU
                d2h = h_table.numel() * 4 + ctypes.sizeof(FkbCounts)
                return
            counter._check(lib.fkb_zero_device(ctx, k, acc.table.data_ptr(), acc.flags.data_ptr(), acc.partials.data_ptr(), st.cuda_stream))
            torch.cuda.synchronize()
            sb, stop, eih = counter.count_fasta_range(h_raw, look, k, acc)
            h2d = h_raw.numel() - look + 64
            reduce_acc()
            if rank == 0:
                counter._check(lib.fkb_finalize_device(ctx, k, acc.table.data_ptr(), acc.flags.data_ptr(), acc.partials.data_ptr(),
                                                       stream_total, d_counts.data_ptr(), st.cuda_stream))
                h_table.copy_(acc.table, non_blocking=True)
                hc = d_counts.cpu()
                ctypes.memmove(ctypes.byref(h_counts), hc.numpy().ctypes.data, ctypes.sizeof(FkbCounts))
                d2h = h_table.numel() * 4 + ctypes.sizeof(FkbCounts)
            torch.cuda.synchronize()

        e2e_step()  # warm-up (allocates the pipeline's device stream and pinned slots)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s, float(h2d)], dtype=torch.float64, device=dev)
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            e2e_s, h2d = float(tm[0]), int(t[1])
        if rank == 0:
            assert h_counts.n_kmers == expect and h_counts.base_total == args.bases
            assert int(h_table.to(torch.int64).sum()) == expect
        e2e = {"value": args.bases / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "path": "raw FASTA bytes in pinned host memory -> H2D in 128 MiB chunks -> device strip (record/line compaction) -> "
                       "count -> [reduce] -> finalize -> table + counts D2H; pageable inputs use the host loader threads instead"}
        # the same call on PAGEABLE host memory -- what the drop-in CLI does with its mmap'ed file: the host loader threads strip into
        # pinned slots (cudaHostRegister of a file mapping is refused: profiles/r02_hostreg_probe.txt), H2D of the stripped stream
        if world == 1:
            import numpy as np
            h_page = np.empty(h_raw.numel(), dtype=np.uint8)
            h_page[:] = h_raw.numpy()

            def pageable_step():
                counter._check(lib.fkb_count_fasta_host(ctx, h_page.ctypes.data, h_page.size, k, h_table.data_ptr(), ctypes.byref(h_counts)))

            pageable_step()  # warm-up: pins the loader's ring
            t0 = time.perf_counter()
            for _ in range(2):
                pageable_step()
            pg_s = (time.perf_counter() - t0) / 2
            assert h_counts.n_kmers == expect and int(h_table.to(torch.int64).sum()) == expect
            e2e["pageable"] = {"value": args.bases / pg_s / 1e9, "unit": UNIT, "ms_per_step": pg_s * 1e3, "steps": 2,
                               "path": "raw FASTA bytes in pageable host memory -> host loader threads (AVX-512 strip into a ring of pinned 4 MiB "
                                       "slots) -> H2D of the stripped stream -> count -> finalize -> table + counts D2H"}
            del h_page

    # ---- secondary workload (rank 0, N = 1): BASELINE.json config 5 -- 3.1 Gbp with N runs and soft-masked lower case, k = 8 ----
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        import hashlib
        k2 = 8
        lay5 = synth.config5(n_bases=args.bases).stripped()
        d5 = counter.synth_fasta_device(lay5)
        acc5 = counter.new_accumulators(k2)
        d_counts5 = torch.zeros(12, dtype=torch.int64, device=dev)
        ev5 = []
        for i in range(3 + 5):
            counter._check(lib.fkb_zero_device(ctx, k2, acc5.table.data_ptr(), acc5.flags.data_ptr(), acc5.partials.data_ptr(), st.cuda_stream))
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(st)
            counter.count_stream_device(d5, k2, acc5)
            e1.record(st)
            counter._check(lib.fkb_finalize_device(ctx, k2, acc5.table.data_ptr(), acc5.flags.data_ptr(), acc5.partials.data_ptr(),
                                                   d5.numel(), d_counts5.data_ptr(), st.cuda_stream))
            e2.record(st)
            if i >= 3:
                ev5.append((e0, e1, e2))
        torch.cuda.synchronize()
        ms5 = sum(a.elapsed_time(c_) for a, _, c_ in ev5) / len(ev5)
        kms5 = sum(a.elapsed_time(b_) for a, b_, _ in ev5) / len(ev5)
        c5 = FkbCounts.from_buffer_copy(d_counts5.cpu().numpy().tobytes())
        par5 = "unchecked (no golden record for this size)"
        gpath = ROOT / "tests" / "golden" / "golden_fullsize.json"
        if gpath.exists() and args.bases == N_BASES:
            g = json.loads(gpath.read_text())["records"]["config5_k8"]
            sha = hashlib.sha256(acc5.table.cpu().numpy().astype("<u4").tobytes()).hexdigest()
            assert sha == g["table_sha256"] and (c5.n_kmers, c5.base_total, list(c5.base_count), c5.node_count, c5.unknown_chars) == \
                (g["n_kmers"], g["base_total"], g["base_count"], g["node_count"], g["unknown_chars"])
            par5 = "ok: sha256(table) + scalars == tests/golden/golden_fullsize.json config5_k8"
        bytes5 = d5.numel() + (4 ** k2) * 4
        secondary = {"config": f"config 5: synthetic {args.bases / 1e9:.3g} Gbp, 24 records, ~5 % of positions in N runs + ~10 % soft-masked lower case, k={k2}",
                     "value": args.bases / (ms5 * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms5, "steps": len(ev5),
                     "kernel": "count_smem_kernel<8> (single pass, 65536 16-bit counters per CTA in shared memory) + edge slivers",
                     "roofline": {"bound": "hbm", "achieved": bytes5 / (kms5 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": bytes5 / (kms5 * 1e-3) / 1e9 / peak, "kernel_ms": kms5, "algorithmic_bytes_per_launch": bytes5},
                     "parity": par5}
        del d5, acc5
        torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        with tempfile.TemporaryDirectory() as d:
            cs, ts, kind = reference_sample_run(args.cpu_sample_bases, k, d)
        cpu_baseline = {"value": args.cpu_sample_bases / cs / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                        "sample": f"{args.cpu_sample_bases} bases of the same shape (24 records, 60-col), k={k}; count phase {cs:.2f} s "
                                  f"(whole run {ts:.2f} s); the reference is single-threaded; host has {os.cpu_count()} cores"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
                "data": "synthetic", "config": workload_config(args), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clock_info, "parity": parity, "secondary": secondary}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    counter.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
