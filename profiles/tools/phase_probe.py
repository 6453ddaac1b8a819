import sys, ctypes, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, ".")
from dataclasses import replace
from findkmer_b200 import synth
from findkmer_b200.engine import KmerCounter
c = KmerCounter(0); lib, ctx = c._lib, c._ctx
if len(sys.argv) > 1: c.set_variant(int(sys.argv[1]))   # 2 = 13-mer buckets, 4 = 16-mer items (k = 11)
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else (388_000_000, 775_000_000, 3_100_000_000)
st = torch.cuda.current_stream()
lib.fkb_set_option(ctx, b"phase_events", 1)
for n in sizes:
    d = c.synth_fasta_device(replace(synth.config4(), n_bases=n).stripped())
    for k in (11,):
        acc = c.new_accumulators(k); acc_ms=[0,0,0]; tot=0
        for it in range(6):
            lib.fkb_zero_device(ctx, k, acc.table.data_ptr(), acc.flags.data_ptr(), acc.partials.data_ptr(), st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); c.count_stream_device(d, k, acc); e1.record(st); torch.cuda.synchronize()
            ms = (ctypes.c_double*3)(); lib.fkb_phase_times(ctx, ctypes.byref(ms))
            if it >= 2:
                acc_ms=[a+b for a,b in zip(acc_ms, ms)]; tot += e0.elapsed_time(e1)
        print(n, k, "total %.3f pass1 %.3f pass2 %.3f rest %.3f" % (tot/4, acc_ms[0]/4, acc_ms[1]/4, tot/4-acc_ms[0]/4-acc_ms[1]/4))
    del d; torch.cuda.empty_cache()
