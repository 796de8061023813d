#!/usr/bin/env python
"""GPU-box diagnostic: does the device-resident PE step slow down after other engine calls?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

def main():
    args = bench.argparse.Namespace(gpus=1, steps=3, warmup=3, genome_mb=float(os.environ.get("GENOME_MB", 3100)),
                                    reads=int(os.environ.get("READS", 5_000_000)), read_len=150, group_width=8, table_depth=0, workload="pe")
    torch.cuda.set_device(0)
    wl = bench.Workload(args, 0, 0)
    st = torch.cuda.current_stream()
    def timed(tag, k=4):
        for _ in range(2): wl.device_step(st.cuda_stream)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(st)
        for _ in range(k): wl.device_step(st.cuda_stream)
        ev1.record(st); torch.cuda.synchronize()
        print(f"{tag}: {ev0.elapsed_time(ev1)/k:.2f} ms/step (wall {(time.perf_counter()-t0)*1e3/k:.2f})", flush=True)
    timed("A fresh")
    b1, o1 = wl.sample_reads(200000, 1); b2, o2 = wl.sample_reads(200000, 2)
    wl.e.map_pe(b1, o1, b2, o2, m=wl.m, b=bench.B, top_k=bench.TOP_K, frag_range=bench.FRAG)
    timed("B after host-path map_pe (full lists)")
    wl.e.export_subindex(0, int(wl.lengths.sum()), want_seq=True, want_counter=True, want_index=True)
    timed("C after export_subindex")
    x = np.ones(4_000_000_000 // 8); x.sum(); del x
    time.sleep(5)
    timed("D after 5 s idle + 4 GB host alloc")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refio
    if refio.have_reference():
        hidx = wl.host_index()
        timed("E after host_index (reference-owned copies of 4 sub-indexes)")
        t, _ = bench.reference_pass(wl, hidx, 20000, os.cpu_count())
        timed("F after an OpenMP reference pass")
        L = refio.ref_lib()
        for h in hidx.values(): L.waltref_index_free(h)
        timed("G after freeing them")

if __name__ == "__main__":
    main()
