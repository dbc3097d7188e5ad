#!/usr/bin/env python
"""Selected metrics + top stall reasons of every kernel launch in an .ncu-rep (ncu --page raw --csv).  usage: ncu_summary.py <report.ncu-rep>"""
import subprocess, sys, csv, io
rep = sys.argv[1]
keys = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active",
"smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
"sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__inst_executed.sum","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__maximum_warps_per_active_cycle_pct","launch__grid_size","launch__block_size",
"dram__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sector_hit_rate.pct","lts__t_bytes.sum","l1tex__lsu_writeback_active_mem_lg.sum","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active","smsp__cycles_active.avg","sm__cycles_elapsed.max"]
out = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:  # one block per captured kernel launch
    d = dict(zip(hdr, zip(units, vals)))
    print(rep, d.get("Kernel Name",("",""))[1][:100])
    for k in keys:
        if k in d: print(f"  {k:75s} {d[k][1]:>18s} {d[k][0]}")
    # stall reasons
    st = [(k, float(v[1].replace(',',''))) for k,v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") or (k.startswith("smsp__average_warp_latency_issue_stalled") and k.endswith(".ratio"))]
    st.sort(key=lambda x:-x[1])
    for k,v in st[:8]: print(f"  stall {k.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_','')[:40]:45s} {v:.3f}")
