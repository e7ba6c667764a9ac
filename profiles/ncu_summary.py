#!/usr/bin/env python
"""Key raw metrics of an ncu report (first profiled launch):  ncu_summary.py report.ncu-rep
   and launch-list aggregation:                               ncu_summary.py --launches launches.csv"""
import csv, subprocess, sys
WANT=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed',
'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__inst_executed.sum','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','launch__block_size','smsp__inst_executed.sum',
'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']
if sys.argv[1]=='--launches':
    rows=list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
    h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
    d={}
    for r in rows[1:]: d.setdefault(r[ki].split('(')[0][-40:],[]).append(float(r[vi].replace(',','')))
    tot=sum(sum(v)/len(v) for v in d.values())
    for k,v in d.items(): print('%-44s n=%3d avg=%10.1f ns  share=%.1f%%'%(k,len(v),sum(v)/len(v),100*sum(v)/len(v)/tot))
else:
    txt=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
    rows=list(csv.reader(txt.splitlines())); h=rows[0]; u=rows[1]; v=rows[2]
    for w in WANT:
        if w in h: i=h.index(w); print('%-82s %s %s'%(w,v[i],u[i]))
