#!/bin/bash
# usage: scripts/ncu_summary.sh rep.ncu-rep  -> key metrics + by-function table + stall totals
REP=$1
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); hdr=r[0]; vals=r[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_active.avg','lts__t_sectors.sum','l1tex__t_sectors.sum']
for k in keys:
    for i,h in enumerate(hdr):
        if h==k: print(k,'=',vals[i],r[1][i])
"
ncu -i $REP --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/_cs.csv
python scripts/ncu_by_function.py /tmp/_cs.csv aom-av1-psy_b200/csrc/tf_kernels.cuh ${2:-25}
ncu -i $REP --page source --csv 2>/dev/null > /tmp/_s.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/_s.csv'))); hdr=rows[1]; tot={}
for r in rows[2:]:
    for h,v in zip(hdr,r):
        if h.startswith('stall') and 'Not Issued' not in h:
            try: tot[h]=tot.get(h,0)+int(v)
            except: pass
s=sum(tot.values()) or 1
print("stalls:", ", ".join(f"{k[6:]} {v/s*100:.1f}%" for k,v in sorted(tot.items(), key=lambda x:-x[1])[:8]))
PY
