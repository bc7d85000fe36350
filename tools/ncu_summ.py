import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
H=rows[0]; V=rows[2] if len(rows)>2 else rows[1]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','launch__grid_size','launch__registers_per_thread']
for h,v in zip(H,V):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h): print(h,v)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
H=rows[1]
ia=H.index('Address'); isrc=H.index('Source'); iall=H.index('Warp Stall Sampling (All Samples)'); iex=H.index('Instructions Executed')
data=[]
for r in rows[2:]:
    try: data.append((int(r[iall]), r[isrc], int(r[iex]), r[ia]))
    except: pass
tot=sum(d[0] for d in data)
from collections import Counter
c=Counter(); ce=Counter()
for smp,s,ex,ad in data:
    op=s.split()[0] if not s.startswith('@') else s.split()[1]
    op=op.split('.')[0]
    c[op]+=smp; ce[op]+=ex
print("total",tot)
for op,v in c.most_common(10): print(f"{op:10s} samples {v:7d} ({100*v/tot:5.1f}%)  executed {ce[op]}")
for d in sorted(data,reverse=True)[:12]: print(d[0], d[3][-5:], d[1][:90], d[2])
