import csv,sys,subprocess,re
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','sm__inst_executed.sum','smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors.sum','smsp__inst_executed_pipe_fp32x2.sum' ]
for r in rows[2:]:
    name=r[hdr.index('Kernel Name')][:90]
    print('---',name)
    for w in want:
        if w in hdr: print(f'   {w:72s} {r[hdr.index(w)][:20]} {units[hdr.index(w)]}')
    for i,h in enumerate(hdr):
        if re.search(r'pipe_f.*(2|x2)|fp32x2|fma2', h): print('   *', h, r[i])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
kern=None; hdr=None; agg={}
for r in rows:
    if r and r[0]=='Kernel Name': kern=r[1][:80]; agg[kern]={}; continue
    if r and r[0]=='Address': hdr=r; continue
    if kern is None or hdr is None or len(r)<len(hdr): continue
    for i,h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            agg[kern][h]=agg[kern].get(h,0)+int(r[i] or 0)
for k,v in agg.items():
    tot=sum(v.values()) or 1
    print(k); print('   ', ', '.join(f'{h[6:]}={100*c/tot:.1f}%' for h,c in sorted(v.items(), key=lambda kv:-kv[1])[:8]))
