import sys,csv,subprocess,collections
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy','sm__pipe_fmalite','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','launch__grid_size','smsp__average_warp_latency_per_inst_issued','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct','sm__inst_executed_pipe_fma.avg.pct','sm__inst_executed_pipe_lsu.avg.pct','sm__inst_executed_pipe_fmaheavy.avg.pct','sm__inst_executed_pipe_fmalite.avg.pct','smsp__warps_eligible.avg.per_cycle_active','smsp__issue_active.avg.per_cycle_active','sm__throughput.avg.pct']
for h,u,v in zip(hdr,units,vals):
    if any(h.startswith(w) for w in want) and 'pct_of_peak_sustained_elapsed' not in h.replace('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','') and '.per_second' not in h:
        print(f"{h} [{u}] = {v}")
print("--- stall reasons (warps per issue-active cycle)")
st=[(float(v),h) for h,v in zip(hdr,vals) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for v,h in sorted(st,reverse=True)[:8]:
    print(f"  {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):28s} {v:.3f}")
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ia=hdr.index("Instructions Executed"); isrc=hdr.index("Source")
ops=collections.Counter(); tot=0
for r in rows[2:]:
    try: n=int(r[ia])
    except: continue
    t=r[isrc].split(); op=t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]]+=n; tot+=n
print("--- executed warp-instructions by opcode (total %d)"%tot)
for op,n in ops.most_common(12): print(f"  {op:10s} {n:12d} {100*n/tot:5.1f}%")
