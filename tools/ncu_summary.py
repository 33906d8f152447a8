"""Summarise ncu raw/source CSV exports (development aid): python tools/ncu_summary.py raw.csv [src.csv ...]"""
import csv, collections, sys
def raw(path):
    rows=list(csv.reader(open(path)))
    hdr=rows[0]; units=rows[1]; data=rows[2:]
    idx={h:i for i,h in enumerate(hdr)}
    base=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_elapsed.avg.per_second']
    stalls=[h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    for d in data:
        print('----', d[idx['Kernel Name']][:50], 'grid', d[idx['launch__grid_size']], 'block', d[idx['launch__block_size']])
        for w in base:
            if w in idx: print(f"   {w:70s} {d[idx[w]]:>16s} {units[idx[w]]}")
        st=sorted([(float(d[idx[h]]),h) for h in stalls], reverse=True)[:6]
        print('      stalls: '+', '.join(f"{h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}={v:.2f}" for v,h in st))
def src(path, top=12):
    rows=list(csv.reader(open(path)))
    print('====', rows[0][1][:60])
    hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
    data=[d for d in rows[2:] if len(d)>idx['# Samples'] and d[idx['# Samples']].isdigit()]
    tot=sum(int(d[idx['# Samples']]) for d in data)
    stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg=collections.Counter(); aggs=collections.defaultdict(collections.Counter); ex=collections.Counter()
    for d in data:
        s=d[idx['Source']].strip().split()
        op=s[1] if s and s[0].startswith('@') else (s[0] if s else '')
        n=int(d[idx['# Samples']]); agg[op]+=n
        v=d[idx['Instructions Executed']]
        if v.isdigit(): ex[op]+=int(v)
        for sc in stall_cols:
            v=d[idx[sc]]
            if v.isdigit() and int(v): aggs[op][sc]+=int(v)
    print('total samples', tot, ' executed warp-instr', sum(ex.values()))
    for op,n in agg.most_common(top):
        print(f"  {op:24s} samples {100*n/max(tot,1):5.1f}%  exec {100*ex[op]/max(sum(ex.values()),1):5.1f}%  ", dict(aggs[op].most_common(3)))
    allst=collections.Counter()
    for op in aggs:
        for k,v in aggs[op].items(): allst[k]+=v
    print('  all stalls:', {k: f"{100*v/max(tot,1):.1f}%" for k,v in allst.most_common(8)})
if __name__=='__main__':
    raw(sys.argv[1])
    for p in sys.argv[2:]: src(p)
