"""Profiling aid: per-kernel stall-reason shares, opcode mix (executed share / stall-sample share) and the hottest SASS lines from
`ncu -i capture.ncu-rep --page source --csv > src.csv`.  Usage: summarise_ncu_source.py src.csv [lines per kernel]"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
# multiple kernels: split on "Kernel Name" rows
blocks=[];cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={'name':r[1],'rows':[]}; blocks.append(cur)
    elif cur is not None: cur['rows'].append(r)
for b in blocks:
    hdr=b['rows'][0]; data=[r for r in b['rows'][1:] if len(r)==len(hdr)]
    I=lambda n: hdr.index(n)
    tot=sum(int(r[I('# Samples')]) for r in data); inst=sum(int(r[I('Instructions Executed')]) for r in data)
    print("==",b['name'][:60],"samples",tot,"warp-inst",inst)
    stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg={s:sum(int(r[I(s)]) for r in data) for s in stalls}
    print("  "+"  ".join(f"{k[6:]}:{v/tot:.2f}" for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:7]))
    op=collections.Counter(); ops=collections.Counter()
    for r in data:
        t=r[I('Source')].split(); o=t[1] if t[0].startswith('@') else t[0]
        op[o.split('.')[0]]+=int(r[I('Instructions Executed')]); ops[o.split('.')[0]]+=int(r[I('# Samples')])
    print("  ops: "+"  ".join(f"{k}:{v/inst:.2f}/{ops[k]/tot:.2f}" for k,v in op.most_common(14)))
    for r in sorted(data,key=lambda r:-int(r[I('# Samples')]))[:int(sys.argv[2]) if len(sys.argv)>2 else 8]:
        print(f"  {int(r[I('# Samples')]):5d} {int(r[I('Instructions Executed')]):8d}  {r[I('Source')].strip()[:100]}")
