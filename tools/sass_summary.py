import csv,sys,collections,re
def summarize(path, top=25):
    rows=list(csv.reader(open(path)))
    hdr=rows[1]; si=hdr.index('Source'); ei=hdr.index('Instructions Executed'); sm=hdr.index('# Samples')
    ops=collections.Counter(); samp=collections.Counter(); tot=0; tots=0
    for r in rows[2:]:
        if len(r)<=ei: continue
        m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si]); 
        if not m: continue
        op=m.group(2); op='.'.join(op.split('.')[:2])
        n=int(r[ei]); s=int(r[sm]); ops[op]+=n; samp[op]+=s; tot+=n; tots+=s
    print(f'{path}: {len(rows)-2} SASS instrs, {tot/1e6:.1f}M warp-instr executed')
    for op,n in ops.most_common(top): print(f'  {op:22s} {n/1e6:8.2f}M {100*n/tot:5.1f}%   samples {100*samp[op]/max(tots,1):5.1f}%')
summarize(sys.argv[1], int(sys.argv[2]) if len(sys.argv)>2 else 25)
