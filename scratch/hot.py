import csv, sys
fn=sys.argv[1]; N=float(sys.argv[2]) if len(sys.argv)>2 else 1e8
rows=list(csv.reader(open(fn)))
hdr=rows[1]; data=rows[2:]
isrc=hdr.index('Source'); ins=hdr.index('Instructions Executed'); ismp=hdr.index('# Samples')
tot=sum(int(r[ins]) for r in data if r[ins].isdigit())
smp=sum(int(r[ismp]) for r in data if r[ismp].isdigit())
print('total warp instr', tot, 'per record', tot/N, 'samples', smp)
thr = float(sys.argv[3]) if len(sys.argv)>3 else 0.004
for i,r in enumerate(data):
    try: n=int(r[ins]); s=int(r[ismp])
    except: continue
    if n>tot*thr or s>smp*thr*2: print(f"{i:5d} {n/N:7.3f} {100*s/smp:5.1f}%  {r[isrc][:100]}")
