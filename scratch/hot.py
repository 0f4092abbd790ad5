import csv, sys
fn = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
rows = list(csv.reader(open(fn)))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); ismp = hdr.index('# Samples')
ith = hdr.index('Avg. Predicated-On Threads Executed')
tot = sum(int(r[iex]) for r in data if r[iex].isdigit())
smp = sum(int(r[ismp]) for r in data if r[ismp].isdigit())
print('total instr', tot, 'per rec', tot / 1e8, 'samples', smp)
for n, r in enumerate(data):
    ex = int(r[iex])
    if ex > tot * thr:
        print(n, r[isrc].strip()[:64].ljust(64), f'{ex / 1e8 * 32:7.2f}', f'{int(r[ismp]) / smp * 100:5.2f}%', r[ith])
