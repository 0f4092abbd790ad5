import torch, time
n=19_085_233
dev=torch.device('cuda',0)
k=torch.randint(-2**62, 2**62, (n,), device=dev, dtype=torch.int64); u=k.clone()
def t(label,f):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=f(); torch.cuda.synchronize(); print(f'{label}: {(time.perf_counter()-t0)*1e3:.2f} ms'); return r
for rep in range(3):
    print('rep',rep)
    owner=t('owner', lambda: ((k * -7046029254386353131) >> 40) % 8)
    counts=t('bincount', lambda: torch.bincount(owner, minlength=8))
    o8=t('to u8', lambda: owner.to(torch.uint8))
    order=t('sort u8', lambda: torch.sort(o8).indices)
    ks=t('gather k', lambda: k[order]); us=t('gather u', lambda: u[order])
    t('clone x2', lambda: (k.clone(), u.clone()))
    t('cpu', lambda: counts.cpu().tolist())
