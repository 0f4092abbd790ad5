"""F1 measurement: SAM text -> columns on the device vs the host reader."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
from woltka_b200.engine import Engine
from woltka_b200 import align

rng = np.random.default_rng(7)
rows = []
for g in range(200_000):
    name = f'S{g % 8:02d}_read{g:08d}'
    first = int(rng.integers(0, 10000))
    for _ in range(int(min(rng.geometric(0.48), 16))):
        sub = f'G{(first + int(rng.integers(0, 20))) % 10000:09d}'
        rows.append(f'{name}\t{16 * int(rng.integers(0, 2))}\t{sub}\t{int(rng.integers(1, 5000000))}\t42\t150M\t*\t0\t0\t'
                    'ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTAC\tIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII')
block = ('\n'.join(rows) + '\n').encode()
reps = 8
text = block * reps
n_lines = len(rows) * reps
eng = Engine(0)
for _ in range(2):
    out = eng.parse_sam(text, True)
eng.sync()
t0 = time.perf_counter()
K = 5
for _ in range(K):
    out = eng.parse_sam(text, True)
eng.sync()
dt = (time.perf_counter() - t0) / K
lines = block.decode().splitlines(keepends=True)
t0 = time.perf_counter()
nq = sum(1 for _ in align.iter_align(iter(lines), 'sam'))
ht = time.perf_counter() - t0
print(json.dumps({
    'what': 'F1 SAM reader: host text -> device columns (H2D of the text inside)',
    'text_bytes': len(text), 'lines': n_lines, 'records': out[0], 'queries': out[1],
    'device_ms': dt * 1e3, 'device_records_per_s': out[0] / dt,
    'device_text_GBps': len(text) / dt / 1e9,
    'host_reader_records_per_s': len(rows) / ht, 'host_reader': 'woltka_b200.align (pure Python, 1 core)'}))
