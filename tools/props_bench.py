#!/usr/bin/env python
"""Device time of the property recipes (include/eggprops.h): `python tools/props_bench.py [ngal]`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from egg_b200 import properties, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 195000
opt = synth.make_opt_lib()
rng = np.random.default_rng(1)
z = torch.from_numpy(rng.uniform(0.05, 9, n).astype(np.float32)).cuda()
m = torch.from_numpy(rng.uniform(7.5, 11.9, n).astype(np.float32)).cuda()
p = torch.from_numpy((rng.random(n) < 0.25).astype(np.uint8)).cuda()
pr = properties.Properties(opt)
cols, nbytes = {}, 0
for k in pr.columns():
    shape = (n, properties.SHAPE[k]) if k in properties.SHAPE else (n,)
    cols[k] = torch.empty(shape, dtype=torch.int64 if k in properties.U64_COLS else torch.float32, device="cuda")
    nbytes += cols[k].numel() * cols[k].element_size()
st = torch.cuda.current_stream()
ptrs = {k: v.data_ptr() for k, v in cols.items()}
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    pr.compute_device(z.data_ptr(), m.data_ptr(), p.data_ptr(), n, ptrs, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print(f"props_kernel: {n} galaxies in {ms:.3f} ms = {n / ms * 1e3:.3e} galaxies/s, {nbytes / n:.0f} B written per galaxy, {nbytes / ms / 1e6:.1f} GB/s")
