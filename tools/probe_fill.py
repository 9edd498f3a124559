#!/usr/bin/env python
"""Write-only HBM ceiling on this GPU: a streaming fill of the C2 output block (7.86 GB) with plain
128-bit stores, and torch's own fill for comparison.  Usage (GPU box): python tools/probe_fill.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from signals_b200 import _lib
L = _lib.lib()
L.sigb_probe_fill.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_int32, ctypes.c_void_p]
n = 480000 * 4096
out = torch.empty(n, dtype=torch.float32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for blocks in (148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 32):
    ms = timed(lambda: L.sigb_probe_fill(ctypes.c_void_p(out.data_ptr()), n, 1.0, blocks, ctypes.c_void_p(st)))
    print(f'k_probe_fill {blocks:5d} CTAs: {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s')
ms = timed(lambda: out.fill_(2.0))
print(f'torch fill_: {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s')
src = torch.empty(n // 2, dtype=torch.float32, device='cuda'); dst = torch.empty_like(src)
ms = timed(lambda: dst.copy_(src))
print(f'torch copy_ (read+write bytes): {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s')

L.sigb_probe_fill_tiled.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
for width in (32, 64, 128):        # floats per tile row; a warp's float4 lanes cover at most 128 floats (512 bytes)
    for rows in (16, 64):
        for blocks in (148 * 3, 148 * 6):
            ms = timed(lambda: L.sigb_probe_fill_tiled(ctypes.c_void_p(out.data_ptr()), 480000, 4096, width, rows, blocks, ctypes.c_void_p(st)))
            print(f'tiled fill: {width * 4:4d}-byte rows x {rows:2d} rows per tile, {blocks:4d} CTAs: {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s')

# adjacent column stripes written in lockstep by neighbouring CTAs (mode 1) against the same schedule with the stripes permuted (mode 2)
L.sigb_probe_fill_lockstep.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
for width in (32, 64):
    tiles = 4096 // width
    for mult in (1, 2, 4):
        blocks = max(1, (148 * mult) // tiles) * tiles
        for mode in (1, 2):
            ms = timed(lambda: L.sigb_probe_fill_lockstep(ctypes.c_void_p(out.data_ptr()), 480000, 4096, width, 16, blocks, mode, ctypes.c_void_p(st)))
            print(f'lockstep fill mode {mode}: {width * 4:4d}-byte rows x 16 rows per tile, {blocks:4d} CTAs: {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s')

# store flavour: streaming (.cs, what the kernels use), plain write-back, .cg, .wt
for how, name in ((0, 'st.global.cs'), (1, 'st.global (write-back)'), (2, 'st.global.cg'), (3, 'st.global.wt')):
    L.sigb_probe_set_store(how)
    ms = timed(lambda: L.sigb_probe_fill(ctypes.c_void_p(out.data_ptr()), n, 1.0, 148 * 32, ctypes.c_void_p(st)))
    print(f'{name:24s} contiguous fill, 4736 CTAs: {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s')
    for width in (32, 64):
        ms = timed(lambda: L.sigb_probe_fill_tiled(ctypes.c_void_p(out.data_ptr()), 480000, 4096, width, 16, 148 * 3, ctypes.c_void_p(st)))
        print(f'{name:24s} tiled fill {width * 4:4d}-byte rows x 16, 444 CTAs: {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s')
L.sigb_probe_set_store(0)

# time-major sweep (mode 3): every warp of the launch writes inside the same narrow band of rows (few 2 MB pages live at a time)
for width in (32, 64):
    for blocks in (148 * 3, 148 * 6):
        ms = timed(lambda: L.sigb_probe_fill_lockstep(ctypes.c_void_p(out.data_ptr()), 480000, 4096, width, 16, blocks, 3, ctypes.c_void_p(st)))
        print(f'time-major tiled fill: {width * 4:4d}-byte rows x 16 rows per tile, {blocks:4d} CTAs: {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s')
