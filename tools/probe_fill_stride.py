import ctypes, os, sys
sys.path.insert(0, '/root/repo')
import torch
from signals_b200 import _lib
L = _lib.lib()
L.sigb_probe_fill_tiled.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
st = torch.cuda.current_stream().cuda_stream
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for C in (4096, 4160, 4032, 8192, 2048, 6144):
    frames = 480000 * 4096 // C // 16 * 16
    out = torch.empty(frames * C, dtype=torch.float32, device='cuda')
    for width in (64,):
        for rows in (16, 144):
            ms = timed(lambda: L.sigb_probe_fill_tiled(ctypes.c_void_p(out.data_ptr()), frames, C, width, rows, 148 * 3, ctypes.c_void_p(st)))
            print(f'C={C} tiled fill {width*4}-byte rows x {rows}: {ms:.3f} ms {frames*C*4/ms/1e6:.0f} GB/s')
    del out
