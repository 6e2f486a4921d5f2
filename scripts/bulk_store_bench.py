import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import torch
import metrocpp_b200 as mb
eng = mb.MetroEngine(0)
lib = C.CDLL(os.path.join(os.path.dirname(mb.__file__), "libmetro_b200_bench.so"))      # micro-benchmarks live outside the product ABI
lib.metro_bulk_store_bench.restype = C.c_int
out = torch.empty(1 << 30, dtype=torch.int64, device="cuda")      # 8 GB
ms = C.c_float()
for nbytes in (64, 128, 256, 1024):
    for blocks in (148 * 3, 148 * 8):
        per = 256
        rc = lib.metro_bulk_store_bench(eng.h, C.c_void_p(out.data_ptr()), C.c_uint64(out.numel() * 8), blocks, per, nbytes, C.byref(ms))
        n = blocks * 256 * per
        print(f"{nbytes:5d} B x {n/1e6:.1f} M copies, {blocks} blocks: rc {rc} {ms.value:.3f} ms = {n*nbytes/ms.value/1e6:.0f} GB/s, {n/ms.value/1e6:.2f} G copies/s, "
              f"{ms.value*1e-3*1.9e9*148/n:.1f} SM-cycles per copy")
