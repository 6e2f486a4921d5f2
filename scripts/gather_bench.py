import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import torch
import metrocpp_b200 as mb
eng = mb.MetroEngine(0)
lib = C.CDLL(os.path.join(os.path.dirname(mb.__file__), "libmetro_b200_bench.so"))      # micro-benchmarks live outside the product ABI
lib.metro_gather_bench.restype = C.c_int
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 6.4
n_sectors = int(gb * 1e9 / 32)
table = torch.ones(n_sectors * 4, dtype=torch.int64, device="cuda")
n_threads = 148 * 2048 * 4
out = torch.empty(n_threads, dtype=torch.int64, device="cuda")
ms = C.c_float()
for ilp in (1, 4, 8):
    per = 416
    rc = lib.metro_gather_bench(eng.h, C.c_void_p(table.data_ptr()), C.c_uint64(n_sectors), C.c_int64(n_threads), per, ilp, C.c_void_p(out.data_ptr()), C.byref(ms))
    n = n_threads * per
    print(f"table {gb} GB ilp {ilp}: rc {rc} {n/1e6:.0f} M sector reads in {ms.value:.3f} ms = {n*32/ms.value/1e6:.0f} GB/s useful, {n/ms.value/1e6:.2f} G probes/s")
