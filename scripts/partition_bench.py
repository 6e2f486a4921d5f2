import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import metrocpp_b200 as mb
eng = mb.MetroEngine(0)
lib = C.CDLL(os.path.join(os.path.dirname(mb.__file__), "libmetro_b200_bench.so"))      # micro-benchmarks live outside the product ABI
lib.metro_partition_bench.restype = C.c_int
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
keys = torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g)
cap = int(n / 256 * 1.03) + 8192
out = torch.empty(256 * cap, dtype=torch.int64, device="cuda")
ms = C.c_float(); counts = (C.c_uint64 * 257)()
for shift in (54, 22):
    rc = lib.metro_partition_bench(eng.h, C.c_void_p(keys.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int64(n), shift, C.c_int64(cap), 5, C.byref(ms), counts)
    c = np.array(counts[:256]); print("shift", shift, "rc", rc, "ms", ms.value, "GB/s", 16 * n / ms.value / 1e6, "overflow", counts[256], "min/max bucket", c.min(), c.max(), "sum ok", c.sum() == n)
# verify one bucket
b = 17; cnt = int(counts[b]); got = out[b * cap: b * cap + cnt].cpu().numpy().astype(np.uint64)
exp = keys.cpu().numpy().astype(np.uint64); exp = exp[((exp >> np.uint64(22)) & np.uint64(255)) == b]
print("bucket multiset equal:", np.array_equal(np.sort(got), np.sort(exp)))
