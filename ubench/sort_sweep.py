"""Radix-sort microbench (config c5): Gkeys/s for pairs and keys-only over n = 2^20 .. 2^28, for the library
named by GSR_LIB_PATH (kernel-variant experiments) or the in-tree libgsr.so."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from godotgaussiansplatting_b200 import _lib
from godotgaussiansplatting_b200.synthetic import radix_keys
L = _lib.lib()
sizes = [int(a) for a in sys.argv[1:]] or [20, 22, 24, 26]
res = {}
for lg in sizes:
    n = 1 << lg
    for kind in ("tile_depth", "uniform32"):
        keys = torch.from_numpy(radix_keys(n, 5, kind).view(np.int32)).cuda()
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        s = C.c_void_p(); _lib.check(L.gsr_sorter_create(0, n, C.byref(s)), "create")
        for with_vals in (True, False):
            t = []
            for it in range(6):
                k = keys.clone(); v = vals.clone() if with_vals else None
                torch.cuda.synchronize()
                _lib.check(L.gsr_sorter_sort_device(s, C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()) if with_vals else None, n, None), "sort")
                ms = C.c_float(); _lib.check(L.gsr_sorter_last_ms(s, C.byref(ms)), "ms")
                if it >= 2: t.append(ms.value)
            if lg <= 22:
                assert bool((k[1:].view(torch.int32).to(torch.int64) & 0xFFFFFFFF >= k[:-1].to(torch.int64) & 0xFFFFFFFF).all())
            res[f"2^{lg} {kind} {'pairs' if with_vals else 'keys'}"] = round(n / np.mean(t) / 1e6, 2)
        L.gsr_sorter_destroy(s)
print(os.environ.get("GSR_LIB_PATH", "libgsr.so"), json.dumps(res))
