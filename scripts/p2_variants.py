"""time pb_poseidon2_permute for each experimental build variant (powdr_b200/_lib/variants/*.so)"""
import ctypes as C, glob, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for so in sorted(glob.glob(os.path.join(ROOT, "powdr_b200/_lib/variants/*.so"))):
    lib = C.CDLL(so)
    h = C.c_void_p()
    assert lib.pb_ctx_create(C.byref(h), 0, None) == 0
    n, reps = 1 << 21, 16
    d = C.c_void_p()
    lib.pb_device_alloc(C.byref(d), C.c_size_t(64 * n))
    lib.pb_memset_zero(h, d, C.c_size_t(64 * n))
    best = 1e9
    for it in range(4):
        lib.pb_ctx_synchronize(h)
        t = time.time()
        lib.pb_poseidon2_permute(h, d, C.c_size_t(n), C.c_int(reps))
        lib.pb_ctx_synchronize(h)
        best = min(best, time.time() - t)
    print(os.path.basename(so), "%.3f ms  %.2f Gperm/s" % (best * 1e3, n * reps / best / 1e9), flush=True)
    lib.pb_device_free(d)
    lib.pb_ctx_destroy(h)
