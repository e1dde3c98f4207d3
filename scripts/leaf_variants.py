"""time pb_merkle_commit (leaf hashing dominated) for each build variant"""
import ctypes as C, glob, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for so in sorted(glob.glob(os.path.join(ROOT, "powdr_b200/_lib/variants/*.so"))):
    lib = C.CDLL(so)
    h = C.c_void_p()
    assert lib.pb_ctx_create(C.byref(h), 0, None) == 0
    log_h, w = 21, 256
    n = 1 << log_h
    d, lay = C.c_void_p(), C.c_void_p()
    lib.pb_device_alloc(C.byref(d), C.c_size_t(4 * n * w))
    lib.pb_device_alloc(C.byref(lay), C.c_size_t(64 * n))
    lib.pb_memset_zero(h, d, C.c_size_t(4 * n * w))
    mats = (C.c_void_p * 1)(d.value)
    ws = (C.c_size_t * 1)(w)
    best = 1e9
    for it in range(3):
        lib.pb_ctx_synchronize(h)
        t = time.time()
        lib.pb_merkle_commit(h, mats, ws, C.c_size_t(1), C.c_size_t(log_h), lay, None)
        lib.pb_ctx_synchronize(h)
        best = min(best, time.time() - t)
    perms = n * (w // 8) + n
    print(os.path.basename(so), "%.2f ms  %.3f Gperm/s" % (best * 1e3, perms / best / 1e9), flush=True)
    lib.pb_device_free(d); lib.pb_device_free(lay); lib.pb_ctx_destroy(h)
