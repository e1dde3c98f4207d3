"""time pb_poseidon2_permute for each experimental build variant (powdr_b200/_lib/variants/*.so)"""
import ctypes as C, glob, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for so in sorted(glob.glob(os.path.join(ROOT, "powdr_b200/_lib/variants/*.so"))):
    lib = C.CDLL(so)
    h = C.c_void_p()
    assert lib.pb_ctx_create(C.byref(h), 0, None) == 0
    # correctness first: Plonky3's default-permutation KAT (tests/test_oracle.py) through this build
    P = 2013265921
    st = np.array([(i << 32) % P for i in range(16)], dtype=np.uint32)
    ds = C.c_void_p()
    lib.pb_device_alloc(C.byref(ds), C.c_size_t(64))
    lib.pb_copy_h2d(h, ds, st.ctypes.data_as(C.c_void_p), C.c_size_t(64))
    lib.pb_poseidon2_permute(h, ds, C.c_size_t(1), C.c_int(1))
    out = np.zeros(16, dtype=np.uint32)
    lib.pb_copy_d2h(h, out.ctypes.data_as(C.c_void_p), ds, C.c_size_t(64))
    got = [int(x) * pow(1 << 32, -1, P) % P for x in out]
    ok = got[:4] == [1906786279, 1737026427, 1959749225, 700325316] and got[15] == 304856115
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
    print(os.path.basename(so), "%.3f ms  %.2f Gperm/s" % (best * 1e3, n * reps / best / 1e9), "KAT ok" if ok else "KAT MISMATCH", flush=True)
    lib.pb_device_free(d)
    lib.pb_ctx_destroy(h)
