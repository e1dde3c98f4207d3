"""ctypes binding of include/powdr_b200.h -- the Python stand-in for the Rust `extern "C"` block + safe wrappers of
/root/reference/openvm/src/cuda_abi.rs:8-64,97-135,174-223.  Raises if the CUDA library is missing: no CPU fallback."""
import ctypes as C
import os

import numpy as np

P = 2013265921
R_MOD_P = (1 << 32) % P
R_INV_MOD_P = pow(1 << 32, -1, P)
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libpowdr_b200.so")

EXPORTS = [
    "pb_ctx_create", "pb_ctx_destroy", "pb_ctx_synchronize", "pb_ctx_set_poseidon2", "pb_host_alloc", "pb_host_free",
    "pb_device_alloc", "pb_device_free", "pb_copy_h2d", "pb_copy_d2h", "pb_memset_zero", "pb_to_monty", "pb_from_monty",
    "pb_lde_batch", "pb_air_compile", "pb_air_free", "pb_air_is_jit", "pb_air_jit_compile_only", "pb_quotient", "pb_constraint_fold", "pb_merkle_commit",
    "pb_merkle_commit_rows8", "pb_poseidon2_permute", "pb_fri_fold", "pb_eval_at_point", "pb_deep_quotient", "pb_prove_segment", "pb_query_words", "pb_query_segment", "pb_last_openings", "pb_last_stage_ms",
    "pb_ctx_set_fri_params", "pb_air_set_interactions", "pb_air_perm_width", "pb_air_logup_compile_only", "pb_allgather_caps", "pb_bus_compile", "pb_bus_free", "pb_bus_apply",
    "pb_shard_columns", "pb_lde_shard", "pb_prove_segment_sharded", "pb_query_segment_sharded", "pb_prove_chips", "pb_chips_sizes", "pb_query_chips", "pb_host_poseidon2_permute",
    "pb_launch_count", "pb_leaf_kernel_profile", "_apc_tracegen", "_apc_apply_derived_expr", "_apc_apply_bus",
]


class LibraryMissing(RuntimeError):
    pass


class PbError(RuntimeError):
    def __init__(self, code, where):
        super().__init__("%s failed with code %d (%s)" % (where, code, "cudaError_t" if code > 0 else "PB_ERR"))
        self.code = code


class Span(C.Structure):
    _fields_ = [("off", C.c_uint32), ("len", C.c_uint32)]


class OriginalAir(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("buffer", C.c_void_p), ("row_block_size", C.c_int)]


class Subst(C.Structure):
    _fields_ = [("air_index", C.c_int), ("col", C.c_int), ("row", C.c_int), ("apc_col", C.c_int)]


class DerivedExprSpec(C.Structure):
    _fields_ = [("col_base", C.c_uint64), ("span", Span)]


class DevInteraction(C.Structure):
    _fields_ = [("bus_id", C.c_uint32), ("num_args", C.c_uint32), ("args_index_off", C.c_uint32)]


class SegmentProof(C.Structure):
    """pb_segment_proof_t (include/powdr_b200.h)"""
    _fields_ = [("trace_root", C.c_uint32 * 8), ("logup_alpha", C.c_uint32 * 4), ("logup_beta", C.c_uint32 * 4),
                ("perm_root", C.c_uint32 * 8), ("cumulative_sum", C.c_uint32 * 4), ("alpha", C.c_uint32 * 4),
                ("quotient_root", C.c_uint32 * 8), ("zeta", C.c_uint32 * 4), ("gamma", C.c_uint32 * 4),
                ("n_fri_layers", C.c_uint32), ("fri_roots", (C.c_uint32 * 8) * 32), ("fri_betas", (C.c_uint32 * 4) * 32),
                ("final_poly", (C.c_uint32 * 4) * 8), ("final_len", C.c_uint32), ("pow_witness", C.c_uint32),
                ("pow_bits", C.c_uint32), ("n_queries", C.c_uint32), ("perm_width", C.c_uint32)]
    VEC = ("trace_root", "logup_alpha", "logup_beta", "perm_root", "cumulative_sum", "alpha", "quotient_root", "zeta", "gamma")
    SCALAR = ("n_fri_layers", "final_len", "pow_witness", "pow_bits", "n_queries", "perm_width")

    def as_dict(self):
        n = self.n_fri_layers
        d = {k: list(getattr(self, k)) for k in self.VEC}
        d.update({k: int(getattr(self, k)) for k in self.SCALAR})
        d["fri_roots"] = [list(self.fri_roots[i]) for i in range(n)]
        d["fri_betas"] = [list(self.fri_betas[i]) for i in range(n)]
        d["final_poly"] = [list(self.final_poly[i]) for i in range(self.final_len)]
        return d


class Chip(C.Structure):
    """pb_chip_t"""
    _fields_ = [("air", C.c_void_p), ("d_trace", C.c_void_p), ("log_n", C.c_size_t), ("width", C.c_size_t)]


class ChipsProof(C.Structure):
    """pb_chips_proof_t (include/powdr_b200.h)"""
    _fields_ = [("main_root", C.c_uint32 * 8), ("perm_root", C.c_uint32 * 8), ("quotient_root", C.c_uint32 * 8),
                ("logup_alpha", C.c_uint32 * 4), ("logup_beta", C.c_uint32 * 4), ("alpha", C.c_uint32 * 4), ("zeta", C.c_uint32 * 4),
                ("gamma", C.c_uint32 * 4), ("n_fri_layers", C.c_uint32), ("fri_roots", (C.c_uint32 * 8) * 32),
                ("fri_betas", (C.c_uint32 * 4) * 32), ("final_poly", (C.c_uint32 * 4) * 8), ("final_len", C.c_uint32),
                ("pow_witness", C.c_uint32), ("pow_bits", C.c_uint32), ("n_queries", C.c_uint32), ("n_chips", C.c_uint32), ("log_max", C.c_uint32)]
    VEC = ("main_root", "perm_root", "quotient_root", "logup_alpha", "logup_beta", "alpha", "zeta", "gamma")
    SCALAR = ("n_fri_layers", "final_len", "pow_witness", "pow_bits", "n_queries", "n_chips", "log_max")
    as_dict = SegmentProof.as_dict


STAGES = ("h2d", "lde", "merkle", "logup_gen", "logup_commit", "quotient", "qlde", "qmerkle", "open", "fri", "pow", "total")


_lib = None


def load_library(path=None):
    """dlopen the in-tree sm_100a library; never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise LibraryMissing("%s not built: run `python -m powdr_b200.build` (nvcc, sm_100a)" % path)
    lib = C.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise LibraryMissing("symbol %s missing from %s" % (name, path))
        getattr(lib, name).restype = C.c_int
    lib.pb_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


def _chk(code, where):
    if code != 0:
        raise PbError(code, where)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class DeviceBuffer:
    """Caller-owned device memory (DeviceBuffer<T> of openvm-cuda-common)."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        _chk(ctx.lib.pb_device_alloc(C.byref(p), C.c_size_t(max(4, self.nbytes))), "pb_device_alloc")
        self.ptr = p.value

    def free(self):
        if self.ptr:
            self.ctx.lib.pb_device_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def upload(self, arr):
        a = np.ascontiguousarray(arr)
        assert a.nbytes <= self.nbytes
        _chk(self.ctx.lib.pb_copy_h2d(self.ctx.h, C.c_void_p(self.ptr), a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes)), "pb_copy_h2d")
        self.ctx.synchronize()
        return self

    def download(self, shape, dtype=np.uint32):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        _chk(self.ctx.lib.pb_copy_d2h(self.ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), C.c_size_t(out.nbytes)), "pb_copy_d2h")
        return out

    def zero(self):
        _chk(self.ctx.lib.pb_memset_zero(self.ctx.h, C.c_void_p(self.ptr), C.c_size_t(self.nbytes)), "pb_memset_zero")
        return self


class Air:
    def __init__(self, ctx, bytecode, spans, width, bus=None):
        """bus: (interactions [(bus_id, num_args, args_index_off)], arg_spans [(off, len)], bytecode) as returned by
        powdr_b200.machine.compile_bus(machine, 1) -- attaches the AIR's bus interactions (LogUp phase of the prover)"""
        self.ctx, self.width, self.n_constraints = ctx, width, len(spans)
        bc = _u32(bytecode)
        sp = (Span * max(1, len(spans)))()
        for i, (o, l) in enumerate(spans):
            sp[i].off, sp[i].len = o, l
        h = C.c_void_p()
        _chk(ctx.lib.pb_air_compile(ctx.h, bc.ctypes.data_as(C.c_void_p), C.c_size_t(bc.size), sp, C.c_size_t(len(spans)),
                                    C.c_uint32(width), C.byref(h)), "pb_air_compile")
        self.h = h
        self.perm_width = 0
        if bus and len(bus[0]):
            ints, isp, ibc = bus
            ibc = _u32(ibc)
            spn = (Span * max(1, len(isp)))()
            for i, (o, l) in enumerate(isp):
                spn[i].off, spn[i].len = o, l
            di = (DevInteraction * len(ints))()
            for i, (b, n, o) in enumerate(ints):
                di[i].bus_id, di[i].num_args, di[i].args_index_off = b, n, o
            _chk(ctx.lib.pb_air_set_interactions(ctx.h, self.h, ibc.ctypes.data_as(C.c_void_p), C.c_size_t(ibc.size), spn, C.c_size_t(len(isp)),
                                                 di, C.c_size_t(len(ints))), "pb_air_set_interactions")
            wp = C.c_size_t()
            _chk(ctx.lib.pb_air_perm_width(self.h, C.byref(wp)), "pb_air_perm_width")
            self.perm_width = int(wp.value)

    @property
    def is_jit(self):
        return bool(self.ctx.lib.pb_air_is_jit(self.h))

    def free(self):
        if self.h:
            self.ctx.lib.pb_air_free(self.h)
            self.h = None


class Context:
    """pb_ctx_t: one per GPU / stream."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        _chk(self.lib.pb_ctx_create(C.byref(h), C.c_int(device), C.c_void_p(stream or 0)), "pb_ctx_create")
        self.h = h
        self.device = device
        self.n_queries = 100

    def close(self):
        if self.h:
            self.lib.pb_ctx_destroy(self.h)
            self.h = None

    def synchronize(self):
        _chk(self.lib.pb_ctx_synchronize(self.h), "pb_ctx_synchronize")

    # ---- memory ----
    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr, monty=True):
        """numpy (canonical) -> device buffer (Montgomery unless monty=False); conversion done on the host."""
        a = _u32(arr)
        if monty:
            a = ((a.astype(np.uint64) * np.uint64(R_MOD_P)) % np.uint64(P)).astype(np.uint32)
        return DeviceBuffer(self, a.nbytes).upload(a)

    def to_host(self, buf, shape, monty=True):
        raw = buf.download(shape)
        if monty:
            raw = ((raw.astype(np.uint64) * np.uint64(R_INV_MOD_P)) % np.uint64(P)).astype(np.uint32)
        return raw

    # ---- stages ----
    def air(self, bytecode, spans, width, bus=None):
        return Air(self, bytecode, spans, width, bus)

    def set_fri_params(self, n_queries, pow_bits):
        _chk(self.lib.pb_ctx_set_fri_params(self.h, C.c_uint32(n_queries), C.c_uint32(pow_bits)), "pb_ctx_set_fri_params")
        self.n_queries = n_queries

    def lde_batch(self, d_trace_ptr, log_n, width, d_lde_ptr, log_blowup=1, shift=31):
        _chk(self.lib.pb_lde_batch(self.h, C.c_void_p(d_trace_ptr), C.c_size_t(log_n), C.c_size_t(width), C.c_uint32(log_blowup),
                                   C.c_uint32(shift), C.c_void_p(d_lde_ptr)), "pb_lde_batch")

    def quotient(self, air, d_lde_ptr, log_n, alpha, d_q_ptr, log_blowup=1, shift=31):
        al = (C.c_uint32 * 4)(*[int(x) for x in alpha])
        _chk(self.lib.pb_quotient(self.h, air.h, C.c_void_p(d_lde_ptr), C.c_size_t(log_n), C.c_uint32(log_blowup), C.c_uint32(shift),
                                  al, C.c_void_p(d_q_ptr)), "pb_quotient")

    def constraint_fold(self, air, d_mat_ptr, height, alpha, d_out_ptr):
        al = (C.c_uint32 * 4)(*[int(x) for x in alpha])
        _chk(self.lib.pb_constraint_fold(self.h, air.h, C.c_void_p(d_mat_ptr), C.c_size_t(height), al, C.c_void_p(d_out_ptr)),
             "pb_constraint_fold")

    def merkle_commit(self, mat_ptrs, widths, log_h, d_layers_ptr, want_root=True):
        n = len(mat_ptrs)
        ptrs = (C.c_void_p * n)(*mat_ptrs)
        ws = (C.c_size_t * n)(*widths)
        root = (C.c_uint32 * 8)()
        _chk(self.lib.pb_merkle_commit(self.h, ptrs, ws, C.c_size_t(n), C.c_size_t(log_h), C.c_void_p(d_layers_ptr),
                                       root if want_root else None), "pb_merkle_commit")
        return list(root) if want_root else None

    def merkle_commit_rows8(self, d_rows_ptr, log_h, d_layers_ptr, want_root=True):
        root = (C.c_uint32 * 8)()
        _chk(self.lib.pb_merkle_commit_rows8(self.h, C.c_void_p(d_rows_ptr), C.c_size_t(log_h), C.c_void_p(d_layers_ptr),
                                             root if want_root else None), "pb_merkle_commit_rows8")
        return list(root) if want_root else None

    def poseidon2_permute(self, d_states_ptr, n, reps=1):
        _chk(self.lib.pb_poseidon2_permute(self.h, C.c_void_p(d_states_ptr), C.c_size_t(n), C.c_int(reps)), "pb_poseidon2_permute")

    def fri_fold(self, d_in_ptr, log_len, shift, beta, d_out_ptr):
        b = (C.c_uint32 * 4)(*[int(x) for x in beta])
        _chk(self.lib.pb_fri_fold(self.h, C.c_void_p(d_in_ptr), C.c_size_t(log_len), C.c_uint32(shift), b, C.c_void_p(d_out_ptr)),
             "pb_fri_fold")

    def eval_at_point(self, d_mat_ptr, log_n, width, shift, zeta, d_ys_ptr):
        z = (C.c_uint32 * 4)(*[int(x) for x in zeta])
        _chk(self.lib.pb_eval_at_point(self.h, C.c_void_p(d_mat_ptr), C.c_size_t(log_n), C.c_size_t(width), C.c_uint32(shift), z,
                                       C.c_void_p(d_ys_ptr)), "pb_eval_at_point")

    def deep_quotient(self, mat_ptrs, widths, log_m, shift, zeta, gamma, d_ys_ptr, d_out_ptr):
        n = len(mat_ptrs)
        ptrs = (C.c_void_p * n)(*mat_ptrs)
        ws = (C.c_size_t * n)(*widths)
        z = (C.c_uint32 * 4)(*[int(x) for x in zeta])
        g = (C.c_uint32 * 4)(*[int(x) for x in gamma])
        _chk(self.lib.pb_deep_quotient(self.h, ptrs, ws, C.c_size_t(n), C.c_size_t(log_m), C.c_uint32(shift), z, g, C.c_void_p(d_ys_ptr),
                                       C.c_void_p(d_out_ptr)), "pb_deep_quotient")

    def prove_segment(self, air, trace_ptr, log_n, width, on_device=False):
        proof = SegmentProof()
        _chk(self.lib.pb_prove_segment(self.h, air.h, C.c_void_p(trace_ptr), C.c_size_t(log_n), C.c_size_t(width),
                                       C.c_uint32(1 if on_device else 0), C.byref(proof)), "pb_prove_segment")
        return proof.as_dict()

    def lde_shard(self, d_trace_ptr, log_n, width, world, blk, d_out_ptr, shift=31):
        _chk(self.lib.pb_lde_shard(self.h, C.c_void_p(d_trace_ptr), C.c_size_t(log_n), C.c_size_t(width), C.c_uint32(shift), C.c_int(world),
                                   C.c_int(blk), C.c_void_p(d_out_ptr)), "pb_lde_shard")

    def prove_segment_sharded(self, air, trace_cols_ptr, log_n, width, comm, on_device=False):
        """one rank of a multi-GPU segment proof; comm: powdr_b200.sharded.Comm.  trace_cols_ptr: this rank's column block."""
        proof = SegmentProof()
        _chk(self.lib.pb_prove_segment_sharded(self.h, air.h, C.c_void_p(trace_cols_ptr or 0), C.c_size_t(log_n), C.c_size_t(width),
                                               C.c_uint32(1 if on_device else 0), C.byref(comm.c), C.byref(proof)), "pb_prove_segment_sharded")
        return proof.as_dict()

    def query_segment_sharded(self, comm, log_n, width, perm_width=0):
        """query openings of the last prove_segment_sharded (collective: every rank calls it, every rank gets the whole set)"""
        wpq = C.c_size_t()
        _chk(self.lib.pb_query_words(C.c_size_t(log_n), C.c_size_t(width), C.c_size_t(perm_width), C.byref(wpq)), "pb_query_words")
        out = np.empty((self.n_queries, wpq.value), dtype=np.uint32)
        _chk(self.lib.pb_query_segment_sharded(self.h, C.byref(comm.c), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size)), "pb_query_segment_sharded")
        return out

    def query_segment(self, log_n, width, perm_width=0):
        """-> (n_queries, words_per_query) uint32 array of openings for the last prove_segment (n_queries from set_fri_params),
        and the opened values (width + 2*perm_width + 8, 4)"""
        wpq = C.c_size_t()
        _chk(self.lib.pb_query_words(C.c_size_t(log_n), C.c_size_t(width), C.c_size_t(perm_width), C.byref(wpq)), "pb_query_words")
        out = np.empty((self.n_queries, wpq.value), dtype=np.uint32)
        _chk(self.lib.pb_query_segment(self.h, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size)), "pb_query_segment")
        ys = np.empty((width + 2 * perm_width + 8, 4), dtype=np.uint32)
        _chk(self.lib.pb_last_openings(self.h, ys.ctypes.data_as(C.c_void_p), C.c_size_t(ys.size)), "pb_last_openings")
        return out, ys

    def prove_chips(self, chips, want_queries=True):
        """chips: [(Air, d_trace_ptr, log_n, width)], device-resident Montgomery traces.  All chips of a segment under one transcript
        (pb_prove_chips).  -> (proof dict, cumulative sums (K, 4), opened values (n_opened, 4), queries (n_queries, words) or None)"""
        K = len(chips)
        arr = (Chip * K)(*[Chip(a.h.value, int(p), int(ln), int(w)) for a, p, ln, w in chips])
        proof = ChipsProof()
        cums = np.zeros((K, 4), dtype=np.uint32)
        _chk(self.lib.pb_prove_chips(self.h, arr, C.c_size_t(K), C.byref(proof), cums.ctypes.data_as(C.c_void_p)), "pb_prove_chips")
        n_open, wpq = C.c_size_t(), C.c_size_t()
        _chk(self.lib.pb_chips_sizes(arr, C.c_size_t(K), C.byref(n_open), C.byref(wpq)), "pb_chips_sizes")
        ys = np.empty((n_open.value, 4), dtype=np.uint32)
        q = np.empty((self.n_queries, wpq.value), dtype=np.uint32) if want_queries and self.n_queries else None
        _chk(self.lib.pb_query_chips(self.h, q.ctypes.data_as(C.c_void_p) if q is not None else None, C.c_size_t(q.size if q is not None else 0),
                                     ys.ctypes.data_as(C.c_void_p), C.c_size_t(ys.size)), "pb_query_chips")
        return proof.as_dict(), cums, ys, q

    def last_stage_ms(self):
        ms = (C.c_float * len(STAGES))()
        _chk(self.lib.pb_last_stage_ms(self.h, ms), "pb_last_stage_ms")
        return dict(zip(STAGES, [float(x) for x in ms]))

    def launch_count(self):
        return int(self.lib.pb_launch_count(self.h))

    def leaf_kernel_profile(self):
        n, ms, by = C.c_int(), C.c_double(), C.c_double()
        _chk(self.lib.pb_leaf_kernel_profile(self.h, C.byref(n), C.byref(ms), C.byref(by)), "pb_leaf_kernel_profile")
        return n.value, ms.value, by.value

    def host_alloc(self, nbytes):
        p = C.c_void_p()
        _chk(self.lib.pb_host_alloc(C.byref(p), C.c_size_t(nbytes)), "pb_host_alloc")
        return p.value

    def host_free(self, ptr):
        self.lib.pb_host_free(C.c_void_p(ptr))

    # ---- stage 0, generated periphery kernel ----
    def bus_compile(self, bus, width, var_bus=3, tuple2_bus=7, bitwise_bus=6):
        """bus = machine.compile_bus(machine, 1) -> opaque handle for bus_apply"""
        ints, isp, ibc = bus
        ibc = _u32(ibc)
        spn = (Span * max(1, len(isp)))()
        for i, (o, l) in enumerate(isp):
            spn[i].off, spn[i].len = o, l
        di = (DevInteraction * max(1, len(ints)))()
        for i, (b, n, o) in enumerate(ints):
            di[i].bus_id, di[i].num_args, di[i].args_index_off = b, n, o
        h = C.c_void_p()
        _chk(self.lib.pb_bus_compile(self.h, ibc.ctypes.data_as(C.c_void_p), C.c_size_t(ibc.size), spn, C.c_size_t(len(isp)), di, C.c_size_t(len(ints)),
                                     C.c_uint32(width), C.c_uint32(var_bus), C.c_uint32(tuple2_bus), C.c_uint32(bitwise_bus), C.byref(h)), "pb_bus_compile")
        return h

    def bus_apply(self, handle, d_trace_ptr, H, num_calls, d_var, var_bins, d_t2, sz0, sz1, d_bw):
        _chk(self.lib.pb_bus_apply(self.h, handle, C.c_void_p(d_trace_ptr), C.c_size_t(H), C.c_int(num_calls), C.c_void_p(d_var), C.c_size_t(var_bins),
                                   C.c_void_p(d_t2), C.c_uint32(sz0), C.c_uint32(sz1), C.c_void_p(d_bw)), "pb_bus_apply")

    def bus_free(self, handle):
        self.lib.pb_bus_free(handle)

    # ---- stage 0 (reference symbols) ----
    def apc_tracegen(self, d_out_ptr, H, d_airs_ptr, d_subs_ptr, n_subs, num_calls):
        _chk(self.lib._apc_tracegen(C.c_void_p(d_out_ptr), C.c_size_t(H), C.c_void_p(d_airs_ptr), C.c_void_p(d_subs_ptr),
                                    C.c_size_t(n_subs), C.c_int(num_calls)), "_apc_tracegen")

    def apc_apply_derived_expr(self, d_out_ptr, H, num_calls, d_specs_ptr, n_cols, d_bc_ptr):
        _chk(self.lib._apc_apply_derived_expr(C.c_void_p(d_out_ptr), C.c_size_t(H), C.c_int(num_calls), C.c_void_p(d_specs_ptr),
                                              C.c_size_t(n_cols), C.c_void_p(d_bc_ptr)), "_apc_apply_derived_expr")

    def apc_apply_bus(self, d_out_ptr, num_calls, d_bc_ptr, bc_len, d_ints_ptr, n_ints, d_spans_ptr, n_spans, var_bus, d_var, var_bins,
                      t2_bus, d_t2, sz0, sz1, bw_bus, d_bw):
        _chk(self.lib._apc_apply_bus(C.c_void_p(d_out_ptr), C.c_int(num_calls), C.c_void_p(d_bc_ptr), C.c_size_t(bc_len),
                                     C.c_void_p(d_ints_ptr), C.c_size_t(n_ints), C.c_void_p(d_spans_ptr), C.c_size_t(n_spans),
                                     C.c_uint32(var_bus), C.c_void_p(d_var), C.c_size_t(var_bins), C.c_uint32(t2_bus), C.c_void_p(d_t2),
                                     C.c_uint32(sz0), C.c_uint32(sz1), C.c_uint32(bw_bus), C.c_void_p(d_bw)), "_apc_apply_bus")
