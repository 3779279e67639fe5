"""ctypes binding of libtnb200.so (the C ABI declared in include/tnb200.h).

The library is the product; this module only loads it and declares signatures.  There is
no CPU fallback: if the shared object is missing or no sm_100 device is present, every
entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtnb200.so")

F64, C128 = 0, 1
CONJ_A, CONJ_B = 1, 2
TRUNC_ABSOLUTE_CUTOFF, TRUNC_NO_RELATIVE = 1, 2
DECOMP_AUTO, DECOMP_SVD, DECOMP_EIGEN, DECOMP_QR = 0, 1, 2, 3
ORTHO_LEFT, ORTHO_RIGHT = 0, 1

STATUS = {0: "OK", 1: "BAD_ARG", 2: "DIM_MISMATCH", 3: "UNSUPPORTED", 4: "CUDA", 5: "NO_CONVERGENCE", 6: "ALLOC"}


class TnbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libtnb200: %s (%s)" % (msg, STATUS.get(code, code)))
        self.code = code


class DimensionMismatch(TnbError):
    """Mirrors Julia's DimensionMismatch (reference src/cuitensor.jl:55,72,78)."""


class BondDims(C.Structure):
    _fields_ = [("chiL", C.c_int64), ("chiR", C.c_int64), ("d1", C.c_int32), ("d2", C.c_int32),
                ("wL", C.c_int32), ("wM", C.c_int32), ("wR", C.c_int32)]


class PlanDesc(C.Structure):
    """tnb_plan_desc (include/tnb200.h): the planner's matricisation of one contraction."""
    _A = C.c_int64 * 12
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("n_m", C.c_int32), ("n_n", C.c_int32), ("n_k", C.c_int32), ("family", C.c_int32),
                ("ext_m", _A), ("a_stride_m", _A), ("c_stride_m", _A),
                ("ext_n", _A), ("b_stride_n", _A), ("c_stride_n", _A),
                ("ext_k", _A), ("a_stride_k", _A), ("b_stride_k", _A),
                ("a_k_major", C.c_int32), ("b_k_major", C.c_int32), ("a_vec", C.c_int32), ("b_vec", C.c_int32),
                ("tile_m", C.c_int32), ("tile_n", C.c_int32), ("tile_k", C.c_int32), ("herm_upper", C.c_int32),
                ("tiles", C.c_int64), ("waves", C.c_double)]


_vp, _i64, _i32, _int, _dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_double
_pi64, _pi32, _pdbl, _pint = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int)
_pbd = C.POINTER(BondDims)

# name -> (restype, argtypes); must list every symbol include/tnb200.h declares
SIGNATURES = {
    "tnb_create": (_int, [C.POINTER(_vp)]),
    "tnb_destroy": (_int, [_vp]),
    "tnb_last_error": (C.c_char_p, [_vp]),
    "tnb_version": (_int, []),
    "tnb_reserve": (_int, [_vp, C.c_size_t]),
    "tnb_workspace_bytes": (C.c_size_t, [_vp]),
    "tnb_launch_count": (C.c_uint64, [_vp]),
    "tnb_plan_describe": (_int, [_int, _int, _pi64, _pi32, _int, _pi64, _pi32, _int, _pi64, _pi32, _int, _int,
                                 C.POINTER(PlanDesc), C.c_char_p, C.c_size_t]),
    "tnb_plan_describe_strided": (_int, [_int, _int, _pi64, _pi32, _pi64, _int, _pi64, _pi32, _pi64, _int, _pi64, _pi32, _pi64,
                                         _int, _int, C.POINTER(PlanDesc), C.c_char_p, C.c_size_t]),
    "tnb_plan_cache_stats": (_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tnb_plan_cache_clear": (_int, [_vp]),
    "tnb_set_autotune": (_int, [_vp, _int]),
    "tnb_kernel_family_counts": (_int, [_vp, C.POINTER(C.c_uint64)]),
    "tnb_set_workspace_limit": (_int, [_vp, C.c_size_t]),
    "tnb_get_workspace_limit": (C.c_size_t, [_vp]),
    "tnb_bond_workspace_bytes": (C.c_size_t, [_int, _int, _pbd, _int, _int, _int, _pi32]),
    "tnb_matrix_workspace_bytes": (C.c_size_t, [_int, _int, _i64, _i64]),
    "tnb_contract": (_int, [_vp, _int, _int, _pi64, _pi32, _vp, _int, _pi64, _pi32, _vp, _int, _pi64, _pi32, _vp,
                            _vp, _vp, _int, _vp]),
    "tnb_permute_axpby": (_int, [_vp, _int, _int, _pi64, _pi32, _vp, _pi32, _vp, _vp, _vp, _vp]),
    "tnb_diag_contract": (_int, [_vp, _int, _int, _pi64, _pi32, _vp, _i32, _vp, _int, _pi32, _vp, _vp]),
    "tnb_scale": (_int, [_vp, _int, _i64, _vp, _vp, _vp]),
    "tnb_dot": (_int, [_vp, _int, _i64, _vp, _vp, _vp, _vp, _vp]),
    "tnb_nrm2": (_int, [_vp, _int, _i64, _vp, _vp, _pdbl, _vp]),
    "tnb_truncate": (_int, [_vp, _vp, _i64, _i64, _i64, _dbl, _int, _pi64, _pdbl, _pdbl, _vp]),
    "tnb_svd_trunc": (_int, [_vp, _int, _i64, _i64, _vp, _i64, _i64, _dbl, _int, _int, _vp, _vp, _vp, _pi64, _pdbl,
                             _vp]),
    "tnb_eigh_trunc": (_int, [_vp, _int, _i64, _vp, _i64, _i64, _dbl, _int, _int, _vp, _vp, _pi64, _pdbl, _vp]),
    "tnb_qr": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _vp]),
    "tnb_heff_apply": (_int, [_vp, _int, _pbd, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tnb_heff_apply_shard": (_int, [_vp, _int, _pbd, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tnb_heff_apply_shard_fused": (_int, [_vp, _int, _pbd, _int, _int, _i64, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp),
                                          C.POINTER(_vp), C.c_uint64, _vp]),
    "tnb_peer_alloc": (_int, [_vp, C.c_size_t, C.POINTER(_vp), C.c_char_p]),
    "tnb_peer_open": (_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "tnb_peer_close": (_int, [_vp, _vp]),
    "tnb_peer_free": (_int, [_vp, _vp]),
    "tnb_peer_status": (_int, [_vp, _vp]),
    "tnb_comm_init": (_int, [_vp, _int, _int, C.POINTER(_vp)]),
    "tnb_comm_finalize": (_int, [_vp]),
    "tnb_comm_barrier": (_int, [_vp, _vp]),
    "tnb_comm_allgather": (_int, [_vp, C.POINTER(_vp), C.c_size_t, C.c_size_t, _vp]),
    "tnb_shard_stage_bytes": (C.c_size_t, [_int, _i64, _i32, _i32, _int]),
    "tnb_env_update_left_shard": (_int, [_vp, _int, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, C.POINTER(_vp), _vp, _vp]),
    "tnb_env_update_right_shard": (_int, [_vp, _int, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, C.POINTER(_vp), _vp, _vp]),
    "tnb_eigsolve_lanczos_shard": (_int, [_vp, _int, _pbd, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _int,
                                          _int, _dbl, _pdbl, _pint, _vp]),
    "tnb_dmrg_bond_step_shard": (_int, [_vp, _int, _pbd, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _i64, _i64, _dbl,
                                        _dbl, _int, _int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _pdbl, _pi64,
                                        _pdbl, _vp]),
    "tnb_heff_apply_shard_host": (_int, [_vp, _int, _pbd, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp]),
    "tnb_heff_apply_host": (_int, [_vp, _int, _pbd, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tnb_env_update_left": (_int, [_vp, _int, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "tnb_env_update_right": (_int, [_vp, _int, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "tnb_eigsolve_lanczos": (_int, [_vp, _int, _pbd, _vp, _vp, _vp, _vp, _vp, _int, _int, _dbl, _pdbl, _pint, _vp]),
    "tnb_noise_term": (_int, [_vp, _int, _pbd, _vp, _vp, _vp, _vp, _vp, _int, _dbl, _int, _vp, _vp]),
    "tnb_factorize_bond": (_int, [_vp, _int, _pbd, _vp, _int, _int, _i64, _i64, _dbl, _vp, _int, _vp, _vp, _pi64,
                                  _pdbl, _vp]),
    "tnb_dmrg_bond_step": (_int, [_vp, _int, _pbd, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _i64, _i64, _dbl, _dbl,
                                  _int, _int, _pdbl, _pi64, _pdbl, _vp]),
    "tnb_dmrg_sweep": (_int, [_vp, _int, _i32, _pi64, _pi32, _pi32, C.POINTER(_vp), _pi64, C.POINTER(_vp), C.POINTER(_vp), _pi64,
                              _int, _i64, _i64, _dbl, _dbl, _int, _int, _int, _pdbl, _pdbl, _pdbl, _pdbl, _vp]),
    "tnb_tebd_apply_gate": (_int, [_vp, _int, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _dbl, _pi64,
                                   _pdbl, _vp]),
    "tnb_tebd_gate_bform": (_int, [_vp, _int, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _vp,
                                   _pi64, _pdbl, _vp]),
}

_lib = None


def load():
    """dlopen the shared library and attach signatures.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libtnb200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(expected at %s).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


FAMILIES = {0: "ldgsts", 1: "smallk", 2: "tma"}
WS_HEFF_APPLY, WS_FACTORIZE_BOND, WS_DMRG_BOND_STEP, WS_SVD, WS_EIGH, WS_QR = range(6)


def bond_workspace_bytes(op, chiL, chiR, d1, d2, wL, wM, wR, dtype=F64, ortho=ORTHO_LEFT, noise=False, krylovdim=3):
    """(arena bytes, H_eff slabs) an entry point needs for one bond -- host arithmetic, no GPU (tnb_bond_workspace_bytes)"""
    bd = BondDims(chiL, chiR, d1, d2, wL, wM, wR)
    ns = C.c_int32(0)
    b = load().tnb_bond_workspace_bytes(op, dtype, C.byref(bd), ortho, 1 if noise else 0, krylovdim, C.byref(ns))
    if b == 0:
        raise TnbError(1, "bond_workspace_bytes: bad argument")
    return int(b), int(ns.value)


def matrix_workspace_bytes(op, m, n, dtype=F64):
    b = load().tnb_matrix_workspace_bytes(op, dtype, m, n)
    if b == 0:
        raise TnbError(1, "matrix_workspace_bytes: bad argument")
    return int(b)


def set_workspace_limit(nbytes):
    """process-wide bound on the pair of H_eff temporaries (0 = default 40 GB); needs no handle"""
    load().tnb_set_workspace_limit(None, int(nbytes))


def get_workspace_limit():
    return int(load().tnb_get_workspace_limit(None))


def plan_describe(dims_a, modes_a, dims_b, modes_b, modes_c, dtype=F64, flags=0, num_sms=148, strides_a=None,
                  strides_b=None, strides_c=None):
    """Dry run of tnb_contract's planner (no GPU, no handle): a dict with the matricised extents M/N/K, the merged
    mode groups with their strides in A, B and C, and the kernel family / tile the launch would use.  modes_* are
    hashable labels (ints or strings); the extents of C follow from the operands.  strides_* (optional, elements per
    mode) describe operands that are strided windows of larger tensors."""
    lib = load()
    labels = {}
    lab = lambda m: labels.setdefault(m, len(labels))
    ma = [lab(m) for m in modes_a]; mb = [lab(m) for m in modes_b]
    ext = dict(zip(modes_a, dims_a)); ext.update(zip(modes_b, dims_b))
    mc = [lab(m) for m in modes_c]
    dims_c = [ext.get(m, 0) for m in modes_c]
    arr = lambda T, v: (T * max(len(v), 1))(*v)
    d = PlanDesc()
    err = C.create_string_buffer(512)
    st = lambda v: arr(C.c_int64, list(v)) if v is not None else None
    rc = lib.tnb_plan_describe_strided(dtype, len(ma), arr(C.c_int64, list(dims_a)), arr(C.c_int32, ma), st(strides_a),
                                       len(mb), arr(C.c_int64, list(dims_b)), arr(C.c_int32, mb), st(strides_b),
                                       len(mc), arr(C.c_int64, dims_c), arr(C.c_int32, mc), st(strides_c), flags, num_sms,
                                       C.byref(d), err, 512)
    if rc != 0:
        raise (DimensionMismatch if rc == 2 else TnbError)(rc, err.value.decode())
    g = lambda name, n: [int(x) for x in getattr(d, name)[:n]]
    return {"M": int(d.M), "N": int(d.N), "K": int(d.K), "family": FAMILIES[d.family],
            "m": {"ext": g("ext_m", d.n_m), "a": g("a_stride_m", d.n_m), "c": g("c_stride_m", d.n_m)},
            "n": {"ext": g("ext_n", d.n_n), "b": g("b_stride_n", d.n_n), "c": g("c_stride_n", d.n_n)},
            "k": {"ext": g("ext_k", d.n_k), "a": g("a_stride_k", d.n_k), "b": g("b_stride_k", d.n_k)},
            "a_k_major": bool(d.a_k_major), "b_k_major": bool(d.b_k_major), "a_vec": int(d.a_vec), "b_vec": int(d.b_vec),
            "tile": (int(d.tile_m), int(d.tile_n), int(d.tile_k)), "herm_upper": bool(d.herm_upper),
            "tiles": int(d.tiles), "waves": float(d.waves)}


class Handle:
    """Owns one tnb_handle_t bound to the current CUDA device."""

    def __init__(self):
        self.lib = load()
        h = _vp()
        rc = self.lib.tnb_create(C.byref(h))
        if rc != 0:
            raise TnbError(rc, "tnb_create failed: a B200 (sm_100) device is required; no CPU fallback")
        self.h = h

    def check(self, rc):
        if rc != 0:
            msg = self.lib.tnb_last_error(self.h).decode()
            raise (DimensionMismatch if rc == 2 else TnbError)(rc, msg)

    def close(self):
        if self.h:
            self.lib.tnb_destroy(self.h)
            self.h = None

    @property
    def launches(self):
        return int(self.lib.tnb_launch_count(self.h))

    def plan_cache_stats(self):
        """{'entries', 'hits', 'misses', 'autotuned'} of the contraction plan cache (``ContractionPlans`` analogue)"""
        v = [C.c_uint64(0) for _ in range(4)]
        self.check(self.lib.tnb_plan_cache_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("entries", "hits", "misses", "autotuned"), [int(x.value) for x in v]))

    def kernel_family_counts(self):
        """contraction calls per kernel family: {'ldgsts', 'smallk', 'tma', 'tma_split_k'}"""
        v = (C.c_uint64 * 4)()
        self.check(self.lib.tnb_kernel_family_counts(self.h, v))
        return {"ldgsts": int(v[0]), "smallk": int(v[1]), "tma": int(v[2]), "tma_split_k": int(v[3])}

    def plan_cache_clear(self):
        self.check(self.lib.tnb_plan_cache_clear(self.h))

    def set_autotune(self, on):
        self.check(self.lib.tnb_set_autotune(self.h, 1 if on else 0))

    def set_workspace_limit(self, nbytes):
        """bound on the pair of temporaries of one matvec / noise term / environment update (0 = default 40 GB)"""
        self.check(self.lib.tnb_set_workspace_limit(self.h, int(nbytes)))

    @property
    def workspace_limit(self):
        return int(self.lib.tnb_get_workspace_limit(self.h))

    @property
    def workspace_bytes(self):
        return int(self.lib.tnb_workspace_bytes(self.h))


_handles = {}


def handle():
    """The library handle of the CURRENT CUDA device (one per device: its workspace arena, scratch scalars and plan
    cache belong to that device; tnb_create binds to the device that is current when it is called).  Calls on one
    handle must be issued from one stream at a time -- the arena is reused in stream order (include/tnb200.h)."""
    import torch
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    h = _handles.get(dev)
    if h is None:
        h = _handles[dev] = Handle()
    return h
