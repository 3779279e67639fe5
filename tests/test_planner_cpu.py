"""Host logic of tnb_contract, driven WITHOUT a GPU through tnb_plan_describe.

The planner turns mode labels into a grouped GEMM: three mixed-radix mode groups (M, N, K) with the strides of every
merged mode in A, B and C.  These tests replay exactly the address arithmetic the kernels execute
(contract_kernel.cuh: element (m, k) of A lives at offM_A(m) + offK_A(k), ...) in NumPy on the flat column-major
buffers and compare the result with the oracle contraction -- so a wrong grouping, merge, stride or mode order is
caught on the CPU.  Mirrors the layout coverage of /root/reference/test/test_cucontract.jl:34-221 (every layout case,
the complete permutation matrices, the rank-14 case).
"""
import itertools

import numpy as np
import pytest

from oracle import tensor as ot


def _tn():
    import __graft_entry__ as g
    g.build()
    from itensorsgpu_b200 import tn
    return tn


def _offsets(ext, strides):
    """all mixed-radix offsets of one group, ext[0] fastest"""
    off = np.zeros(1, dtype=np.int64)
    for e, s in zip(ext, strides):      # the new digit is slower than everything before it
        off = (np.arange(e, dtype=np.int64)[:, None] * s + off[None, :]).reshape(-1)
    return off


def _replay(plan, a_flat, b_flat, c_len):
    """C_flat from the plan's address arithmetic (what the tile kernels compute, as one dense matrix product)"""
    om_a, om_c = _offsets(plan["m"]["ext"], plan["m"]["a"]), _offsets(plan["m"]["ext"], plan["m"]["c"])
    on_b, on_c = _offsets(plan["n"]["ext"], plan["n"]["b"]), _offsets(plan["n"]["ext"], plan["n"]["c"])
    ok_a, ok_b = _offsets(plan["k"]["ext"], plan["k"]["a"]), _offsets(plan["k"]["ext"], plan["k"]["b"])
    assert len(om_a) == plan["M"] and len(on_b) == plan["N"] and len(ok_a) == plan["K"]
    Am = a_flat[om_a[:, None] + ok_a[None, :]]
    Bm = b_flat[ok_b[:, None] + on_b[None, :]]
    dest = om_c[:, None] + on_c[None, :]
    assert len(np.unique(dest)) == dest.size == c_len, "every output element written exactly once"
    c = np.zeros(c_len, dtype=np.result_type(a_flat, b_flat))
    c[dest.reshape(-1)] = (Am @ Bm).reshape(-1)
    return c


def _check(tn, dims, la, lb, lc=None, cplx=False, seed=0):
    rng = np.random.default_rng(seed)
    mk = (lambda s: rng.standard_normal(s) + 1j * rng.standard_normal(s)) if cplx else rng.standard_normal
    A = mk(tuple(dims[l] for l in la)); B = mk(tuple(dims[l] for l in lb))
    want, lout = ot.contract(A, la, B, lb)
    if lc is None:
        lc = lout
    want = ot.permute(want, lout, lc) if len(lc) else want
    plan = tn._lib.plan_describe([dims[l] for l in la], la, [dims[l] for l in lb], lb, lc,
                                 dtype=tn._lib.C128 if cplx else tn._lib.F64)
    got = _replay(plan, ot.flat(A), ot.flat(B), int(np.prod([dims[l] for l in lc], dtype=np.int64)))
    assert ot.rel_err(got, ot.flat(want)) < 1e-13
    return plan


def test_all_mode_orders_rank3_times_rank3():
    """the complete 6 x 6 permutation matrix of test_cucontract.jl:170-196, output in the NDTensors order"""
    tn = _tn()
    dims = {"i": 3, "j": 4, "k": 5, "l": 6, "a": 7}
    for pa in itertools.permutations(("i", "j", "a")):
        for pb in itertools.permutations(("a", "k", "l")):
            _check(tn, dims, pa, pb)


def test_all_output_orders_and_two_contracted_modes():
    tn = _tn()
    dims = {"i": 3, "j": 4, "a": 5, "b": 2, "k": 6}
    la, lb = ("a", "i", "b", "j"), ("k", "b", "a")
    for lc in itertools.permutations(("i", "j", "k")):
        _check(tn, dims, la, lb, lc)
        _check(tn, dims, la, lb, lc, cplx=True)


def test_layout_cases_scalar_vector_matrix_outer():
    """test_cucontract.jl:34-169: rank-0 results, matrix*vector, outer products, extent-1 modes"""
    tn = _tn()
    dims = {"i": 4, "j": 5, "k": 1, "l": 3}
    p = _check(tn, dims, ("i", "j"), ("i", "j"))                      # full contraction -> scalar
    assert (p["M"], p["N"], p["K"]) == (1, 1, 20) and p["k"]["ext"] == [20]     # i and j fuse: adjacent in both
    p = _check(tn, dims, ("i", "j"), ("j", "i"))                      # ... not adjacent in B: stay apart
    assert p["K"] == 20 and len(p["k"]["ext"]) == 2
    p = _check(tn, dims, ("i",), ("j",))                              # outer product
    assert (p["M"], p["N"], p["K"]) == (4, 5, 1)
    _check(tn, dims, ("i", "j"), ("j",))                              # matrix * vector
    _check(tn, dims, ("j",), ("l", "j"))                              # vector * matrix
    p = _check(tn, dims, ("i", "k", "j"), ("k", "j", "l"))            # extent-1 contracted mode is dropped
    assert p["k"]["ext"] == [5]
    _check(tn, dims, ("k", "i"), ("l", "k"))                          # K = 1 overall
    _check(tn, dims, ("i", "j"), ("l",), lc=("l", "j", "i"))          # outer product with a permuted output


def test_rank14_extent2_goes_through_the_same_planner():
    """the reference needs a cuBLAS fallback above rank 12 (cudense.jl:170-236, test_cucontract.jl:197-221)"""
    tn = _tn()
    labs = ["m%d" % i for i in range(14)]
    dims = {l: 2 for l in labs}
    la = labs[:9]                                   # m0..m8
    lb = labs[5:]                                   # m5..m13: contracts m5..m8
    p = _check(tn, dims, la, lb)
    assert (p["M"], p["N"], p["K"]) == (32, 32, 16)
    assert p["m"]["ext"] == [32] and p["n"]["ext"] == [32] and p["k"]["ext"] == [16]    # contiguous runs fuse
    lb2 = list(reversed(lb))
    p = _check(tn, dims, la, lb2)                   # reversed B: the contracted modes cannot fuse (opposite orders in
    assert len(p["k"]["ext"]) == 4                  # A and B); B's free modes still do (C follows B's order)
    assert p["n"]["ext"] == [32]
    p = _check(tn, dims, la, lb2, lc=labs[:5] + labs[9:])      # ... unless the output wants them in A-style order
    assert len(p["n"]["ext"]) == 5


def test_heff_steps_at_the_benchmark_shape_map_to_the_documented_gemms():
    """step 1 is M = d^2 chi, N = chi w, K = chi; step 4 is K = chi w; both K-major, TMA-eligible,
    big tile; step 2 (K = w d) takes the streaming kernel.  Plans only -- nothing is allocated."""
    tn = _tn()
    chi, d, w = 4096, 2, 5
    P = tn._lib.plan_describe
    p1 = P((chi, d, d, chi), ("l", "s1", "s2", "r"), (chi, chi, w), ("l", "lp", "a"), ("s1", "s2", "r", "lp", "a"))
    assert (p1["M"], p1["N"], p1["K"]) == (d * d * chi, chi * w, chi)
    assert p1["family"] == "tma" and p1["a_k_major"] and p1["b_k_major"] and p1["tile"] == (64, 128, 16)
    assert p1["tiles"] == (d * d * chi // 64) * (chi * w // 128)
    assert 2.0 * p1["M"] * p1["N"] * p1["K"] == pytest.approx(2.749e12, rel=1e-3)          # DESIGN 4.1 / VERDICT
    p4 = P((chi, chi, d, d, w), ("r", "lp", "s1p", "s2p", "c"), (chi, chi, w), ("r", "rp", "c"), ("lp", "s1p", "s2p", "rp"))
    assert (p4["M"], p4["N"], p4["K"]) == (chi * d * d, chi, chi * w)
    assert p4["family"] == "tma" and len(p4["k"]["ext"]) == 2                               # K over two modes (r, c)
    p2 = P((d, d, chi, chi, w), ("s1", "s2", "r", "lp", "a"), (w, d, d, w), ("a", "s1", "s1p", "b"),
           ("s2", "r", "lp", "s1p", "b"))
    assert p2["K"] == w * d and p2["N"] == d * w and p2["family"] == "smallk"
    # 8-way shard of step 4: 1024 tiles on 296 CTA slots (VERDICT weak 7)
    p4s = P((chi, chi // 8, d, d, w), ("r", "lp", "s1p", "s2p", "c"), (chi, chi, w), ("r", "rp", "c"),
            ("lp", "s1p", "s2p", "rp"))
    assert p4s["tiles"] == 1024 and p4s["waves"] == pytest.approx(1024 / 296)


def test_small_problems_take_the_small_tile_and_complex_tiles_are_half_as_wide():
    tn = _tn()
    P = tn._lib.plan_describe
    p = P((256, 2, 256), ("l", "s", "r"), (256, 5, 256), ("l", "a", "lp"), ("s", "r", "a", "lp"))
    assert p["tile"] == (64, 64, 16)               # 512 x 1280 in 64x128 tiles = 80 < 148 SMs
    p = P((2048, 2, 2048), ("l", "s", "r"), (2048, 5, 2048), ("l", "a", "lp"), ("s", "r", "a", "lp"), dtype=tn._lib.C128)
    assert p["tile"] == (64, 64, 8) and p["a_vec"] == 1
    # an odd leading extent forbids 16-byte copies of f64 along that direction
    p = P((7, 9), ("i", "j"), (7, 11), ("i", "k"), ("j", "k"))
    assert p["a_vec"] == 1 and p["b_vec"] == 1 and p["family"] == "ldgsts"


def test_herm_upper_is_honoured_only_for_c_ascending_square_results():
    tn = _tn()
    P = tn._lib.plan_describe
    HU = 4
    p = P((512, 300), ("i", "k"), (512, 300), ("j", "k"), ("i", "j"), flags=HU | tn._lib.CONJ_B)
    assert p["herm_upper"] and p["M"] == p["N"] == 512
    p = P((512, 300), ("i", "k"), (512, 300), ("j", "k"), ("j", "i"), flags=HU)        # transposed output: ignored
    assert not p["herm_upper"]
    with pytest.raises(tn.TnbError):
        P((512, 300), ("i", "k"), (256, 300), ("j", "k"), ("i", "j"), flags=HU)


def test_planner_errors_match_the_reference_exceptions():
    """DimensionMismatch / ArgumentError analogues (src/cuitensor.jl:55,72,78; include/tnb200.h status codes)"""
    tn = _tn()
    P = tn._lib.plan_describe
    with pytest.raises(tn._lib.DimensionMismatch):
        lib = tn.load()
        import ctypes as C
        d = tn._lib.PlanDesc()
        err = C.create_string_buffer(256)
        rc = lib.tnb_plan_describe(0, 2, (C.c_int64 * 2)(3, 4), (C.c_int32 * 2)(0, 1), 2, (C.c_int64 * 2)(5, 6),
                                   (C.c_int32 * 2)(1, 2), 2, (C.c_int64 * 2)(3, 6), (C.c_int32 * 2)(0, 2), 0, 148,
                                   C.byref(d), err, 256)
        assert rc == 2 and b"extent" in err.value
        raise tn._lib.DimensionMismatch(rc, err.value.decode())
    with pytest.raises(tn.TnbError, match="repeated mode"):
        P((3, 3), ("i", "i"), (3,), ("j",), ("j",))
    with pytest.raises(tn.TnbError, match="only one tensor"):
        P((3, 4), ("i", "j"), (4,), ("j",), ())                      # i would be dropped from the output
    with pytest.raises(tn.TnbError, match="neither input"):
        P((3,), ("i",), (4,), ("j",), ("i", "j", "z"))
    with pytest.raises(tn.TnbError, match="batch mode"):
        P((3, 4), ("i", "j"), (3, 4), ("i", "j"), ("i",))
    with pytest.raises(tn.TnbError, match="dtype"):
        P((3,), ("i",), (3,), ("i",), (), dtype=7)


def test_random_contractions_property():
    """hypothesis-style sweep (seeded): random ranks, extents, label placements, output orders, real and complex"""
    tn = _tn()
    rng = np.random.default_rng(1234)
    for trial in range(150):
        nk, nm, nn = rng.integers(0, 4), rng.integers(0, 4), rng.integers(0, 4)
        labs_k = ["k%d" % i for i in range(nk)]; labs_m = ["m%d" % i for i in range(nm)]; labs_n = ["n%d" % i for i in range(nn)]
        dims = {l: int(rng.integers(1, 5)) for l in labs_k + labs_m + labs_n}
        la = list(rng.permutation(labs_k + labs_m)); lb = list(rng.permutation(labs_k + labs_n))
        lc = list(rng.permutation(labs_m + labs_n))
        _check(tn, dims, la, lb, lc, cplx=bool(trial & 1), seed=trial)


# ---------------------------------------------------------------------------------------------------------------------
# Strided windows: the fixed-workspace (tnb_set_workspace_limit) and multi-GPU entry points never copy a slab of L, R, A
# or of the output vector -- they hand the planner a base pointer inside the stored tensor plus per-mode strides
# (csrc/heff.cu: heff_core / heff_apply_any / env_update_impl / heff_shard_fused_host_tail).  The cases below replay
# exactly those call sites (same extents, mode labels, strides and base offsets) on the CPU.
# ---------------------------------------------------------------------------------------------------------------------
def _replay_window(plan, a_flat, b_flat, c_flat, conj_b=False):
    """like _replay, but C is a window of a larger buffer: writes land in place, everything else stays untouched"""
    om_a, om_c = _offsets(plan["m"]["ext"], plan["m"]["a"]), _offsets(plan["m"]["ext"], plan["m"]["c"])
    on_b, on_c = _offsets(plan["n"]["ext"], plan["n"]["b"]), _offsets(plan["n"]["ext"], plan["n"]["c"])
    ok_a, ok_b = _offsets(plan["k"]["ext"], plan["k"]["a"]), _offsets(plan["k"]["ext"], plan["k"]["b"])
    Bm = b_flat[ok_b[:, None] + on_b[None, :]]
    dest = (om_c[:, None] + on_c[None, :]).reshape(-1)
    assert len(np.unique(dest)) == dest.size and dest.min() >= 0 and dest.max() < c_flat.size
    c_flat[dest] = (a_flat[om_a[:, None] + ok_a[None, :]] @ (np.conj(Bm) if conj_b else Bm)).reshape(-1)
    return dest


@pytest.mark.parametrize("cplx", [False, True])
def test_heff_slab_of_the_output_bond_as_strided_windows(cplx):
    """heff_apply_any: slab j of l' -- step 1 reads L[:, l'_j, :] as a window of the stored L, step 4 writes
    out[l'_j, :, :, :] as a window of the full output vector.  Steps 2+3 are dense and done by the oracle here."""
    tn = _tn()
    from oracle import dmrg as od
    rng = np.random.default_rng(5)
    cl, cr, d, w, G = 12, 10, 2, 3, 3
    clp = cl // G
    mk = (lambda s: rng.standard_normal(s) + 1j * rng.standard_normal(s)) if cplx else rng.standard_normal
    L, R, W1, W2, phi = mk((cl, cl, w)), mk((cr, cr, w)), mk((w, d, d, w)), mk((w, d, d, w)), mk((cl, d, d, cr))
    want = od.heff_apply(L, W1, W2, R, phi)
    out = np.full(cl * d * d * cr, np.nan, dtype=want.dtype)
    dt = tn._lib.C128 if cplx else tn._lib.F64
    for j in range(G):
        # step 1: T1[s1,s2,r,l'_c,a] = phi[l,s1,s2,r] L[l,l'_c,a], L window: strides {1, cl, cl*lp_stored}, base j*clp*cl
        p1 = tn._lib.plan_describe((cl, d, d, cr), ("l", "s1", "s2", "r"), (cl, clp, w), ("l", "lp", "a"),
                                   ("s1", "s2", "r", "lp", "a"), dtype=dt, strides_b=(1, cl, cl * cl))
        t1 = _replay(p1, ot.flat(phi), ot.flat(L)[j * clp * cl:], d * d * cr * clp * w)
        T1 = ot.unflat(t1, (d, d, cr, clp, w))
        ref1, _ = ot.contract(phi, ("l", "s1", "s2", "r"), L[:, j * clp:(j + 1) * clp, :], ("l", "lp", "a"))
        assert ot.rel_err(T1, ref1) < 1e-13
        # steps 2, 3 (dense): -> T3[r, l'_c, s1', s2', c]
        T2, _ = ot.contract(T1, ("s1", "s2", "r", "lp", "a"), W1, ("a", "s1", "s1p", "b"))       # (s2,r,lp,s1p,b)
        T3, l3 = ot.contract(T2, ("s2", "r", "lp", "s1p", "b"), W2, ("b", "s2", "s2p", "c"))     # (r,lp,s1p,s2p,c)
        assert l3 == ("r", "lp", "s1p", "s2p", "c")
        # step 4: out[l'_c,s1',s2',r'] = T3 R[r,r',c], C window: strides {1, cl, cl*d, cl*d*d}, base j*clp
        p4 = tn._lib.plan_describe((cr, clp, d, d, w), ("r", "lp", "s1p", "s2p", "c"), (cr, cr, w), ("r", "rp", "c"),
                                   ("lp", "s1p", "s2p", "rp"), dtype=dt, strides_c=(1, cl, cl * d, cl * d * d))
        assert p4["m"]["c"][0] == 1 and p4["n"]["c"] == [cl * d * d]
        _replay_window(p4, ot.flat(T3), ot.flat(R), out[j * clp:])
    assert not np.any(np.isnan(out))                         # the G windows tile the output vector exactly
    assert ot.rel_err(ot.unflat(out, (cl, d, d, cr)), want) < 1e-12


def test_environment_update_chunks_and_host_tail_windows():
    """env_update_impl (left): chunk j of the bra bond l' -- E window {1, cl, cl*cl} at j*cc*cl in the first contraction,
    conj(A) window {1, cl, cl*d} at j*cc in the last one, accumulated over chunks; heff_shard_fused_host_tail: R window
    over r' and an output window strided in both l' (rank slab) and r' (chunk)."""
    tn = _tn()
    from oracle import dmrg as od
    rng = np.random.default_rng(6)
    cl, cr, d, wl, wr, G = 8, 6, 2, 3, 2, 2
    cc = cl // G
    E, A, W = rng.standard_normal((cl, cl, wl)), rng.standard_normal((cl, d, cr)), rng.standard_normal((wl, d, d, wr))
    want = od.env_left_update(E, A, W)
    acc = np.zeros(cr * cr * wr)
    for j in range(G):
        p1 = tn._lib.plan_describe((cl, cc, wl), ("l", "lp", "a"), (cl, d, cr), ("l", "s", "r"), ("lp", "a", "s", "r"),
                                   strides_a=(1, cl, cl * cl))
        T1 = ot.unflat(_replay(p1, ot.flat(E)[j * cc * cl:], ot.flat(A), cc * wl * d * cr), (cc, wl, d, cr))
        T2, _ = ot.contract(T1, ("lp", "a", "s", "r"), W, ("a", "s", "sp", "b"))
        T2 = ot.permute(T2, ("lp", "r", "sp", "b"), ("lp", "r", "sp", "b"))
        p3 = tn._lib.plan_describe((cc, cr, d, wr), ("lp", "r", "sp", "b"), (cc, d, cr), ("lp", "sp", "rp"), ("r", "rp", "b"),
                                   strides_b=(1, cl, cl * d), flags=tn._lib.CONJ_B)
        part = np.zeros(cr * cr * wr)
        _replay_window(p3, ot.flat(T2), ot.flat(A)[j * cc:], part, conj_b=True)
        acc += part                                           # beta = 1 accumulation over the chunks
    assert ot.rel_err(ot.unflat(acc, (cr, cr, wr)), want) < 1e-12

    # host tail: rank `rank` of `world` owns l' slab [rank*clp, +clp); step 4 is cut over r' into pieces [q0, q1)
    cl, cr, d, w, world, rank = 8, 9, 2, 3, 2, 1
    clp = cl // world
    T3, R = rng.standard_normal((cr, clp, d, d, w)), rng.standard_normal((cr, cr, w))
    ref, _ = ot.contract(T3, ("r", "lp", "s1p", "s2p", "c"), R, ("r", "rp", "c"))
    full = np.full(cl * d * d * cr, np.nan)
    col = cl * d * d
    for q0, q1 in ((0, 3), (3, 6), (6, 9)):
        p = tn._lib.plan_describe((cr, clp, d, d, w), ("r", "lp", "s1p", "s2p", "c"), (cr, q1 - q0, w), ("r", "rp", "c"),
                                  ("lp", "s1p", "s2p", "rp"), strides_b=(1, cr, cr * cr), strides_c=(1, cl, cl * d, cl * d * d))
        _replay_window(p, ot.flat(T3), ot.flat(R)[q0 * cr:], full[rank * clp + q0 * col:])
    got = ot.unflat(full, (cl, d, d, cr))
    assert ot.rel_err(got[rank * clp:(rank + 1) * clp], ref) < 1e-13
    assert np.all(np.isnan(got[:rank * clp]))                 # the other rank's slab is not touched
