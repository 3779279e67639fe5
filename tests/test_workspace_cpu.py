"""Workspace planning of the DMRG tier, checked WITHOUT a GPU through the host-only query entry points
(tnb_bond_workspace_bytes / tnb_matrix_workspace_bytes / tnb_set_workspace_limit with a null handle).

The fixed-workspace execution of config C5 (BASELINE.json configs[4]: chi = 8192, dense MPO bond ~30) rests on host
arithmetic: how many slabs of the output bond H_eff is cut into, and how large the arena gets.  These tests pin that
arithmetic to the numbers measured on the B200 (profiles/r02_c5_chi8192.json) and to SURVEY.md section 8(a3)/(a5).
Supersedes the reference's dead out-of-core attempt (/root/reference/src/tensor/dense.jl:50-193).
"""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GB = 1e9


def _lib():
    import __graft_entry__ as g
    g.build()
    from itensorsgpu_b200 import tn
    return tn._lib


@pytest.fixture()
def lib():
    L = _lib()
    L.set_workspace_limit(0)
    yield L
    L.set_workspace_limit(0)


def test_c3_heff_temporaries_are_two_times_2_68_gb(lib):
    """the temporaries of one H_eff*phi at chi = 4096, d = 2, w = 5 are 335.5 M elements each"""
    b, slabs = lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, 4096, 4096, 2, 2, 5, 5, 5)
    assert slabs == 1
    assert b == 2 * 4096 * 4096 * 4 * 5 * 8
    bc, _ = lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, 4096, 4096, 2, 2, 5, 5, 5, dtype=lib.C128)
    assert bc == 2 * b


def test_c5_runs_in_a_fixed_workspace_and_matches_the_measured_arena(lib):
    """chi = 8192, w = 30: 2 x 64 GB unchunked; under the default 40 GiB limit H_eff is cut into 4 slabs of l'.  The
    arena the B200 run ended with (profiles/r02_c5_chi8192.json: workspace_GB) is the bond-step requirement plus the
    allocator's 1/8 slack -- reproduced here to the byte."""
    dims = (8192, 8192, 2, 2, 30, 30, 30)
    assert lib.get_workspace_limit() == 40 << 30
    b, slabs = lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, *dims)
    assert slabs == 4
    assert b == 2 * 8192 * 8192 * 4 * 30 * 8 // 4 and b / GB < 33
    step, _ = lib.bond_workspace_bytes(lib.WS_DMRG_BOND_STEP, *dims, noise=False, krylovdim=3)
    prof = json.load(open(os.path.join(ROOT, "profiles", "r02_c5_chi8192.json")))
    assert prof["heff_launches"] == 4 * slabs
    arena = step + (step >> 3) + (1 << 20)               # ws_require's growth rule (api.cu)
    assert arena == int(round(prof["workspace_GB"] * GB))
    assert arena / GB < 60                                # VERDICT next-round item 8
    noisy, _ = lib.bond_workspace_bytes(lib.WS_DMRG_BOND_STEP, *dims, noise=True, krylovdim=3)
    assert noisy - step == 16384 * 16384 * 8              # the perturbation of rho, pinned at the arena front
    assert noisy <= arena                                 # ... fits the slack: the second rule did not regrow the arena


def test_workspace_limit_controls_the_slab_count(lib):
    dims = (8192, 8192, 2, 2, 30, 30, 30)
    seen = []
    for gib, want in ((200, 1), (100, 2), (40, 4), (20, 8), (10, 16)):
        lib.set_workspace_limit(gib << 30)
        b, slabs = lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, *dims)
        assert slabs == want and b <= (gib << 30)
        seen.append(b)
    assert seen == sorted(seen, reverse=True)
    lib.set_workspace_limit(1)                            # absurd limit: slabs stop at 32 columns of l' each
    _, slabs = lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, *dims)
    assert slabs == 8192 // 32
    lib.set_workspace_limit(0)
    assert lib.get_workspace_limit() == 40 << 30
    # an odd bond cannot be halved: one slab whatever the limit
    lib.set_workspace_limit(1 << 20)
    _, slabs = lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, 4097, 4096, 2, 2, 5, 5, 5)
    assert slabs == 1


def test_bond_step_covers_its_stages(lib):
    """the bond step pins phi, then needs max(Lanczos stage, factorize stage)"""
    for dims in ((64, 48, 2, 2, 5, 5, 5), (200, 200, 3, 3, 5, 5, 5), (4096, 4096, 2, 2, 5, 5, 5)):
        heff, _ = lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, *dims)
        fac, _ = lib.bond_workspace_bytes(lib.WS_FACTORIZE_BOND, *dims)
        step, _ = lib.bond_workspace_bytes(lib.WS_DMRG_BOND_STEP, *dims)
        phi = dims[0] * dims[1] * dims[2] * dims[3] * 8
        assert step >= phi + max(heff + 4 * phi, fac)
        assert step <= phi + max(heff + 4 * phi, fac) + (1 << 20)
        more, _ = lib.bond_workspace_bytes(lib.WS_DMRG_BOND_STEP, *dims, krylovdim=6)
        assert more >= step


def test_matrix_queries_and_bad_arguments(lib):
    svd = lib.matrix_workspace_bytes(lib.WS_SVD, 8192, 8192)
    eig = lib.matrix_workspace_bytes(lib.WS_EIGH, 8192, 8192)
    qr = lib.matrix_workspace_bytes(lib.WS_QR, 8192, 4096)
    assert svd > eig > 8192 * 8192 * 8 and qr > 8192 * 4096 * 8
    assert lib.matrix_workspace_bytes(lib.WS_EIGH, 8192, 8192, dtype=lib.C128) > eig
    # monotone in the size
    sizes = [lib.matrix_workspace_bytes(lib.WS_EIGH, n, n) for n in (64, 256, 1024, 4096, 16384)]
    assert sizes == sorted(sizes)
    # n = 16384 (C5: chi*d) is above the old divide-and-conquer size limit: must be a finite plan
    assert sizes[-1] / GB < 20
    with pytest.raises(lib.TnbError):
        lib.matrix_workspace_bytes(lib.WS_EIGH, 100, 99)
    with pytest.raises(lib.TnbError):
        lib.matrix_workspace_bytes(99, 100, 100)
    with pytest.raises(lib.TnbError):
        lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, 0, 4, 2, 2, 5, 5, 5)
    with pytest.raises(lib.TnbError):
        lib.bond_workspace_bytes(lib.WS_HEFF_APPLY, 4, 4, 2, 2, 5, 5, 5, dtype=9)


def test_shard_staging_buffer_matches_the_c5_multi_gpu_run(lib):
    """tnb_shard_stage_bytes: the peer-mapped staging buffer of the sharded environment updates / noise term.  At C5
    on 8 GPUs it is the full-size noise-term temporary, 64.4 GB per GPU -- the '64 GB of it the noise-term staging
    buffer' of BASELINE.md section 4 (profiles/r02_c5_chi8192_n8.json: 142 GB of HBM in use per GPU)."""
    L = lib.load()
    b = L.tnb_shard_stage_bytes(lib.F64, 8192, 2, 30, 8)
    assert b == 8192 * 8192 * 2 * 2 * 30 * 8 + 4096
    assert 64.0 < b / GB < 64.5
    # C3: the noise term dominates the two environment regions as well
    b3 = L.tnb_shard_stage_bytes(lib.F64, 4096, 2, 5, 8)
    assert b3 == 4096 * 4096 * 4 * 5 * 8 + 4096
    assert L.tnb_shard_stage_bytes(lib.C128, 4096, 2, 5, 2) == 2 * (b3 - 4096) + 4096
