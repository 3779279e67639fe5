"""The committed golden fixture (tests/golden/hotpath_small.npz) against the oracle that generated it and against the
numbers the reference's own tests hold (test/test_cutruncate.jl:9-17).  Runs on the CPU; guards the bar the GPU parity
tests are held to: if a change in oracle/ moves any stored result, this fails before the GPU suite can drift with it."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_fixture_is_what_the_oracle_produces_today():
    m = _gen()
    stored = dict(np.load(m.PATH))
    assert len(stored) >= 60
    assert m.compare(stored, m.build()) == []


def test_fixture_agrees_with_the_reference_held_numbers():
    g = np.load(_gen().PATH)
    # KAT 1 (zeros -> (0, 0, keep 1)) and KAT 2 (absolute cutoff): reference and CPU rule agree
    assert list(g["tr0"]) == [0.0, 0.0, 1.0]
    assert np.allclose(g["tr1"], g["ref_truncate_kat2"], atol=1e-15)
    # KAT 3: the reference's GPU rule keeps 1 / docut 0.45; the CPU rule (the parity target) keeps 2 / docut 0.25 with
    # the same truncation error -- stored side by side so the divergence stays visible
    assert g["tr2"][0] == g["ref_truncate_kat3_gpu_rule"][0] == 0.1
    assert list(g["tr2"][1:]) == [0.25, 2.0] and list(g["ref_truncate_kat3_gpu_rule"][1:]) == [0.45, 1.0]
    # maxdim-bound case of SURVEY 8 a15 (iii): CPU rule reports the DISCARDED weight 0.3
    assert np.allclose(g["tr3"], [0.3, 0.25, 2.0])
    # internal consistency of the stored bond step: Lanczos energy == bond-step energy; 3 matvecs (krylovdim 3, maxiter 1)
    assert g["b_lanczos_nmv"] == 3
    assert g["b_svdL_energy"] == g["b_lanczos_energy"] == g["b_eigR_energy"]
    assert g["b_svdL_keep"] == 12 and g["b_eigR_keep"] == 12
    assert abs(g["b_svdL_truncerr"] - g["b_eigR_truncerr"]) < 1e-8           # noise 1e-8 moves it at that order only
    assert np.all(np.diff(g["s_S"]) <= 0) and np.all(np.diff(g["s_D"]) <= 0)   # descending spectra
