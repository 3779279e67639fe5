"""Oracle: spectrum truncation.  TEST INFRASTRUCTURE.

Two functions:

* ``truncate``  -- the **CPU rule** ([EXT] NDTensors ``truncate!``): walk from the
  smallest weight, discard while the running *sum* of discarded weights stays within
  ``cutoff*scale`` (or each weight <= cutoff for absolute cutoff).  This is the parity
  target (north star: "match the reference ITensors.jl CPU path").
* ``truncate_gpu_reference`` -- a line-by-line transliteration of the reference's GPU
  ``truncate!`` (``src/tensor/cutruncate.jl:1-93``), kept only to pin the oracle
  against the reference's three known-answer vectors (``test/test_cutruncate.jl:9-17``).
  Its semantics diverge from the CPU rule.

Both return ``(truncerr, docut, n_keep)``; weights must be sorted descending.
"""
import numpy as np


def truncate(P, maxdim=None, mindim=1, cutoff=0.0, use_absolute_cutoff=False,
             use_relative_cutoff=True):
    """[EXT] NDTensors ``truncate!(P::Vector{Float64})`` -- CPU rule.

    The return triple and the ``docut`` rule are identical to the reference's GPU
    version (``src/tensor/cutruncate.jl:78-92``).
    """
    P = np.array(P, dtype=np.float64)
    origm = len(P)
    maxdim = origm if maxdim is None else min(int(maxdim), origm)
    mindim = max(min(int(mindim), maxdim), 1)
    cutoff = max(float(cutoff), 0.0)
    if origm == 0:
        return 0.0, 0.0, 0
    if P[0] <= 0.0:
        return 0.0, 0.0, 1
    if origm == 1:
        return 0.0, float(P[0] / 2), 1
    # zero out trailing negative weight
    for n in range(origm - 1, -1, -1):
        if P[n] >= 0.0:
            break
        P[n] = 0.0
    n = origm
    truncerr = 0.0
    while n > maxdim:
        truncerr += P[n - 1]
        n -= 1
    if use_absolute_cutoff:
        while n > mindim and P[n - 1] <= cutoff:
            truncerr += P[n - 1]
            n -= 1
    else:
        scale = 1.0
        if use_relative_cutoff:
            scale = float(np.sum(P))
            if scale == 0.0:
                scale = 1.0
        while n > mindim and (truncerr + P[n - 1] <= cutoff * scale):
            truncerr += P[n - 1]
            n -= 1
        truncerr /= scale
    if n < 1:
        n = 1
    docut = 0.0
    if n < origm:
        docut = (P[n - 1] + P[n]) / 2
        if abs(P[n - 1] - P[n]) < 1e-3 * P[n - 1]:
            docut += 1e-3 * P[n - 1]
    return float(truncerr), float(docut), int(n)


def _first_negative_index(err):
    """cutruncate.jl:35-38 / 53-56: err./abs(err) gives +-1; a sign-bit flag turns that into
    -2.0 for negative entries, 0 otherwise; ``iamax`` returns the FIRST index of maximum
    magnitude (1-based), i.e. the first below-threshold entry, or 1 if there is none."""
    with np.errstate(divide="ignore", invalid="ignore"):
        s = err / np.abs(err)
    flags = np.where(np.signbit(s), 2.0, 0.0)
    v = np.abs(s * flags)
    v = np.where(np.isnan(v), 0.0, v)
    return int(np.argmax(v)) + 1


def truncate_gpu_reference(P, maxdim=None, mindim=1, cutoff=0.0, absoluteCutoff=False,
                           doRelCutoff=True):
    """Transliteration of ``src/tensor/cutruncate.jl:1-93`` (line numbers in comments)."""
    P = np.array(P, dtype=np.float64)
    origm = len(P)
    maxdim = origm if maxdim is None else min(int(maxdim), origm)          # :3
    mindim = min(int(mindim), maxdim)                                       # :4
    docut = 0.0
    maxP = float(np.max(P))                                                 # :10
    if maxP == 0.0:                                                         # :11-14
        return 0.0, 0.0, 1
    if origm == 1:                                                          # :15-18
        return 0.0, maxP / 2, 1
    rP = np.where(np.signbit(P), 0.0, P)                                    # :23
    n = origm
    truncerr = 0.0
    if n > maxdim:                                                          # :26-29
        truncerr = float(np.sum(rP[: n - maxdim]))
        n = maxdim
    if absoluteCutoff:                                                      # :32-41
        cut_ind = _first_negative_index(rP - cutoff) - 1
        n = min(maxdim, origm - cut_ind)
        n = max(n, mindim)
        truncerr += float(np.sum(rP[cut_ind:]))
    else:
        scale = 1.0
        if doRelCutoff:                                                     # :45-48
            scale = float(np.sum(P))
            scale = scale if scale > 0.0 else 1.0
        cut_ind = _first_negative_index(rP + truncerr - cutoff * scale) - 1  # :53-56
        if cut_ind > 0:                                                     # :57-65
            truncerr += float(np.sum(rP[cut_ind:]))
        else:                                                               # :66-75
            truncerr += float(np.sum(rP[:maxdim]))
        n = min(maxdim, origm - cut_ind)
        n = max(n, mindim)
        truncerr = 0.0 if scale == 0.0 else truncerr / scale
    if n < 1:                                                               # :78-80
        n = 1
    if n < origm:                                                           # :81-87
        docut = (P[n - 1] + P[n]) / 2
        if abs(P[n - 1] - P[n]) < 1e-3 * P[n - 1]:
            docut += 1e-3 * P[n - 1]
    return float(truncerr), float(docut), int(n)
