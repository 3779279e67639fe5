"""itensorsgpu.jl_b200 -- B200-native (sm_100a) replacement for the ITensorsGPU.jl hot path.

Only what the path needs: ``csrc/`` (CUDA kernels + the C ABI, built into
``lib/libtnb200.so``), ``_lib`` (ctypes binding), ``ops`` (one host function per entry
point on flat device tensors) and ``itensor`` (the host-side mirror of the reference's
ITensor-level interface: ``cu``, ``cpu``, ``*``, ``+``, ``svd``, ``eigen``, ``qr``, ``dmrg``,
``apply``).  The directory name is not an importable identifier; import it through the
repo-root shim:  ``from itensorsgpu_b200 import tn``.
"""
from . import _lib, ops, itensor, mps, shard, tebd, io  # noqa: F401
from ._lib import TnbError, DimensionMismatch, handle, load  # noqa: F401
from .ops import DTensor  # noqa: F401
from .itensor import (Index, ITensor, cu, cpu, cuITensor, randomCuITensor, prime, dag, noprime, norm, dot, permute,  # noqa: F401
                      svd, eigen, qr, davidson, commonind, delta, diagITensor)
from .mps import (MPS, MPO, Sweeps, dmrg, apply, inner, orthogonalize, add, truncate, contract, contract_mpo, add_mpo, truncate_mpo, cuMPS, cuMPO, randomCuMPS, productCuMPS,  # noqa: F401
                  randomCuMPO, heisenberg_mpo, tfim_mpo)
from .io import save_chain, load_chain, save_itensor, load_itensor  # noqa: F401
