"""On-disk format for MPS / MPO / ITensors of the hot path: checkpoint a long sweep with
``cpu(psi)`` -> disk and resume with disk -> ``cu(psi)``.

The reference lists HDF5 in ``Project.toml:12`` but never calls it; [EXT] ITensors.jl writes an MPS as an HDF5 group
``{type="MPS", version, length, llim, rlim, "MPS[n]" -> ITensor{inds -> IndexSet{index_n -> Index{id, dim, plev, tags}},
storage -> Dense{data}}}``.  No HDF5 library exists in this image (h5py / libhdf5 absent, no network), so the SAME
logical schema is written into a NumPy ``.npz`` container (a zip of ``.npy`` members, readable from Julia with NPZ.jl):

  ``meta.json``            {"format": "itensorsgpu_b200", "version": 1, "type": "MPS" | "MPO" | "ITensor", "length": N,
                            "llim", "rlim", "sites": [{"dims": [...], "inds": [{"id","dim","plev","tags"}...] | null,
                            "storage": "Dense" | "Diag", "eltype": "Float64" | "ComplexF64"}], "extra": {...}}
  ``site_<n>``             flat COLUMN-MAJOR data vector of site n -- byte-identical to the ``CuVector`` inside
                           ``Dense`` (``src/tensor/cudense.jl:252-254``) and to what crosses the C ABI

so a converter to / from ITensors' HDF5 files is a loop over groups with no data reshuffling.  Files are written to a
temporary name and renamed, so an interrupted checkpoint never leaves a truncated file behind.
"""
import io as _io
import json
import os
import zipfile

import numpy as np

from . import _lib
from .ops import DTensor

FORMAT, VERSION = "itensorsgpu_b200", 1


def _flat_host(t):
    if isinstance(t, DTensor):
        return t.data.cpu().numpy(), t.dims
    a = np.asarray(t)
    return np.ascontiguousarray(a.ravel(order="F")), a.shape


def _eltype(v):
    return "ComplexF64" if np.iscomplexobj(v) else "Float64"


def _write(path, meta, arrays):
    tmp = path + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_STORED) as z:
        z.writestr("meta.json", json.dumps(meta))
        for name, a in arrays.items():
            buf = _io.BytesIO()
            np.save(buf, a, allow_pickle=False)
            z.writestr(name + ".npy", buf.getvalue())
    os.replace(tmp, path)


def _read(path):
    with zipfile.ZipFile(path, "r") as z:
        meta = json.loads(z.read("meta.json"))
        if meta.get("format") != FORMAT or meta.get("version", 0) > VERSION:
            raise _lib.TnbError(1, "%s: not a %s v<=%d file" % (path, FORMAT, VERSION))
        arrays = {n[:-4]: np.load(_io.BytesIO(z.read(n)), allow_pickle=False) for n in z.namelist() if n.endswith(".npy")}
    return meta, arrays


def save_chain(path, chain, extra=None):
    """Write an ``MPS`` or ``MPO`` (on the GPU or on the host).  ``extra``: JSON-serialisable user data (sweep number,
    energy, ...).  Returns the number of bytes of tensor data written."""
    from .mps import MPS, MPO
    if not isinstance(chain, (MPS, MPO)):
        raise TypeError("save_chain: MPS or MPO expected")
    sites, arrays, nbytes = [], {}, 0
    for n, t in enumerate(chain.tensors):
        v, dims = _flat_host(t)
        if v.dtype not in (np.float64, np.complex128):
            v = v.astype(np.complex128 if np.iscomplexobj(v) else np.float64)
        arrays["site_%d" % n] = v
        nbytes += v.nbytes
        sites.append({"dims": [int(d) for d in dims], "inds": None, "storage": "Dense", "eltype": _eltype(v)})
    meta = {"format": FORMAT, "version": VERSION, "type": "MPS" if isinstance(chain, MPS) else "MPO", "length": len(chain),
            "llim": getattr(chain, "llim", None), "rlim": getattr(chain, "rlim", None), "sites": sites, "extra": extra or {}}
    _write(path, meta, arrays)
    return nbytes


def load_chain(path, device=True):
    """Read an MPS / MPO.  ``device=True`` uploads the site tensors (``cu``); returns (chain, extra)."""
    from .mps import MPS, MPO
    meta, arrays = _read(path)
    if meta["type"] not in ("MPS", "MPO"):
        raise _lib.TnbError(1, "%s holds a %s, not an MPS / MPO" % (path, meta["type"]))
    ts = []
    for n, sd in enumerate(meta["sites"]):
        v = arrays["site_%d" % n]
        dims = tuple(sd["dims"])
        if v.size != int(np.prod(dims, dtype=np.int64)):
            raise _lib.DimensionMismatch(2, "%s: site %d has %d elements for dims %s" % (path, n, v.size, dims))
        ts.append(v.reshape(dims, order="F"))
    out = MPS(ts, llim=meta["llim"], rlim=meta["rlim"]) if meta["type"] == "MPS" else MPO(ts)
    return (out.cu() if device else out), meta.get("extra", {})


def save_itensor(path, T, extra=None):
    """Write one ITensor (Dense or Diag storage) with its index metadata."""
    from .itensor import ITensor, DiagStore, UniformDiagStore
    if not isinstance(T, ITensor):
        raise TypeError("save_itensor: ITensor expected")
    if isinstance(T.store, UniformDiagStore):
        v, storage = np.full(T.store.k, T.store.value), "Diag"
    elif isinstance(T.store, DiagStore):
        v, storage = T.store.vec.cpu().numpy(), "Diag"
    else:
        v, _ = _flat_host(T.store)
        storage = "Dense"
    inds = [{"id": i.id, "dim": i.dim, "plev": i.plev, "tags": i.tags} for i in T.inds]
    meta = {"format": FORMAT, "version": VERSION, "type": "ITensor", "length": 1, "llim": None, "rlim": None,
            "sites": [{"dims": [i.dim for i in T.inds], "inds": inds, "storage": storage, "eltype": _eltype(v)}], "extra": extra or {}}
    _write(path, meta, {"site_0": v})
    return v.nbytes


def load_itensor(path, device=True):
    import torch
    from .itensor import ITensor, Index, DiagStore
    meta, arrays = _read(path)
    if meta["type"] != "ITensor":
        raise _lib.TnbError(1, "%s holds a %s, not an ITensor" % (path, meta["type"]))
    sd = meta["sites"][0]
    inds = [Index(i["dim"], i["tags"], i["plev"], i["id"]) for i in sd["inds"]]
    v = arrays["site_0"]
    if sd["storage"] == "Diag":
        if not device:
            return ITensor(np.diag(v), inds), meta.get("extra", {})
        return ITensor(DiagStore(torch.from_numpy(v).cuda()), inds), meta.get("extra", {})
    a = v.reshape(tuple(sd["dims"]), order="F")
    return ITensor(DTensor.from_numpy(a) if device else a, inds), meta.get("extra", {})
