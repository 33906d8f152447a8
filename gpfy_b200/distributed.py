"""Data-parallel ELBO: rows of (X, Y) are sharded over ranks, parameters are replicated, and ONE
all-reduce of the packed [data term, gradients] buffer joins the ranks.

The reference has no multi-device code; what is sharded here is the `jnp.mean(var_exp)` over the N
points of optimization.py:153-156 and its reverse pass.  KL (variational.py:225-240) and every
N-independent chain rule are computed identically on each rank and never reduced.

One process per GPU.  `torch.distributed` is the rendezvous (it carries the NCCL unique id and, in the
CPU tests, the gloo all-reduce of the same packed buffer); the device all-reduce itself is enqueued by
the C-ABI (`gpfy_allreduce_packed`) on the compute stream, directly behind the kernels that wrote the
buffer, on a communicator this module creates with NCCL's own C API.
"""
import ctypes
import glob
import os

import numpy as np
import torch

from . import _lib


def shard_bounds(n_total, rank, world):
    """Contiguous row block [lo, hi) of rank `rank`: rows differ by at most one between ranks."""
    n_total, rank, world = int(n_total), int(rank), int(world)
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _NcclUniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * 128)]


def _load_nccl():
    """The NCCL the host framework ships (torch's bundled copy), made visible to libgpfy_b200.so."""
    cands = []
    try:
        import nvidia.nccl  # type: ignore
        for p in nvidia.nccl.__path__:
            cands += glob.glob(os.path.join(p, "lib", "libnccl.so*"))
    except Exception:
        pass
    cands += ["libnccl.so.2"]
    err = None
    for c in cands:
        try:
            return ctypes.CDLL(c, mode=ctypes.RTLD_GLOBAL)
        except OSError as e:  # pragma: no cover
            err = e
    raise _lib.GpfyError(f"NCCL library not found: {err}")


class PackedAllReduce:
    """The single exchange step.  backend 'nccl': device buffers through gpfy_allreduce_packed;
    backend 'gloo': host buffers through torch.distributed (CPU tests of the host-side logic)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.comm = None
        self.backend = dist.get_backend(group) if dist.is_initialized() else "none"
        if self.world > 1 and self.backend == "nccl":
            self._init_nccl()

    def _init_nccl(self):
        nccl = _load_nccl()
        uid = _NcclUniqueId()
        if self.rank == 0:
            rc = nccl.ncclGetUniqueId(ctypes.byref(uid))
            if rc != 0:
                raise _lib.GpfyError(f"ncclGetUniqueId failed: {rc}")
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).to(dev)
        self.dist.broadcast(t, src=0, group=self.group)
        raw = bytes(t.cpu().numpy().tobytes())
        ctypes.memmove(ctypes.byref(uid), raw, 128)
        comm = ctypes.c_void_p()
        nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _NcclUniqueId, ctypes.c_int]
        rc = nccl.ncclCommInitRank(ctypes.byref(comm), self.world, uid, self.rank)
        if rc != 0:
            raise _lib.GpfyError(f"ncclCommInitRank failed: {rc}")
        self.comm, self._nccl = comm, nccl

    def __call__(self, packed):
        """In-place sum over ranks; stream-ordered for device buffers."""
        if self.world == 1:
            return packed
        if packed.is_cuda:
            if self.comm is None:
                raise _lib.GpfyError("device buffers need the nccl backend")
            stream = ctypes.c_void_p(torch.cuda.current_stream(packed.device).cuda_stream)
            _lib.check(_lib.load().gpfy_allreduce_packed(self.comm, stream, ctypes.c_void_p(packed.data_ptr()), packed.numel()))
        else:
            self.dist.all_reduce(packed, op=self.dist.ReduceOp.SUM, group=self.group)
        return packed

    def close(self):
        if self.comm is not None:
            self._nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self._nccl.ncclCommDestroy(self.comm)
            self.comm = None


def finish_elbo(summed, n_total, dataset_size, kl, dkl_mu, dkl_sigma, layout_slices):
    """ELBO and its gradient w.r.t. the constrained hot-path parameters from the all-reduced sums:
        elbo = scale / N * sum_n var_exp_n - KL,  scale = dataset_size if > 0 else N   (optimization.py:154-156)
    `summed` is a host float64 vector in GpfyElboLayout order, `layout_slices` maps names to slices
    (see `layout_slices`).  KL terms come from gpfy_prior_kl_valgrad (device) or the caller."""
    summed = np.asarray(summed, dtype=np.float64)
    scale = float(dataset_size) if dataset_size > 0 else float(n_total)
    c = scale / float(n_total)
    g = {k: c * summed[sl] for k, sl in layout_slices.items() if k != "data_term"}
    value = c * float(summed[layout_slices["data_term"]][0]) - float(kl)
    g["mu"] = g["mu"] - np.asarray(dkl_mu, dtype=np.float64).ravel()
    g["sigma"] = g["sigma"] - np.asarray(dkl_sigma, dtype=np.float64).ravel()
    return value, g


def layout_slices(layout, M, P, F, d, dim, lsize, scalar_w, nw=None):
    L = layout
    return {
        "data_term": slice(L.data_term, L.data_term + 1),
        "mu": slice(L.d_mu, L.d_mu + M * P),
        "sigma": slice(L.d_sigma, L.d_sigma + P * M * M),
        "eig": slice(L.d_eig, L.d_eig + F),
        "vhat": slice(L.d_vhat, L.d_vhat + M * dim),
        "lcat": slice(L.d_lcat, L.d_lcat + lsize),
        "w": slice(L.d_w, L.d_w + (nw if nw is not None else (1 if scalar_w else d))),
        "b": slice(L.d_b, L.d_b + 1),
        "v": slice(L.d_v, L.d_v + 1),
        "sigma2": slice(L.d_sigma2, L.d_sigma2 + 1),
    }
