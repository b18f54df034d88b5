"""Multi-rank plumbing of the hot path: one process (or thread) per GPU / tile.

The reference decomposes the cell grid with ``mpp_define_layout`` / ``mpp_define_domains``
(driver/icebergs_driver.F90:157-164, src/icebergs_framework.F90:915-930) and moves bergs between
PEs in ``send_bergs_to_other_pes`` (F:2997).  Here every rank owns a :class:`~icebergs_b200.api.Domain`
tile; the library packs leavers on the device and ships them over NCCL (``make_domain``: one
process per GPU under ``torchrun``) or, for ranks that share a process, through an in-process
group (``LocalGroup``: one host thread per rank, device-to-device copies).
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _cdefs as D
from . import api


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(D.KID_NCCL_UNIQUE_ID_BYTES)
    rc = api.lib().kid_nccl_unique_id(buf, D.KID_NCCL_UNIQUE_ID_BYTES)
    if rc:
        raise api.KidFatal(rc, (api.lib().kid_last_error(None) or b"kid_nccl_unique_id failed").decode())
    return buf.raw


def broadcast_unique_id(rank: int) -> bytes:
    """Rank 0 creates the NCCL unique id; everybody gets it through the process group that
    torchrun set up (any backend: nccl on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(D.KID_NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def nccl_comm(uid: bytes, nranks: int, rank: int, device: int) -> int:
    comm = C.c_void_p()
    rc = api.lib().kid_nccl_init(C.byref(comm), uid, len(uid), nranks, rank, device)
    if rc:
        raise api.KidFatal(rc, (api.lib().kid_last_error(None) or b"kid_nccl_init failed").decode())
    return comm.value


def make_domain(gni, gnj, rank, world, halo=4, device=0, cyclic_x=True) -> api.Domain:
    """Tile + NCCL communicator of this rank (torch.distributed must be initialised)."""
    if world == 1:
        return api.Domain.single(gni, gnj, halo=halo, cyclic_x=cyclic_x, device=device)
    uid = broadcast_unique_id(rank)
    comm = nccl_comm(uid, world, rank, device)
    return api.Domain.decomposed(gni, gnj, rank, world, halo=halo, cyclic_x=cyclic_x, device=device, comm=comm,
                                 comm_kind=D.KID_COMM_NCCL)


def tiles(gni, gnj, world, halo=4, cyclic_x=True):
    """The tiles of every rank (no communicator): what mpp_define_domains hands each PE."""
    return [api.Domain.decomposed(gni, gnj, r, world, halo=halo, cyclic_x=cyclic_x) for r in range(world)]


def split_by_owner(cols: dict, domains) -> list:
    """Distributes berg columns (with ine/jne) over the ranks that own their cells."""
    owner = np.array([domains[0].owner_rank(int(i), int(j)) for i, j in zip(cols["ine"], cols["jne"])])
    return [{k: np.ascontiguousarray(v[owner == r]) for k, v in cols.items()} for r in range(len(domains))]


class LocalGroup:
    """Ranks as host threads of this process (several tiles on one GPU, or GPUs without NCCL).

    ``run(fn)`` calls ``fn(rank)`` on one thread per rank -- every library call that communicates
    (icebergs_init / icebergs_run / step_resident / set_forcing) must be made by all ranks, as with MPI."""

    def __init__(self, nranks: int, devices=None):
        self.nranks = nranks
        self.devices = list(devices) if devices is not None else [0] * nranks
        g = C.c_void_p()
        rc = api.lib().kid_local_comm_create(C.byref(g), nranks)
        if rc:
            raise api.KidFatal(rc, "kid_local_comm_create failed")
        self._g = g

    def domain(self, gni, gnj, rank, halo=4, cyclic_x=True) -> api.Domain:
        return api.Domain.decomposed(gni, gnj, rank, self.nranks, halo=halo, cyclic_x=cyclic_x,
                                     device=self.devices[rank], comm=self._g.value, comm_kind=D.KID_COMM_LOCAL)

    def run(self, fn):
        out, err = [None] * self.nranks, [None] * self.nranks

        def work(r):
            try:
                out[r] = fn(r)
            except BaseException as e:  # noqa: BLE001 -- re-raised on the caller's thread
                err[r] = e

        th = [threading.Thread(target=work, args=(r,)) for r in range(self.nranks)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out

    def close(self):
        if self._g:
            api.lib().kid_local_comm_destroy(self._g)
            self._g = None
