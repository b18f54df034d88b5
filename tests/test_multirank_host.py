"""Host side of the multi-rank path on CPU: two gloo ranks agree on the NCCL bootstrap id,
the tiles partition the grid like mpp_define_layout, bergs are routed to the rank that owns
their cell (send_bergs_to_other_pes F:2997) and the exchange volumes balance."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icebergs_b200 import api, parallel
from icebergs_b200 import synthetic as S

GNI, GNJ, WORLD = 96, 48, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = parallel.broadcast_unique_id(rank)
        doms = parallel.tiles(GNI, GNJ, world)
        me = doms[rank]
        # every rank seeds its own tile (bench.py does the same), then moves the bergs 3 cells east
        # and 1 north: the ones that leave the tile are what send_bergs_to_other_pes ships
        g = S.Grid(GNI, GNJ, me.isc, me.iec, me.jsc, me.jec)
        cols, _ = g.seed_bergs(5000, stream=rank)
        assert ((cols["ine"] >= me.isc) & (cols["ine"] <= me.iec) & (cols["jne"] >= me.jsc) & (cols["jne"] <= me.jec)).all()
        ine, jne = cols["ine"] + 3, cols["jne"] + 1
        owner = np.array([me.owner_rank(int(i), int(j)) for i, j in zip(ine, jne)])
        send = torch.tensor([int(np.sum(owner == r)) for r in range(world)], dtype=torch.int64)
        lost = int(np.sum(owner < 0))
        allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allc, send)                 # what comm_allgather_counts does with ncclAllGather
        counts = torch.stack(allc).numpy()          # counts[src, dst]
        q.put((rank, uid, counts.tolist(), lost, int(send.sum()) + lost))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_route_bergs_consistently():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, WORLD, port, q)) for r in range(WORLD)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=240) for _ in range(WORLD))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (_, uid0, c0, lost0, tot0), (_, uid1, c1, lost1, tot1) = res
    assert uid0 == uid1 and len(uid0) == 128 and any(uid0)
    assert c0 == c1                                        # both ranks see the same count matrix
    counts = np.array(c0)
    assert tot0 == 5000 and tot1 == 5000                   # every berg has exactly one destination (or left the model)
    assert counts[0, 1] > 0 and counts[1, 0] > 0           # east of rank 0 is rank 1; east of rank 1 wraps to rank 0
    assert counts[0, 0] + counts[0, 1] + lost0 == 5000


def test_tiles_partition_the_grid_and_neighbours_are_symmetric():
    for world, gni, gnj in [(2, 96, 48), (4, 96, 48), (8, 1440, 720), (6, 90, 50)]:
        doms = parallel.tiles(gni, gnj, world)
        cover = np.zeros((gnj, gni), dtype=np.int32)
        for d in doms:
            cover[d.jsc - 1:d.jec, d.isc - 1:d.iec] += 1
        assert (cover == 1).all()
        for r, d in enumerate(doms):
            for i, j in [(d.isc, d.jsc), (d.iec, d.jec), (d.isc, d.jec)]:
                assert d.owner_rank(i, j) == r
            if d.pe_E >= 0:
                assert doms[d.pe_E].pe_W == r and d.owner_rank(d.iec + 1, d.jsc) == d.pe_E
            if d.pe_N >= 0:
                assert doms[d.pe_N].pe_S == r and d.owner_rank(d.isc, d.jec + 1) == d.pe_N
            else:
                assert d.owner_rank(d.isc, d.jec + 1) == -1
            assert d.owner_rank(d.isc + gni, d.jsc) == r    # one period off (cyclic x)


def test_owner_of_a_cell_beyond_the_folded_edge():
    """FOLD_NORTH_EDGE (F:933): cell (i, gnj+k) is cell (gni+1-i, gnj+1-k), so a berg that steps over the northern edge of a
    tripolar grid belongs to the rank that owns the mirrored column (send_bergs_to_other_pes F:3138-3147); without the fold
    the northern edge is open (NULL_PE)"""
    gni, gnj = 90, 24
    for world in (1, 2, 4, 6):
        doms = [api.Domain.decomposed(gni, gnj, r, world, halo=4) for r in range(world)] if world > 1 else [api.Domain.single(gni, gnj)]
        d = doms[0]
        assert d.owner_rank(10, gnj + 1) == -1
        d.c.fold_north = 1
        for i in (1, 10, 45, 46, 89, 90, 91, 0):
            for k in (1, 2):
                want = d.owner_rank(gni + 1 - i, gnj + 1 - k)
                assert want >= 0
                assert d.owner_rank(i, gnj + k) == want, (world, i, k)
        # the two pivots of the fold: columns gni/2 | gni/2+1 and gni | 1 face each other
        assert d.owner_rank(gni // 2, gnj + 1) == d.owner_rank(gni // 2 + 1, gnj)
        assert d.owner_rank(gni, gnj + 1) == d.owner_rank(1, gnj)
