"""CPU, world_size 2 and 3 over gloo: the N>1 host logic -- tile ownership, compact slot layout, padding to a
regular gather, gather to rank 0 and frame assembly -- with the oracle standing in for each rank's GPU.
The re-assembled frame must be byte-identical to the single-rank frame (no cross-tile arithmetic)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dod_raytracer_b200 import capi, distributed
from oracle_api import CLS_CYLINDER, CLS_PLANE, CLS_SPHERE, CLS_TREE, HIT_DT, Oracle
from scenes import LIGHT0, teapot_scene

ALL = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE
W, H = 200, 117  # partial edge tiles in both directions


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tile, outdir):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc, scene = Oracle(), teapot_scene(full=True)
    # what this rank's GPU would produce: compact results of its own tiles, padded slots = miss / 0
    full = orc.trace_primary(scene, W, H, ALL)
    fvis = orc.trace_shadow(scene, W, H, ALL, full, LIGHT0)
    frame = distributed.rank_frame(W, H, ALL, rank, world, tile)
    m = capi.frame_pixel_map(frame)
    spr = distributed.slots_per_rank(W, H, world, tile)
    assert len(m) <= spr
    hits = np.zeros(spr, HIT_DT)
    hits["prim"] = 0xFFFFFFFF
    hits["t"] = np.inf
    vis = np.zeros(spr, np.uint8)
    ok = m != 0xFFFFFFFF
    hits[: len(m)][ok] = full[m[ok]]
    vis[: len(m)][ok] = fvis[m[ok]]
    g_hits = distributed.gather_to_rank0(torch.from_numpy(hits.view(np.uint8).reshape(spr, 16)), world, rank)
    g_vis = distributed.gather_to_rank0(torch.from_numpy(vis), world, rank)
    if rank == 0:
        gh = g_hits.numpy().reshape(world, spr * 16).view(HIT_DT).reshape(world, spr)
        a_hits, a_vis = distributed.assemble_host(W, H, world, tile, gh, g_vis.numpy())
        np.save(os.path.join(outdir, "ok.npy"),
                np.array([a_hits.tobytes() == full.tobytes(), a_vis.tobytes() == fvis.tobytes()]))
    else:
        assert g_hits is None and g_vis is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,tile", [(2, (32, 32)), (3, (16, 8))])
def test_tile_split_gather_assemble(world, tile, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), tile, str(tmp_path)), nprocs=world, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.all()


def test_every_pixel_has_exactly_one_owner_at_bench_sizes():
    for (w, h) in ((3840, 2160), (7680, 4320)):
        for world in (2, 4, 8):
            spr = distributed.slots_per_rank(w, h, world)
            total = 0
            for r in range(world):
                n = capi.frame_local_pixels(distributed.rank_frame(w, h, ALL, r, world))
                assert n <= spr
                total += n
            tiles = -(-w // 32) * -(-h // 32)
            assert total == tiles * 32 * 32
