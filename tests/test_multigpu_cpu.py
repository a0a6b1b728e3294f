"""world_size-2 gloo test of the multi-GPU plumbing (image sharding + final frame gather), on CPU."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.multiprocessing as mp

ROOT = str(Path(__file__).resolve().parent.parent)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_images: int, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from v3d_b200 import parallel

    r, _, w = parallel.init(backend="gloo")
    mine = parallel.shard_images(n_images, r, w)
    counts = [len(parallel.shard_images(n_images, i, w)) for i in range(w)]
    frames = torch.stack([torch.full((2, 4, 4, 3), img, dtype=torch.uint8) for img in mine]) if mine else \
        torch.zeros((0, 2, 4, 4, 3), dtype=torch.uint8)
    out = parallel.gather_frames(frames, counts)
    t = parallel.max_over_ranks(float(rank + 1), "cpu")
    parallel.barrier()
    q.put((rank, mine, out[:, 0, 0, 0, 0].tolist(), t))
    torch.distributed.destroy_process_group()


def test_image_sharding_and_frame_gather_gloo():
    world, n_images = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]           # contiguous blocks, remainder first
    for _, _, gathered, tmax in res:
        assert gathered == [0, 1, 2, 3, 4]                           # image order preserved on every rank
        assert tmax == 2.0                                            # max over ranks


def test_shard_partition_properties():
    from v3d_b200 import parallel

    for n in (1, 7, 8, 18):
        for w in (1, 2, 4, 8):
            parts = [parallel.shard_images(n, r, w) for r in range(w)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
