"""world_size-2 gloo test of the multi-GPU host logic (runs on CPU): contiguous clip sharding covers the
batch exactly once with no data-path collective, per-rank parameter streams are independent, and the
max-over-ranks timing reduction bench.py uses works."""
import os
import random
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vision_collision_detection_b200 import create_video_transforms
from vision_collision_detection_b200.videos import shard_clips


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard_clips(n_clips, rank, world))
    tf = create_video_transforms(mode="train", enable_custom_augmentation=True, crop_size=32)
    random.seed(1234 + rank)
    recs = tf.sample_params(len(mine), 48, 64)
    flips = torch.tensor([float(r["flip"]) for r in recs] + [0.0] * (n_clips - len(mine)))
    owned = torch.zeros(n_clips)
    owned[mine] = 1
    dist.all_reduce(owned)                       # test-only collective: every clip owned exactly once
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # the timing reduction of bench.py
    gathered = [torch.zeros_like(flips) for _ in range(world)]
    dist.all_gather(gathered, flips)
    if rank == 0:
        out.put((owned.tolist(), t.item(), [g.tolist() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_cpu():
    world, n_clips = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    owned, tmax, flips = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert owned == [1.0] * n_clips
    assert tmax == 11.0
    assert flips[0] != flips[1]                  # ranks draw from different random streams
