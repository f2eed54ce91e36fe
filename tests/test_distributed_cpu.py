"""CPU (gloo, world_size 2): the host-side data-parallel logic -- bucket planning, launch order and the
averaged result of GradSync -- plus clip sharding.  No CUDA involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def layout_of(shapes):
    off, out = 0, {}
    for name, n in shapes:
        out[name] = (off, n, torch.Size([n]))
        off += (n + 3) // 4 * 4
    return out, off


SHAPES = [("feature_extractor.head.0.weight", 1728), ("feature_extractor.head.0.bias", 64),
          ("motion_estimator.flow_net.0.weight", 93312), ("temporal_aggregator.attention.0.weight", 110592),
          ("residual_blocks.0.layers.0.0.weight", 18432), ("residual_blocks.0.lff.weight", 14336),
          ("residual_blocks.1.layers.0.0.weight", 18432), ("residual_blocks.1.lff.weight", 14337),
          ("gff.0.weight", 36864), ("upsampler.conv.weight", 6912), ("upsampler.conv.bias", 12)]
READY_ORDER = ["upsampler.", "gff.", "residual_blocks.1.", "residual_blocks.0.", "temporal_aggregator.",
               "motion_estimator.", "feature_extractor."]


def test_bucket_plan_covers_buffer_contiguously():
    from nerve_cl_b200.distributed import plan_buckets
    layout, total = layout_of(SHAPES)
    buckets = plan_buckets(layout, 64 << 10)
    assert buckets[0][1] >= layout["upsampler.conv.bias"][0] + 12          # first bucket = the tail
    assert buckets[-1][0] == 0
    for (a0, b0, _), (a1, b1, _) in zip(buckets[:-1], buckets[1:]):
        assert b1 == a0 and a1 < b1                                        # contiguous, descending
    names = [n for _, _, ns in buckets for n in ns]
    assert sorted(names) == sorted(n for n, _ in SHAPES)


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerve_cl_b200.distributed import GradSync, shard_clips
    layout, total = layout_of(SHAPES)
    flat = torch.full((total,), float(rank + 1))
    sync = GradSync(None, bucket_bytes=64 << 10)
    sync.begin(flat, layout)
    launched_after = []
    for prefix in READY_ORDER:
        sync.on_ready(prefix)
        launched_after.append(len(sync.launch_log))
    sync.finish()
    ok = bool(torch.allclose(flat, torch.full((total,), (1 + world) / 2.0)))
    results[rank] = (ok, launched_after, sync.launch_log, shard_clips(7, rank, world))
    dist.destroy_process_group()


def test_gradsync_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    (ok0, la0, log0, sh0), (ok1, la1, log1, sh1) = results[0], results[1]
    assert ok0 and ok1                                    # mean of rank values everywhere
    assert log0 == log1                                   # same bucket order on both ranks
    assert la0[0] <= la0[-1] and la0[2] >= 1              # buckets launch during "backward", not only at the end
    assert sorted(sh0 + sh1) == list(range(7)) and sh0 == [0, 2, 4, 6]
