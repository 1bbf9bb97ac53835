"""CPU (gloo, world_size 2): the host-side logic of the multi-GPU paths -- bucket planning and the
bucketed asynchronous gradient all-reduce of the data-parallel training step, and the image
sharding arithmetic of multi-GPU inference (no collective on that path)."""
import os
import socket

import numpy as np
import pytest


def test_plan_buckets_contiguous_and_ordered():
    from disyolo_b200.parallel import plan_buckets
    from oracle import dis_oracle as O
    t = O.layer_table()
    spans, off = [], 0
    for n in range(53, 83):
        cnt = t[n]['k'] ** 2 * t[n]['cin'] * t[n]['cout'] + (2 if t[n]['bn'] else 1) * t[n]['cout']
        spans.append((n, off, cnt))
        off += cnt
    assert off == 21070737
    buckets = plan_buckets(spans, 25 << 20)
    assert buckets[0][0] == 82 and buckets[-1][1] == 1          # the ranges tile 82..1 (frozen layers included)
    covered = 0
    prev_lo = 83
    for hi, lo, o, c in buckets:
        assert hi == prev_lo - 1 and lo <= hi
        prev_lo = lo
        covered += c
        assert c * 4 >= (25 << 20) or lo == 1
    assert covered == off
    # descending layers <-> descending offsets: every bucket ends where the previous one starts
    for (h1, l1, o1, c1), (h2, l2, o2, c2) in zip(buckets, buckets[1:]):
        assert o2 + c2 == o1


def test_plan_buckets_covers_locked_layers_between_trainable_ones():
    """A custom lock pattern with frozen layers BETWEEN trainable ones (59-61 locked, 53-58 and 62-82
    trainable): whatever the bucket boundaries, every layer 82..1 is visited exactly once, so a locked
    layer's input gradient is still propagated (dy_train_backward plans overwrite / accumulate statically)."""
    from disyolo_b200.parallel import plan_buckets
    spans, off = [], 0
    for n in list(range(53, 59)) + list(range(62, 83)):
        spans.append((n, off, 1000 + n))
        off += 1000 + n
    for bucket_bytes in (4 * 1000, 4 * 3500, 4 * 9000, 1 << 30):
        buckets = plan_buckets(spans, bucket_bytes)
        visited = []
        for hi, lo, o, c in buckets:
            visited += list(range(hi, lo - 1, -1))
        assert visited == list(range(82, 0, -1)), bucket_bytes
        assert sum(c for _, _, _, c in buckets) == off
        # a bucket's slice holds exactly the trainable layers of its range
        for hi, lo, o, c in buckets:
            want = sum(cnt for n, _, cnt in spans if lo <= n <= hi)
            assert c == want


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from disyolo_b200.parallel import BucketedAllReduce, plan_buckets
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    spans = [(n, (n - 1) * 1000, 1000) for n in range(1, 11)]
    buckets = plan_buckets(spans, bucket_bytes=3000 * 4)
    flat = torch.full((10000,), float(rank + 1))
    flat[rank::7] += 0.5
    ar = BucketedAllReduce(flat, buckets)
    for i in range(len(buckets)):
        ar.bucket_ready(i)
    ar.wait()
    q.put((rank, flat.numpy().copy(), buckets))
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r, flat, buckets = q.get(timeout=120)
        res[r] = flat
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = np.full(10000, 3.0, np.float32)
    want[0::7] += 0.5
    want[1::7] += 0.5
    assert np.array_equal(res[0], want) and np.array_equal(res[1], want)
    assert [b[:2] for b in buckets] == [(82, 8), (7, 5), (4, 2), (1, 1)]


def test_inference_sharding_is_collective_free():
    """Inference shards by image: rank r of N processes its own 64-image batch; the aggregate
    metric is N * 64 * steps / max-over-ranks time (bench.py)."""
    import bench
    assert bench.PER_GPU_BATCH == 64 and bench.IMAGE == 576
    fl = bench.layer_flops(576)
    assert abs(sum(fl.values()) - 132.68385792e9) < 1e3
    assert abs(sum(fl[n] for n in range(2, 83)) + fl[1] - 132.68385792e9) < 1e3
