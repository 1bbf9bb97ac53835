"""Data-parallel training: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch), a
bucketed gradient all-reduce issued in reverse layer order on a side stream so that it overlaps the
backward pass of the earlier layers.  The reference is single-GPU (yolo/config.py:18); this is the
data-parallel extension north_star asks for.  BatchNorm statistics stay per
replica (the reference's semantics at the per-GPU batch); moving averages are averaged on demand.
"""
import numpy as np


def plan_buckets(spans, bucket_bytes=25 << 20, elem_bytes=4):
    """spans: [(layer, offset, count)] of the trainable layers (ascending layer = ascending offset).
    Returns buckets [(layer_hi, layer_lo, offset, count)] in the order backward produces them
    (descending layers); each bucket is one contiguous slice of the flat gradient vector.

    The layer ranges tile 82..1 without gaps: a locked layer that sits between two trainable ones still
    has to propagate its input gradient, and dy_train_backward plans statically (in descending layer
    order) which consumer overwrites a gradient buffer and which accumulates into it, so every layer
    must be visited exactly once per step."""
    spans = sorted([s for s in spans if s[2] > 0], key=lambda s: s[0], reverse=True)
    groups, cur = [], []
    size = 0
    for layer, off, cnt in spans:
        cur.append((layer, off, cnt))
        size += cnt * elem_bytes
        if size >= bucket_bytes:
            groups.append(cur)
            cur, size = [], 0
    if cur:
        groups.append(cur)
    out = []
    for i, g in enumerate(groups):
        off = min(x[1] for x in g)
        end = max(x[1] + x[2] for x in g)
        if sum(x[2] for x in g) != end - off:
            raise ValueError('bucket is not contiguous in the flat gradient vector')
        hi = 82 if i == 0 else groups[i - 1][-1][0] - 1            # right below the previous bucket's lowest layer
        lo = 1 if i == len(groups) - 1 else g[-1][0]              # down to this bucket's lowest trainable layer
        out.append((hi, lo, off, end - off))
    return out


class BucketedAllReduce(object):
    """Issues one async all-reduce per bucket as soon as the bucket's gradients are complete."""

    def __init__(self, flat, buckets, group=None, comm_stream=None):
        import torch.distributed as dist
        self.dist, self.flat, self.buckets, self.group, self.comm_stream = dist, flat, buckets, group, comm_stream
        self.works = []

    def bucket_ready(self, i):
        """Call right after bucket i's gradients have been enqueued on the current stream."""
        import torch
        hi, lo, off, cnt = self.buckets[i]
        view = self.flat[off:off + cnt]
        if self.flat.is_cuda and self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                self.works.append(self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group,
                                                       async_op=True))
        else:
            self.works.append(self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        import torch
        for w in self.works:
            w.wait()
        self.works = []
        if self.flat.is_cuda and self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


class DataParallelTrainer(object):
    """engine: an Engine (bf16 tensor-core or fp32) with load_weights() done.  Every rank must hold identical
    weights."""

    def __init__(self, engine, bucket_mb=25, group=None):
        import torch
        import torch.distributed as dist
        self.eng, self.torch, self.dist, self.group = engine, torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if not hasattr(engine, 'grad_flat'):
            engine.train_init()
        spans = []
        for layer in range(1, 83):
            off, cnt = engine.layer_span(layer)
            if cnt > 0:
                spans.append((layer, off, cnt))
        self.buckets = plan_buckets(spans, int(bucket_mb) << 20)
        self.comm_stream = torch.cuda.Stream(device=engine.device) if self.world > 1 else None
        # measurement only (bench.py): the same bucketed step without the collectives; step time with minus
        # step time without = the all-reduce time that is NOT hidden behind the backward pass
        self.skip_allreduce = False

    def step(self, images, labels, true_boxes, true_masks, perm_prop, perm_gt, det_thresh, lr):
        """One data-parallel training step on this rank's shard; returns the 8 loss scalars averaged
        over ranks."""
        eng, t = self.eng, self.torch
        losses = eng.train_forward(images, labels, true_boxes, true_masks, perm_prop, perm_gt, det_thresh)
        if self.world == 1:
            eng.train_backward(82, 1)
            eng.train_apply(lr, 1.0)
            return losses
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(t.cuda.current_stream(eng.device))
        ar = BucketedAllReduce(eng.grad_flat, self.buckets, self.group, self.comm_stream)
        for i, (hi, lo, off, cnt) in enumerate(self.buckets):
            eng.train_backward(hi, lo)
            if not self.skip_allreduce:
                ar.bucket_ready(i)
        ar.wait()
        eng.train_apply(lr, 1.0 / self.world)
        lt = t.from_numpy(np.asarray(losses, np.float32)).to(eng.device)
        self.dist.all_reduce(lt, op=self.dist.ReduceOp.SUM, group=self.group)
        return (lt / self.world).cpu().numpy()
