"""Whole-image evaluation pipeline: what the reference's test driver does per image around
sess.run(net.evaluation) (calculate_test_map.py:208-269) -- letterbox (image_read, :149-176), network +
decode + NMS + mask assembly, then per detection correct_yolo_boxes / crop / resize / threshold / paste
and the merged semantic mask -- with every step on the GPU and only uint8 images going up and the
boxes + merged masks (optionally the boolean instance masks) coming back.

A batch whose frames all have the same shape (the usual case: one camera) takes ONE batched letterbox
launch and ONE post-processing launch pair (dy_letterbox_batch / dy_postprocess_batch) and one D2H copy per
result array; mixed shapes fall back to one launch per image.

Three streams and `depth` slots: the uint8 H2D copy of batch k+1 and the D2H of batch k-1 overlap the
convolutions of batch k.  The network's activation buffers are shared, so letterbox / forward /
post-processing of successive batches are serialised on one compute stream.  (Measured: running the
letterbox / post-processing kernels of the neighbouring batches on their own streams next to the
convolutions is SLOWER -- 1,019 img/s, 4,132 with a high-priority network stream, against 4,406 serialised:
their thousands of small blocks delay the one-CTA-per-SM persistent conv grids.)
"""
import ctypes as C

import numpy as np

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class ImagePipeline(object):
    def __init__(self, engine, max_h, max_w, depth=2, want_instance_masks=False):
        import torch
        self.t, self.eng, self.lib = torch, engine, engine.lib
        self.depth, self.max_h, self.max_w = int(depth), int(max_h), int(max_w)
        self.want_full = bool(want_instance_masks)
        dev, B, S, md, sm = engine.device, engine.max_batch, engine.image_size, engine.max_detection, engine.mask_size
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.slots = []
        for _ in range(self.depth):
            s = dict(
                u8=torch.empty((B, max_h * max_w * 3), dtype=torch.uint8, device=dev),
                batch=torch.empty((B, S, S, 3), dtype=torch.float32, device=dev),
                win_host=torch.zeros((B, 4), dtype=torch.float32).pin_memory(),
                win=torch.empty((B, 4), dtype=torch.float32, device=dev),
                det_raw=torch.empty((B, md, 6), dtype=torch.float32, device=dev),
                det_box=torch.empty((B, md, 6), dtype=torch.float32, device=dev),
                det_count=torch.empty((B,), dtype=torch.int32, device=dev),
                masks=torch.empty((B, md, sm, sm), dtype=torch.float32, device=dev),
                boxes=torch.empty((B, md, 4), dtype=torch.int32, device=dev),
                valid=torch.empty((B, md), dtype=torch.uint8, device=dev),
                merged=torch.empty((B, max_h * max_w), dtype=torch.uint8, device=dev),
                full=(torch.empty((B, md * max_h * max_w), dtype=torch.uint8, device=dev) if self.want_full else None),
                h_boxes=torch.empty((B, md, 4), dtype=torch.int32).pin_memory(),
                h_valid=torch.empty((B, md), dtype=torch.uint8).pin_memory(),
                h_det=torch.empty((B, md, 6), dtype=torch.float32).pin_memory(),
                h_count=torch.empty((B,), dtype=torch.int32).pin_memory(),
                h_merged=torch.empty((B, max_h * max_w), dtype=torch.uint8).pin_memory(),
                h_full=(torch.empty((B, md * max_h * max_w), dtype=torch.uint8).pin_memory() if self.want_full else None),
                ev_h2d=torch.cuda.Event(), ev_comp=torch.cuda.Event(), ev_d2h=torch.cuda.Event(), busy=False, shapes=None)
            self.slots.append(s)
        self.next = 0
        self.h2d_bytes = self.d2h_bytes = 0

    def submit(self, images, det_thresh):
        """images: list (<= max_batch) of uint8 [h,w,3] RGB arrays / CPU tensors (pinned memory makes the copy
        asynchronous).  Returns a ticket for result()."""
        t, eng, lib = self.t, self.eng, self.lib
        sl = self.slots[self.next]
        if sl['busy']:
            raise _lib.DisYoloError('all pipeline slots are in flight: call result() first')
        B, S, md, sm = len(images), eng.image_size, eng.max_detection, eng.mask_size
        shapes = []
        with t.cuda.stream(self.h2d):
            for b, im in enumerate(images):
                im = im if isinstance(im, t.Tensor) else t.from_numpy(np.ascontiguousarray(im, np.uint8))
                h, w = int(im.shape[0]), int(im.shape[1])
                if h > self.max_h or w > self.max_w:
                    raise ValueError('image larger than the pipeline was sized for')
                shapes.append((h, w))
                sl['u8'][b, :h * w * 3].copy_(im.reshape(-1), non_blocking=True)
                self.h2d_bytes += h * w * 3
            sl['ev_h2d'].record(self.h2d)
        cs = C.c_void_p(self.comp.cuda_stream)
        same = len(set(shapes)) == 1                      # one shape: batched launches, compact result layout
        with t.cuda.stream(self.comp):
            self.comp.wait_event(sl['ev_h2d'])
            self.comp.wait_event(sl['ev_d2h'])            # the slot's previous results have left the device
            wh = sl['win_host']
            if same:
                h, w = shapes[0]
                _lib.check(lib.dy_letterbox_batch(_p(sl['u8']), self.max_h * self.max_w * 3, B, h, w, S, _p(sl['batch']),
                                                  C.c_void_p(wh.data_ptr()), cs), 'dy_letterbox_batch')
            else:
                for b, (h, w) in enumerate(shapes):
                    _lib.check(lib.dy_letterbox(_p(sl['u8'][b]), h, w, S, _p(sl['batch'][b]),
                                                C.c_void_p(wh[b].data_ptr()), cs), 'dy_letterbox')
            sl['win'][:B].copy_(wh[:B], non_blocking=True)
            _lib.check(lib.dy_forward(eng.h, _p(sl['batch']), B, _p(sl['win']), float(det_thresh), _p(sl['det_raw']),
                                      _p(sl['det_box']), _p(sl['det_count']), _p(sl['masks']), cs), 'dy_forward')
            if same:
                # [B,h,w] / [B,md,h,w] stacks at the start of the (flat) result buffers
                _lib.check(lib.dy_postprocess_batch(_p(sl['det_box']), _p(sl['det_count']), B, md, _p(sl['masks']), sm,
                                                    h, w, S, _p(sl['boxes']), _p(sl['valid']),
                                                    _p(sl['full']) if self.want_full else None, _p(sl['merged']), cs),
                           'dy_postprocess_batch')
            else:
                for b, (h, w) in enumerate(shapes):
                    _lib.check(lib.dy_postprocess(_p(sl['det_box'][b]), _p(sl['det_count'][b:b + 1]), md,
                                                  _p(sl['masks'][b]), sm, h, w, S, _p(sl['boxes'][b]), _p(sl['valid'][b]),
                                                  _p(sl['full'][b]) if self.want_full else None, _p(sl['merged'][b]), cs),
                               'dy_postprocess')
            sl['ev_comp'].record(self.comp)
        with t.cuda.stream(self.d2h):
            self.d2h.wait_event(sl['ev_comp'])
            sl['h_boxes'][:B].copy_(sl['boxes'][:B], non_blocking=True)
            sl['h_valid'][:B].copy_(sl['valid'][:B], non_blocking=True)
            sl['h_det'][:B].copy_(sl['det_box'][:B], non_blocking=True)
            sl['h_count'][:B].copy_(sl['det_count'][:B], non_blocking=True)
            self.d2h_bytes += B * (md * 16 + md + md * 24 + 4)
            if same:
                h, w = shapes[0]
                sl['h_merged'].view(-1)[:B * h * w].copy_(sl['merged'].view(-1)[:B * h * w], non_blocking=True)
                self.d2h_bytes += B * h * w
                if self.want_full:
                    sl['h_full'].view(-1)[:B * md * h * w].copy_(sl['full'].view(-1)[:B * md * h * w], non_blocking=True)
                    self.d2h_bytes += B * md * h * w
            else:
                for b, (h, w) in enumerate(shapes):
                    sl['h_merged'][b, :h * w].copy_(sl['merged'][b, :h * w], non_blocking=True)
                    self.d2h_bytes += h * w
                    if self.want_full:
                        sl['h_full'][b, :md * h * w].copy_(sl['full'][b, :md * h * w], non_blocking=True)
                        self.d2h_bytes += md * h * w
            sl['ev_d2h'].record(self.d2h)
        sl['busy'], sl['shapes'], sl['compact'] = True, shapes, same
        ticket = self.next
        self.next = (self.next + 1) % self.depth
        return ticket

    def result(self, ticket):
        """Blocks until the batch's results are in host memory.  Returns one dict per image:
        boxes [n,4] int32 (x1,y1,x2,y2 in original pixels), valid [n], classes [n], scores [n],
        merged [h,w] uint8 (class+1 of the last detection covering each pixel), masks [n,h,w] bool or None.
        The arrays are views of pinned buffers that the slot's next submit() overwrites."""
        sl = self.slots[ticket]
        if not sl['busy']:
            raise _lib.DisYoloError('ticket is not in flight')
        sl['ev_d2h'].synchronize()
        md = self.eng.max_detection
        out = []
        for b, (h, w) in enumerate(sl['shapes']):
            n = int(sl['h_count'][b])
            det = sl['h_det'][b, :n].numpy()
            if sl['compact']:
                merged = sl['h_merged'].view(-1)[b * h * w:(b + 1) * h * w]
                full = sl['h_full'].view(-1)[b * md * h * w:(b + 1) * md * h * w] if self.want_full else None
            else:
                merged = sl['h_merged'][b, :h * w]
                full = sl['h_full'][b, :md * h * w] if self.want_full else None
            out.append(dict(boxes=sl['h_boxes'][b, :n].numpy(), valid=sl['h_valid'][b, :n].numpy().astype(bool),
                            classes=det[:, 4].astype(np.int32), scores=det[:, 5],
                            merged=merged.numpy().reshape(h, w),
                            masks=(full.numpy().reshape(md, h, w)[:n].astype(bool) if self.want_full else None)))
        sl['busy'] = False
        return out
