"""Thin object wrapper over the C ABI (include/disyolo.h).  PyTorch tensors are used only as
device / pinned-host buffers; every compute call goes through libdisyolo_b200.so."""
import ctypes as C

import numpy as np

from . import _lib

_TABLE = None


def layer_table():
    """[{id, cin, cout, k, s, bn, res, size_div}] for convolutional1..82 -- the shapes of the
    reference's builder calls (yolo3_net_pos.py:159-412); ``cin`` includes concatenated channels."""
    global _TABLE
    if _TABLE is not None:
        return _TABLE
    t = []

    def add(n, cin, cout, k, s=1, bn=True, res=False, div=1):
        t.append(dict(id=n, cin=cin, cout=cout, k=k, s=s, bn=bn, res=res, size_div=div))

    add(1, 3, 32, 3, div=1)
    add(2, 32, 64, 3, 2, div=2)
    add(3, 64, 32, 1, div=2); add(4, 32, 64, 3, res=True, div=2)
    add(5, 64, 128, 3, 2, div=4)
    for n in (6, 8):
        add(n, 128, 64, 1, div=4); add(n + 1, 64, 128, 3, res=True, div=4)
    add(10, 128, 256, 3, 2, div=8)
    for i in range(8):
        add(11 + 2 * i, 256, 128, 1, div=8); add(12 + 2 * i, 128, 256, 3, res=True, div=8)
    add(27, 256, 512, 3, 2, div=16)
    for i in range(8):
        add(28 + 2 * i, 512, 256, 1, div=16); add(29 + 2 * i, 256, 512, 3, res=True, div=16)
    add(44, 512, 1024, 3, 2, div=32)
    for i in range(4):
        add(45 + 2 * i, 1024, 512, 1, div=32); add(46 + 2 * i, 512, 1024, 3, res=True, div=32)
    for n, (ci, co, k) in zip(range(53, 59), [(1024, 512, 1), (512, 1024, 3)] * 3):
        add(n, ci, co, k, div=32)
    add(59, 1024, 24, 1, bn=False, div=32)
    add(60, 512, 256, 1, div=32)
    add(61, 768, 256, 1, div=16)
    for n, (ci, co, k) in zip(range(62, 67), [(256, 512, 3), (512, 256, 1)] * 3):
        add(n, ci, co, k, div=16)
    add(67, 512, 24, 1, bn=False, div=16)
    add(68, 256, 128, 1, div=16)
    add(69, 384, 128, 1, div=8)
    for n, (ci, co, k) in zip(range(70, 75), [(128, 256, 3), (256, 128, 1)] * 3):
        add(n, ci, co, k, div=8)
    add(75, 256, 24, 1, bn=False, div=8)
    add(76, 128, 64, 1, div=8)
    add(77, 192, 64, 1, div=4); add(78, 64, 128, 3, div=4); add(79, 128, 32, 1, div=4)
    add(80, 96, 32, 1, div=2); add(81, 32, 64, 3, div=2); add(82, 64, 9, 1, bn=False, div=2)
    _TABLE = t
    return t


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Engine(object):
    """One network on one GPU."""

    def __init__(self, image_size=576, max_batch=1, precision='bf16', device=0, anchors=None,
                 num_classes=3, k_map=3, alpha=0.1, bn_eps=1e-5, iou_threshold=0.3,
                 max_detection=30, lock=None):
        import torch
        self.torch = torch
        self.lib = _lib.lib()
        _lib.require_gpu()
        cfg = _lib.DyConfig()
        cfg.num_classes = num_classes
        a = np.asarray(anchors if anchors is not None else
                       [[31, 23], [62, 58], [143, 91], [213, 186], [61, 337], [194, 432],
                        [474, 248], [551, 93], [478, 454]], np.float32).reshape(-1)
        if a.size != 18:
            raise ValueError('anchors must be 9 x (w,h)')
        for i in range(18):
            cfg.anchors[i] = float(a[i])
        cfg.image_size = int(image_size)
        cfg.k_map = int(k_map)
        cfg.alpha = float(alpha)
        cfg.bn_eps = float(bn_eps)
        cfg.iou_threshold = float(iou_threshold)
        cfg.max_detection = int(max_detection)
        cfg.max_batch = int(max_batch)
        cfg.precision = {'bf16': _lib.PRECISION_BF16, 'fp32': _lib.PRECISION_FP32}[precision]
        cfg.device = int(device)
        lk = lock if lock is not None else [1] * 52 + [0] * 30
        for i in range(82):
            cfg.lock[i] = int(bool(lk[i]))
        self.cfg = cfg
        self.image_size, self.max_batch, self.max_detection = int(image_size), int(max_batch), int(max_detection)
        self.precision, self.k_map, self.num_classes = precision, int(k_map), int(num_classes)
        self.device = torch.device('cuda', int(device))
        h = C.c_void_p()
        _lib.check(self.lib.dy_create(C.byref(cfg), C.byref(h)), 'dy_create')
        self.h = h
        self._pinned = {}

    def close(self):
        if getattr(self, 'h', None):
            self.lib.dy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights --------------------------------------------------------------------------
    def load_weights(self, weights):
        for name, arr in weights.items():
            a = np.ascontiguousarray(arr, np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            _lib.check(self.lib.dy_load_weights(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim),
                       'dy_load_weights(%s)' % name)
        _lib.check(self.lib.dy_finalize_weights(self.h), 'dy_finalize_weights')

    # ---- helpers ----------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, x, dtype=None):
        t = self.torch
        if not isinstance(x, t.Tensor):
            x = t.from_numpy(np.ascontiguousarray(x))
        if dtype is not None and x.dtype != dtype:
            x = x.to(dtype)
        return x.to(self.device).contiguous()

    @property
    def mask_size(self):
        return self.image_size // 2

    @property
    def num_candidates(self):
        s = self.image_size
        return 3 * ((s // 8) ** 2 + (s // 16) ** 2 + (s // 32) ** 2)

    # ---- device-resident forward ------------------------------------------------------------
    def forward(self, images, windows, det_thresh, want_masks=True, out=None):
        """images [B,S,S,3] fp32 cuda, windows [B,4] fp32 cuda -> dict of cuda tensors."""
        t = self.torch
        images = self._dev(images, t.float32)
        windows = self._dev(windows, t.float32)
        B = images.shape[0]
        md, sm = self.max_detection, self.mask_size
        if out is None:
            out = dict(det_raw=t.empty((B, md, 6), dtype=t.float32, device=self.device),
                       det_box=t.empty((B, md, 6), dtype=t.float32, device=self.device),
                       det_count=t.empty((B,), dtype=t.int32, device=self.device),
                       masks=t.empty((B, md, sm, sm), dtype=t.float32, device=self.device) if want_masks else None)
        _lib.check(self.lib.dy_forward(self.h, _ptr(images), B, _ptr(windows), float(det_thresh),
                                       _ptr(out['det_raw']), _ptr(out['det_box']), _ptr(out['det_count']),
                                       _ptr(out.get('masks')), self._stream()), 'dy_forward')
        return out

    def capture_graph(self, images, windows, det_thresh, out):
        """Capture one forward (conv1..82 + decode + NMS + masks, ~86 launches) into a CUDA graph over
        fixed device buffers; `graph.replay()` then re-runs it with one launch.  Used for the batch-1
        latency path (launch overhead dominates there)."""
        t = self.torch
        s = t.cuda.Stream(device=self.device)
        s.wait_stream(t.cuda.current_stream(self.device))
        with t.cuda.stream(s):
            for _ in range(2):
                self.forward(images, windows, det_thresh, out=out)     # warm-up: one-time attribute calls
        t.cuda.current_stream(self.device).wait_stream(s)
        t.cuda.synchronize(self.device)
        g = t.cuda.CUDAGraph()
        with t.cuda.graph(g):
            self.forward(images, windows, det_thresh, out=out)
        return g

    def forward_network(self, images):
        images = self._dev(images, self.torch.float32)
        _lib.check(self.lib.dy_forward_network(self.h, _ptr(images), images.shape[0], self._stream()),
                   'dy_forward_network')

    def profile_layers(self, images):
        """Per-layer device milliseconds of one network pass (index 1..82)."""
        images = self._dev(images, self.torch.float32)
        ms = np.zeros(83, np.float32)
        _lib.check(self.lib.dy_forward_profile(self.h, _ptr(images), images.shape[0],
                                               ms.ctypes.data_as(C.c_void_p), self._stream()), 'dy_forward_profile')
        return ms

    # ---- host-buffer forward (the reference-facing call) ------------------------------------
    def pinned(self, key, shape, dtype):
        t = self.torch
        buf = self._pinned.get(key)
        n = int(np.prod(shape))
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = t.empty((n,), dtype=dtype).pin_memory()
            self._pinned[key] = buf
        return buf[:n].view(*shape)

    def forward_host(self, images, windows, det_thresh, want_masks=True):
        """images / windows: numpy or pinned torch CPU tensors.  Returns pinned CPU tensors
        (det_raw, det_box, det_count, masks); masks rows beyond det_count[b] are unspecified."""
        t = self.torch
        B = int(images.shape[0])
        S, md, sm = self.image_size, self.max_detection, self.mask_size

        def stage(key, x, shape):
            if isinstance(x, t.Tensor) and x.is_pinned() and x.dtype == t.float32 and x.is_contiguous():
                return x
            buf = self.pinned(key, shape, t.float32)
            buf.copy_(x if isinstance(x, t.Tensor) else t.from_numpy(np.ascontiguousarray(x, np.float32)))
            return buf
        img = stage('img', images, (B, S, S, 3))
        win = stage('win', windows, (B, 4))
        raw = self.pinned('raw', (B, md, 6), t.float32)
        box = self.pinned('box', (B, md, 6), t.float32)
        cnt = self.pinned('cnt', (B,), t.int32)
        msk = self.pinned('msk', (B, md, sm, sm), t.float32) if want_masks else None
        _lib.check(self.lib.dy_forward_host(self.h, _ptr(img), B, _ptr(win), float(det_thresh), _ptr(raw), _ptr(box),
                                            _ptr(cnt), _ptr(msk)), 'dy_forward_host')
        return raw, box, cnt, msk

    def forward_host_begin(self, images, windows, det_thresh, want_masks=True, masks=None):
        """Pipelined form: enqueue H2D + forward of one batch, return a ticket (dy_forward_host_begin /
        dy_forward_host_begin_u8).  images: fp32 [B,S,S,3] in [0,1] (the reference's feed) or uint8
        [B,S,S,3] (the letterboxed image before image_read's `/ 255.`; divided on the device, bit-identical
        input, a quarter of the bytes).  masks: 'full' (reference layout [B,max_det,S/2,S/2]), 'cropped'
        (only det_mask[y1:y2, x1:x2] per detection -- what calculate_test_map.py:247-252 reads) or 'none';
        default follows want_masks."""
        t = self.torch
        B = int(images.shape[0])
        S = self.image_size
        mode = masks if masks is not None else ('full' if want_masks else 'none')
        mode_id = {'none': _lib.MASKS_NONE, 'full': _lib.MASKS_FULL, 'cropped': _lib.MASKS_CROPPED}[mode]
        u8 = (images.dtype == t.uint8) if isinstance(images, t.Tensor) else (np.asarray(images).dtype == np.uint8)
        dt = t.uint8 if u8 else t.float32

        def stage(key, x, shape, dtype):
            if isinstance(x, t.Tensor) and x.is_pinned() and x.dtype == dtype and x.is_contiguous():
                return x
            buf = self.pinned(key, shape, dtype)
            buf.copy_(x if isinstance(x, t.Tensor) else
                      t.from_numpy(np.ascontiguousarray(x, np.uint8 if dtype == t.uint8 else np.float32)))
            return buf
        if getattr(self, '_inflight', 0) >= 3:
            # refuse BEFORE touching a staging buffer: all three belong to batches whose H2D may still run
            raise _lib.DisYoloError('all three pipeline slots are in flight: call forward_host_end first')
        n = getattr(self, '_begin_count', 0)
        img = stage('img%s%d' % ('u8' if u8 else '', n % 3), images, (B, S, S, 3), dt)
        win = stage('win%d' % (n % 3), windows, (B, 4), t.float32)
        ticket = C.c_int32(-1)
        fn = self.lib.dy_forward_host_begin_u8 if u8 else self.lib.dy_forward_host_begin
        _lib.check(fn(self.h, _ptr(img), B, _ptr(win), float(det_thresh), mode_id, C.byref(ticket)),
                   'dy_forward_host_begin')
        self._begin_count = n + 1
        self._inflight = getattr(self, '_inflight', 0) + 1
        return (ticket.value, B, mode, (img, win))

    def forward_host_end(self, ticket):
        """-> (det_raw, det_box, det_count, masks) pinned CPU tensors.  masks is [B,max_det,S/2,S/2] ('full'),
        None ('none') or, for a 'cropped' ticket, the pair (offsets int64 [B*max_det+1], crops fp32 [total]):
        crop (b,d) = crops[offsets[b*max_det+d] : offsets[b*max_det+d+1]] viewed (y2-y1, x2-x1), see
        crop_view / expand_masks."""
        t = self.torch
        tid, B, mode, _keep = ticket
        if mode is True or mode is False:           # tickets of the round-1 form (want_masks flag)
            mode = 'full' if mode else 'none'
        md, sm = self.max_detection, self.mask_size
        raw = self.pinned('raw%d' % tid, (B, md, 6), t.float32)
        box = self.pinned('box%d' % tid, (B, md, 6), t.float32)
        cnt = self.pinned('cnt%d' % tid, (B,), t.int32)
        if mode == 'cropped':
            off = self.pinned('off%d' % tid, (B * md + 1,), t.int64)
            crops = self.pinned('crop%d' % tid, (B * md * sm * sm,), t.float32)
            _lib.check(self.lib.dy_forward_host_end_cropped(self.h, tid, _ptr(raw), _ptr(box), _ptr(cnt), _ptr(off),
                                                            _ptr(crops), crops.numel()), 'dy_forward_host_end_cropped')
            self._inflight = max(0, getattr(self, '_inflight', 0) - 1)
            return raw, box, cnt, (off, crops[:int(off[B * md])])
        msk = self.pinned('msk%d' % tid, (B, md, sm, sm), t.float32) if mode == 'full' else None
        _lib.check(self.lib.dy_forward_host_end(self.h, tid, _ptr(raw), _ptr(box), _ptr(cnt), _ptr(msk)),
                   'dy_forward_host_end')
        self._inflight = max(0, getattr(self, '_inflight', 0) - 1)
        return raw, box, cnt, msk

    def crop_rect(self, box_row):
        """(y1, x1, y2, x2) = np.around(box * S/2) clamped to the map: the reference consumer's crop
        (calculate_test_map.py:247-252) and the extent of a 'cropped' map."""
        sm = self.mask_size
        r = np.around(np.asarray(box_row[:4], np.float32) * np.float32(sm)).astype(np.int64)
        y1, x1, y2, x2 = [int(min(max(v, 0), sm)) for v in r]
        return y1, x1, y2, x2

    def crop_view(self, box, offsets, crops, b, d):
        """The crop of detection d of image b as a 2-D view [(y2-y1), (x2-x1)] of `crops`."""
        y1, x1, y2, x2 = self.crop_rect(box[b, d])
        o = int(offsets[b * self.max_detection + d])
        return crops[o:o + (y2 - y1) * (x2 - x1)].view(y2 - y1, x2 - x1)

    def expand_masks(self, box, cnt, offsets, crops):
        """Cropped result -> the reference's per-image det_mask arrays [n,S/2,S/2] (sigmoid(0) = 0.5 outside
        the box, yolo3_net_pos.py:925-928): host-side convenience for callers that want the dense layout."""
        sm, out = self.mask_size, []
        for b in range(int(cnt.shape[0])):
            n = int(cnt[b])
            m = np.full((n, sm, sm), 0.5, np.float32)
            for d in range(n):
                y1, x1, y2, x2 = self.crop_rect(box[b, d])
                m[d, y1:y2, x1:x2] = self.crop_view(box, offsets, crops, b, d).numpy()
            out.append(m)
        return out

    # ---- training step (bf16 tensor-core engine or fp32 verification engine) ----------------------------------------------------------
    def set_loss_params(self, object_scale=2.0, noobject_scale=1.0, class_scale=1.0, coord_scale=1.0,
                        mask_scale=5.0, ignore_thresh=0.5):
        """cfg.OBJECT_SCALE ... cfg.IGNORE_THRESH (yolo/config.py:49-57) of the training losses."""
        _lib.check(self.lib.dy_set_loss_params(self.h, float(object_scale), float(noobject_scale), float(class_scale),
                                               float(coord_scale), float(mask_scale), float(ignore_thresh)),
                   'dy_set_loss_params')

    def train_init(self):
        """Allocate the training state; returns the number of trainable scalars."""
        _lib.check(self.lib.dy_train_init(self.h), 'dy_train_init')
        self.n_train = int(self.lib.dy_train_param_count(self.h))
        self.grad_flat = self.torch.zeros((self.n_train,), dtype=self.torch.float32, device=self.device)
        return self.n_train

    def layer_span(self, layer):
        off, cnt = C.c_int64(), C.c_int64()
        _lib.check(self.lib.dy_train_layer_span(self.h, layer, C.byref(off), C.byref(cnt)), 'dy_train_layer_span')
        return off.value, cnt.value

    def train_forward(self, images, labels, true_boxes, true_masks, perm_prop, perm_gt, det_thresh):
        """labels = [yolo3, yolo2, yolo1] (stride 8/16/32).  Returns the 8 loss scalars
        (total, object, noobject, class, xy, wh, mask, l2) as a numpy array."""
        t = self.torch
        images = self._dev(images, t.float32)
        B = images.shape[0]

        def dev(x, np_dtype, dtype, flat=True):
            # device tensors pass through untouched (inputs resident in HBM); host arrays are uploaded
            if not isinstance(x, t.Tensor):
                x = np.ascontiguousarray(x).astype(np_dtype, copy=False)
            x = self._dev(x, dtype)
            return x.reshape(B, -1) if flat else x
        lab = [dev(l, np.float32, t.float32) for l in labels]
        tb = dev(true_boxes, np.float32, t.float32)
        tm = dev(true_masks, np.uint8, t.uint8, flat=False)
        pp = dev(perm_prop, np.int32, t.int32, flat=False)
        pg = dev(perm_gt, np.int32, t.int32, flat=False)
        if pp.shape != (B, self.max_detection) or pg.shape != (B, 20):
            raise ValueError('perm_prop must be [B,max_detection], perm_gt [B,20]')
        losses = np.zeros(8, np.float32)
        self._train_keep = (images, lab, tb, tm, pp, pg)
        _lib.check(self.lib.dy_train_forward(self.h, _ptr(images), B, _ptr(lab[0]), _ptr(lab[1]), _ptr(lab[2]),
                                             _ptr(tb), _ptr(tm), _ptr(pp), _ptr(pg), float(det_thresh),
                                             losses.ctypes.data_as(C.c_void_p), self._stream()), 'dy_train_forward')
        self._train_B = B
        return losses

    def train_backward(self, layer_hi=82, layer_lo=1, grad=None):
        g = self.grad_flat if grad is None else grad
        _lib.check(self.lib.dy_train_backward(self.h, self._train_B, layer_hi, layer_lo, _ptr(g), self._stream()),
                   'dy_train_backward')
        return g

    def train_apply(self, lr, grad_scale=1.0, grad=None):
        g = self.grad_flat if grad is None else grad
        _lib.check(self.lib.dy_train_apply(self.h, _ptr(g), float(lr), float(grad_scale), self._stream()),
                   'dy_train_apply')

    def train_tensor(self, layer, which='dy'):
        """Parity tap: 'z' = pre-BN conv output, 'dy' = gradient w.r.t. the layer output (NHWC)."""
        h, w, c = self.layer_shape(layer)
        out = self.torch.empty((self._train_B, h, w, c), dtype=self.torch.float32, device=self.device)
        _lib.check(self.lib.dy_train_get_tensor(self.h, layer, 0 if which == 'z' else 1, self._train_B, _ptr(out),
                                                self._stream()), 'dy_train_get_tensor')
        return out

    def get_weights(self, name, shape):
        out = np.zeros(int(np.prod(shape)), np.float32)
        _lib.check(self.lib.dy_get_weights(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size),
                   'dy_get_weights')
        return out.reshape(shape)

    # ---- parity taps ------------------------------------------------------------------------
    def layer_shape(self, layer):
        h, w, c = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(self.lib.dy_layer_shape(self.h, layer, C.byref(h), C.byref(w), C.byref(c)), 'dy_layer_shape')
        return h.value, w.value, c.value

    def activation(self, layer, B):
        h, w, c = self.layer_shape(layer)
        out = self.torch.empty((B, h, w, c), dtype=self.torch.float32, device=self.device)
        _lib.check(self.lib.dy_get_activation(self.h, layer, B, _ptr(out), self._stream()), 'dy_get_activation')
        return out

    def yolo(self, scale, B):
        g = self.image_size // (8, 16, 32)[scale]
        out = self.torch.empty((B, g, g, 3, 5 + self.num_classes), dtype=self.torch.float32, device=self.device)
        _lib.check(self.lib.dy_get_yolo(self.h, scale, B, _ptr(out), self._stream()), 'dy_get_yolo')
        return out

    def mask_pos(self, B):
        sm = self.mask_size
        out = self.torch.empty((B, sm, sm, self.k_map ** 2), dtype=self.torch.float32, device=self.device)
        _lib.check(self.lib.dy_get_mask_pos(self.h, B, _ptr(out), self._stream()), 'dy_get_mask_pos')
        return out

    # ---- stand-alone stages -----------------------------------------------------------------
    def decode(self, yolos, windows):
        t = self.torch
        y = [self._dev(v, t.float32) for v in yolos]
        windows = self._dev(windows, t.float32)
        B, n0 = y[0].shape[0], self.num_candidates
        box = t.empty((B, n0, 4), dtype=t.float32, device=self.device)
        cls = t.empty((B, n0), dtype=t.int32, device=self.device)
        score = t.empty((B, n0), dtype=t.float32, device=self.device)
        _lib.check(self.lib.dy_decode(self.h, _ptr(y[0]), _ptr(y[1]), _ptr(y[2]), B, _ptr(windows), _ptr(box),
                                      _ptr(cls), _ptr(score), self._stream()), 'dy_decode')
        return box, cls, score

    def detect(self, yolos, windows, det_thresh):
        t = self.torch
        y = [self._dev(v, t.float32) for v in yolos]
        windows = self._dev(windows, t.float32)
        B, md = y[0].shape[0], self.max_detection
        raw = t.empty((B, md, 6), dtype=t.float32, device=self.device)
        box = t.empty((B, md, 6), dtype=t.float32, device=self.device)
        cnt = t.empty((B,), dtype=t.int32, device=self.device)
        _lib.check(self.lib.dy_detect(self.h, _ptr(y[0]), _ptr(y[1]), _ptr(y[2]), B, _ptr(windows), float(det_thresh),
                                      _ptr(raw), _ptr(box), _ptr(cnt), self._stream()), 'dy_detect')
        return raw, box, cnt

    def nms(self, box, cls, score, det_thresh):
        t = self.torch
        box, cls, score = self._dev(box, t.float32), self._dev(cls, t.int32), self._dev(score, t.float32)
        B, N, md = box.shape[0], box.shape[1], self.max_detection
        idx = t.empty((B, md), dtype=t.int32, device=self.device)
        cnt = t.empty((B,), dtype=t.int32, device=self.device)
        raw = t.empty((B, md, 6), dtype=t.float32, device=self.device)
        _lib.check(self.lib.dy_nms(self.h, _ptr(box), _ptr(cls), _ptr(score), B, N, float(det_thresh), _ptr(idx),
                                   _ptr(cnt), _ptr(raw), self._stream()), 'dy_nms')
        return idx, cnt, raw

    def assemble_masks(self, score_maps, det_box, det_count, layout='nhwc', out=None):
        t = self.torch
        score_maps = self._dev(score_maps, t.float32)
        det_box, det_count = self._dev(det_box, t.float32), self._dev(det_count, t.int32)
        B, md, sm = det_box.shape[0], self.max_detection, self.mask_size
        if out is None:
            out = t.empty((B, md, sm, sm), dtype=t.float32, device=self.device)
        _lib.check(self.lib.dy_assemble_masks(self.h, _ptr(score_maps), 0 if layout == 'nhwc' else 1, B,
                                              _ptr(det_box), _ptr(det_count), _ptr(out), self._stream()),
                   'dy_assemble_masks')
        return out

    def postproc_profile(self, yolos, score_maps, windows, det_thresh, masks_out, layout='nhwc', reps=20):
        """Device ms per launch of (decode, nms, finalize, mask assembly); see dy_postproc_profile."""
        t = self.torch
        y = [self._dev(v, t.float32) for v in yolos]
        score_maps, windows = self._dev(score_maps, t.float32), self._dev(windows, t.float32)
        ms = np.zeros(4, np.float32)
        _lib.check(self.lib.dy_postproc_profile(self.h, _ptr(y[0]), _ptr(y[1]), _ptr(y[2]), _ptr(score_maps),
                                                0 if layout == 'nhwc' else 1, y[0].shape[0], _ptr(windows),
                                                float(det_thresh), _ptr(masks_out), int(reps),
                                                ms.ctypes.data_as(C.c_void_p), self._stream()), 'dy_postproc_profile')
        return dict(decode=float(ms[0]), nms=float(ms[1]), finalize=float(ms[2]), masks=float(ms[3]))

    # ---- either side of the hot path ---------------------------------------
    def letterbox(self, image_rgb, out=None):
        """image_read (calculate_test_map.py:149-176) on the GPU: image_rgb [h,w,3] uint8 (numpy or cuda
        tensor) -> (new_image [S,S,3] fp32 cuda, window [4] float32 numpy)."""
        t = self.torch
        img = image_rgb if isinstance(image_rgb, t.Tensor) else t.from_numpy(np.ascontiguousarray(image_rgb, np.uint8))
        img = img.to(self.device).contiguous()
        h, w = int(img.shape[0]), int(img.shape[1])
        S = self.image_size
        if out is None:
            out = t.empty((S, S, 3), dtype=t.float32, device=self.device)
        window = np.zeros(4, np.float32)
        _lib.check(self.lib.dy_letterbox(_ptr(img), h, w, S, _ptr(out), window.ctypes.data_as(C.c_void_p),
                                         self._stream()), 'dy_letterbox')
        return out, window

    def postprocess(self, det_box, det_count, masks, image_h, image_w, want_full=True):
        """The per-detection loop of calculate_test_map.py:233-269 for one image: det_box [max_det,6],
        det_count [1] / scalar tensor, masks [max_det,S,S] (cuda) -> dict(boxes [max_det,4] int32 (x1,y1,x2,y2),
        valid [max_det] uint8, full_masks [max_det,h,w] uint8 or None, merged [h,w] uint8)."""
        t = self.torch
        md, S = int(masks.shape[0]), int(masks.shape[1])
        det_box, masks = self._dev(det_box, t.float32), self._dev(masks, t.float32)
        det_count = self._dev(det_count, t.int32).reshape(-1)
        boxes = t.empty((md, 4), dtype=t.int32, device=self.device)
        valid = t.empty((md,), dtype=t.uint8, device=self.device)
        full = t.empty((md, image_h, image_w), dtype=t.uint8, device=self.device) if want_full else None
        merged = t.empty((image_h, image_w), dtype=t.uint8, device=self.device)
        _lib.check(self.lib.dy_postprocess(_ptr(det_box), _ptr(det_count), md, _ptr(masks), S, int(image_h),
                                           int(image_w), self.image_size, _ptr(boxes), _ptr(valid), _ptr(full),
                                           _ptr(merged), self._stream()), 'dy_postprocess')
        return dict(boxes=boxes, valid=valid, full_masks=full, merged=merged)

    def mask_overlaps(self, masks1, masks2):
        """compute_overlaps_masks (utils/voc_eval_mask.py:38-56) on the GPU: masks1 [n1,h,w], masks2 [n2,h,w]
        (bool / uint8; numpy or cuda tensors, instance-major) -> IoU [n1,n2] fp32 cuda tensor."""
        t = self.torch
        a = masks1 if isinstance(masks1, t.Tensor) else t.from_numpy(np.ascontiguousarray(masks1).astype(np.uint8))
        b = masks2 if isinstance(masks2, t.Tensor) else t.from_numpy(np.ascontiguousarray(masks2).astype(np.uint8))
        a = a.to(self.device).to(t.uint8).contiguous()
        b = b.to(self.device).to(t.uint8).contiguous()
        n1, n2 = int(a.shape[0]), int(b.shape[0])
        if n1 == 0 or n2 == 0:
            return t.zeros((n1, n2), dtype=t.float32, device=self.device)
        P = int(a[0].numel())
        if int(b[0].numel()) != P:
            raise ValueError('mask sets have different extents')
        out = t.empty((n1, n2), dtype=t.float32, device=self.device)
        _lib.check(self.lib.dy_mask_overlaps(_ptr(a), n1, _ptr(b), n2, P, _ptr(out), self._stream()),
                   'dy_mask_overlaps')
        return out

    def assign_labels(self, boxes, nbox, place, flip=None, max_box=20):
        """The label assignment of defect_train.get (utils/train_data.py:134-178) on the GPU.
        boxes [B,max_box,5] (x1,y1,x2,y2,class; original pixels), nbox [B], place [B,4] (sx,sy,dx,dy),
        flip [B] (1/2/3) -> (yolo3, yolo2, yolo1, true_boxes) cuda tensors in the training step's feed layout."""
        t = self.torch
        boxes, place = self._dev(boxes, t.float32), self._dev(place, t.float32)
        nbox = self._dev(nbox, t.int32)
        flip = self._dev(flip, t.int32) if flip is not None else None
        B, S, d = int(boxes.shape[0]), self.image_size, 5 + self.num_classes
        ys = [t.empty((B, S // m, S // m, 3, d), dtype=t.float32, device=self.device) for m in (8, 16, 32)]
        tb = t.empty((B, max_box, 5), dtype=t.float32, device=self.device)
        _lib.check(self.lib.dy_assign_labels(self.h, _ptr(boxes), _ptr(nbox), _ptr(place), _ptr(flip), B, int(max_box),
                                             _ptr(ys[0]), _ptr(ys[1]), _ptr(ys[2]), _ptr(tb), self._stream()),
                   'dy_assign_labels')
        return ys[0], ys[1], ys[2], tb


# ---- the training data pipeline on the device (utils/train_data.py) ---------------------------------
class TrainData(object):
    """Device side of defect_train.get() for ONE image (utils/train_data.py:44-276): polygons -> masks -> boxes,
    random scale / crop / flip of image and masks, noise / light / blur, / 255.  The caller makes the reference's
    random draws (scale_crop, flip, bnl and the draws inside the three effects) and passes them in; label
    assignment is Engine.assign_labels.  All arrays are CUDA tensors; nothing is computed on the host."""

    def __init__(self, image_size=576, device=0):
        import torch
        self.t, self.lib, self.S = torch, _lib.lib(), int(image_size)
        _lib.require_gpu()
        self.device = torch.device('cuda', int(device))

    def _st(self):
        return C.c_void_p(self.t.cuda.current_stream(self.device).cuda_stream)

    def _u8(self, x):
        t = self.t
        x = x if isinstance(x, t.Tensor) else t.from_numpy(np.ascontiguousarray(x).astype(np.uint8, copy=False))
        return x.to(self.device).to(t.uint8).contiguous()

    def polygon_masks(self, polygons, h, w):
        """load_mask (:321-338).  polygons: per instance a list of {'type', 'all_points_x', 'all_points_y'} (VIA
        shape_attributes) -> uint8 cuda tensor [n_inst, h, w]."""
        t = self.t
        verts, poly, inst = [], [], [0]
        for each_instance in polygons:
            for p in each_instance:
                poly.append((len(verts), len(p['all_points_x']), 1 if p['type'] == 'out' else 0))
                verts += list(zip(p['all_points_x'], p['all_points_y']))
            inst.append(len(poly))
        n = len(polygons)
        out = t.empty((n, h, w), dtype=t.uint8, device=self.device)
        if n == 0:
            return out
        v = t.tensor(verts if verts else [(0, 0)], dtype=t.float64, device=self.device)
        pl = t.tensor(poly if poly else [(0, 0, 0)], dtype=t.int32, device=self.device)
        it = t.tensor(inst, dtype=t.int32, device=self.device)
        _lib.check(self.lib.dy_polygon_masks(_ptr(v), _ptr(pl), _ptr(it), n, int(h), int(w), _ptr(out), self._st()),
                   'dy_polygon_masks')
        return out

    def mask_boxes(self, masks):
        """extract_bboxes (:358-374) for every mask: int32 [n,4] (x1,y1,x2,y2), zeros for an empty mask."""
        t = self.t
        masks = self._u8(masks)
        n, h, w = [int(v) for v in masks.shape]
        out = t.empty((n, 4), dtype=t.int32, device=self.device)
        _lib.check(self.lib.dy_mask_boxes(_ptr(masks), n, h, w, _ptr(out), self._st()), 'dy_mask_boxes')
        return out

    def place_image(self, image, new_w, new_h, dx, dy, flip=1):
        """apply_random_scale_and_crop(mode='image') + flip: uint8 [h,w,3] -> uint8 [S,S,3]."""
        t = self.t
        image = self._u8(image)
        h, w = int(image.shape[0]), int(image.shape[1])
        out = t.empty((self.S, self.S, 3), dtype=t.uint8, device=self.device)
        _lib.check(self.lib.dy_augment_image(_ptr(image), h, w, self.S, int(new_w), int(new_h), int(dx), int(dy),
                                             int(flip), _ptr(out), self._st()), 'dy_augment_image')
        return out

    def place_masks(self, masks, new_w, new_h, dx, dy, flip=1):
        """resize_mask (:403-421): uint8 / bool [n,h,w] -> uint8 (0/1) [n,S,S]."""
        t = self.t
        masks = self._u8(masks)
        n, h, w = [int(v) for v in masks.shape]
        out = t.empty((n, self.S, self.S), dtype=t.uint8, device=self.device)
        if n:
            _lib.check(self.lib.dy_augment_masks(_ptr(masks), n, h, w, self.S, int(new_w), int(new_h), int(dx), int(dy),
                                                 int(flip), _ptr(out), self._st()), 'dy_augment_masks')
        return out

    def salt_pepper(self, image, salt_rc, pepper_rc):
        """add_salt_pepper_noise (:494-509), in place on a uint8 [S,S,3] cuda tensor; *_rc int32 [n,2] (row, col)."""
        t = self.t
        s = t.as_tensor(np.ascontiguousarray(salt_rc, np.int32)).to(self.device).contiguous()
        p = t.as_tensor(np.ascontiguousarray(pepper_rc, np.int32)).to(self.device).contiguous()
        _lib.check(self.lib.dy_salt_pepper(_ptr(image), int(image.shape[0]), _ptr(s), int(s.shape[0]), _ptr(p),
                                           int(p.shape[0]), self._st()), 'dy_salt_pepper')
        return image

    def change_light(self, image, coeff):
        """change_light (:511-521), in place on a uint8 [...,3] cuda tensor."""
        _lib.check(self.lib.dy_change_light(_ptr(image), int(image.numel() // 3), float(coeff), self._st()),
                   'dy_change_light')
        return image

    def motion_blur3(self, image, kernel):
        """linearmotion_blur3C (:452-481) with the 3x3 line kernel pyblur builds (numpy [3,3] float32)."""
        k = np.ascontiguousarray(kernel, np.float32).reshape(9)
        out = self.t.empty_like(image)
        _lib.check(self.lib.dy_motion_blur3(_ptr(image), int(image.shape[0]), k.ctypes.data_as(C.c_void_p), _ptr(out),
                                            self._st()), 'dy_motion_blur3')
        return out

    def to_unit_float(self, image):
        """image.astype(np.float32) / 255.0 (:399-400)."""
        out = self.t.empty(image.shape, dtype=self.t.float32, device=self.device)
        _lib.check(self.lib.dy_u8_to_unit_float(_ptr(image), _ptr(out), int(image.numel()), self._st()),
                   'dy_u8_to_unit_float')
        return out

    def image_read(self, image, new_w, new_h, dx, dy, flip, bnl, salt_rc=None, pepper_rc=None, coeff=None, kernel=None):
        """image_read (:376-401): place + flip, then bnl 1 none / 2 salt & pepper / 3 light / 4 motion blur, / 255."""
        im = self.place_image(image, new_w, new_h, dx, dy, flip)
        if bnl == 2:
            im = self.salt_pepper(im, salt_rc, pepper_rc)
        elif bnl == 3:
            im = self.change_light(im, coeff)
        elif bnl == 4:
            im = self.motion_blur3(im, kernel)
        return self.to_unit_float(im)


def set_option(name, value):
    """dy_set_option: planning overrides of the conv engine ('tc_resident', 'tc_halo'; -1 = auto)."""
    _lib.check(_lib.lib().dy_set_option(name.encode(), int(value)), 'dy_set_option')


def conv_layer(x, w, stride, scale, shift, act, alpha=0.1, residual=None, precision='bf16'):
    """One conv / conv_bn / res_conv_bn of the reference through the library (dy_conv_layer).
    x [B,H,W,cin] cuda fp32, w HWIO numpy, scale/shift numpy [cout]."""
    import torch
    lib = _lib.lib()
    _lib.require_gpu()
    x = x.contiguous().float()
    B, H, W, cin = x.shape
    w = np.ascontiguousarray(w, np.float32)
    k, cout = w.shape[0], w.shape[3]
    scale = np.ascontiguousarray(scale, np.float32)
    shift = np.ascontiguousarray(shift, np.float32)
    out = torch.empty((B, H // stride, W // stride, cout), dtype=torch.float32, device=x.device)
    if residual is not None:
        residual = residual.contiguous().float()
    with torch.cuda.device(x.device):
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.dy_conv_layer({'bf16': 0, 'fp32': 1}[precision], _ptr(x), B, H, W, cin,
                                     w.ctypes.data_as(C.c_void_p), k, stride, cout,
                                     scale.ctypes.data_as(C.c_void_p), shift.ctypes.data_as(C.c_void_p),
                                     int(bool(act)), float(alpha), _ptr(residual), _ptr(out), st), 'dy_conv_layer')
    return out


def conv_backward(x, dz, w, want_dx=True, want_dw=True, stride=1):
    """Backward of one conv through the tensor-core training engine (dy_conv_backward), stride 1 or 2 (TF 'SAME').
    x [B,H,W,cin], dz [B,H/stride,W/stride,cout] cuda fp32, w HWIO numpy -> (dx [B,H,W,cin], dw [k,k,cin,cout])."""
    import torch
    lib = _lib.lib()
    _lib.require_gpu()
    x, dz = x.contiguous().float(), dz.contiguous().float()
    B, H, W, cin = x.shape
    w = np.ascontiguousarray(w, np.float32)
    k, cout = w.shape[0], w.shape[3]
    dx = torch.empty_like(x) if want_dx else None
    dw = torch.zeros((k, k, cin, cout), dtype=torch.float32, device=x.device) if want_dw else None
    with torch.cuda.device(x.device):
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.dy_conv_backward(_ptr(x), _ptr(dz), B, H, W, cin, w.ctypes.data_as(C.c_void_p), k, int(stride),
                                        cout, _ptr(dx), _ptr(dw), st), 'dy_conv_backward')
    return dx, dw
