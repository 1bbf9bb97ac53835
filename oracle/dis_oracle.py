"""CPU oracle for the DIS-YOLO hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in NumPy (with torch-CPU used for the fp32 convolution
primitive only), of the algorithm in the reference's ``yolo/yolo3_net_pos.py``.  It is the
checker that the CUDA path is compared against.  It is NOT part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  The product (``dis-yolo_b200``) never imports anything from ``oracle/``.

PARITY STATUS: **parity unpinned** for everything whose arithmetic lives inside TensorFlow 1.x
(conv2d, batch_normalization, sigmoid/exp/softmax, non_max_suppression, top_k, set ops,
round).  TensorFlow is an un-vendored, un-pinned dependency of the reference ("TensorFlow > 1.0",
README.md:5) that cannot be installed in this environment, and the reference ships no tests,
golden vectors or checkpoints.  The oracle therefore restates TensorFlow's documented
semantics at each call site.  What IS pinned: the pieces of reference code that import here
(letterbox, box un-letterbox, sigmoid, mask-IoU, VOC AP) were run to generate
``tests/golden/ref_kat.json`` (see ``tests/golden/make_golden.py``), and the hand-derived
bin-edge known answers of SURVEY.md section 8(c).

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# configuration defaults  (yolo/config.py:21-22,38,41-72)
# --------------------------------------------------------------------------------------
ANCHORS = np.array([[31, 23], [62, 58], [143, 91], [213, 186], [61, 337], [194, 432],
                    [474, 248], [551, 93], [478, 454]], dtype=F32)      # config.py:22
NUM_CLASS = 3            # config.py:21
ALPHA = 0.1              # config.py:38
K_MAP = 3                # config.py:43
OBJ_THRESHOLD = 0.25     # config.py:60
IOU_THRESHOLD = 0.3      # config.py:63
MAX_DETECTION = 30       # config.py:72
BN_EPS = 1e-5            # yolo3_net_pos.py:75
BN_DECAY = 0.997         # yolo3_net_pos.py:74


# --------------------------------------------------------------------------------------
# layer table  (yolo3_net_pos.py:153-412)
# --------------------------------------------------------------------------------------
def layer_table():
    """Return {layer_no: dict(cin, cout, k, s, bn, res)} for convolutional1..82.

    ``bn``  True  -> conv_bn (+leaky)              yolo3_net_pos.py:132-146
            False -> biased linear conv            yolo3_net_pos.py:109-130 (is_act=False)
    ``res`` True  -> res_conv_bn (add after act)   yolo3_net_pos.py:148-151
    """
    t = {}

    def add(n, cin, cout, k, s=1, bn=True, res=False):
        t[n] = dict(cin=cin, cout=cout, k=k, s=s, bn=bn, res=res)

    add(1, 3, 32, 3)                                   # :159-161
    add(2, 32, 64, 3, 2)                               # :165-167
    add(3, 64, 32, 1); add(4, 32, 64, 3, res=True)     # :169-176
    add(5, 64, 128, 3, 2)                              # :180-182
    for n in (6, 8):                                   # :184-200
        add(n, 128, 64, 1); add(n + 1, 64, 128, 3, res=True)
    add(10, 128, 256, 3, 2)                            # :204-206
    for i in range(8):                                 # :208-218
        add(11 + 2 * i, 256, 128, 1); add(12 + 2 * i, 128, 256, 3, res=True)
    add(27, 256, 512, 3, 2)                            # :222-224
    for i in range(8):                                 # :226-236
        add(28 + 2 * i, 512, 256, 1); add(29 + 2 * i, 256, 512, 3, res=True)
    add(44, 512, 1024, 3, 2)                           # :240-242
    for i in range(4):                                 # :244-254
        add(45 + 2 * i, 1024, 512, 1); add(46 + 2 * i, 512, 1024, 3, res=True)
    # head 1 (stride 32)  :258-281
    add(53, 1024, 512, 1); add(54, 512, 1024, 3); add(55, 1024, 512, 1)
    add(56, 512, 1024, 3); add(57, 1024, 512, 1); add(58, 512, 1024, 3)
    add(59, 1024, 24, 1, bn=False)
    # head 2 (stride 16)  :285-316
    add(60, 512, 256, 1)
    add(61, 768, 256, 1); add(62, 256, 512, 3); add(63, 512, 256, 1)
    add(64, 256, 512, 3); add(65, 512, 256, 1); add(66, 256, 512, 3)
    add(67, 512, 24, 1, bn=False)
    # head 3 (stride 8)   :320-351
    add(68, 256, 128, 1)
    add(69, 384, 128, 1); add(70, 128, 256, 3); add(71, 256, 128, 1)
    add(72, 128, 256, 3); add(73, 256, 128, 1); add(74, 128, 256, 3)
    add(75, 256, 24, 1, bn=False)
    # mask subnet, stride 2 (the only active variant)  :381-412
    add(76, 128, 64, 1)
    add(77, 192, 64, 1); add(78, 64, 128, 3); add(79, 128, 32, 1)
    add(80, 96, 32, 1); add(81, 32, 64, 3); add(82, 64, 9, 1, bn=False)
    return t


def default_lock_flags():
    """lock=True for layers 1-52, False for 53-82 (yolo3_net_pos.py:155-156 and each call)."""
    return {n: (n <= 52) for n in range(1, 83)}


def flops_per_image(size=576):
    """2*M*K*N summed over the 82 convs (132.684 GFLOP @576)."""
    t = layer_table()
    hw = _output_sizes(size)
    tot = 0
    for n, L in t.items():
        tot += 2 * hw[n] * hw[n] * L['k'] * L['k'] * L['cin'] * L['cout']
    return tot


def _output_sizes(size):
    s = {}
    s[1] = size
    for n in range(2, 5): s[n] = size // 2
    for n in range(5, 10): s[n] = size // 4
    for n in range(10, 27): s[n] = size // 8
    for n in range(27, 44): s[n] = size // 16
    for n in range(44, 61): s[n] = size // 32
    for n in range(61, 69): s[n] = size // 16
    for n in range(69, 77): s[n] = size // 8
    for n in range(77, 80): s[n] = size // 4
    for n in range(80, 83): s[n] = size // 2
    return s


# --------------------------------------------------------------------------------------
# primitive ops  [TF-semantics]
# --------------------------------------------------------------------------------------
def same_pad(in_size, k, s):
    """TensorFlow 'SAME' padding: total=max((ceil(in/s)-1)*s+k-in,0), before=total//2."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w, stride):
    """tf.nn.conv2d(x NHWC, w HWIO, strides=[1,s,s,1], padding='SAME')  (yolo3_net_pos.py:125,142).

    fp32 via torch-CPU conv2d with explicit (possibly asymmetric) TF padding.
    """
    import torch
    import torch.nn.functional as Fnn
    k = w.shape[0]
    pt, pb = same_pad(x.shape[1], k, stride)
    pl, pr = same_pad(x.shape[2], k, stride)
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=F32)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(w, dtype=F32)).permute(3, 2, 0, 1).contiguous()
    xt = Fnn.pad(xt, (pl, pr, pt, pb))
    y = Fnn.conv2d(xt, wt, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def conv2d_same_numpy(x, w, stride):
    """Independent im2col NumPy restatement of the same op (used to cross-check conv2d_same)."""
    k = w.shape[0]
    B, H, W, C = x.shape
    pt, pb = same_pad(H, k, stride)
    pl, pr = same_pad(W, k, stride)
    xp = np.zeros((B, H + pt + pb, W + pl + pr, C), F32)
    xp[:, pt:pt + H, pl:pl + W] = x
    Ho, Wo = -(-H // stride), -(-W // stride)
    cols = np.zeros((B, Ho, Wo, k, k, C), F32)
    for kh in range(k):
        for kw in range(k):
            cols[:, :, :, kh, kw] = xp[:, kh:kh + stride * (Ho - 1) + 1:stride,
                                       kw:kw + stride * (Wo - 1) + 1:stride]
    out = cols.reshape(B * Ho * Wo, k * k * C).astype(np.float64) @ \
        w.reshape(k * k * C, -1).astype(np.float64)
    return out.reshape(B, Ho, Wo, -1).astype(F32)


def leaky_relu(x, alpha=ALPHA):
    """tf.maximum(alpha*x, x)   (yolo3_net_pos.py:68-69)."""
    return np.maximum(F32(alpha) * x, x)


def batch_norm_infer(x, gamma, beta, mean, var, eps=BN_EPS):
    """tf.nn.batch_normalization with moving stats (yolo3_net_pos.py:81,101):
    (x-mean)*rsqrt(var+eps)*gamma + beta, evaluated as x*inv + (beta-mean*inv)."""
    inv = (gamma / np.sqrt(var + F32(eps))).astype(F32)
    return x * inv + (beta - mean * inv).astype(F32)


def batch_norm_train(x, gamma, beta, eps=BN_EPS):
    """Batch statistics over (N,H,W) (yolo3_net_pos.py:88-98); returns (out, mean, var)."""
    mean = x.mean(axis=(0, 1, 2), dtype=np.float64)
    var = x.var(axis=(0, 1, 2), dtype=np.float64)
    inv = gamma.astype(np.float64) / np.sqrt(var + eps)
    out = x * inv.astype(F32) + (beta - mean * inv).astype(F32)
    return out.astype(F32), mean.astype(F32), var.astype(F32)


def upsample2(x):
    """tf.image.resize_nearest_neighbor to 2x (align_corners=False): out[y,x]=in[y//2,x//2]
    (yolo3_net_pos.py:290,325,386,401)."""
    return x.repeat(2, axis=1).repeat(2, axis=2)


def sigmoid(x):
    x = np.asarray(x, F32)
    with np.errstate(over='ignore'):
        return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


# --------------------------------------------------------------------------------------
# weights  (variable names: train_yolo3_mask.py:87-103)
# --------------------------------------------------------------------------------------
def vname(n, what):
    base = 'yolo/convolutional%d/' % n
    return base + {'w': 'weights', 'b': 'biases', 'gamma': 'BatchNorm/gamma',
                   'beta': 'BatchNorm/beta', 'mean': 'BatchNorm/moving_mean',
                   'var': 'BatchNorm/moving_variance'}[what]


def make_weights(flavour='lively', seed=0):
    """Seeded weights keyed by the reference's TF variable names, HWIO fp32.

    'faithful': the reference's own initialisers -- locked layers truncated-normal sigma=0.001
        with identity BN (yolo3_net_pos.py:77-80,112-113,135-136), unlocked layers
        Xavier-uniform with zero bias (:118-123,138-140).  Numerically degenerate in inference
        mode, kept for completeness.
    'lively': He-scaled weights, random BN statistics, non-zero head biases, so that decode,
        NMS and mask assembly are actually exercised.
    """
    rng = np.random.default_rng(seed)
    t = layer_table()
    lock = default_lock_flags()
    W = {}
    for n in range(1, 83):
        L = t[n]
        k, cin, cout = L['k'], L['cin'], L['cout']
        shape = (k, k, cin, cout)
        if flavour == 'faithful':
            if lock[n]:
                w = rng.standard_normal(shape) * 0.001
                w = np.clip(w, -0.002, 0.002)
            else:
                lim = np.sqrt(6.0 / (k * k * cin + k * k * cout))
                w = rng.uniform(-lim, lim, shape)
            W[vname(n, 'w')] = w.astype(F32)
            if L['bn']:
                W[vname(n, 'gamma')] = np.ones(cout, F32)
                W[vname(n, 'beta')] = np.zeros(cout, F32)
                W[vname(n, 'mean')] = np.zeros(cout, F32)
                W[vname(n, 'var')] = np.ones(cout, F32)
            else:
                W[vname(n, 'b')] = np.zeros(cout, F32)
        else:
            gain = 1.0 if not L['res'] else 0.3
            std = gain * np.sqrt(2.0 / ((1.0 + ALPHA * ALPHA) * k * k * cin))
            W[vname(n, 'w')] = (rng.standard_normal(shape) * std).astype(F32)
            if L['bn']:
                W[vname(n, 'gamma')] = rng.uniform(0.7, 1.3, cout).astype(F32)
                W[vname(n, 'beta')] = (rng.standard_normal(cout) * 0.2).astype(F32)
                W[vname(n, 'mean')] = (rng.standard_normal(cout) * 0.2).astype(F32)
                W[vname(n, 'var')] = rng.uniform(0.6, 1.6, cout).astype(F32)
            else:
                if cout == 24:
                    # head: [tx,ty,tw,th,obj,c0,c1,c2] x 3 anchors
                    # activations reach |x|~10 by the heads: damp so that logits are O(1)
                    w = W[vname(n, 'w')].reshape(k, k, cin, 3, 8) * F32(0.1)
                    w[..., 2:4] *= 0.25          # keep exp(t_wh) tame
                    W[vname(n, 'w')] = w.reshape(shape)
                    b = np.zeros((3, 8), F32)
                    b[:, 4] = -2.6               # sparse objectness
                    b[:, 0:2] = rng.standard_normal((3, 2)) * 0.3
                    b[:, 2:4] = rng.standard_normal((3, 2)) * 0.2 - 0.3
                    b[:, 5:] = rng.standard_normal((3, 3)) * 0.5
                    W[vname(n, 'b')] = b.reshape(24).astype(F32)
                else:
                    W[vname(n, 'w')] *= F32(0.35)
                    W[vname(n, 'b')] = (rng.standard_normal(cout) * 0.5).astype(F32)
    return W


# --------------------------------------------------------------------------------------
# the network  (yolo3_net_pos.py:153-463)
# --------------------------------------------------------------------------------------
def _layer(n, x, W, shortcut=None, training=False, lock=None, stats=None):
    L = layer_table()[n]
    y = conv2d_same(x, W[vname(n, 'w')], L['s'])
    if L['bn']:
        if training and lock is not None and not lock[n]:
            y, m, v = batch_norm_train(y, W[vname(n, 'gamma')], W[vname(n, 'beta')])
            if stats is not None:
                stats[n] = (m, v)
        else:
            y = batch_norm_infer(y, W[vname(n, 'gamma')], W[vname(n, 'beta')],
                                 W[vname(n, 'mean')], W[vname(n, 'var')])
        y = leaky_relu(y)
    else:
        y = y + W[vname(n, 'b')]
    if L['res']:
        y = y + shortcut                      # add AFTER activation (:148-151)
    return y.astype(F32)


def forward_network(images, W, training=False, lock=None, acts=None, stats=None):
    """build_network (yolo3_net_pos.py:153-412).  images: [B,H,W,3] fp32 RGB/255.

    Returns (yolos, mask_pos) with yolos = [yolov3_3 (stride 8), yolov3_2 (16), yolov3_1 (32)]
    (list order of :353), each [B,g,g,3,8]; mask_pos [B,H/2,W/2,9].
    If ``acts`` is a dict it receives every layer's output (NHWC fp32) keyed by layer number.
    """
    if lock is None:
        lock = default_lock_flags()

    def run(n, x, shortcut=None):
        y = _layer(n, x, W, shortcut, training, lock, stats)
        if acts is not None:
            acts[n] = y
        return y

    net = run(1, images)
    net = run(2, net)
    sc = net; net = run(3, net); net = run(4, net, sc); skip2 = net
    net = run(5, net)
    for n in (6, 8):
        sc = net; net = run(n, net); net = run(n + 1, net, sc)
    skip3 = net
    net = run(10, net)
    for i in range(8):
        sc = net; net = run(11 + 2 * i, net); net = run(12 + 2 * i, net, sc)
    skip4 = net
    net = run(27, net)
    for i in range(8):
        sc = net; net = run(28 + 2 * i, net); net = run(29 + 2 * i, net, sc)
    skip5 = net
    net = run(44, net)
    for i in range(4):
        sc = net; net = run(45 + 2 * i, net); net = run(46 + 2 * i, net, sc)
    for n in range(53, 58):
        net = run(n, net)
    y1 = run(59, run(58, net))
    net = run(60, net)
    net = np.concatenate([skip5, upsample2(net)], axis=-1)      # [skip, up]  :291
    for n in range(61, 66):
        net = run(n, net)
    y2 = run(67, run(66, net))
    net = run(68, net)
    net = np.concatenate([skip4, upsample2(net)], axis=-1)      # :326
    for n in range(69, 74):
        net = run(n, net)
    y3 = run(75, run(74, net))
    net = run(76, net)
    net = np.concatenate([skip3, upsample2(net)], axis=-1)      # :387
    net = run(77, net); net = run(78, net); net = run(79, net)
    net = np.concatenate([skip2, upsample2(net)], axis=-1)      # :402
    net = run(80, net); net = run(81, net)
    mask_pos = run(82, net)

    def r5(y):
        return y.reshape(y.shape[0], y.shape[1], y.shape[2], 3, 8)
    return [r5(y3), r5(y2), r5(y1)], mask_pos


# --------------------------------------------------------------------------------------
# decode  (yolo3_net_pos.py:465-514)
# --------------------------------------------------------------------------------------
def interpret_output(yolos, anchors=ANCHORS):
    """Returns dict with lists (per scale, stride 8/16/32) of conf_logits [B,g,g,3,1],
    class_logits [B,g,g,3,C], pred_coords [B,g,g,3,4] (sig xy | raw twh),
    pred_norm_coords [B,g,g,3,4] (xc,yc,w,h in [0,1]) and anchors_pwh [3,2]."""
    net_h = yolos[2].shape[1] * 32                          # :474-476
    net_w = yolos[2].shape[2] * 32
    out = dict(net=(net_h, net_w), conf=[], cls=[], coord=[], norm=[], anchors=[])
    for i in range(3):
        p = yolos[i].astype(F32)
        gh, gw = p.shape[1], p.shape[2]
        cxy = sigmoid(p[..., :2])                           # :487
        cell_x = np.arange(gw, dtype=F32)[None, None, :, None]
        cell_y = np.arange(gh, dtype=F32)[None, :, None, None]
        bx = (cell_x + cxy[..., 0]) / F32(gw)               # :493,504  offset[...,0]=x
        by = (cell_y + cxy[..., 1]) / F32(gh)
        a = anchors[3 * i:3 * i + 3].astype(F32)            # :495-496
        bw = np.exp(p[..., 2], dtype=F32) * a[:, 0] / F32(net_w)    # :502,505
        bh = np.exp(p[..., 3], dtype=F32) * a[:, 1] / F32(net_h)
        out['conf'].append(p[..., 4:5])
        out['cls'].append(p[..., 5:])
        out['coord'].append(np.concatenate([cxy, p[..., 2:4]], -1))
        out['norm'].append(np.stack([bx, by, bw, bh], -1).astype(F32))
        out['anchors'].append(a)
    return out


def clip_boxes(boxes, window):
    """clip_boxes_graph (yolo3_net_pos.py:940-952).  boxes [N,4] y1,x1,y2,x2; window [4]."""
    wy1, wx1, wy2, wx2 = [F32(v) for v in window]
    y1 = np.maximum(np.minimum(boxes[:, 0], wy2), wy1)
    x1 = np.maximum(np.minimum(boxes[:, 1], wx2), wx1)
    y2 = np.maximum(np.minimum(boxes[:, 2], wy2), wy1)
    x2 = np.maximum(np.minimum(boxes[:, 3], wx2), wx1)
    return np.stack([y1, x1, y2, x2], 1).astype(F32)


def decode_candidates(pred, i_img, window):
    """The per-image candidate arrays of filter_detections (yolo3_net_pos.py:523-555):
    flatten each scale in (y,x,anchor) order, concat scales 8->16->32.
    Returns (box [N0,4] y1x1y2x2 clipped, classid [N0] int32, score [N0] f32)."""
    confs, clss, boxes = [], [], []
    for j in range(3):
        confs.append(sigmoid(pred['conf'][j][i_img]).reshape(-1))            # :528-529
        lg = pred['cls'][j][i_img].astype(F32)
        e = np.exp(lg - lg.max(-1, keepdims=True), dtype=F32)                # :532 softmax
        sm = (e / e.sum(-1, keepdims=True, dtype=F32)).astype(F32)
        clss.append(sm.reshape(-1, sm.shape[-1]))
        boxes.append(pred['norm'][j][i_img].reshape(-1, 4))
    conf = np.concatenate(confs)
    cls = np.concatenate(clss)
    nb = np.concatenate(boxes).astype(F32)
    classid = np.argmax(cls, -1).astype(np.int32)                            # :545
    score = (conf * cls[np.arange(len(cls)), classid]).astype(F32)           # :546-548
    xc, yc, w, h = nb[:, 0], nb[:, 1], nb[:, 2], nb[:, 3]
    half = F32(2.0)
    box = np.stack([yc - h / half, xc - w / half, yc + h / half, xc + w / half], 1).astype(F32)
    box = clip_boxes(box, window)                                            # :552-555
    return box, classid, score


# --------------------------------------------------------------------------------------
# NMS  [TF-semantics: tf.image.non_max_suppression]
# --------------------------------------------------------------------------------------
def iou_tf(b1, b2):
    """fp32 IoU exactly as TF's non_max_suppression_op computes it: corner min/max
    normalisation, 0 if either area <= 0, inter/(a1+a2-inter); no FMA contraction."""
    y1a, y2a = min(b1[0], b1[2]), max(b1[0], b1[2])
    x1a, x2a = min(b1[1], b1[3]), max(b1[1], b1[3])
    y1b, y2b = min(b2[0], b2[2]), max(b2[0], b2[2])
    x1b, x2b = min(b2[1], b2[3]), max(b2[1], b2[3])
    aa = F32(F32(y2a - y1a) * F32(x2a - x1a))
    ab = F32(F32(y2b - y1b) * F32(x2b - x1b))
    if aa <= 0 or ab <= 0:
        return F32(0)
    iy1, ix1 = max(y1a, y1b), max(x1a, x1b)
    iy2, ix2 = min(y2a, y2b), min(x2a, x2b)
    inter = F32(max(F32(iy2 - iy1), F32(0)) * max(F32(ix2 - ix1), F32(0)))
    return F32(inter / F32(F32(aa + ab) - inter))


def nms_tf(boxes, scores, idx_tiebreak, max_out, iou_thr):
    """Greedy NMS: visit by (score desc, idx asc) [tie order is OUR definition, SURVEY 7];
    keep if IoU <= thr (suppress when IoU > thr, strict) against every kept box; stop at
    max_out.  Returns positions into ``boxes`` in selection order."""
    boxes = np.asarray(boxes, F32)
    order = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), int(idx_tiebreak[i])))
    thr = F32(iou_thr)
    keep = []
    for i in order:
        if len(keep) >= max_out:
            break
        ok = True
        for j in reversed(keep):
            if iou_tf(boxes[i], boxes[j]) > thr:
                ok = False
                break
        if ok:
            keep.append(i)
    return keep


def nms_tf_vectorised(boxes, scores, idx_tiebreak, max_out, iou_thr):
    """Same result as nms_tf but vectorised per selection (for large candidate counts)."""
    boxes = np.asarray(boxes, F32)
    n = len(scores)
    if n == 0:
        return []
    order = np.lexsort((np.asarray(idx_tiebreak), -np.asarray(scores, F32).astype(np.float64)))
    b = boxes[order]
    y1 = np.minimum(b[:, 0], b[:, 2]); y2 = np.maximum(b[:, 0], b[:, 2])
    x1 = np.minimum(b[:, 1], b[:, 3]); x2 = np.maximum(b[:, 1], b[:, 3])
    area = ((y2 - y1).astype(F32) * (x2 - x1).astype(F32)).astype(F32)
    alive = np.ones(n, bool)
    keep = []
    thr = F32(iou_thr)
    for p in range(n):
        if not alive[p]:
            continue
        keep.append(int(order[p]))
        if len(keep) >= max_out:
            break
        iy1 = np.maximum(y1[p], y1); ix1 = np.maximum(x1[p], x1)
        iy2 = np.minimum(y2[p], y2); ix2 = np.minimum(x2[p], x2)
        inter = (np.maximum((iy2 - iy1).astype(F32), F32(0)) *
                 np.maximum((ix2 - ix1).astype(F32), F32(0))).astype(F32)
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = (inter / ((area[p] + area).astype(F32) - inter).astype(F32)).astype(F32)
        iou = np.where((area[p] <= 0) | (area <= 0), F32(0), iou)
        sup = iou > thr
        sup[:p + 1] = False
        alive &= ~sup
    return keep


# --------------------------------------------------------------------------------------
# filter_detections  (yolo3_net_pos.py:517-628)
# --------------------------------------------------------------------------------------
def select_detections(box, classid, score, thresh, iou_thr=IOU_THRESHOLD,
                      max_detection=MAX_DETECTION, cand_index=None):
    """Threshold (:558, strict >), per-class NMS with max_output_size=max_detection (:566-587),
    intersection with keep (ascending indices, :590-592), top-k by score with ties to the lower
    index (:608-612).  ``box/classid/score`` may be the full candidate arrays or a pre-compacted
    subset; ``cand_index`` carries the original candidate index for tie-breaking.
    Returns (rows [n,6] = y1,x1,y2,x2,classid,score ; kept candidate indices [n])."""
    box = np.asarray(box, F32); score = np.asarray(score, F32)
    classid = np.asarray(classid)
    if cand_index is None:
        cand_index = np.arange(len(score))
    cand_index = np.asarray(cand_index)
    keep = np.nonzero(score > F32(thresh))[0]
    sel = []
    for c in np.unique(classid[keep]):
        ixs = keep[classid[keep] == c]
        fn = nms_tf if len(ixs) <= 512 else nms_tf_vectorised
        k = fn(box[ixs], score[ixs], cand_index[ixs], max_detection, iou_thr)
        sel.extend(ixs[k].tolist())
    sel = np.array(sorted(sel, key=lambda i: int(cand_index[i])), dtype=np.int64)
    if len(sel):
        order = sorted(range(len(sel)), key=lambda p: (-float(score[sel[p]]), int(cand_index[sel[p]])))
        sel = sel[order[:max_detection]]
    rows = np.zeros((len(sel), 6), F32)
    if len(sel):
        rows[:, :4] = box[sel]
        rows[:, 4] = classid[sel].astype(F32)
        rows[:, 5] = score[sel]
    return rows, cand_index[sel] if len(sel) else np.zeros(0, np.int64)


def filter_detections(pred, windows, thresh, iou_thr=IOU_THRESHOLD, max_detection=MAX_DETECTION):
    """-> det_out [B, max_detection, 6], zero padded (yolo3_net_pos.py:615-628)."""
    B = pred['conf'][0].shape[0]
    out = np.zeros((B, max_detection, 6), F32)
    for i in range(B):
        box, cid, sc = decode_candidates(pred, i, windows[i])
        rows, _ = select_detections(box, cid, sc, thresh, iou_thr, max_detection)
        out[i, :len(rows)] = rows
    return out


# --------------------------------------------------------------------------------------
# position-sensitive mask assembly  (yolo3_net_pos.py:862-938)
# --------------------------------------------------------------------------------------
def bin_edges(pb, k=K_MAP):
    """grid_x / grid_y of assemble_kmask_from_box (:881-897) for a rounded box
    pb=[y1,x1,y2,x2] (float32 integers).  tf.cast truncates, tf.round is half-to-even."""
    y1, x1, y2, x2 = [F32(v) for v in pb]
    sub_w = F32(F32(x2 - x1) / F32(k))
    sub_h = F32(F32(y2 - y1) / F32(k))
    gx = [int(x1)] + [int(np.rint(F32(x1 + F32(F32(j) * sub_w)))) for j in range(1, k)] + [int(x2)]
    gy = [int(y1)] + [int(np.rint(F32(y1 + F32(F32(j) * sub_h)))) for j in range(1, k)] + [int(y2)]
    return gx, gy


def assemble_masks(boxes, score_map, k=K_MAP):
    """val_test for one image (:869-933).  boxes [m,6] (zero rows allowed), score_map [S,S,k*k].
    Returns (det_box [n,6], det_mask [n,S,S] fp32) or (det_box [0,6], 0.0) when n == 0."""
    S = score_map.shape[1]                                     # :873 (tf.shape(pred_masks)[1])
    pb = np.rint(boxes[:, :4].astype(F32) * F32(S)).astype(F32)      # :876 half-to-even
    keep = np.nonzero(((pb[:, 2] - pb[:, 0]) > 0) & ((pb[:, 3] - pb[:, 1]) > 0))[0]   # :877-878
    props = boxes[keep]
    pb = pb[keep]
    if props.size == 0:
        return props, F32(0.0)                                 # :933
    masks = np.zeros((len(keep), score_map.shape[0], S), F32)  # logits; outside box -> 0
    for n in range(len(keep)):
        gx, gy = bin_edges(pb[n], k)
        for by in range(k):
            for bx in range(k):
                ys, ye, xs, xe = gy[by], gy[by + 1], gx[bx], gx[bx + 1]
                if ye > ys and xe > xs:
                    masks[n, ys:ye, xs:xe] = score_map[ys:ye, xs:xe, by * k + bx]   # :910-926
    return props, sigmoid(masks)                               # :928  (outside box: 0.5)


def val_test(box_out, mask_pos, k=K_MAP):
    det_box, det_mask = [], []
    for i in range(box_out.shape[0]):
        b, m = assemble_masks(box_out[i], mask_pos[i], k)
        det_box.append(b); det_mask.append(m)
    return det_box, det_mask


# --------------------------------------------------------------------------------------
# the equivalent of sess.run(net.evaluation)  (calculate_test_map.py:214-218)
# --------------------------------------------------------------------------------------
def evaluate(images, windows, thresh, W, anchors=ANCHORS, iou_thr=IOU_THRESHOLD,
             max_detection=MAX_DETECTION, k=K_MAP, acts=None):
    yolos, mask_pos = forward_network(images, W, acts=acts)
    pred = interpret_output(yolos, anchors)
    det = filter_detections(pred, windows, thresh, iou_thr, max_detection)
    det_box, det_mask = val_test(det, mask_pos, k)
    return dict(yolos=yolos, mask_pos=mask_pos, pred=pred, detections=det,
                det_box=det_box, det_mask=det_mask)


# --------------------------------------------------------------------------------------
# host-side helpers of the callers (used as KAT targets; validation_map.py:200-226)
# --------------------------------------------------------------------------------------
def correct_yolo_boxes(x1, y1, x2, y2, image_h, image_w, net_h, net_w):
    """utils/validation_map.py:200-217 / calculate_test_map.py:121-138."""
    if (float(net_w) / image_w) < (float(net_h) / image_h):
        new_w = net_w
        new_h = (image_h * net_w) // image_w
    else:
        new_h = net_h
        new_w = (image_w * net_h) // image_h
    x_offset, x_scale = float((net_w - new_w) // 2) / net_w, float(new_w) / net_w
    y_offset, y_scale = float((net_h - new_h) // 2) / net_h, float(new_h) / net_h

    def fix(v, off, sc, lim):
        return max(min(int(np.around((v - off) / sc * lim).astype(np.int32)), lim), 0)
    return (fix(x1, x_offset, x_scale, image_w), fix(y1, y_offset, y_scale, image_h),
            fix(x2, x_offset, x_scale, image_w), fix(y2, y_offset, y_scale, image_h))


def letterbox_window(img_h, img_w, size):
    """Clip window of utils/val_data.py:36-63 for an img_h x img_w image letterboxed to size."""
    if (float(size) / img_w) < (float(size) / img_h):
        img_h = (img_h * size) // img_w
        img_w = size
    else:
        img_w = (img_w * size) // img_h
        img_h = size
    top = (size - img_h) // 2
    left = (size - img_w) // 2
    return np.array([top / size, left / size, (img_h + top) / size, (img_w + left) / size], F32)
