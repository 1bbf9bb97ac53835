"""CPU oracle for the DIS-YOLO TRAINING step  --  TEST INFRASTRUCTURE ONLY (see dis_oracle.py).

Restates, with torch-CPU fp32 tensors so that torch.autograd supplies the reference gradients, what
`sess.run([net.total_loss, optimizer])` computes in the reference (train_yolo3_mask.py:216):
  forward in training mode  yolo/yolo3_net_pos.py:71-151,153-463 (batch-stat BN for unlocked layers)
  loss_yolo                 :631-747
  loss_mask                 :750-860   (the unseeded tf.random_shuffle of :781-782 is replaced by
                                        permutations passed in by the caller)
  total_loss                :61        (+ l2_regularizer(1e-4) over unlocked weights and biases, :38)
  Adam                      train_yolo3_mask.py:55 (lr 1e-4, beta1 .9, beta2 .999, eps 1e-8, TF form)
PARITY STATUS: parity unpinned (TensorFlow semantics restated; the reference has no training test).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as Fnn

from . import dis_oracle as O

OBJECT_SCALE, NOOBJECT_SCALE, CLASS_SCALE, COORD_SCALE, MASK_SCALE = 2.0, 1.0, 1.0, 1.0, 5.0   # config.py:49-54
IGNORE_THRESH = 0.5        # config.py:57
L2_SCALE = 1e-4            # yolo3_net_pos.py:38
MAX_BOX = 20               # config.py:69
NP_DT = np.float32         # set to np.float64 to obtain a high-precision reference of the same graph
EMULATE_BF16 = False       # True: round where the tensor-core engine stores bf16 (conv operands, z, y),
                           # straight-through in the backward pass -- "identical inputs" oracle for the
                           # mixed-precision training step (leaky masks then agree with the engine's)


def _rb(x):
    """bf16 round-to-nearest-even with a straight-through gradient."""
    if not EMULATE_BF16:
        return x
    return x + (x.detach().float().bfloat16().to(x.dtype) - x.detach())


# inputs of convolutional53..82: n -> (src0, src1); src1's output is 2x-upsampled and concatenated AFTER
# src0's (yolo3_net_pos.py:258-412; concat order [skip, up], :291,326,387,402)
TOPOLOGY_53_82 = {53: (52, 0), 54: (53, 0), 55: (54, 0), 56: (55, 0), 57: (56, 0), 58: (57, 0), 59: (58, 0),
                  60: (57, 0), 61: (43, 60), 62: (61, 0), 63: (62, 0), 64: (63, 0), 65: (64, 0), 66: (65, 0),
                  67: (66, 0), 68: (65, 0), 69: (26, 68), 70: (69, 0), 71: (70, 0), 72: (71, 0), 73: (72, 0),
                  74: (73, 0), 75: (74, 0), 76: (73, 0), 77: (9, 76), 78: (77, 0), 79: (78, 0), 80: (4, 79),
                  81: (80, 0), 82: (81, 0)}


def _conv_same(x, w, stride):
    """x NCHW, w HWIO -> NCHW; TF 'SAME' padding."""
    k = w.shape[0]
    pt, pb = O.same_pad(x.shape[2], k, stride)
    pl, pr = O.same_pad(x.shape[3], k, stride)
    return Fnn.conv2d(Fnn.pad(x, (pl, pr, pt, pb)), w.permute(3, 2, 0, 1), stride=stride)


def conv_backward(x, dz, w, stride=1):
    """What TF's autodiff emits for tf.nn.conv2d (yolo3_net_pos.py:125,142): Conv2DBackpropInput and
    Conv2DBackpropFilter.  x [B,H,W,cin], dz [B,Ho,Wo,cout], w HWIO (numpy) -> (dx NHWC, dw HWIO),
    float64 on the CPU through torch.autograd of the same 'SAME'-padded convolution."""
    xt = torch.from_numpy(np.ascontiguousarray(x, np.float64)).permute(0, 3, 1, 2).requires_grad_(True)
    wt = torch.from_numpy(np.ascontiguousarray(w, np.float64)).requires_grad_(True)
    y = _conv_same(xt, wt, stride)
    y.backward(torch.from_numpy(np.ascontiguousarray(dz, np.float64)).permute(0, 3, 1, 2))
    return xt.grad.permute(0, 2, 3, 1).numpy(), wt.grad.numpy()


def forward_train(images, P, lock, stats_out=None, acts_out=None, z_out=None):
    """Training-mode forward.  images [B,H,W,3] numpy; P: dict name -> torch tensor (leaf tensors
    with requires_grad for trainables).  Unlocked BN layers use batch moments over (N,H,W)
    (yolo3_net_pos.py:88-98); locked ones the moving statistics (:76-81).
    Returns (yolos [3 x [B,g,g,3,8]], mask_pos [B,S,S,9]) as torch tensors (NHWC)."""
    t = O.layer_table()
    x0 = torch.from_numpy(np.ascontiguousarray(images, NP_DT)).permute(0, 3, 1, 2)

    def run(n, x, shortcut=None):
        L = t[n]
        y = _conv_same(_rb(x) if n == 1 else x, _rb(P[O.vname(n, 'w')]), L['s'])
        if L['bn']:
            g, b = P[O.vname(n, 'gamma')].view(1, -1, 1, 1), P[O.vname(n, 'beta')].view(1, -1, 1, 1)
            if lock[n]:
                m, v = P[O.vname(n, 'mean')].view(1, -1, 1, 1), P[O.vname(n, 'var')].view(1, -1, 1, 1)
            else:
                y = _rb(y)                 # the engine keeps the pre-BN output z in bf16
                if z_out is not None:
                    z_out[n] = y.detach().permute(0, 2, 3, 1).numpy().copy()
                m = y.mean(dim=(0, 2, 3), keepdim=True)
                v = ((y - m) ** 2).mean(dim=(0, 2, 3), keepdim=True)
                if stats_out is not None:
                    stats_out[n] = (m.detach().reshape(-1).numpy().copy(), v.detach().reshape(-1).numpy().copy())
            y = (y - m) * torch.rsqrt(v + O.BN_EPS) * g + b
            y = torch.maximum(O.ALPHA * y, y)
        else:
            y = y + P[O.vname(n, 'b')].view(1, -1, 1, 1)
        if L['res']:
            y = y + shortcut
        if L['bn']:
            y = _rb(y)                     # bf16 activations; the biased linear heads stay fp32
        if acts_out is not None:
            if y.requires_grad:
                y.retain_grad()
            acts_out[n] = y
        return y

    up = lambda z: z.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    net = run(1, x0); net = run(2, net)
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
    net = torch.cat([skip5, up(net)], 1)
    for n in range(61, 66):
        net = run(n, net)
    y2 = run(67, run(66, net))
    net = run(68, net)
    net = torch.cat([skip4, up(net)], 1)
    for n in range(69, 74):
        net = run(n, net)
    y3 = run(75, run(74, net))
    net = run(76, net)
    net = torch.cat([skip3, up(net)], 1)
    net = run(77, net); net = run(78, net); net = run(79, net)
    net = torch.cat([skip2, up(net)], 1)
    net = run(80, net); net = run(81, net)
    mp = run(82, net)

    def r5(y):
        y = y.permute(0, 2, 3, 1)
        return y.reshape(y.shape[0], y.shape[1], y.shape[2], 3, 8)
    return [r5(y3), r5(y2), r5(y1)], mp.permute(0, 2, 3, 1)


def loss_yolo(yolos, true_boxes, labels, anchors=O.ANCHORS):
    """yolo3_net_pos.py:631-747.  yolos: stride 8/16/32 maps (torch); labels: [yolo3, yolo2, yolo1]
    numpy [B,g,g,3,8]; true_boxes numpy [B,1,1,1,20,5].  Returns dict of scalar tensors."""
    B = yolos[0].shape[0]
    net = yolos[2].shape[1] * 32
    tb = torch.from_numpy(np.ascontiguousarray(true_boxes, NP_DT))
    out = dict(obj=0., noobj=0., cls=0., xy=0., wh=0.)
    for i in range(3):
        p = yolos[i]
        g = p.shape[1]
        lab = torch.from_numpy(np.ascontiguousarray(labels[i], NP_DT))
        a = torch.from_numpy(anchors[3 * i:3 * i + 3].astype(NP_DT))
        cxy = torch.sigmoid(p[..., :2])
        cell_x = torch.arange(g, dtype=tb.dtype).view(1, 1, g, 1)
        cell_y = torch.arange(g, dtype=tb.dtype).view(1, g, 1, 1)
        cell = torch.stack([cell_x.expand(1, g, g, 3), cell_y.expand(1, g, g, 3)], -1)
        box_xy = (cell + cxy) / float(g)
        box_wh = torch.exp(p[..., 2:4]) * a.view(1, 1, 1, 3, 2) / float(net)
        # ignore mask (:657-680): best IoU of each predicted box with the <=20 true boxes
        pxy, pwh = box_xy.unsqueeze(4), box_wh.unsqueeze(4)
        pmin, pmax = pxy - pwh / 2., pxy + pwh / 2.
        txy, twh = tb[..., 0:2], tb[..., 2:4]
        tmin, tmax = txy - twh / 2., txy + twh / 2.
        iwh = torch.clamp(torch.minimum(pmax, tmax) - torch.maximum(pmin, tmin), min=0.)
        inter = iwh[..., 0] * iwh[..., 1]
        union = torch.clamp(pwh[..., 0] * pwh[..., 1] + twh[..., 0] * twh[..., 1] - inter, min=1e-10)
        best = torch.clamp(inter / union, 0., 1.).max(dim=4)[0]
        ignore = (best < IGNORE_THRESH).float().unsqueeze(4).detach()
        obj = lab[..., 4:5]
        noobj = 1. - obj
        conf = p[..., 4:5]
        bce = Fnn.binary_cross_entropy_with_logits(conf, obj, reduction='none')
        out['obj'] = out['obj'] + (obj * bce * OBJECT_SCALE).sum(dim=(1, 2, 3, 4)).mean()
        out['noobj'] = out['noobj'] + (ignore * noobj * bce * NOOBJECT_SCALE).sum(dim=(1, 2, 3, 4)).mean()
        tcls = lab[..., 5:].argmax(-1)
        ce = Fnn.cross_entropy(p[..., 5:].reshape(-1, 3), tcls.reshape(-1), reduction='none').view(tcls.shape)
        out['cls'] = out['cls'] + (obj[..., 0] * ce * CLASS_SCALE).sum(dim=(1, 2, 3)).mean()
        tcoord = lab[..., 0:4]
        true_cxy = tcoord[..., 0:2] * float(g) - cell
        true_twh = torch.clamp(torch.log(tcoord[..., 2:4] * float(net) / a.view(1, 1, 1, 3, 2)), -1e2, 1e2)
        whs = (2. - tcoord[..., 2] * tcoord[..., 3]).unsqueeze(4)
        dxy = obj * (cxy - true_cxy)
        dwh = obj * (p[..., 2:4] - true_twh)
        out['xy'] = out['xy'] + (dxy ** 2 * whs ** 2 * COORD_SCALE).sum(dim=(1, 2, 3, 4)).mean()
        out['wh'] = out['wh'] + (dwh ** 2 * whs ** 2 * COORD_SCALE).sum(dim=(1, 2, 3, 4)).mean()
    return out


def overlaps(b1, b2):
    """overlaps_graph (:954-975): pairwise IoU of y1x1y2x2 boxes, no epsilon."""
    b1 = np.asarray(b1, np.float32)[:, None, :]
    b2 = np.asarray(b2, np.float32)[None, :, :]
    y1 = np.maximum(b1[..., 0], b2[..., 0]); x1 = np.maximum(b1[..., 1], b2[..., 1])
    y2 = np.minimum(b1[..., 2], b2[..., 2]); x2 = np.minimum(b1[..., 3], b2[..., 3])
    inter = np.maximum(x2 - x1, 0) * np.maximum(y2 - y1, 0)
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    with np.errstate(divide='ignore', invalid='ignore'):
        return (inter / (a1 + a2 - inter)).astype(np.float32)


def mask_rois(det_i, true_boxes_i, perm_prop, perm_gt):
    """Steps 1-4 of loss_mask for one image (:757-796).  det_i [max_det,6]; true_boxes_i [20,5];
    perm_prop / perm_gt: the permutations standing in for tf.random_shuffle.
    Returns (positive rois [n,4] y1x1y2x2, gt index per roi [n] into the trimmed gt list,
    indices of the kept gt rows)."""
    props = det_i[:, :4]
    props = props[np.abs(props).sum(1) > 0]
    gt = true_boxes_i[:, :4]
    keep = np.nonzero(np.abs(gt).sum(1) > 0)[0]
    gt = gt[keep]
    half = np.float32(2.0)
    gtb = np.stack([gt[:, 1] - gt[:, 3] / half, gt[:, 0] - gt[:, 2] / half,
                    gt[:, 1] + gt[:, 3] / half, gt[:, 0] + gt[:, 2] / half], 1).astype(np.float32)
    pp = [i for i in perm_prop if i < len(props)]
    pg = [i for i in perm_gt if i < len(gtb)]
    rois = np.concatenate([props[pp][:7], gtb[pg][:3]], 0).astype(np.float32)
    if len(rois) == 0 or len(gtb) == 0:
        return np.zeros((0, 4), np.float32), np.zeros(0, np.int64), keep
    ov = overlaps(rois, gtb)
    ov = np.nan_to_num(ov, nan=-1.0)
    pos = np.nonzero(ov.max(1) >= 0.5)[0]
    return rois[pos], ov[pos].argmax(1), keep


def loss_mask(detections, mask_pos, true_boxes, true_masks, perms, k=3):
    """yolo3_net_pos.py:750-860.  detections numpy [B,max_det,6]; mask_pos torch [B,S,S,9];
    true_masks bool numpy [B,20,H,W]; perms: list (per image) of (perm_prop, perm_gt)."""
    B, S = mask_pos.shape[0], mask_pos.shape[1]
    total = 0.
    for i in range(B):
        rois, assign, keep = mask_rois(detections[i], true_boxes[i, 0, 0, 0], perms[i][0], perms[i][1])
        if len(rois) == 0:
            continue
        gm = true_masks[i][keep].astype(np.float32)
        f = gm.shape[1] // S
        gm = np.rint(gm[:, ::f, ::f])                     # legacy bilinear resize at an integer factor
        pr = np.rint(rois * np.float32(S)).astype(np.float32)
        per_roi = []
        for r in range(len(rois)):
            gx, gy = O.bin_edges(pr[r], k)
            chan = torch.zeros((S, S), dtype=torch.long)
            inside = torch.zeros((S, S), dtype=mask_pos.dtype)
            for by in range(k):
                for bx in range(k):
                    ys, ye, xs, xe = gy[by], gy[by + 1], gx[bx], gx[bx + 1]
                    if ye > ys and xe > xs:
                        chan[ys:ye, xs:xe] = by * k + bx
                        inside[ys:ye, xs:xe] = 1.
            logit = torch.gather(mask_pos[i], 2, chan.unsqueeze(-1))[..., 0] * inside
            tgt = torch.from_numpy(gm[assign[r]].astype(NP_DT))
            bce = Fnn.binary_cross_entropy_with_logits(logit, tgt, reduction='none')
            per_roi.append((inside * bce).sum() / inside.sum())
        total = total + MASK_SCALE * torch.stack(per_roi).mean()
    return total / B


def l2_loss(P, lock):
    t = O.layer_table()
    tot = 0.
    for n in range(1, 83):
        if lock[n]:
            continue
        tot = tot + L2_SCALE * 0.5 * (P[O.vname(n, 'w')] ** 2).sum()
        if not t[n]['bn']:
            tot = tot + L2_SCALE * 0.5 * (P[O.vname(n, 'b')] ** 2).sum()
    return tot


def trainable_names(lock):
    t = O.layer_table()
    names = []
    for n in range(1, 83):
        if lock[n]:
            continue
        names.append(O.vname(n, 'w'))
        names += [O.vname(n, 'gamma'), O.vname(n, 'beta')] if t[n]['bn'] else [O.vname(n, 'b')]
    return names


def train_step(images, W, lock, labels, true_boxes, true_masks, perms, det_thresh=0.25, lr=1e-4, adam=None,
               step=1, windows=None, apply=True):
    """One training step.  W: dict of numpy weights (updated copy returned).  Returns
    (losses dict of floats, grads dict of numpy, new W, adam state, batch stats)."""
    P = {k: torch.from_numpy(np.array(v, NP_DT)) for k, v in W.items()}
    names = trainable_names(lock)
    for nm in names:
        P[nm].requires_grad_(True)
    stats = {}
    acts = {}
    zs = {}
    yolos, mp = forward_train(images, P, lock, stats, acts, zs)
    ly = loss_yolo(yolos, true_boxes, labels)
    B = images.shape[0]
    if windows is None:
        windows = np.tile(np.array([[0, 0, 1, 1]], np.float32), (B, 1))
    pred = O.interpret_output([y.detach().numpy().astype(np.float32) for y in yolos])
    det = O.filter_detections(pred, windows, det_thresh)
    lm = loss_mask(det, mp, true_boxes, true_masks, perms)
    reg = l2_loss(P, lock)
    conf = ly['obj'] + ly['noobj']
    total = conf + ly['cls'] + ly['xy'] + ly['wh'] + lm + reg
    total.backward()
    losses = dict(total=float(total), obj=float(ly['obj']), noobj=float(ly['noobj']), cls=float(ly['cls']),
                  xy=float(ly['xy']), wh=float(ly['wh']), mask=float(lm), l2=float(reg))
    grads = {nm: (P[nm].grad.numpy().copy() if P[nm].grad is not None else np.zeros_like(W[nm])) for nm in names}
    newW = {k: np.array(v, np.float32) for k, v in W.items()}
    if adam is None:
        adam = {nm: (np.zeros_like(W[nm], np.float32), np.zeros_like(W[nm], np.float32)) for nm in names}
    if apply:
        b1, b2, eps = 0.9, 0.999, 1e-8
        lr_t = lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        for nm in names:
            m, v = adam[nm]
            g = grads[nm]
            m = b1 * m + (1 - b1) * g
            v = b2 * v + (1 - b2) * g * g
            adam[nm] = (m.astype(np.float32), v.astype(np.float32))
            newW[nm] = (newW[nm] - lr_t * m / (np.sqrt(v) + eps)).astype(np.float32)
        for n, (m, v) in stats.items():     # moving averages (:92-95), decay 0.997
            newW[O.vname(n, 'mean')] = (W[O.vname(n, 'mean')] * O.BN_DECAY + m * (1 - O.BN_DECAY)).astype(np.float32)
            newW[O.vname(n, 'var')] = (W[O.vname(n, 'var')] * O.BN_DECAY + v * (1 - O.BN_DECAY)).astype(np.float32)
    act_grads = {n: a.grad.permute(0, 2, 3, 1).numpy().copy() for n, a in acts.items() if a.grad is not None}
    return losses, grads, newW, adam, dict(stats=stats, detections=det, act_grads=act_grads, z=zs,
                                           acts={n: a.detach().permute(0, 2, 3, 1).numpy().copy() for n, a in acts.items()},
                                           yolos=[y.detach().numpy() for y in yolos], mask_pos=mp.detach().numpy())


def make_labels(rng, B, size, n_boxes=4, anchors=O.ANCHORS, num_class=3):
    """Synthetic labels in the format of utils/train_data.py:44-52,146-178,258-265: returns
    (labels [yolo3, yolo2, yolo1], true_boxes [B,1,1,1,20,5], true_masks bool [B,20,size,size])."""
    base = size // 32
    yolos = [np.zeros((B, base * m, base * m, 3, 5 + num_class), np.float32) for m in (4, 2, 1)]   # stride 8,16,32
    tb = np.zeros((B, 1, 1, 1, MAX_BOX, 5), np.float32)
    tm = np.zeros((B, MAX_BOX, size, size), bool)
    for b in range(B):
        for j in range(n_boxes):
            w, h = rng.uniform(0.15, 0.6, 2) * size
            xc, yc = rng.uniform(w / 2, size - w / 2), rng.uniform(h / 2, size - h / 2)
            cls = int(rng.integers(0, num_class))
            tb[b, 0, 0, 0, j] = [xc / size, yc / size, w / size, h / size, cls]
            x1, y1, x2, y2 = int(xc - w / 2), int(yc - h / 2), int(xc + w / 2), int(yc + h / 2)
            yy, xx = np.mgrid[0:size, 0:size]
            tm[b, j] = ((xx - xc) / (w / 2)) ** 2 + ((yy - yc) / (h / 2)) ** 2 <= 1.0       # filled ellipse
            # best anchor by IoU of centred boxes (train_data.py:149-165)
            inter = np.minimum(w, anchors[:, 0]) * np.minimum(h, anchors[:, 1])
            iou = inter / (w * h + anchors[:, 0] * anchors[:, 1] - inter)
            a = int(np.argmax(iou))
            scale = 2 - a // 3                       # anchors 0-2 -> stride 8 map (index 0 in `yolos`)
            lab = yolos[a // 3]
            g = lab.shape[1]
            xi, yi = int(xc * g / size), int(yc * g / size)
            if lab[b, yi, xi, a % 3, 4] == 1:
                continue
            lab[b, yi, xi, a % 3, 0:4] = [xc / size, yc / size, w / size, h / size]
            lab[b, yi, xi, a % 3, 4] = 1
            lab[b, yi, xi, a % 3, 5 + cls] = 1
    return yolos, tb, tm
