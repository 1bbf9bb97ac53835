"""GPU: the training step (training-mode forward with batch-statistics BN, loss_yolo, loss_mask, L2,
backward, Adam, moving averages) against the torch-autograd oracle on the same seeded inputs.

The oracle runs the same graph in float64 (ground truth).  Batch-statistics BN backward is
ill-conditioned in fp32 (g - mean(g) - xhat*mean(g*xhat) cancels): measured on B200, this fp32
engine is 2e-3..6e-3 (norm-wise) from the float64 gradients below the first BN layer, the torch
fp32 restatement of the same graph 3e-3..9e-3.  Tolerances: losses 1e-4 relative; gradients of the
head convs (no BN in the path) 1e-4, every other tensor 1.5e-2 norm-wise against float64;
parameters after an Adam step: within 2.5*lr everywhere (Adam's first step is lr*sign(g)) and
within 5e-6 for 98 % of the weight entries."""
import numpy as np
import pytest

from oracle import dis_oracle as O
from oracle import dis_oracle_train as T
from tests.util import bf16_round, rel_err

pytestmark = pytest.mark.gpu


def _setup(B=3, size=160, seed=0, thresh=0.1):
    rng = np.random.default_rng(seed)
    W = O.make_weights('lively', seed + 1)
    img = rng.random((B, size, size, 3), dtype=np.float32)
    labels, tb, tm = T.make_labels(rng, B, size)
    perm_prop = np.stack([rng.permutation(30) for _ in range(B)]).astype(np.int32)
    perm_gt = np.stack([rng.permutation(20) for _ in range(B)]).astype(np.int32)
    return W, img, labels, tb, tm, perm_prop, perm_gt, thresh


def test_train_step_matches_autograd_oracle():
    import disyolo_b200 as dy
    W, img, labels, tb, tm, pp, pg, thresh = _setup()
    B, size = img.shape[0], img.shape[1]
    lock = O.default_lock_flags()
    eng = dy.Engine(image_size=size, max_batch=B, precision='fp32')
    eng.load_weights(W)
    n = eng.train_init()
    assert n == 21070737                       # stage-1 trainables
    perms = [(pp[b].tolist(), pg[b].tolist()) for b in range(B)]
    adam, Wo = None, W
    T.NP_DT = np.float64
    for step in (1, 2):
        losses = eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
        ol, og, Wn, adam, aux = T.train_step(img, Wo, lock, labels, tb, tm, perms, det_thresh=thresh, lr=1e-4,
                                             adam=adam, step=step)
        want = np.array([ol[k] for k in ('total', 'obj', 'noobj', 'cls', 'xy', 'wh', 'mask', 'l2')])
        print('step', step, 'losses', losses, 'oracle', want, 'detections', (aux['detections'][..., 5] > 0).sum(1))
        assert np.allclose(losses, want, rtol=1e-4, atol=1e-5)
        g = eng.train_backward(82, 1).cpu().numpy()
        worst, errs = 0.0, {}
        for layer in range(53, 83):
            off, cnt = eng.layer_span(layer)
            L = O.layer_table()[layer]
            nw = L['k'] ** 2 * L['cin'] * L['cout']
            names = ['w', 'gamma', 'beta'] if L['bn'] else ['w', 'b']
            sizes = [nw] + [L['cout']] * (len(names) - 1)
            assert cnt == sum(sizes)
            o = off
            for nm, sz in zip(names, sizes):
                ref = og[O.vname(layer, nm)].reshape(-1)
                e = rel_err(g[o:o + sz], ref)
                errs['%d%s' % (layer, nm)] = e
                worst = max(worst, e)
                o += sz
        print('gradient rel errs:', ' '.join('%s:%.1e' % kv for kv in errs.items()))
        print('worst gradient rel err %.3g' % worst)
        for key, e in errs.items():
            head = key.startswith(('59', '67', '75', '82'))
            assert e < (1e-3 if head else 1.5e-2), 'grad %s rel err %.3g' % (key, e)
        eng.train_apply(1e-4)
        Wgpu = {}
        for layer in range(53, 83):
            L = O.layer_table()[layer]
            for nm in (['w', 'gamma', 'beta', 'mean', 'var'] if L['bn'] else ['w', 'b']):
                name = O.vname(layer, nm)
                got = eng.get_weights(name, Wn[name].shape)
                Wgpu[name] = got
                d = np.abs(got - Wn[name])
                if nm in ('mean', 'var'):
                    assert np.all(d <= 2e-6 + 1e-4 * np.abs(Wn[name])), name
                else:
                    assert d.max() <= 2.5e-4, name
                    if nm == 'w':
                        assert np.mean(d > 5e-6) < 0.02, (name, float(np.mean(d > 5e-6)))
        # continue both sides from the SAME parameters so that step 2 is again a clean comparison
        Wo = dict(Wn)
        Wo.update(Wgpu)
    T.NP_DT = np.float32
    eng.close()


def test_loss_decreases_over_steps():
    """Property test: a few Adam steps on a fixed batch reduce the total loss."""
    import disyolo_b200 as dy
    W, img, labels, tb, tm, pp, pg, thresh = _setup(seed=3)
    eng = dy.Engine(image_size=img.shape[1], max_batch=img.shape[0], precision='fp32')
    eng.load_weights(W)
    eng.train_init()
    hist = []
    for _ in range(6):
        hist.append(float(eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)[0]))
        eng.train_backward()
        eng.train_apply(1e-3)
    print('loss history', hist)
    assert hist[-1] < hist[0]
    eng.close()


def test_session_training_protocol():
    """The reference's training call: sess.run([net.total_loss, optimizer], feed_dict)
    (train_yolo3_mask.py:146-149,216) through the YOLONet / Session / AdamOptimizer mirror."""
    import disyolo_b200.yolo.config as cfg
    from disyolo_b200.yolo.yolo3_net_pos import YOLONet, Session, AdamOptimizer
    cfg.BATCH_SIZE, cfg.IMAGE_SIZE = 2, 128
    try:
        rng = np.random.default_rng(5)
        net = YOLONet(True)
        opt = AdamOptimizer(learning_rate=1e-3).minimize(net.total_loss)
        sess = Session(net, seed=0)
        sess.restore(O.make_weights('lively', 2))
        img = rng.random((2, 128, 128, 3), dtype=np.float32)
        labels, tb, tm = T.make_labels(rng, 2, 128)
        feed = {net.images: img, net.yolo1: labels[2], net.yolo2: labels[1], net.yolo3: labels[0],
                net.true_boxes: tb, net.true_masks: tm, net.is_training: True,
                net.det_thresh: [cfg.OBJ_THRESHOLD], net.clip_window: np.tile([[0., 0., 1., 1.]], (2, 1))}
        hist = []
        for _ in range(4):
            loss, _none = sess.run([net.total_loss, opt], feed_dict=feed)
            hist.append(loss)
        assert all(np.isfinite(hist)) and hist[-1] < hist[0]
        assert set(sess.last_losses) == {'total', 'object', 'noobject', 'class', 'xy', 'wh', 'mask', 'l2'}
        # evaluation with the updated weights still works on the same net (validation loop, :163-176)
        det_box, det_mask = sess.run(net.evaluation, feed_dict={net.is_training: False, net.det_thresh: [0.1],
                                                                net.clip_window: feed[net.clip_window],
                                                                net.images: img})
        assert len(det_box) == 2
    finally:
        cfg.BATCH_SIZE, cfg.IMAGE_SIZE = 2, 576


def _dp_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    import disyolo_b200 as dy
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    W, img, labels, tb, tm, pp, pg, thresh = _setup(B=2, size=128, seed=10 + rank)
    W = O.make_weights('lively', 1)                    # identical weights on every rank
    eng = dy.Engine(image_size=128, max_batch=2, precision='fp32', device=rank)
    eng.load_weights(W)
    tr = dy.DataParallelTrainer(eng, bucket_mb=8)
    # reference for this rank: its own gradients before averaging
    eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
    g_local = eng.train_backward(82, 1).clone()
    glist = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(glist, g_local)
    losses = tr.step(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh, 1e-4)
    torch.cuda.synchronize()
    g_sum = eng.grad_flat.clone()                      # after the bucketed all-reduce: sum over ranks
    want = sum(glist)
    err = float((g_sum - want).norm() / want.norm())
    w53 = eng.get_weights('yolo/convolutional53/weights', (1, 1, 1024, 512))
    q.put((rank, err, float(np.abs(w53).sum()), [float(v) for v in losses], len(tr.buckets)))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_two_gpus():
    """NCCL data parallelism: bucketed all-reduce overlapped with backward gives sum-over-ranks
    gradients (wgrad atomics make each rank's gradients reproducible only to fp32 rounding) and
    identical parameters on every rank."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from tests.test_parallel_cpu import _free_port
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res.sort()
    print('dp results', res)
    assert res[0][1] < 1e-4 and res[1][1] < 1e-4            # all-reduced gradient == sum of rank gradients
    assert res[0][2] == res[1][2]                           # identical parameters after the step
    assert res[0][3] == res[1][3] and res[0][4] >= 2        # averaged losses agree; > 1 bucket


# ---------------------------------------------------------------------------------------------
# tensor-core (bf16) training engine
# ---------------------------------------------------------------------------------------------
def _grad_table(eng, g, og):
    errs, cos = {}, {}
    for layer in range(53, 83):
        off, cnt = eng.layer_span(layer)
        L = O.layer_table()[layer]
        nw = L['k'] ** 2 * L['cin'] * L['cout']
        names = ['w', 'gamma', 'beta'] if L['bn'] else ['w', 'b']
        sizes = [nw] + [L['cout']] * (len(names) - 1)
        o = off
        for nm, sz in zip(names, sizes):
            ref = og[O.vname(layer, nm)].reshape(-1).astype(np.float64)
            got = g[o:o + sz].astype(np.float64)
            errs['%d%s' % (layer, nm)] = rel_err(got, ref)
            cos['%d%s' % (layer, nm)] = float(got @ ref / max(np.linalg.norm(got) * np.linalg.norm(ref), 1e-30))
            o += sz
    return errs, cos


def _local_layer_oracle(n, x_in, w_bf, z_eng, dy_eng, gamma, beta, bias, bn):
    """float64 restatement of ONE layer's training-mode forward and backward on the engine's own
    inputs (x_in = the engine's input activations, z_eng = its stored pre-BN output, dy_eng = its
    gradient w.r.t. the layer output).  Returns dict(z, y, dz, dw, dx, dgamma, dbeta / dbias)."""
    import torch
    out = {}
    xt = torch.from_numpy(np.ascontiguousarray(x_in, np.float64)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(w_bf, np.float64))
    z = T._conv_same(xt, wt, 1)
    out['z'] = z.permute(0, 2, 3, 1).numpy()
    dy = torch.from_numpy(np.ascontiguousarray(dy_eng, np.float64)).permute(0, 3, 1, 2)
    if bn:
        ze = torch.from_numpy(np.ascontiguousarray(z_eng, np.float64)).permute(0, 3, 1, 2).requires_grad_(True)
        g = torch.from_numpy(np.asarray(gamma, np.float64)).requires_grad_(True)
        b = torch.from_numpy(np.asarray(beta, np.float64)).requires_grad_(True)
        m = ze.mean(dim=(0, 2, 3), keepdim=True)
        v = ((ze - m) ** 2).mean(dim=(0, 2, 3), keepdim=True)
        y = (ze - m) * torch.rsqrt(v + O.BN_EPS) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        y = torch.maximum(O.ALPHA * y, y)
        y.backward(dy)
        out['y'] = y.detach().permute(0, 2, 3, 1).numpy()
        out['dz'] = ze.grad.permute(0, 2, 3, 1).numpy()
        out['dgamma'], out['dbeta'] = g.grad.numpy(), b.grad.numpy()
        out['mean'], out['var'] = m.detach().reshape(-1).numpy(), v.detach().reshape(-1).numpy()
    else:
        out['y'] = out['z'] + np.asarray(bias, np.float64)
        out['dz'] = np.ascontiguousarray(dy_eng, np.float64)
        out['dbias'] = out['dz'].sum(axis=(0, 1, 2))
    out['dx'], out['dw'] = T.conv_backward(x_in, out['dz'], w_bf)
    return out


def test_train_step_bf16_tensor_core_engine():
    """The training step on the tcgen05 engine (bf16 operands / activations / activation gradients,
    fp32 accumulation, fp32 BN statistics, master weights and Adam).

    End to end only the LOSSES are comparable with the float64 oracle (3e-2): with batch-statistics
    BN at random init the network is chaotic in its rounding -- two bf16 forwards that differ only
    in fp32 accumulation order (this engine vs a float64 oracle rounding to bf16 at the same points,
    T.EMULATE_BF16) agree to 2e-5 after conv1 and are fully decorrelated (5e-3, the bf16 noise
    floor) by layer 40; that flips leaky-ReLU masks and moves weight gradients by 0.4-1.2 norm-wise
    (measured between the two float64 oracles themselves).  The kernels are therefore checked layer
    by layer on IDENTICAL inputs: for every trained layer the float64 oracle recomputes z, y, dz,
    dW, d gamma / d beta / d bias and the input gradient from the engine's own x, z, dy.
    Tolerances (norm-wise): z, y 4e-3 (one bf16 rounding); d gamma, d beta, d bias 1e-2; dW 2e-2 (bf16 dz);
    input gradients 1.5e-2 (bf16 dz, bf16 weights, bf16 store, accumulation over consumers)."""
    import disyolo_b200 as dy
    W, img, labels, tb, tm, pp, pg, thresh = _setup()
    B, size = img.shape[0], img.shape[1]
    lock = O.default_lock_flags()
    eng = dy.Engine(image_size=size, max_batch=B, precision='bf16')
    eng.load_weights(W)
    assert eng.train_init() == 21070737
    perms = [(pp[b].tolist(), pg[b].tolist()) for b in range(B)]
    keys = ('total', 'obj', 'noobj', 'cls', 'xy', 'wh', 'mask', 'l2')
    T.NP_DT = np.float64
    try:
        losses = eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
        ol64 = T.train_step(img, W, lock, labels, tb, tm, perms, det_thresh=thresh, lr=1e-4, apply=False)[0]
        print('bf16 losses', losses, 'float64 oracle', [ol64[k] for k in keys])
        assert np.allclose(losses, [ol64[k] for k in keys], rtol=3e-2, atol=1e-4)
        g = eng.train_backward(82, 1).cpu().numpy()
        assert np.all(np.isfinite(g))
        tab = O.layer_table()
        topo = T.TOPOLOGY_53_82          # n -> (src0, src1)
        act = {n: eng.activation(n, B).cpu().numpy() for n in list(range(53, 83)) + [52, 43, 26, 9, 4]}
        dyt = {n: eng.train_tensor(n, 'dy').cpu().numpy() for n in range(53, 83)}
        up2 = lambda a: a.repeat(2, axis=1).repeat(2, axis=2)
        dx_sum, errs, stats = {}, {}, {}
        for n in range(82, 52, -1):
            L = tab[n]
            s0, s1 = topo[n]
            x_in = act[s0] if s1 == 0 else np.concatenate([act[s0], up2(act[s1])], axis=3)
            w_bf = bf16_round(W[O.vname(n, 'w')])
            z_eng = eng.train_tensor(n, 'z').cpu().numpy() if L['bn'] else None
            r = _local_layer_oracle(n, x_in, w_bf, z_eng, dyt[n], W.get(O.vname(n, 'gamma')), W.get(O.vname(n, 'beta')),
                                    W.get(O.vname(n, 'b')), L['bn'])
            off, cnt = eng.layer_span(n)
            nw = L['k'] ** 2 * L['cin'] * L['cout']
            if L['bn']:
                errs['%dz' % n] = rel_err(z_eng, r['z'])
                errs['%dgamma' % n] = rel_err(g[off + nw:off + nw + L['cout']], r['dgamma'])
                errs['%dbeta' % n] = rel_err(g[off + nw + L['cout']:off + cnt], r['dbeta'])
                stats[n] = (r['mean'], r['var'])
            else:
                errs['%db' % n] = rel_err(g[off + nw:off + cnt], r['dbias'])
            errs['%dy' % n] = rel_err(act[n], r['y'])
            errs['%dw' % n] = rel_err(g[off:off + nw], r['dw'].reshape(-1))
            c0 = act[s0].shape[3]
            if s0 >= 53:
                dx_sum[s0] = dx_sum.get(s0, 0) + r['dx'][..., :c0]
            if s1 >= 53:
                d1 = r['dx'][..., c0:]
                dx_sum[s1] = dx_sum.get(s1, 0) + (d1[:, 0::2, 0::2] + d1[:, 1::2, 0::2] + d1[:, 0::2, 1::2] + d1[:, 1::2, 1::2])
        for p_, ref in dx_sum.items():
            errs['%ddy' % p_] = rel_err(dyt[p_], ref)
        print('local rel errs:', ' '.join('%s:%.1e' % kv for kv in sorted(errs.items(), key=lambda kv: kv[0])))
        for key, e in errs.items():
            if key.endswith('dy'):
                tol = 1.5e-2
            elif key.endswith(('z', 'y')):
                tol = 4e-3
            elif key.endswith('w'):
                tol = 2e-2       # sum_p x_p dz_p with sum_p dz_p = 0 (BN backward): the mean of x cancels in the
                                 # signal but not in dz's bf16 rounding noise (measured: 1.3e-2 on conv61)
            else:
                tol = 1e-2
            assert e < tol, 'layer tensor %s: rel err %.3g (tol %.1e)' % (key, e, tol)
        # ---- Adam on the fp32 master weights, from the engine's own gradients ----
        before = {n: eng.get_weights(O.vname(n, 'w'), W[O.vname(n, 'w')].shape) for n in (53, 58, 61, 75, 81, 82)}
        eng.train_apply(1e-4)
        lr_t = 1e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
        for n, w0 in before.items():
            off, cnt = eng.layer_span(n)
            gw = g[off:off + w0.size].reshape(w0.shape).astype(np.float64) + 1e-4 * w0     # + L2 gradient (:38)
            want = w0 - lr_t * (0.1 * gw) / (np.sqrt(0.001 * gw * gw) + 1e-8)
            got = eng.get_weights(O.vname(n, 'w'), w0.shape)
            assert np.abs(got - want).max() < 2e-6, (n, float(np.abs(got - want).max()))
        for n in (53, 81):                           # moving averages (:92-95) from the batch moments of z
            for nm, i in (('mean', 0), ('var', 1)):
                name = O.vname(n, nm)
                want = W[name] * O.BN_DECAY + stats[n][i] * (1 - O.BN_DECAY)
                assert np.allclose(eng.get_weights(name, W[name].shape), want, rtol=1e-4, atol=1e-6), name
        # ---- the next forward runs on the re-packed bf16 operands of the updated master weights ----
        Wn = dict(W)
        for n in range(53, 83):
            for nm in (['w', 'gamma', 'beta', 'mean', 'var'] if tab[n]['bn'] else ['w', 'b']):
                Wn[O.vname(n, nm)] = eng.get_weights(O.vname(n, nm), W[O.vname(n, nm)].shape)
        l2 = eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
        o2 = T.train_step(img, Wn, lock, labels, tb, tm, perms, det_thresh=thresh, lr=1e-4, apply=False)[0]
        print('step 2 losses', l2, [o2[k] for k in keys])
        assert np.allclose(l2, [o2[k] for k in keys], rtol=3e-2, atol=1e-4)
    finally:
        T.NP_DT = np.float32
        eng.close()


def test_loss_decreases_bf16():
    import disyolo_b200 as dy
    W, img, labels, tb, tm, pp, pg, thresh = _setup(seed=3)
    eng = dy.Engine(image_size=img.shape[1], max_batch=img.shape[0], precision='bf16')
    eng.load_weights(W)
    eng.train_init()
    hist = []
    for _ in range(6):
        hist.append(float(eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)[0]))
        eng.train_backward()
        eng.train_apply(1e-3)
    print('bf16 loss history', hist)
    assert np.all(np.isfinite(hist)) and hist[-1] < hist[0]
    # evaluation through the same net after training steps (validation loop, train_yolo3_mask.py:163-176)
    import torch
    out = eng.forward(torch.from_numpy(img).cuda(), torch.tensor([[0, 0, 1, 1.]] * img.shape[0]).cuda(), 0.1)
    torch.cuda.synchronize()
    assert out['det_count'].shape[0] == img.shape[0]
    eng.close()


def test_evaluate_after_training_is_ordered_after_the_step():
    """The reference's validation loop (train_yolo3_mask.py:146-176): sess.run(net.evaluation) right after
    sess.run([total_loss, optimizer]).  The host-buffer pipeline runs on its own stream: it must wait for
    the backward pass / Adam / repacking still queued on the caller's stream.  Evaluating immediately
    must equal evaluating after an explicit device synchronisation."""
    import torch
    import disyolo_b200 as dy
    W, img, labels, tb, tm, pp, pg, thresh = _setup(B=2, size=160, seed=3)
    win = np.tile(np.array([[0, 0, 1, 1]], np.float32), (2, 1))

    def run(sync):
        eng = dy.Engine(image_size=160, max_batch=2, precision='bf16')
        eng.load_weights(W)
        eng.train_init()
        outs = []
        for _ in range(3):
            eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
            eng.train_backward(82, 1)
            eng.train_apply(1e-3)
            if sync:
                torch.cuda.synchronize()
            raw, box, cnt, msk = eng.forward_host(img, win, 0.1)
            outs.append((raw.numpy().copy(), cnt.numpy().copy(),
                         [msk[b, :int(cnt[b])].numpy().copy() for b in range(2)]))
        eng.close()
        return outs
    a, b = run(False), run(True)
    for (ra, ca, ma), (rb, cb, mb) in zip(a, b):
        assert np.array_equal(ca, cb) and np.array_equal(ra, rb)
        for x, y in zip(ma, mb):
            assert np.array_equal(x, y)


def test_refinalize_after_training_keeps_and_restores_state():
    """dy_finalize_weights after dy_train_init: (1) with nothing newly loaded it must NOT revert the trained
    weights to the pre-training host copies; (2) variables loaded afterwards (Saver.restore) replace the
    training masters, restart Adam, and training continues from THEM."""
    import disyolo_b200 as dy
    from disyolo_b200 import _lib
    W, img, labels, tb, tm, pp, pg, thresh = _setup(B=2, size=128, seed=4)
    name = 'yolo/convolutional81/weights'
    for precision in ('bf16', 'fp32'):
        eng = dy.Engine(image_size=128, max_batch=2, precision=precision)
        eng.load_weights(W)
        eng.train_init()
        for _ in range(2):
            eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
            eng.train_backward(82, 1)
            eng.train_apply(1e-3)
        trained = eng.get_weights(name, W[name].shape)
        assert not np.array_equal(trained, W[name])
        win = np.tile(np.array([[0, 0, 1, 1]], np.float32), (2, 1))
        before = eng.forward_host(img, win, 0.1)
        before = [x.numpy().copy() for x in before[:3]]
        _lib.check(eng.lib.dy_finalize_weights(eng.h), 'dy_finalize_weights')          # (1)
        assert np.array_equal(eng.get_weights(name, W[name].shape), trained)
        after = [x.numpy().copy() for x in eng.forward_host(img, win, 0.1)[:3]]
        for x, y in zip(before, after):
            assert np.array_equal(x, y)
        # (2) restore the original checkpoint: the masters follow, the next step equals a fresh engine's first step
        eng.load_weights(W)
        assert np.array_equal(eng.get_weights(name, W[name].shape), W[name])
        l_restored = eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
        eng.train_backward(82, 1)
        eng.train_apply(1e-3)
        w_restored = eng.get_weights(name, W[name].shape)
        eng.close()
        fresh = dy.Engine(image_size=128, max_batch=2, precision=precision)
        fresh.load_weights(W)
        fresh.train_init()
        l_fresh = fresh.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
        fresh.train_backward(82, 1)
        fresh.train_apply(1e-3)
        w_fresh = fresh.get_weights(name, W[name].shape)
        fresh.close()
        assert np.allclose(l_restored, l_fresh, rtol=1e-5, atol=1e-6), (precision, l_restored, l_fresh)
        # Adam's first step is lr * g / (|g| + eps): entries whose gradient is ~0 amplify summation-order noise
        diff = np.abs(w_restored - w_fresh)
        assert diff.max() <= 2.5e-3 and np.mean(diff < 1e-6) > 0.99, (precision, diff.max(), np.mean(diff < 1e-6))


def test_partial_checkpoint_restore_and_loss_params():
    """Stage 1 of the reference: global_variables_initializer, then assign_from_checkpoint_fn(...,
    ignore_missing_vars=True) on a checkpoint WITHOUT convolutional76..82 (train_yolo3_mask.py:63,85-107):
    the mask subnet keeps its initial value.  And the loss scales are the cfg values, not constants."""
    import disyolo_b200.yolo.config as cfg
    from disyolo_b200.yolo.yolo3_net_pos import YOLONet, Session
    from disyolo_b200.weights import init_weights
    cfg.BATCH_SIZE, cfg.IMAGE_SIZE = 2, 128
    old_mask_scale = cfg.MASK_SCALE
    try:
        rng = np.random.default_rng(8)
        W = O.make_weights('lively', 2)
        coco = {k: v for k, v in W.items() if int(k.split('convolutional')[1].split('/')[0]) <= 75}
        net = YOLONet(True, precision='fp32')
        sess = Session(net, seed=5)
        sess.restore(coco)
        init = init_weights('reference', seed=5, lock=net.lock)
        assert len(sess.initialized_from_init) == 6 * 5 + 2          # 76..81: w + 4 BN, 82: w + bias
        got = net.engine.get_weights('yolo/convolutional80/weights', init['yolo/convolutional80/weights'].shape)
        assert np.array_equal(got, init['yolo/convolutional80/weights'])
        got = net.engine.get_weights('yolo/convolutional53/weights', W['yolo/convolutional53/weights'].shape)
        assert np.array_equal(got, W['yolo/convolutional53/weights'])
        img = rng.random((2, 128, 128, 3), dtype=np.float32)
        labels, tb, tm = T.make_labels(rng, 2, 128)
        feed = {net.images: img, net.yolo1: labels[2], net.yolo2: labels[1], net.yolo3: labels[0],
                net.true_boxes: tb, net.true_masks: tm, net.is_training: True,
                net.det_thresh: [0.05], net.clip_window: np.tile([[0., 0., 1., 1.]], (2, 1))}
        sess.rng = np.random.default_rng(77)   # same RoI permutations for both evaluations
        sess.run(net.total_loss, feed_dict=feed)
        base = dict(sess.last_losses)
        net.engine.set_loss_params(object_scale=4.0, noobject_scale=1.0, class_scale=1.0, coord_scale=1.0,
                                   mask_scale=10.0, ignore_thresh=0.5)
        sess.rng = np.random.default_rng(77)
        sess.run(net.total_loss, feed_dict=feed)
        scaled = dict(sess.last_losses)
        assert np.isclose(scaled['object'], 2.0 * base['object'], rtol=1e-5)
        assert np.isclose(scaled['mask'], 2.0 * base['mask'], rtol=1e-5)
        assert np.isclose(scaled['class'], base['class'], rtol=1e-6)
        cfg.MAX_BOX_PER_IMAGE = 10
        with pytest.raises(ValueError):
            YOLONet(True, precision='fp32')
    finally:
        cfg.BATCH_SIZE, cfg.IMAGE_SIZE, cfg.MAX_BOX_PER_IMAGE = 2, 576, 20
        cfg.MASK_SCALE = old_mask_scale


def test_fully_unlocked_net_trains_on_the_tensor_core_engine():
    """The reference's stage 2 ("lock=False for all layers", yolo3_net_pos.py:155-156; 61.66 M trainables,
    SURVEY a16) on the bf16 engine: convolutional1 (CUDA-core weight gradient on bf16 operands), the five stride-2
    convs (wgrad over the space-to-depth copies, dgrad as parity-block GEMMs) and every backbone layer take part
    in the backward pass.  Checked like the stage-1 test, layer by layer on IDENTICAL inputs: the float64 oracle
    recomputes dW and the input gradient of layers 1, 2, 5, 10, 27, 44 (and a residual block) from the engine's own
    x, dz; dz itself comes from the engine's dy through the float64 BN backward."""
    import torch
    import disyolo_b200 as dy
    W, img, labels, tb, tm, pp, pg, thresh = _setup(B=2, size=128, seed=6)
    B, size = 2, 128
    eng = dy.Engine(image_size=size, max_batch=B, precision='bf16', lock=[0] * 82)
    eng.load_weights(W)
    assert eng.train_init() == 61655665
    losses = eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
    assert np.all(np.isfinite(losses))
    g = eng.train_backward(82, 1).cpu().numpy()
    assert np.all(np.isfinite(g))
    tab = O.layer_table()

    def bn_bwd(z, dyv, gamma, beta):
        ze = torch.from_numpy(np.ascontiguousarray(z, np.float64)).requires_grad_(True)
        m = ze.mean(dim=(0, 1, 2), keepdim=True)
        v = ((ze - m) ** 2).mean(dim=(0, 1, 2), keepdim=True)
        y = (ze - m) * torch.rsqrt(v + O.BN_EPS) * torch.from_numpy(gamma.astype(np.float64)) + \
            torch.from_numpy(beta.astype(np.float64))
        y = torch.maximum(O.ALPHA * y, y)
        y.backward(torch.from_numpy(np.ascontiguousarray(dyv, np.float64)))
        return ze.grad.numpy()
    errs = {}
    for n, src in ((1, 0), (2, 1), (5, 4), (10, 9), (27, 26), (44, 43), (4, 3)):
        L = tab[n]
        x_in = bf16_round(img) if src == 0 else eng.activation(src, B).cpu().numpy()
        z = eng.train_tensor(n, 'z').cpu().numpy()
        dyv = eng.train_tensor(n, 'dy').cpu().numpy()
        dz = bf16_round(bn_bwd(z, dyv, W[O.vname(n, 'gamma')], W[O.vname(n, 'beta')]).astype(np.float32))
        w_bf = bf16_round(W[O.vname(n, 'w')])
        dx_ref, dw_ref = T.conv_backward(x_in, dz, w_bf, stride=L['s'])
        off, cnt = eng.layer_span(n)
        nw = L['k'] ** 2 * L['cin'] * L['cout']
        errs['%ddw' % n] = rel_err(g[off:off + nw].reshape(dw_ref.shape), dw_ref)
        if src in (1, 3):
            # these producers have exactly one consumer: their dy IS this layer's input gradient
            errs['%ddx' % n] = rel_err(eng.train_tensor(src, 'dy').cpu().numpy(), dx_ref)
    print('stage-2 local checks:', ' '.join('%s:%.2g' % kv for kv in errs.items()))
    for k, e in errs.items():
        assert e < (2e-2 if k.endswith('dw') else 1.5e-2), (k, e)
    # and the step runs end to end: Adam, moving averages, re-packing (incl. the stride-2 dgrad operands)
    hist = [float(losses[0])]
    eng.train_apply(1e-4)
    for _ in range(3):
        hist.append(float(eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)[0]))
        eng.train_backward()
        eng.train_apply(1e-4)
    print('stage-2 loss history', hist)
    assert np.all(np.isfinite(hist)) and hist[-1] < hist[0]
    eng.close()


def test_losses_with_degenerate_ground_truth_box():
    """overlaps_graph without epsilon (yolo3_net_pos.py:954-975): image 1's ground truth is ONE zero-area box, so every
    RoI / ground-truth IoU is 0 or 0/0 and the image contributes no positive RoI to loss_mask; the eight scalars
    still match the float64 oracle (the label tensors are whatever they are: both sides read the same arrays)."""
    import disyolo_b200 as dy
    W, img, labels, tb, tm, pp, pg, thresh = _setup()
    B, size = img.shape[0], img.shape[1]
    tb = tb.copy()
    tb[1] = 0.0
    tb[1, 0, 0, 0, 0] = [0.4, 0.6, 0.0, 0.0, 1.0]          # xc, yc, w = h = 0: kept (non-zero row), zero area
    eng = dy.Engine(image_size=size, max_batch=B, precision='fp32')
    eng.load_weights(W)
    eng.train_init()
    perms = [(pp[b].tolist(), pg[b].tolist()) for b in range(B)]
    T.NP_DT = np.float64
    try:
        losses = eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
        ol, og, Wn, adam, aux = T.train_step(img, W, O.default_lock_flags(), labels, tb, tm, perms, det_thresh=thresh,
                                             lr=1e-4, adam=None, step=1)
    finally:
        T.NP_DT = np.float32
    want = np.array([ol[k] for k in ('total', 'obj', 'noobj', 'cls', 'xy', 'wh', 'mask', 'l2')])
    print('degenerate-gt losses', losses, 'oracle', want)
    assert np.all(np.isfinite(losses))
    assert np.allclose(losses, want, rtol=1e-4, atol=1e-5)
    eng.close()
