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
from tests.util import rel_err

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
