"""Weights of convolutional1..82 keyed by the reference's TensorFlow variable names
(train_yolo3_mask.py:87-103): ``yolo/convolutional{N}/weights`` (HWIO fp32),
``.../BatchNorm/{gamma,beta,moving_mean,moving_variance}``, ``.../biases``.

``init_weights('reference')`` reproduces the initialisers of the reference's builders
(yolo3_net_pos.py:77-80,112-123,135-140); ``init_weights('lively')`` is a synthetic init whose
activations stay O(1) through all 82 layers so that decode / NMS / mask assembly see real work
(used by bench.py and smoke(); the reference init is numerically degenerate in inference mode).
"""
import numpy as np

from .engine import layer_table


def variable_names(n, bn):
    base = 'yolo/convolutional%d/' % n
    if bn:
        return [base + 'weights', base + 'BatchNorm/gamma', base + 'BatchNorm/beta',
                base + 'BatchNorm/moving_mean', base + 'BatchNorm/moving_variance']
    return [base + 'weights', base + 'biases']


def init_weights(flavour='lively', seed=0, lock=None):
    rng = np.random.default_rng(seed)
    out = {}
    for L in layer_table():
        n, k, cin, cout = L['id'], L['k'], L['cin'], L['cout']
        base = 'yolo/convolutional%d/' % n
        shape = (k, k, cin, cout)
        locked = (n <= 52) if lock is None else bool(lock[n - 1])
        if flavour == 'reference':
            if locked:     # tf.truncated_normal(stddev=0.001): resample beyond 2 sigma
                w = rng.standard_normal(shape) * 0.001
                bad = np.abs(w) > 0.002
                while bad.any():
                    w[bad] = rng.standard_normal(int(bad.sum())) * 0.001
                    bad = np.abs(w) > 0.002
            else:          # xavier_initializer(): uniform(+-sqrt(6/(fan_in+fan_out)))
                lim = np.sqrt(6.0 / (k * k * cin + k * k * cout))
                w = rng.uniform(-lim, lim, shape)
            out[base + 'weights'] = w.astype(np.float32)
            if L['bn']:
                out[base + 'BatchNorm/gamma'] = np.ones(cout, np.float32)
                out[base + 'BatchNorm/beta'] = np.zeros(cout, np.float32)
                out[base + 'BatchNorm/moving_mean'] = np.zeros(cout, np.float32)
                out[base + 'BatchNorm/moving_variance'] = np.ones(cout, np.float32)
            else:
                out[base + 'biases'] = np.zeros(cout, np.float32)
            continue
        # 'lively': the oracle's generator statement by statement (oracle/dis_oracle.py make_weights): the same
        # draws in the same order, the same float32 roundings -- bit-identical weights
        # (tests/test_oracle_golden.py::test_product_and_oracle_lively_weights_are_bit_identical)
        f32 = np.float32
        gain = 0.3 if L['res'] else 1.0
        std = gain * np.sqrt(2.0 / ((1.0 + 0.1 * 0.1) * k * k * cin))
        w = (rng.standard_normal(shape) * std).astype(f32)
        if L['bn']:
            out[base + 'weights'] = w
            out[base + 'BatchNorm/gamma'] = rng.uniform(0.7, 1.3, cout).astype(f32)
            out[base + 'BatchNorm/beta'] = (rng.standard_normal(cout) * 0.2).astype(f32)
            out[base + 'BatchNorm/moving_mean'] = (rng.standard_normal(cout) * 0.2).astype(f32)
            out[base + 'BatchNorm/moving_variance'] = rng.uniform(0.6, 1.6, cout).astype(f32)
        elif cout == 24:
            w = w.reshape(k, k, cin, 3, 8) * f32(0.1)     # head logits O(1) on activations that reach |x| ~ 10
            w[..., 2:4] *= 0.25                            # keep exp(t_wh) tame
            b = np.zeros((3, 8), f32)
            b[:, 4] = -2.6                                 # sparse objectness
            b[:, 0:2] = rng.standard_normal((3, 2)) * 0.3
            b[:, 2:4] = rng.standard_normal((3, 2)) * 0.2 - 0.3
            b[:, 5:] = rng.standard_normal((3, 3)) * 0.5
            out[base + 'weights'] = w.reshape(shape)
            out[base + 'biases'] = b.reshape(24).astype(f32)
        else:
            w *= f32(0.35)
            out[base + 'weights'] = w
            out[base + 'biases'] = (rng.standard_normal(cout) * 0.5).astype(f32)
    return out


def save_npz(path, weights):
    np.savez(path, **{k.replace('/', '|'): v for k, v in weights.items()})


def load_npz(path):
    with np.load(path) as z:
        return {k.replace('|', '/'): z[k] for k in z.files}
