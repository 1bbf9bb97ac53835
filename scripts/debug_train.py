import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import disyolo_b200 as dy
from oracle import dis_oracle as O, dis_oracle_train as T
from tests.test_gpu_train import _setup
from tests.util import rel_err
W, img, labels, tb, tm, pp, pg, thresh = _setup()
B, size = img.shape[0], img.shape[1]
lock = O.default_lock_flags()
eng = dy.Engine(image_size=size, max_batch=B, precision='fp32'); eng.load_weights(W); eng.train_init()
perms = [(pp[b].tolist(), pg[b].tolist()) for b in range(B)]
losses = eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
ol, og, Wn, adam, aux = T.train_step(img, W, lock, labels, tb, tm, perms, det_thresh=thresh, apply=False)
eng.train_backward(82, 1)
for n in range(82, 52, -1):
    if n in aux['act_grads']:
        got = eng.train_tensor(n, 'dy').cpu().numpy()
        print(n, 'dy rel err %.2e' % rel_err(got, aux['act_grads'][n]), 'norm %.3e' % np.linalg.norm(aux['act_grads'][n]))
for n in (80, 57, 73):
    got = eng.train_tensor(n, 'dy').cpu().numpy(); ref = aux['act_grads'][n]
    d = np.abs(got - ref)
    H = ref.shape[1]
    inner = d[:, 1:-1, 1:-1].max(); border = max(d[:, 0].max(), d[:, -1].max(), d[:, :, 0].max(), d[:, :, -1].max())
    print(n, 'max abs diff interior %.3e border %.3e; ref absmax %.3e' % (inner, border, np.abs(ref).max()))
    e_int = rel_err(got[:, 1:-1, 1:-1], ref[:, 1:-1, 1:-1]) if H > 2 else -1
    print('   rel err interior only %.3e; per-image rel err' % e_int, [float('%.2e' % rel_err(got[b], ref[b])) for b in range(B)])
    print('   mean ratio got/ref over large entries', float(np.mean((got / np.where(np.abs(ref) > 1e-4, ref, np.nan))[np.abs(ref) > 1e-4])))
print('--- float64 oracle ---')
T.NP_DT = np.float64
ol64, og64, _, _, aux64 = T.train_step(img, W, lock, labels, tb, tm, perms, det_thresh=thresh, apply=False)
for n in (81, 80, 74, 73, 58, 57, 53):
    got = eng.train_tensor(n, 'dy').cpu().numpy()
    print(n, 'mine vs f64 %.2e | f32 oracle vs f64 %.2e' % (rel_err(got, aux64['act_grads'][n]), rel_err(aux['act_grads'][n], aux64['act_grads'][n])))
