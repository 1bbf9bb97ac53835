"""Per-layer device time (ms, batch 64 @576) under each planning mode of the conv engine."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import disyolo_b200 as dy
from disyolo_b200.engine import set_option
B = 64
W = dy.init_weights('lively', 0)
img = torch.from_numpy(np.random.default_rng(0).random((B, 576, 576, 3), dtype=np.float32)).cuda()
modes = {'auto': (-1, -1, -1, -1), 'auto-coop': (-1, -1, -1, 0), 'base': (0, 0, 0, 0), 'tma': (0, 0, 1, 1),
         'res+tma': (1, 0, 1, 1), 'res+halo+tma': (1, 1, 1, 1), 'halo+tma': (0, 1, 1, 1), 'res+halo': (1, 1, 0, 0),
         'halo': (0, 1, 0, 0), 'res': (1, 0, 0, 0)}
res = {}
for name, (r, h, s, t) in modes.items():
    set_option('tc_resident', r); set_option('tc_halo', h); set_option('tc_staged', s); set_option('tc_tma_epi', t)
    eng = dy.Engine(image_size=576, max_batch=B, precision='bf16')
    eng.load_weights(W)
    ms = np.zeros(83)
    eng.profile_layers(img)
    for _ in range(3):
        ms += eng.profile_layers(img)
    res[name] = ms / 3
    eng.close()
    del eng
    torch.cuda.empty_cache()
names = list(modes)
print('layer ' + ' '.join('%11s' % n for n in names) + '   best')
for n in range(1, 83):
    row = [res[k][n] for k in names]
    b = int(np.argmin(row[2:])) + 2
    print('%5d ' % n + ' '.join('%11.4f' % v for v in row) + '   ' + names[b])
print('total ' + ' '.join('%11.3f' % res[k][1:].sum() for k in names))
print('best-of total %.3f' % sum(min(res[k][n] for k in names[2:]) for n in range(1, 83)))
