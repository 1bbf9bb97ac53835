"""Per-layer device time (ms, batch 64 @576) under named option sets (dy_set_option): A/B of planner switches.
usage: ab_opts.py "name:opt=v,opt=v" "name2:..."   (name 'auto' with no options = defaults)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import disyolo_b200 as dy
from disyolo_b200.engine import set_option
B = 64
W = dy.init_weights('lively', 0)
img = torch.from_numpy(np.random.default_rng(0).random((B, 576, 576, 3), dtype=np.float32)).cuda()
specs = sys.argv[1:] or ['auto:']
res, names, used = {}, [], set()
for spec in specs:
    name, _, opts = spec.partition(':')
    kv = [o.split('=') for o in opts.split(',') if o]
    for k, v in kv:
        set_option(k, int(v)); used.add(k)
    eng = dy.Engine(image_size=576, max_batch=B, precision='bf16')
    eng.load_weights(W)
    eng.profile_layers(img)
    ms = np.median(np.stack([eng.profile_layers(img) for _ in range(5)]), axis=0)
    res[name] = ms; names.append(name)
    eng.close(); del eng
    torch.cuda.empty_cache()
    for k, v in kv:
        set_option(k, -1 if k not in ('tc_skip_epilogue',) else 0)
print('layer ' + ' '.join('%12s' % n for n in names))
for n in range(1, 83):
    print('%5d ' % n + ' '.join('%12.4f' % res[k][n] for k in names))
print('total ' + ' '.join('%12.3f' % res[k][1:].sum() for k in names))
