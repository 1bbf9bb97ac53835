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
res, names, used, fwd = {}, [], set(), {}
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
    win = torch.tensor([[0, 0, 1, 1]], dtype=torch.float32).repeat(B, 1).cuda()
    out = eng.forward(img, win, 0.25)
    for _ in range(3):
        eng.forward(img, win, 0.25, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.forward(img, win, 0.25, out=out)
    e1.record(); torch.cuda.synchronize()
    fwd.setdefault(name, []).append(e0.elapsed_time(e1) / 20)
    del out
    eng.close(); del eng
    torch.cuda.empty_cache()
    for k, v in kv:
        set_option(k, -1 if k not in ('tc_skip_epilogue',) else 0)
print('layer ' + ' '.join('%12s' % n for n in names))
for n in range(1, 83):
    print('%5d ' % n + ' '.join('%12.4f' % res[k][n] for k in names))
print('total ' + ' '.join('%12.3f' % res[k][1:].sum() for k in names))
print('step  ' + ' '.join('%12.3f' % min(fwd[k]) for k in names) + '   (ms per whole dy_forward, 20 back to back, no per-layer events)')
