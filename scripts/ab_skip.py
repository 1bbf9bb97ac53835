"""Per-layer time with the conv epilogue body skipped (load + MMA side only) vs normal."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import disyolo_b200 as dy
from disyolo_b200.engine import set_option
B = 64
W = dy.init_weights('lively', 0)
img = torch.from_numpy(np.random.default_rng(0).random((B, 576, 576, 3), dtype=np.float32)).cuda()
res = {}
for name, v in (('normal', 0), ('skip_epi', 1), ('ld_only', 2), ('ld_math_sts', 3), ('no_store', 4)):
    set_option('tc_skip_epilogue', v)
    eng = dy.Engine(image_size=576, max_batch=B, precision='bf16'); eng.load_weights(W)
    eng.profile_layers(img)
    ms = np.zeros(83)
    for _ in range(3): ms += eng.profile_layers(img)
    res[name] = ms / 3
    eng.close(); del eng; torch.cuda.empty_cache()
for n in range(1, 83):
    print('%3d normal %.4f  skip_epi %.4f  ld_only %.4f  ld_math_sts %.4f  no_store %.4f' % (n, res['normal'][n], res['skip_epi'][n], res['ld_only'][n], res['ld_math_sts'][n], res['no_store'][n]))
print('total', res['normal'][1:].sum(), res['skip_epi'][1:].sum())
