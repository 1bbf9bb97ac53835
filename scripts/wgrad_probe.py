"""Bring-up aid for the MN-major UMMA descriptors of wgrad_tc_kernel: runs one small wgrad under a
few (LBO, SBO) conventions and prints the error of each against torch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import disyolo_b200.engine as E
from oracle import dis_oracle_train as T
from tests.util import bf16_round, rel_err

for (cin, cout, k) in [(128, 256, 1), (32, 32, 1), (64, 64, 3)]:
    rng = np.random.default_rng(0)
    B, H = 2, 12
    x = bf16_round(rng.standard_normal((B, H, H, cin)).astype(np.float32))
    dz = bf16_round(rng.standard_normal((B, H, H, cout)).astype(np.float32))
    w = np.zeros((k, k, cin, cout), np.float32)
    _, ref = T.conv_backward(x, dz, w)
    awa = 64 if cin % 64 == 0 else 32
    awb = 64 if cout % 64 == 0 else 32
    box_a, box_b = 64 * awa * 2, 64 * awb * 2
    grp_a, grp_b = 8 * awa * 2, 8 * awb * 2
    # (the alternatives tried during bring-up -- LBO/SBO swapped, LBO=1 -- read outside the stage and fault;
    # pass other values through dy_set_option("wgrad_lbo_a", ...) to experiment)
    for name, (la, sa, lb, sb) in dict(default=(0, 0, 0, 0)).items():
        for o, v in zip(('wgrad_lbo_a', 'wgrad_sbo_a', 'wgrad_lbo_b', 'wgrad_sbo_b'), (la, sa, lb, sb)):
            E.set_option(o, v)
        _, dw = E.conv_backward(torch.from_numpy(x).cuda(), torch.from_numpy(dz).cuda(), w, want_dx=False)
        torch.cuda.synchronize()
        print('cin %d cout %d k %d  %-8s rel err %.3g' % (cin, cout, k, name, rel_err(dw.cpu().numpy(), ref)))
for o in ('wgrad_lbo_a', 'wgrad_sbo_a', 'wgrad_lbo_b', 'wgrad_sbo_b'):
    E.set_option(o, 0)
