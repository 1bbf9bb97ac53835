"""Does the very first forward of a fresh engine differ from later ones?  Per-layer hashes."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import disyolo_b200 as dy
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
img = torch.from_numpy(np.random.default_rng(1000).random((B, 576, 576, 3), dtype=np.float32)).cuda()
win = torch.tensor([[0, 0, 1, 1]], dtype=torch.float32).repeat(B, 1).cuda()
W = dy.init_weights('lively', 0)
PROBE = [4, 9, 10, 26, 27, 43, 44, 58, 82]
def layers(e):
    return [hashlib.md5(e.activation(n, B).cpu().numpy().tobytes()).hexdigest()[:6] for n in PROBE]
for trial in range(4):
    eng = dy.Engine(image_size=576, max_batch=B, precision='bf16')
    eng.load_weights(W)
    o = eng.forward(img, win, 0.25); torch.cuda.synchronize()
    c1, h1 = int(o['det_count'].sum().item()), layers(eng)
    r1 = o['det_raw'].cpu().numpy().copy()
    o = eng.forward(img, win, 0.25); torch.cuda.synchronize()
    c2, h2 = int(o['det_count'].sum().item()), layers(eng)
    r2 = o['det_raw'].cpu().numpy().copy()
    diff = [PROBE[n] for n in range(len(PROBE)) if h1[n] != h2[n]]
    print('trial', trial, 'dets', c1, c2, 'layers differing between forward 1 and 2:', diff[:20])
    print('   det_raw equal:', np.array_equal(r1, r2), 'rows differing:', int((np.abs(r1 - r2).max(axis=2) > 0).sum()))
    eng.close()
