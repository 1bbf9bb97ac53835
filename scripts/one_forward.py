"""One warm-up + N network passes at batch 64 (for ncu captures; never a bench number)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import disyolo_b200 as dy
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = dy.Engine(image_size=576, max_batch=B, precision='bf16')
eng.load_weights(dy.init_weights('lively', 0))
img = torch.from_numpy(np.random.default_rng(0).random((B, 576, 576, 3), dtype=np.float32)).cuda()
win = torch.tensor([[0, 0, 1, 1]], dtype=torch.float32).repeat(B, 1).cuda()
for _ in range(1 + n):
    eng.forward(img, win, 0.25)
torch.cuda.synchronize()
print('done')
