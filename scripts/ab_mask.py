"""A/B of the mask-assembly kernel's store policy and per-kernel post-processing times (batch 64 @576)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import disyolo_b200 as dy
from disyolo_b200.engine import set_option
B = 64
eng = dy.Engine(image_size=576, max_batch=B, precision='bf16')
eng.load_weights(dy.init_weights('lively', 0))
img = torch.from_numpy(np.random.default_rng(1000).random((B, 576, 576, 3), dtype=np.float32)).cuda()
win = torch.tensor([[0, 0, 1, 1]], dtype=torch.float32).repeat(B, 1).cuda()
out = eng.forward(img, win, 0.25)
torch.cuda.synchronize()
dets = int(out['det_count'].sum().item())
yol = [eng.yolo(s, B) for s in range(3)]
mp = eng.mask_pos(B)
pl = mp.permute(0, 3, 1, 2).contiguous()
fill = torch.empty_like(out['masks'][:, :max(1, dets // B)])
for _ in range(3):
    fill.fill_(0.5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    fill.fill_(0.5)
e1.record(); torch.cuda.synchronize()
print('fill_ of %d MB: %.1f us' % (fill.numel() * 4 // 2**20, e0.elapsed_time(e1) / 20 * 1e3))
for stream, wl in ((1, 1), (1, 0), (0, 1), (0, 0), (1, 1), (1, 0)):
    set_option('mask_streaming_stores', stream)
    set_option('mask_work_list', wl)
    r = eng.postproc_profile(yol, pl, win, 0.25, out['masks'], layout='planar', reps=20)
    print('streaming=%d work_list=%d dets=%d  %s  mask %.0f GB/s' % (stream, wl, dets, {k: round(v * 1e3, 1) for k, v in r.items()},
                                                                    dets * 288 * 288 * 4 / (r['masks'] * 1e-3) / 1e9))
