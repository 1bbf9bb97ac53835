#!/usr/bin/env python
"""bench.py -- DIS-YOLO hot path on B200: images/s @576^2 forward + NMS + masks.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm (rank 0 only)

A "step" = one pass of the hot path (82 convs -> decode -> NMS -> top-k -> position-sensitive mask
assembly) over one batch of 64 synthetic 576x576 images per GPU (BASELINE.json configs[1]); with
N > 1 every rank runs its own shard of the batch, no collective on the data path (configs[2],
weak scaling).  One JSON line is printed by rank 0.

  value    images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e      the same metric through the host-buffer C-ABI call (dy_forward_host_begin_u8 / _end_cropped):
           pinned host uint8 images H2D (divided by 255 on the device exactly like image_read), forward,
           D2H of counts / boxes / the box-cropped mask maps the reference's consumer reads -- inside the
           timing; e2e_reference_layout = the same through fp32 images and full [n,S,S] maps
  roofline dominant kernel = the tcgen05 conv kernel (81 launches per step): algorithmic FLOPs of
           layers 2..82 / summed per-layer device time measured live with CUDA events
  cpu_baseline  the oracle (NumPy/torch-CPU restatement of the reference graph; TensorFlow 1.x
           cannot be installed) on the box's host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAGE = 576
PER_GPU_BATCH = 64
THRESH = 0.25
METRIC = 'images/s @576^2 fwd+NMS+masks'


def layer_flops(size):
    """Algorithmic 2*M*K*N per layer (no credit for padded K/N); index = layer number."""
    from disyolo_b200 import layer_table
    fl = {}
    for L in layer_table():
        h = size // L['size_div']
        fl[L['id']] = 2.0 * h * h * L['k'] * L['k'] * L['cin'] * L['cout']
    return fl


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'],
                    source='MEASURED_PEAKS.json')
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    In-process NVML polling every ~2 ms (a 64-image step is ~11 ms, too short for `nvidia-smi -lms`);
    falls back to one `nvidia-smi --query-gpu` line per 100 ms if NVML cannot be loaded."""

    REASONS = (('hw_slowdown', 'nvmlClocksEventReasonHwSlowdown', 0x8),
               ('hw_thermal_slowdown', 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
               ('sw_thermal_slowdown', 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
               ('sw_power_cap', 'nvmlClocksEventReasonSwPowerCap', 0x4))

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nv, self.h = index, [], None, None, None
        self.stop_flag = False
        self.thread = None
        self.mx = None

    def _physical_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            try:
                return int(vis.split(',')[self.index])
            except Exception:
                pass
        return self.index

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.nv = nv
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self._physical_index()), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read_smi, daemon=True)
        self.thread.start()

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), mhz, bits))
            except Exception:
                pass
            time.sleep(0.002)

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(',')]
            try:
                mhz, self.mx = float(f[0]), float(f[1])
            except Exception:
                continue
            bits = 0
            for (name, _, bit), v in zip(self.REASONS, f[3:7]):
                if v.lower().startswith('active'):
                    bits |= bit
            self.rows.append((time.time(), mhz, bits))

    def stop(self, t0, t1):
        if self.thread is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no NVML and no nvidia-smi'], samples=0)
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        sm, reasons = [], set()
        for t, mhz, bits in self.rows:
            if t < t0 or t > t1:
                continue
            sm.append(mhz)
            for name, attr, bit in self.REASONS:
                if bits & (getattr(self.nv, attr, bit) if self.nv else bit):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=self.mx, reasons=sorted(reasons),
                    samples=len(sm), source='nvml' if self.nv else 'nvidia-smi')


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


# --------------------------------------------------------------------------------------------
# reference arm: the oracle on the host cores
# --------------------------------------------------------------------------------------------
CPU_SAMPLE_BATCH = 4     # images per CPU step: a bounded sample of the 64-image step


def cpu_reference_images_per_s(n_runs, warm, weights, seed=0, batch=CPU_SAMPLE_BATCH):
    """Oracle (forward + decode + NMS + mask assembly) on the host cores, `batch` images per step, all
    torch / BLAS threads; returns (images/s from the median step, total timed seconds, threads)."""
    import numpy as np
    import torch
    from oracle import dis_oracle as O
    rng = np.random.default_rng(seed)
    img = rng.random((batch, IMAGE, IMAGE, 3), dtype=np.float32)
    win = np.tile(np.array([[0, 0, 1, 1]], np.float32), (batch, 1))
    for _ in range(warm):
        O.evaluate(img, win, THRESH, weights)
    ts = []
    for _ in range(n_runs):
        t = time.perf_counter()
        O.evaluate(img, win, THRESH, weights)
        ts.append(time.perf_counter() - t)
    ts.sort()
    return batch / ts[len(ts) // 2], sum(ts), torch.get_num_threads()


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    import torch
    import disyolo_b200 as dy
    torch.set_num_threads(os.cpu_count() or 1)     # torchrun exports OMP_NUM_THREADS=1: use every host core
    W = dy.init_weights('lively', 0)
    t0 = time.perf_counter()
    # images per step sized so that the K timed steps stay around a minute of host work (~0.35 s per image
    # on 16 cores): 4 images per step up to K = 40, 1 image per step from K = 160
    batch = max(1, min(CPU_SAMPLE_BATCH, 160 // max(1, args.steps)))
    ips, total, cores = cpu_reference_images_per_s(args.steps, max(1, min(args.warmup, 1)), W, batch=batch)
    ms = 1000.0 * batch / ips
    line = dict(metric=METRIC, value=ips, unit='images/s', impl='reference', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic',
                config=dict(workload='DIS-YOLO inference 576x576 (BASELINE configs[1]), lively random-init weights',
                            per_step='%d images (bounded sample of the batch-64 step)' % batch,
                            image=IMAGE, thresh=THRESH),
                cpu_baseline=dict(value=ips, unit='images/s', cores=cores, kind='port',
                                  sample='%d steps x %d images 576x576 through oracle.evaluate (median step)'
                                         % (args.steps, batch)),
                e2e=dict(value=ips, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, wall_s=time.perf_counter() - t0)
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import disyolo_b200 as dy

    rank, local, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        dist = None
        torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    B = args.batch
    peaks = measured_peaks()
    # pinned staging memory on the GPU's own NUMA node (first touch follows the CPU affinity)
    from disyolo_b200.hostmem import bind_to_gpu_numa
    affinity0 = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
    numa = bind_to_gpu_numa(local) if not args.no_numa else dict(bound=False, node=None, cpus=0)

    W = dy.init_weights('lively', 0)
    eng = dy.Engine(image_size=IMAGE, max_batch=B, precision='bf16', device=local)
    eng.load_weights(W)
    rng = np.random.default_rng(1000 + rank)
    # synthetic letterboxed images as image_read holds them before its `/ 255.`: uint8 RGB; the fp32 feed of
    # the reference protocol is their float64 quotient rounded to float32 (calculate_test_map.py:175)
    u8_host = torch.from_numpy(rng.integers(0, 256, (B, IMAGE, IMAGE, 3), dtype=np.uint8)).pin_memory()
    img_host = (u8_host.to(torch.float64) / 255.0).to(torch.float32).pin_memory()
    win_host = torch.tensor([[0, 0, 1, 1]], dtype=torch.float32).repeat(B, 1).pin_memory()
    img = img_host.to(dev)
    win = win_host.to(dev)
    md, sm = eng.max_detection, eng.mask_size
    out = dict(det_raw=torch.empty((B, md, 6), device=dev), det_box=torch.empty((B, md, 6), device=dev),
               det_count=torch.empty((B,), dtype=torch.int32, device=dev),
               masks=torch.empty((B, md, sm, sm), device=dev))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        eng.forward(img, win, THRESH, out=out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # K = 20 steps are ~0.2 s: too short for the power limiter to settle.  Extra untimed steps bring the GPU to its
    # sustained (power-capped) state FIRST, so that the K timed steps are a sample of steady-state operation and
    # not of a cold burst (every rank runs the same number: ~1.5 s worth, measured on the warm-up steps above).
    tw0 = time.perf_counter()
    eng.forward(img, win, THRESH, out=out)
    torch.cuda.synchronize()
    est = max(time.perf_counter() - tw0, 1e-3)
    steady_steps = 0 if args.no_steady else int(min(400, max(0, 1.5 / est)))
    if dist is not None:
        t = torch.tensor([steady_steps], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steady_steps = int(t.item())
    for _ in range(steady_steps):
        eng.forward(img, win, THRESH, out=out)
    eng.lib.dy_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start = time.time()
    e0.record()
    for _ in range(args.steps):
        eng.forward(img, win, THRESH, out=out)
    e1.record()
    barrier()
    t_end = time.time()
    launches = int(eng.lib.dy_launch_count(0))
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    dets = int(out['det_count'].sum().item())

    # ---- roofline of the dominant kernel (tcgen05 conv), measured live per layer ----
    fl = layer_flops(IMAGE)
    # (right after the timed loop, same thermal / power state; per-layer median of 5 passes)
    ms_layers = np.median(np.stack([eng.profile_layers(img) for _ in range(5)]), axis=0)
    tc_ms = float(ms_layers[2:].sum())
    tc_flops = sum(fl[n] for n in range(2, 83)) * B
    achieved_tf = tc_flops / (tc_ms / 1e3) / 1e12
    net_ms = float(ms_layers[1:].sum())
    per_layer = {str(n): dict(ms=round(float(ms_layers[n]), 4),
                              tflops=round(fl[n] * B / (float(ms_layers[n]) / 1e3) / 1e12, 1))
                 for n in range(1, 83)}

    # ---- e2e: host buffers through the C-ABI call ----
    # steady-state serving loop, 3 batches in flight (dy_forward_host_begin* / _end*): every step includes its
    # own pinned-host -> device image copy and the device -> host copy of its results
    from collections import deque

    def e2e_loop(images, mode, n_steps):
        for _ in range(2):      # warm-up through the same pipelined calls (allocates all pinned result sets)
            tks = [eng.forward_host_begin(images, win_host, THRESH, masks=mode) for _ in range(3)]
            for tkk in tks:
                eng.forward_host_end(tkk)
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        flight = deque(eng.forward_host_begin(images, win_host, THRESH, masks=mode) for _ in range(min(2, n_steps)))
        begun = len(flight)
        for i in range(n_steps):
            if begun < n_steps:
                flight.append(eng.forward_host_begin(images, win_host, THRESH, masks=mode))
                begun += 1
            raw, box, cnt, msk = eng.forward_host_end(flight.popleft())
            d2h += B * 4 + 2 * B * md * 24
            if mode == 'cropped':
                d2h += (B * md + 1) * 8 + int(msk[1].numel()) * 4
            else:
                d2h += int(cnt.sum().item()) * sm * sm * 4
        torch.cuda.synchronize()
        secs = max_over_ranks(time.perf_counter() - t0)
        return world * B * n_steps / secs, d2h // n_steps

    n_e2e = max(6, args.steps)
    e2e_value, d2h = e2e_loop(u8_host, 'cropped', n_e2e)
    h2d = B * IMAGE * IMAGE * 3 + B * 16
    ref_value, ref_d2h = e2e_loop(img_host, 'full', max(6, min(args.steps, 20)))
    e2e_ref = dict(value=ref_value, unit='images/s', h2d_bytes_per_step=B * IMAGE * IMAGE * 3 * 4 + B * 16,
                   d2h_bytes_per_step=ref_d2h,
                   mode='fp32 [B,S,S,3] images in, det_count[b] full fp32 [S/2,S/2] maps per image out '
                        '(the reference feed / fetch layouts verbatim)')

    # ---- e2e of the whole per-image sequence of the reference's test driver (calculate_test_map.py:208-269):
    # uint8 images up, letterbox + network + post-processing to original-image masks on the GPU, boxes and
    # merged masks down (ImagePipeline, two batches in flight) ----
    pipe_line = None
    if not args.no_pipeline:
        ph, pw = 720, 1280
        pipe = dy.ImagePipeline(eng, ph, pw, depth=2)
        prng = np.random.default_rng(2000 + rank)
        frames = [torch.from_numpy(prng.integers(0, 256, (ph, pw, 3), dtype=np.uint8)).pin_memory() for _ in range(8)]
        frames = [frames[i % 8] for i in range(B)]
        for _ in range(2):
            pipe.result(pipe.submit(frames, THRESH))
        barrier()
        n_p = max(6, args.steps)
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        tp0 = time.perf_counter()
        tk = pipe.submit(frames, THRESH)
        for i in range(n_p):
            nxt = pipe.submit(frames, THRESH) if i + 1 < n_p else None
            res = pipe.result(tk)
            tk = nxt
        torch.cuda.synchronize()
        pipe_s = max_over_ranks(time.perf_counter() - tp0)
        pipe_line = dict(value=world * B * n_p / pipe_s, unit='images/s', h2d_bytes_per_step=pipe.h2d_bytes // n_p,
                         d2h_bytes_per_step=pipe.d2h_bytes // n_p, steps=n_p,
                         mode='uint8 %dx%d frames in, letterbox + forward + un-letterboxed boxes and merged masks '
                              'out (ImagePipeline, 2 batches in flight)' % (ph, pw))
        del pipe

    # post-processing kernels, each timed alone (20 back-to-back launches between two CUDA events inside
    # the library).  decode and mask assembly are the HBM-bound ones; NMS / top-k work on a few KB and are
    # latency-bound (reported as times only).
    yol = [eng.yolo(s, B) for s in range(3)]
    mp = eng.mask_pos(B)
    mp = mp.permute(0, 3, 1, 2).contiguous()       # the engine's own score-map layout: planar [B,9,S,S]
    pp_ms = eng.postproc_profile(yol, mp, win, THRESH, out['masks'], layout='planar', reps=20)
    n0 = eng.num_candidates
    decode_bytes = B * n0 * 8 * 4
    mask_bytes = dets * sm * sm * 4
    extra = [dict(kernel='decode_kernel (sigmoid/exp/softmax + threshold + compaction)', bound='hbm',
                  ms=pp_ms['decode'], algorithmic_bytes=decode_bytes,
                  achieved=decode_bytes / (pp_ms['decode'] / 1e3) / 1e9, peak=peaks['hbm_gbs'], unit='GB/s',
                  frac=decode_bytes / (pp_ms['decode'] / 1e3) / 1e9 / peaks['hbm_gbs']),
             dict(kernel='mask_kernel (position-sensitive mask assembly)', bound='hbm', ms=pp_ms['masks'],
                  algorithmic_bytes=mask_bytes,
                  achieved=mask_bytes / (pp_ms['masks'] / 1e3) / 1e9 if pp_ms['masks'] > 0 else None,
                  peak=peaks['hbm_gbs'], unit='GB/s',
                  frac=mask_bytes / (pp_ms['masks'] / 1e3) / 1e9 / peaks['hbm_gbs'] if pp_ms['masks'] > 0 else None),
             dict(kernel='nms_kernel + finalize_kernel', bound='latency', ms=pp_ms['nms'] + pp_ms['finalize'],
                  nms_ms=pp_ms['nms'], finalize_ms=pp_ms['finalize'])]

    # ---- batch-1 latency (p50) ----
    lat = None
    if args.latency and rank == 0:
        e1b = dy.Engine(image_size=IMAGE, max_batch=1, precision='bf16', device=local)
        e1b.load_weights(W)
        i1, w1 = img[:1].contiguous(), win[:1].contiguous()
        o1 = dict(det_raw=out['det_raw'][:1], det_box=out['det_box'][:1], det_count=out['det_count'][:1],
                  masks=out['masks'][:1])
        for _ in range(10):
            e1b.forward(i1, w1, THRESH, out=o1)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.latency):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            e1b.forward(i1, w1, THRESH, out=o1)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        lat = dict(p50_ms=ts[len(ts) // 2], p90_ms=ts[int(len(ts) * 0.9)], iters=args.latency, mode='eager launches')
        try:
            g = e1b.capture_graph(i1, w1, THRESH, o1)
            for _ in range(10):
                g.replay()
            torch.cuda.synchronize()
            tg = []
            for _ in range(args.latency):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.replay()
                b.record()
                b.synchronize()
                tg.append(a.elapsed_time(b))
            tg.sort()
            lat['graph_p50_ms'] = tg[len(tg) // 2]
            lat['graph_p90_ms'] = tg[int(len(tg) * 0.9)]
        except Exception as e:      # graph capture is an optimisation of the latency path only
            lat['graph_error'] = str(e)[:200]
        e1b.close()

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if affinity0 is not None:
            os.sched_setaffinity(0, affinity0)          # the CPU baseline uses every host core again
        torch.set_num_threads(os.cpu_count() or 1)
        ips, total, cores = cpu_reference_images_per_s(8, 1, W)
        cpu = dict(value=ips, unit='images/s', cores=cores, kind='port',
                   sample='8 steps x %d images 576x576 through oracle.evaluate (median step), %.1f s CPU wall'
                          % (CPU_SAMPLE_BATCH, total))

    # ---- BASELINE configs[3]: the data-parallel training step, every rank (same torchrun launch) ----
    train = None
    if not args.no_train:
        eng.close()                                   # free the batch-64 inference engine first
        torch.cuda.empty_cache()
        train = train_leg(16, 'bf16', max(5, min(args.steps, 20)), 3, rank, local, world, dist)
        # the reference's stage 2: every layer unlocked (yolo3_net_pos.py:155-156), 61.66 M trainables
        st2 = train_leg(16, 'bf16', max(3, min(args.steps, 10)), 3, rank, local, world, dist, lock=[0] * 82)
        train['stage2'] = dict(value=st2['value'], unit='images/s', ms_per_step=st2['ms_per_step'],
                               trainable_params=st2['trainable_params'], buckets=st2['buckets'],
                               allreduce_exposed_ms=st2.get('allreduce_exposed_ms'),
                               gpu_launches_per_step=st2['gpu_launches_per_step'])
    # ---- BASELINE configs[4]: stress (rank 0, N=1) ----
    stress = None
    if not args.no_stress and rank == 0 and world == 1:
        try:
            eng.close()
        except Exception:
            pass
        torch.cuda.empty_cache()
        stress = stress_leg(local, peaks)

    # DRAM bytes of the dominant kernel: not measurable without a profiler, so the figure is the one ncu measured on
    # this code (profiles/traffic.json, written by scripts/traffic_from_launches.py from the committed launch list of
    # this very command) -- or null
    traffic, traffic_src = args.traffic, ('--traffic' if args.traffic is not None else None)
    if traffic is None:
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                tj = json.load(f)
            traffic, traffic_src = tj['conv_tc_dram_bytes_per_step'], tj['source']
        except Exception:
            traffic = None
    if rank == 0:
        line = dict(
            metric=METRIC, value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_step, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
            data='synthetic',
            config=dict(workload='DIS-YOLO bf16 inference batch %d/GPU at 576x576 (BASELINE configs[1]/[2])' % B,
                        per_gpu_batch=B, image=IMAGE, det_thresh=THRESH, max_detection=md,
                        weights='lively random init seed 0',
                        cache='inputs (255 MB) and activations (GBs) exceed the 126 MB L2; no flush needed',
                        detections_per_step=dets, steady_state_warmup_steps=steady_steps),
            clocks=clocks,
            e2e=dict(value=e2e_value, unit='images/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     steps=n_e2e,
                     mode='3 batches in flight: dy_forward_host_begin_u8 / dy_forward_host_end_cropped -- uint8 '
                          'letterboxed images in (divided by 255 on the device, bit-identical input), boxes + '
                          'counts + the box-cropped mask maps det_mask[y1:y2,x1:x2] out', numa=numa),
            e2e_reference_layout=e2e_ref,
            e2e_pipeline=pipe_line,
            gpu_launches=launches,
            roofline=dict(bound='tensor', kernel='conv_tc_kernel (80 launches/step: layers 2..81, convolutional82 evaluated inside 81\'s epilogue)',
                          achieved=achieved_tf, peak=peaks['tf_sustained'], unit='TFLOP/s',
                          frac=achieved_tf / peaks['tf_sustained'], frac_of_burst=achieved_tf / peaks['tf_burst'],
                          peak_source=peaks['source'] + ' (sustained bf16; kernel timed inside a long step)',
                          algorithmic_flops_per_step=tc_flops, kernel_ms_per_step=tc_ms,
                          network_ms_per_step=net_ms, traffic=traffic, traffic_source=traffic_src),
            roofline_extra=extra, latency_batch1=lat, cpu_baseline=cpu, train=train, stress=stress,
            per_layer=per_layer)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def synth_train_batch(B, rank):
    """Synthetic training feed of BASELINE configs[3] in defect_train.get()'s formats (utils/train_data.py:44-52,
    146-178, 258-265): 6 boxes per image, best-anchor label assignment, rectangular bool masks, fixed RoI permutations."""
    import numpy as np
    rng = np.random.default_rng(7 + rank)
    img = rng.random((B, IMAGE, IMAGE, 3), dtype=np.float32)
    base = IMAGE // 32
    labels = [np.zeros((B, base * m, base * m, 3, 8), np.float32) for m in (4, 2, 1)]
    tb = np.zeros((B, 20, 5), np.float32)
    tm = np.zeros((B, 20, IMAGE, IMAGE), np.uint8)
    anchors = np.array([[31, 23], [62, 58], [143, 91], [213, 186], [61, 337], [194, 432], [474, 248], [551, 93],
                        [478, 454]], np.float32)
    for b in range(B):
        for j in range(6):
            w, h = rng.uniform(0.1, 0.6, 2) * IMAGE
            xc, yc = rng.uniform(w / 2, IMAGE - w / 2), rng.uniform(h / 2, IMAGE - h / 2)
            cls = int(rng.integers(0, 3))
            tb[b, j] = [xc / IMAGE, yc / IMAGE, w / IMAGE, h / IMAGE, cls]
            tm[b, j, int(yc - h / 2):int(yc + h / 2), int(xc - w / 2):int(xc + w / 2)] = 1
            inter = np.minimum(w, anchors[:, 0]) * np.minimum(h, anchors[:, 1])
            a = int(np.argmax(inter / (w * h + anchors[:, 0] * anchors[:, 1] - inter)))
            lab = labels[a // 3]
            g = lab.shape[1]
            lab[b, int(yc * g / IMAGE), int(xc * g / IMAGE), a % 3] = [xc / IMAGE, yc / IMAGE, w / IMAGE, h / IMAGE, 1,
                                                                      cls == 0, cls == 1, cls == 2]
    pp = np.stack([rng.permutation(30) for _ in range(B)]).astype(np.int32)
    pg = np.stack([rng.permutation(20) for _ in range(B)]).astype(np.int32)
    return img, labels, tb, tm, pp, pg


def train_leg(B, precision, steps, warm, rank, local, world, dist, e2e_steps=0, lock=None):
    """BASELINE configs[3]: training step (forward + losses + backward + all-reduce + Adam), batch B per GPU at
    576x576, data parallel (bucketed NCCL all-reduce overlapped with backward).  Returns the result dict."""
    import torch
    import disyolo_b200 as dy
    eng = dy.Engine(image_size=IMAGE, max_batch=B, precision=precision, device=local, lock=lock)
    eng.load_weights(dy.init_weights('lively', 0))
    tr = dy.DataParallelTrainer(eng, bucket_mb=25)
    img_h, labels_h, tb_h, tm_h, pp_h, pg_h = synth_train_batch(B, rank)
    # inputs resident in HBM before the timed region (the host-buffer protocol is timed separately below)
    img = torch.from_numpy(img_h).cuda()
    labels = [torch.from_numpy(l).cuda() for l in labels_h]
    tb, tm = torch.from_numpy(tb_h).cuda(), torch.from_numpy(tm_h).cuda()
    pp, pg = torch.from_numpy(pp_h).cuda(), torch.from_numpy(pg_h).cuda()

    def timed(n):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = tr.step(img, labels, tb, tm, pp, pg, THRESH, 1e-4)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n, out

    for _ in range(warm):
        losses = tr.step(img, labels, tb, tm, pp, pg, THRESH, 1e-4)
    eng.lib.dy_launch_count(1)
    ms_step, losses = timed(steps)
    launches = int(eng.lib.dy_launch_count(0))
    res = dict(metric='training images/s @576^2 (fwd+losses+bwd+allreduce+Adam)', value=world * B / (ms_step / 1e3),
               unit='images/s', ms_per_step=ms_step, steps=steps, per_gpu_batch=B, n_gpus=world,
               dtype='bf16' if precision == 'bf16' else 'f32',
               stage='1 (layers 53-82 trainable)' if lock is None else 'custom lock (%d layers trainable)' % lock.count(0),
               trainable_params=eng.n_train, buckets=len(tr.buckets),
               gradient_bytes=int(eng.n_train) * 4, losses=[float(v) for v in losses],
               gpu_launches_per_step=launches // max(1, steps))
    if world > 1:
        # the same bucketed step with the collectives left out: the difference is the all-reduce time that the
        # overlap with the backward pass did NOT hide (ranks diverge from here on; this engine is discarded)
        tr.skip_allreduce = True
        ms_nocomm, _ = timed(max(3, steps // 2))
        tr.skip_allreduce = False
        res['allreduce_exposed_ms'] = max(0.0, ms_step - ms_nocomm)
        res['ms_per_step_without_allreduce'] = ms_nocomm
    if e2e_steps:
        # end to end: the reference's feed_dict protocol -- host numpy images / labels / masks uploaded every step
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            tr.step(img_h, labels_h, tb_h, tm_h, pp_h, pg_h, THRESH, 1e-4)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        h2d = img_h.nbytes + sum(l.nbytes for l in labels_h) + tb_h.nbytes + tm_h.nbytes + pp_h.nbytes + pg_h.nbytes
        res['e2e'] = dict(value=world * B * e2e_steps / e2e_s, unit='images/s', h2d_bytes_per_step=int(h2d),
                          d2h_bytes_per_step=32, steps=e2e_steps,
                          mode='host numpy feed_dict every step (pageable), losses read back')
    eng.close()
    del tr, eng
    torch.cuda.empty_cache()
    return res


def stress_leg(local, peaks, B=4, size=1152, max_det=1000, thresh=0.05):
    """BASELINE configs[4]: 1152 x 1152 inputs, low score threshold, >= 1,000 boxes per image through NMS and
    position-sensitive mask assembly.  The network pass is timed on synthetic images; decode / NMS / top-k / masks on
    synthetic head maps whose boxes are small enough that >= 1,000 survive NMS (random-init head maps overlap too
    much for that), the 576 x 576 x 9 score maps being the network's own."""
    import numpy as np
    import torch
    import disyolo_b200 as dy
    eng = dy.Engine(image_size=size, max_batch=B, precision='bf16', device=local, max_detection=max_det)
    eng.load_weights(dy.init_weights('lively', 0))
    rng = np.random.default_rng(26)
    img = torch.from_numpy(rng.random((B, size, size, 3), dtype=np.float32)).cuda()
    win = torch.tensor([[0, 0, 1, 1]], dtype=torch.float32).repeat(B, 1).cuda()
    for _ in range(2):
        eng.forward_network(img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.forward_network(img)
    e1.record()
    torch.cuda.synchronize()
    net_ms = e0.elapsed_time(e1) / 5
    heads = []
    for s in (8, 16, 32):
        g = size // s
        y = rng.standard_normal((B, g, g, 3, 8)).astype(np.float32)
        y[..., 2:4] = y[..., 2:4] * 0.5 - 1.5          # small boxes: many survive the IoU > 0.3 suppression
        y[..., 4] = y[..., 4] * 1.5                     # dense objectness
        y[..., 5:] *= 2.0
        heads.append(torch.from_numpy(y).cuda())
    mp = eng.mask_pos(B).permute(0, 3, 1, 2).contiguous()          # planar [B,9,S,S], the engine's own layout
    sm = eng.mask_size
    masks = torch.empty((B, max_det, sm, sm), dtype=torch.float32, device='cuda')
    raw, box, cnt = eng.detect(heads, win, thresh)
    pp = eng.postproc_profile(heads, mp, win, thresh, masks, layout='planar', reps=5)
    dets = cnt.cpu().numpy()
    mask_bytes = int(dets.sum()) * sm * sm * 4
    res = dict(workload='stress: %dx%d, batch %d, det_thresh %.2f, max_detection %d' % (size, size, B, thresh, max_det),
               detections_per_image=[int(v) for v in dets], min_detections=int(dets.min()),
               candidates_per_image=eng.num_candidates,
               network_ms=net_ms, network_images_per_s=B / (net_ms / 1e3),
               decode_ms=pp['decode'], nms_ms=pp['nms'], finalize_ms=pp['finalize'], masks_ms=pp['masks'],
               mask_bytes=mask_bytes, mask_gbs=mask_bytes / (pp['masks'] / 1e3) / 1e9,
               mask_frac_of_hbm=mask_bytes / (pp['masks'] / 1e3) / 1e9 / peaks['hbm_gbs'],
               hbm_peak_note='MEASURED_PEAKS hbm_gbs is a read+write copy figure; a store-dominated kernel can exceed it',
               decode_bytes=B * eng.num_candidates * 32,
               decode_frac_of_hbm=B * eng.num_candidates * 32 / (pp['decode'] / 1e3) / 1e9 / peaks['hbm_gbs'])
    eng.close()
    del masks, eng
    torch.cuda.empty_cache()
    return res


def run_train(args):
    """`--workload train`: the training leg alone, printed as its own JSON line."""
    import torch
    rank, local, world = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        dist = None
    B = args.batch if args.batch != PER_GPU_BATCH else 16
    steps, warm = max(1, args.steps), max(3, args.warmup)
    res = train_leg(B, args.train_precision, steps, warm, rank, local, world, dist, e2e_steps=max(2, min(steps, 5)),
                    lock=[0] * 82 if args.train_stage == 2 else None)
    if rank == 0:
        e2e = res.pop('e2e')
        print(json.dumps(dict(metric=res.pop('metric'), value=res.pop('value'), unit='images/s', n_gpus=world,
                              steps=steps, warmup=warm, ms_per_step=res.pop('ms_per_step'), higher_is_better=True,
                              scaling='weak', vs_baseline=None, dtype=res['dtype'], data='synthetic',
                              config=dict(workload='DIS-YOLO training step, batch %d/GPU at 576x576, stage 1 '
                                                   '(layers 53-82 trainable), data parallel' % B,
                                          trainable_params=res['trainable_params'], buckets=res['buckets']),
                              e2e=e2e, losses=res['losses'], gpu_launches=res['gpu_launches_per_step'] * steps,
                              train=res)))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=PER_GPU_BATCH)
    ap.add_argument('--latency', type=int, default=200, help='batch-1 latency iterations (0 = skip)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-pipeline', action='store_true', help='skip the whole-image pipeline e2e leg')
    ap.add_argument('--workload', default='inference', choices=['inference', 'train'])
    ap.add_argument('--train-precision', default='bf16', choices=['bf16', 'fp32'],
                    help='training engine: bf16 = tcgen05 dgrad/wgrad (mixed precision), fp32 = verification engine')
    ap.add_argument('--train-stage', type=int, default=1, choices=[1, 2],
                    help='--workload train: 1 = layers 53-82 trainable (reference stage 1), 2 = every layer unlocked')
    ap.add_argument('--traffic', type=float, default=None,
                    help='DRAM bytes per step of the conv kernel (sum over its launches) from an ncu launch list of '
                         'THIS build (profiles/); omitted -> roofline.traffic is null (never a stale constant)')
    ap.add_argument('--no-numa', action='store_true', help='do not bind the process to the GPU NUMA node')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step leg (BASELINE configs[3])')
    ap.add_argument('--no-steady', action='store_true', help='skip the steady-state warm-up (profiler runs)')
    ap.add_argument('--no-stress', action='store_true', help='skip the 1152^2 stress leg (BASELINE configs[4])')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        import __graft_entry__ as g
        rank, _, _ = dist_env()
        if rank == 0:
            g.build()
        if os.environ.get('DY_OPTS'):          # A/B aid: DY_OPTS="wgrad_fuse_kw=0,tc_halo=0" -> dy_set_option
            from disyolo_b200.engine import set_option
            for kv in os.environ['DY_OPTS'].split(','):
                k, v = kv.split('=')
                set_option(k.strip(), int(v))
        if args.workload == 'train':
            run_train(args)
        else:
            run_ours(args)


if __name__ == '__main__':
    main()
