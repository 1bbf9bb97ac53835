"""CPU: the C-ABI library builds, loads and exports every symbol include/disyolo.h declares; the
host-side mirror of the reference interface is consistent; no compute call succeeds without a GPU
(there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'disyolo.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dy_[a-z_0-9]+)\s*\(', src)))


def test_header_symbols_exported(built_lib):
    h = ctypes.CDLL(built_lib)
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(h, s), 'missing export %s' % s


def test_ctypes_binding_covers_header(built_lib):
    from disyolo_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.lib()
    assert b'sm_100a' in lib.dy_version()
    assert ctypes.sizeof(_lib.DyConfig) == 4 + 72 + 4 * 9 + 84      # struct dy_config layout


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from disyolo_b200 import _lib
    lib = _lib.lib()
    assert lib.dy_device_count() == 0
    cfg = _lib.DyConfig()
    cfg.num_classes, cfg.image_size, cfg.k_map, cfg.max_batch, cfg.max_detection = 3, 64, 3, 1, 30
    h = ctypes.c_void_p()
    rc = lib.dy_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and b'no CPU fallback' in lib.dy_last_error()
    import disyolo_b200 as dy
    with pytest.raises(_lib.DisYoloError):
        dy.Engine(image_size=64)


def test_invalid_config_is_reported(built_lib):
    from disyolo_b200 import _lib
    lib = _lib.lib()
    cfg = _lib.DyConfig()
    cfg.num_classes, cfg.image_size, cfg.k_map, cfg.max_batch, cfg.max_detection = 3, 100, 3, 1, 30
    h = ctypes.c_void_p()
    assert lib.dy_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b'multiple of 32' in lib.dy_last_error()
    assert lib.dy_create(None, ctypes.byref(h)) == -1


def test_config_surface_matches_reference(golden):
    import disyolo_b200.yolo.config as cfg
    g = golden['config']
    for name, val in g.items():
        got = getattr(cfg, name)
        if name == 'ANCHORS':
            assert np.array_equal(np.asarray(got), np.asarray(val, np.float32))
        else:
            assert got == val, name


def test_layer_tables_agree():
    from disyolo_b200 import layer_table
    from oracle import dis_oracle as O
    t, o = layer_table(), O.layer_table()
    assert [L['id'] for L in t] == list(range(1, 83))
    sizes = O._output_sizes(576)
    for L in t:
        r = o[L['id']]
        assert (L['cin'], L['cout'], L['k'], L['s'], L['bn'], L['res']) == \
               (r['cin'], r['cout'], r['k'], r['s'], r['bn'], r['res'])
        assert 576 // L['size_div'] == sizes[L['id']]


def test_weight_names_and_shapes():
    import disyolo_b200 as dy
    from oracle import dis_oracle as O
    w = dy.init_weights('reference', 1)
    ow = O.make_weights('faithful', 1)
    assert sorted(w) == sorted(ow)
    for k in w:
        assert w[k].shape == ow[k].shape and w[k].dtype == np.float32
    assert np.abs(w['yolo/convolutional7/weights']).max() <= 0.002     # truncated normal, locked layer
    assert np.all(w['yolo/convolutional59/biases'] == 0)
    lv = dy.init_weights('lively', 0)
    assert sorted(lv) == sorted(w)


def test_npz_roundtrip(tmp_path):
    import disyolo_b200 as dy
    w = {k: v for k, v in dy.init_weights('reference', 0).items() if 'convolutional82' in k}
    p = str(tmp_path / 'w.npz')
    dy.save_npz(p, w)
    r = dy.load_npz(p)
    assert sorted(r) == sorted(w) and all(np.array_equal(r[k], w[k]) for k in w)


def test_yolonet_surface_without_gpu(built_lib):
    """Attribute names of the reference class exist; construction itself needs a GPU."""
    import torch
    from disyolo_b200.yolo import yolo3_net_pos as m
    for attr in ('YOLONet', 'Session'):
        assert hasattr(m, attr)
    if not torch.cuda.is_available():
        from disyolo_b200 import _lib
        with pytest.raises(_lib.DisYoloError):
            m.YOLONet(False)


def test_every_set_option_name_is_documented_in_the_header():
    """dy_set_option switches (csrc/net.cu) <-> the list in include/disyolo.h: a switch nobody can look up is a trap."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, 'dis-yolo_b200', 'csrc', 'net.cu')).read()
    body = src[src.index('int dy_set_option('):]
    body = body[:body.index('\n}\n')]
    names = set(re.findall(r'n == "([a-z0-9_]+)"', body))
    assert len(names) >= 20, names
    hdr = open(os.path.join(root, 'include', 'disyolo.h')).read()
    doc = hdr[hdr.index('Tuning / test / measurement overrides'):hdr.index('int dy_set_option(')]
    missing = sorted(n for n in names if '"%s"' % n not in doc)
    assert not missing, 'undocumented dy_set_option names: %s' % missing


def test_committed_traffic_figure_matches_the_committed_launch_list():
    """bench.py reports roofline.traffic from profiles/traffic.json; it must be what the committed ncu launch list of
    the bench command says (scripts/traffic_from_launches.py), not a stale constant."""
    import csv
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tj = json.load(open(os.path.join(root, 'profiles', 'traffic.json')))
    rows = list(csv.reader(open(os.path.join(root, 'profiles', 'r2_launches.csv'), errors='replace')))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot, launches = 0.0, 0
    for r in rows[hi + 1:]:
        if len(r) <= vi or 'conv_tc_kernel' not in r[ki]:
            continue
        if r[mi] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            tot += float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
        if r[mi] == 'gpu__time_duration.sum':
            launches += 1
    assert launches == 2 * tj['conv_tc_launches_per_step'] == 160
    assert abs(tot / 2 - tj['conv_tc_dram_bytes_per_step']) < 1e-6 * tot
