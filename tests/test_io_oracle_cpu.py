"""CPU: the I/O oracle (oracle/dis_oracle_io.py) against outputs of the reference's own functions run
in the build container (tests/golden/make_golden.py -> ref_kat.json)."""
import hashlib

import numpy as np
import pytest

from oracle import dis_oracle_io as IO


def test_letterbox_oracle_bit_exact_with_reference_image_read(golden):
    """defect_val.image_read (utils/val_data.py:36-63) was run on seeded images; the oracle must
    reproduce it bit for bit (both call cv2.resize INTER_LINEAR on the float32 image)."""
    pytest.importorskip('cv2')
    for case in golden['letterbox_synthetic']:
        rgb = np.random.default_rng(case['seed']).integers(0, 256, (case['h'], case['w'], 3), dtype=np.uint8)
        img, win = IO.letterbox(rgb, 576)
        assert [float(v) for v in win] == case['window']
        f32 = np.ascontiguousarray(img, np.float32)
        for y, x, v in case['samples']:
            assert [float(t) for t in f32[y, x]] == v
        assert hashlib.sha256(f32.tobytes()).hexdigest() == case['sha256_f32']


def test_correct_yolo_boxes_oracle_matches_reference(golden):
    for case in golden['correct_yolo_boxes']:
        got = IO.correct_yolo_boxes(*case['args'])
        assert [int(v) for v in got] == [int(v) for v in case['out']]


def test_postprocess_oracle_semantics():
    """Later detections overwrite earlier ones in the merged mask; degenerate boxes are skipped."""
    pytest.importorskip('cv2')
    S = 64
    m = np.zeros((3, S, S), np.float32)
    m[0, 8:40, 8:40] = 0.9
    m[1, 20:60, 20:60] = 0.9
    m[2] = 0.9
    det = np.array([[0.125, 0.125, 0.625, 0.625, 0, 0.9], [0.3125, 0.3125, 0.9375, 0.9375, 2, 0.8],
                    [0.5, 0.5, 0.5, 0.9, 1, 0.7]], np.float32)
    boxes, valid, full, merged = IO.postprocess(det, m, 128, 128, 576)
    assert valid.tolist() == [True, True, False]
    assert boxes[0].tolist() == [16, 16, 80, 80] and boxes[1].tolist() == [40, 40, 120, 120]
    assert merged[20, 20] == 1 and merged[100, 100] == 3 and merged[60, 60] == 3 and merged[5, 5] == 0
    assert full[0].sum() > 0 and full[2].sum() == 0
