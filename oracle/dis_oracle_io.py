"""CPU oracle for the steps either side of the hot path  --  TEST INFRASTRUCTURE ONLY (see dis_oracle.py).

  letterbox    image_read            calculate_test_map.py:149-176, utils/val_data.py:36-63
  postprocess  the per-detection loop calculate_test_map.py:233-269 (= utils/validation_map.py:137-166)
               with correct_yolo_boxes (:121-138)
Both call cv2.resize(..., INTER_LINEAR) exactly as the reference does (OpenCV is the reference's own
dependency and is present in this image).  PARITY STATUS: letterbox is PINNED against the reference's
defect_val.image_read run here on seeded images (tests/golden/ref_kat.json, key 'letterbox_synthetic');
correct_yolo_boxes is pinned by the golden key 'correct_yolo_boxes'; the loop itself is a restatement.
"""
import numpy as np


def letterbox(image_rgb, image_size):
    """image_rgb [h,w,3] uint8 -> (new_image [S,S,3] float64, window [4] float32) -- image_read."""
    import cv2
    window = np.array([0., 0., 1., 1.], dtype=np.float32)
    imgh, imgw, _ = image_rgb.shape
    if (float(image_size) / imgw) < (float(image_size) / imgh):
        imgh = (imgh * image_size) // imgw
        imgw = image_size
    else:
        imgw = (imgw * image_size) // imgh
        imgh = image_size
    image = cv2.resize(image_rgb.astype(np.float32), (imgw, imgh), interpolation=cv2.INTER_LINEAR)
    top = (image_size - imgh) // 2
    left = (image_size - imgw) // 2
    window[0] = top / image_size
    window[1] = left / image_size
    window[2] = (imgh + top) / image_size
    window[3] = (imgw + left) / image_size
    new_image = np.ones((image_size, image_size, 3)) * 127.
    new_image[(image_size - imgh) // 2:(image_size + imgh) // 2, (image_size - imgw) // 2:(image_size + imgw) // 2, :] = image
    return new_image / 255.0, window


def correct_yolo_boxes(x1, y1, x2, y2, image_h, image_w, net_h, net_w):
    """calculate_test_map.py:121-138, evaluated in float64 (what NumPy 1.x does for float32 scalar op
    Python float, the reference's environment)."""
    if (float(net_w) / image_w) < (float(net_h) / image_h):
        new_w = net_w
        new_h = (image_h * net_w) // image_w
    else:
        new_h = net_h
        new_w = (image_w * net_h) // image_h
    x_offset, x_scale = float((net_w - new_w) // 2) / net_w, float(new_w) / net_w
    y_offset, y_scale = float((net_h - new_h) // 2) / net_h, float(new_h) / net_h

    def c(v, off, sc, ext):
        return int(max(min(int(np.around((float(v) - off) / sc * ext)), ext), 0))
    return c(x1, x_offset, x_scale, image_w), c(y1, y_offset, y_scale, image_h), \
        c(x2, x_offset, x_scale, image_w), c(y2, y_offset, y_scale, image_h)


def postprocess(det_box, det_mask, image_h, image_w, net_size):
    """det_box [n,6] (y1,x1,y2,x2 normalised, class, score), det_mask [n,S,S] ->
    (boxes [n,4] int (x1,y1,x2,y2), valid [n] bool, full_masks [n,h,w] bool, merged [h,w] uint8)."""
    import cv2
    n = len(det_box)
    boxes = np.zeros((n, 4), np.int32)
    valid = np.zeros((n,), bool)
    full = np.zeros((n, image_h, image_w), bool)
    merged = np.zeros((image_h, image_w), np.uint8)
    for k in range(n):
        y1n, x1n, y2n, x2n = [np.float32(v) for v in det_box[k, :4]]
        x1, y1, x2, y2 = correct_yolo_boxes(x1n, y1n, x2n, y2n, image_h, image_w, net_size, net_size)
        boxes[k] = (x1, y1, x2, y2)
        if (y2 - y1) * (x2 - x1) <= 0:
            continue
        pred_mask = det_mask[k]
        size = pred_mask.shape[0]
        cy1, cx1, cy2, cx2 = [int(np.around(v * np.float32(size)).astype(np.int32)) for v in (y1n, x1n, y2n, x2n)]
        crop = pred_mask[cy1:cy2, cx1:cx2]
        if crop.size == 0:          # cv2.resize would raise on an empty crop; treated as a skipped detection
            continue
        mask = cv2.resize(np.ascontiguousarray(crop, np.float32), (x2 - x1, y2 - y1), interpolation=cv2.INTER_LINEAR)
        mask = mask > 0.5
        full[k, y1:y2, x1:x2] = mask
        valid[k] = True
        merged[full[k]] = int(det_box[k, 4]) + 1
    return boxes, valid, full, merged
