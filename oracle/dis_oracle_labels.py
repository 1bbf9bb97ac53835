"""CPU oracle for the training label assignment  --  TEST INFRASTRUCTURE ONLY (see dis_oracle.py).

NumPy restatement of the label part of defect_train.get: utils/train_data.py:134-178 (box transform, best-IoU
anchor, first box wins a cell), :189-228 (horizontal / vertical flip of the label grids), :258-262
(normalisation by image_size).  PARITY STATUS: unpinned -- utils/train_data.py imports pyblur / skimage
(absent) and the assignment sits inside get(), behind dataset file I/O; it is restated from the source.
"""
import numpy as np

ANCHORS = np.array([[31, 23], [62, 58], [143, 91], [213, 186], [61, 337], [194, 432], [474, 248], [551, 93],
                    [478, 454]], np.float64)


def assign_labels(boxes, nbox, place, flip, net, num_class=3, max_box=20, anchors=ANCHORS):
    """boxes [B,max_box,5] (x1,y1,x2,y2,class) in original pixels -> (yolo3, yolo2, yolo1, true_boxes)."""
    B = len(boxes)
    base = net // 32
    out = [np.zeros((B, base * m, base * m, 3, 5 + num_class), np.float32) for m in (4, 2, 1)]
    true_boxes = np.zeros((B, max_box, 5), np.float32)
    for b in range(B):
        sx, sy, dx, dy = [float(v) for v in place[b]]
        yolos = [o[b] for o in out]
        bbox = np.zeros((max_box, 5), np.float32)
        bbox[:nbox[b]] = boxes[b, :nbox[b]]
        for index in range(nbox[b]):
            cls_ind = int(bbox[index, 4])
            x1, y1, x2, y2 = [float(v) for v in bbox[index, :4]]
            x1 = max(min(x1 * sx + dx, net - 1), 0)
            y1 = max(min(y1 * sy + dy, net - 1), 0)
            x2 = max(min(x2 * sx + dx, net - 1), 0)
            y2 = max(min(y2 * sy + dy, net - 1), 0)
            bx = [(x2 + x1) / 2.0, (y2 + y1) / 2.0, x2 - x1, y2 - y1]
            bbox[index, :4] = bx
            anchors_min = -np.asarray(anchors / 2., dtype='float32')
            anchors_max = -anchors_min
            anchors_areas = anchors_max[:, 1] * anchors_max[:, 0] * 4
            half = np.asarray(bx[2:4]) / np.array([2, 2])
            box_half = np.repeat(np.asarray(half, dtype='float32').reshape((1, 2)), 9, axis=0)
            box_min, box_max = -box_half, box_half
            box_areas = box_half[:, 0] * box_half[:, 1] * 4
            inter_box = np.maximum(np.minimum(box_max, anchors_max) - np.maximum(box_min, anchors_min), 0.)
            inter = inter_box[:, 0] * inter_box[:, 1]
            with np.errstate(invalid='ignore', divide='ignore'):
                iou = inter / (box_areas + anchors_areas - inter)
            if np.max(iou) > 0:
                k = int(np.argmax(iou))
                yolo = yolos[k // 3]
                gh, gw = yolo.shape[0], yolo.shape[1]
                x_ind, y_ind = int(bx[0] * gw / net), int(bx[1] * gh / net)
                if yolo[y_ind, x_ind, k % 3, 4] == 1:
                    continue
                yolo[y_ind, x_ind, k % 3, 0:4] = bx
                yolo[y_ind, x_ind, k % 3, 4] = 1
                yolo[y_ind, x_ind, k % 3, 5 + cls_ind] = 1.
        f = int(flip[b]) if flip is not None else 1
        idx = np.arange(nbox[b])
        if f == 2:
            bbox[idx, 0] = net - 1 - bbox[idx, 0]
            for s in range(3):
                y = yolos[s][:, ::-1].copy()
                obj = y[..., 4] == 1
                y[obj, 0] = (net - 1 - y[obj, 0].astype(np.float64)).astype(np.float32)
                yolos[s][...] = y
        elif f == 3:
            bbox[idx, 1] = net - 1 - bbox[idx, 1]
            for s in range(3):
                y = yolos[s][::-1].copy()
                obj = y[..., 4] == 1
                y[obj, 1] = (net - 1 - y[obj, 1].astype(np.float64)).astype(np.float32)
                yolos[s][...] = y
        bbox[:, 0:4] = bbox[:, 0:4] / net
        for s in range(3):
            yolos[s][..., 0:4] = yolos[s][..., 0:4] / net
        true_boxes[b] = bbox
    return out[0], out[1], out[2], true_boxes
