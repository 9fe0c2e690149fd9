"""Drop-in for the decode functions of the reference's eval/eval_coco2014_multi_modes.py:
process_single_scale (:263-443) and process_multi_scale (:58-261), same signatures and return
value (canvas, candidate (N,4) f64, subset (M,20) f64).

Everything after `model.predict` -- blob up-sampling, 4-scale average, sigma=3 smoothing, peak
test, PAF line-integral scoring, greedy matching, person assembly and pruning -- runs in the
sm_100a kernels behind include/rmpe_b200.h (rmpe_decode_batch_host).  Only the canvas drawing
(cv2.circle / fillConvexPoly, reference :417-441) stays on OpenCV; pass draw=False to skip it.
`decode_frames` decodes many frames per call.
"""
import math

import cv2
import numpy as np

from .. import batch as _batch
from .. import _lib as L
from ..util import padRightDownCorner

orderCOCO = [0, 1, 15, 14, 17, 16, 5, 2, 6, 3, 7, 4, 11, 8, 12, 9, 13, 10]

# 1-based part pairs per limb and the (x, y) channel pair of each limb in the 57-channel stack
limbSeq = [[2, 3], [2, 6], [3, 4], [4, 5], [6, 7], [7, 8], [2, 9], [9, 10],
           [10, 11], [2, 12], [12, 13], [13, 14], [2, 1], [1, 15], [15, 17],
           [1, 16], [16, 18], [3, 17], [6, 18]]
mapIdx = [[31, 32], [39, 40], [33, 34], [35, 36], [41, 42], [43, 44], [19, 20], [21, 22],
          [23, 24], [25, 26], [27, 28], [29, 30], [47, 48], [49, 50], [53, 54], [51, 52],
          [55, 56], [37, 38], [45, 46]]

colors = [[255, 0, 0], [255, 85, 0], [255, 170, 0], [255, 255, 0], [170, 255, 0], [85, 255, 0],
          [0, 255, 0], [0, 255, 85], [0, 255, 170], [0, 255, 255], [0, 170, 255], [0, 85, 255],
          [0, 0, 255], [85, 0, 255], [170, 0, 255], [255, 0, 255], [255, 0, 170], [255, 0, 85]]


def _raise_status(status):
    if status & L.ST_FOUND_GT2:
        # the reference indexes subset_idx[found] with found == 2 here (:192-195)
        raise IndexError("list assignment index out of range")
    if status & (L.ST_PEAK_OVERFLOW | L.ST_CAND_OVERFLOW | L.ST_PERSON_OVERFLOW):
        raise OverflowError("decode capacity exceeded (status 0x%x): raise max_peaks/max_cand/max_persons" % status)


def decode_frames(frames, params, model_params=None, **caps):
    """frames: list of dict(H, W, scales=[(paf, heat, pad_down, pad_right), ...]) (see
    batch.make_frames).  Returns the list of result dicts of batch.decode_batch_host."""
    stride = 8 if model_params is None else int(model_params['stride'])
    return _batch.decode_batch_host(frames, thre1=params['thre1'], thre2=params['thre2'], stride=stride, **caps)


_PEAK_RADIUS = 4
_LIMB_HALF_WIDTH = 4
_LIMB_BLEND = 0.6        # weight of the freshly painted limb over the canvas so far


def _limb_polygon(p0, p1):
    """Outline of the ellipse the reference paints for a limb between peaks p0 and p1 (x, y): centred on the
    truncated midpoint, half the limb long, _LIMB_HALF_WIDTH wide, rotated to whole degrees."""
    (xa, ya), (xb, yb) = p0, p1
    centre = (int((xa + xb) / 2.0), int((ya + yb) / 2.0))
    half_len = int(math.hypot(xa - xb, ya - yb) / 2)
    tilt = int(math.degrees(math.atan2(ya - yb, xa - xb)))
    return cv2.ellipse2Poly(centre, (half_len, _LIMB_HALF_WIDTH), tilt, 0, 360, 1)


def draw_canvas(canvas, candidate, subset, n_peaks):
    """Overlay equal to the reference's drawing tail (eval...:417-441; D7, outside the GPU path): a filled dot on every
    peak in its part colour, then limb by limb (the 17 body limbs, not the ear-shoulder pair) and person by person a
    filled ellipse blended 60:40 over everything painted before it."""
    part_of_row = np.repeat(np.arange(18), np.asarray(n_peaks[:18], dtype=int))
    for row, part in enumerate(part_of_row):
        cv2.circle(canvas, (int(candidate[row, 0]), int(candidate[row, 1])), _PEAK_RADIUS, colors[part], thickness=-1)
    for limb, (part_a, part_b) in enumerate(limbSeq[:17]):
        for person in subset:
            ia, ib = int(person[part_a - 1]), int(person[part_b - 1])
            if ia < 0 or ib < 0:
                continue
            painted = canvas.copy()
            cv2.fillConvexPoly(painted, _limb_polygon(candidate[ia, :2], candidate[ib, :2]), colors[limb])
            canvas = cv2.addWeighted(canvas, 1.0 - _LIMB_BLEND, painted, _LIMB_BLEND, 0)
    return canvas


def _finish(input_image, oriImg, r, draw):
    _raise_status(r["status"])
    candidate, subset = r["candidate"], r["subset"]
    canvas = None
    if draw:
        canvas = cv2.imread(input_image) if isinstance(input_image, str) else oriImg.copy()
        canvas = draw_canvas(canvas, candidate, subset, r["n_peaks"])
    return canvas, candidate, subset


def process_single_scale(input_image, model, params, model_params, draw=True, **caps):
    oriImg = cv2.imread(input_image) if isinstance(input_image, str) else np.asarray(input_image)  # B,G,R order
    input_img = np.transpose(np.float32(oriImg[:, :, :, np.newaxis]), (3, 0, 1, 2))  # (1, H, W, 3)
    output_blobs = model.predict(input_img)
    heatmap = np.squeeze(output_blobs[1], axis=0)   # output 1 is heatmaps
    paf = np.squeeze(output_blobs[0], axis=0)       # output 0 is PAFs
    frame = dict(H=oriImg.shape[0], W=oriImg.shape[1], scales=[(paf, heatmap, 0, 0)])
    r = decode_frames([frame], params, model_params, **caps)[0]
    return _finish(input_image, oriImg, r, draw)


def process_multi_scale(input_image, model, params, model_params, draw=True, **caps):
    oriImg = cv2.imread(input_image) if isinstance(input_image, str) else np.asarray(input_image)
    scale_search = list(params['scale_search'])
    multiplier = [x * model_params['boxsize'] / oriImg.shape[0] for x in scale_search]
    scales = []
    for scale in multiplier:
        imageToTest = cv2.resize(oriImg, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
        imageToTest_padded, pad = padRightDownCorner(imageToTest, model_params['stride'], model_params['padValue'])
        input_img = np.transpose(np.float32(imageToTest_padded[:, :, :, np.newaxis]), (3, 0, 1, 2))
        output_blobs = model.predict(input_img)
        scales.append((np.squeeze(output_blobs[0], axis=0), np.squeeze(output_blobs[1], axis=0), pad[2], pad[3]))
    frame = dict(H=oriImg.shape[0], W=oriImg.shape[1], scales=scales)
    r = decode_frames([frame], params, model_params, **caps)[0]
    return _finish(input_image, oriImg, r, draw)


# ------------------------------------------------------------------------------------------
# result writer (reference write_json, eval/eval_coco2014_multi_modes.py:514-548): the step right
# after the decode path.  Pure host formatting of (candidate, subset) -- no arithmetic on maps.
# ------------------------------------------------------------------------------------------
def coco_keypoint_records(candidate_set, subset_set, image_id_set, category_id=1):
    """List of COCO keypoint result dicts exactly as the reference builds them: 17 parts in
    orderCOCO with the neck skipped, (int x, int y, 2) or (0, 0, 0) for a missing part, score =
    subset[-2] (total score, not the mean)."""
    output_data = []
    for i in range(len(subset_set)):
        cand = np.asarray(candidate_set[i], dtype=np.float64).reshape(-1, 4)
        for person in np.asarray(subset_set[i], dtype=np.float64).reshape(-1, 20):
            keypoints = []
            for part in range(18):
                part_idx = orderCOCO[part]
                if part_idx == 1:      # skip neck for coco eval
                    continue
                idx = int(person[part_idx])
                if idx == -1:
                    keypoints += [0, 0, 0]
                else:
                    keypoints += [cand[idx, 0].astype(int), cand[idx, 1].astype(int), 2]
            output_data.append({"image_id": image_id_set[i], "category_id": category_id,
                                "keypoints": keypoints, "score": person[-2]})
    return output_data


def write_json(candidate_set, subset_set, image_id_set, json_file):
    """Same signature as the reference: `json_file` is an open text file, closed on return."""
    import json

    def _plain(o):
        if isinstance(o, np.generic):
            return o.item()
        raise TypeError(type(o))

    with json_file as outfile:
        json.dump(coco_keypoint_records(candidate_set, subset_set, image_id_set), outfile, default=_plain)


def gather_records(local_records, n_images):
    """Optional multi-GPU gather of the per-rank record lists: every rank decodes the
    images i % world == rank; rank order is restored by image position.  local_records: list of
    (position, [records of that image])."""
    from .. import shard as _shard
    merged = _shard.gather_results([dict(index=int(p), records=r) for p, r in local_records], n_images)
    return [rec for m in merged for rec in m["records"]]
