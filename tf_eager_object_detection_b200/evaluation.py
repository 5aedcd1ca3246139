"""Detection wire / on-disk formats and PASCAL VOC AP — "next" row f4 (host side, rank 0, after the all-gather).

Consumes the `[B, K, 6]` detection records `(x1, y1, x2, y2, score, class)` + counts that `post_ops_prediction_batched` /
`distributed.allgather_detections` produce, and mirrors
  * the VOC per-class result files of `object_detection/evaluation/pascal_eval_files_utils.py:109-122`,
  * the COCO result json of `scripts/eval_coco.py:157-168` (category-id remap :22-62),
  * `voc_ap` / `voc_eval` of `object_detection/evaluation/detectron_pascal_evaluation_utils.py:53-222`.
This is CPU post-processing in the reference too (numpy); nothing here touches the GPU."""
import json
import os
import pickle
import xml.etree.ElementTree as ET

import numpy as np

__all__ = ['PASCAL_CLASSES', 'coco_category_ids', 'records_to_host', 'voc_result_lines', 'write_voc_results',
           'coco_results', 'write_coco_results', 'voc_ap', 'voc_match', 'voc_eval', 'parse_rec', 'get_prediction_files']

PASCAL_CLASSES = ('__background__', 'aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow',
                  'diningtable', 'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train',
                  'tvmonitor')                                     # pascal_eval_files_utils.py:9-13

_COCO_UNUSED = (12, 26, 29, 30, 45, 66, 68, 69, 71, 83)            # ids absent from the 80-class set


def coco_category_ids():
    """label (1..80) -> COCO category id, label 0 = background -> 0 (scripts/eval_coco.py:40-62)."""
    return [0] + [i for i in range(1, 91) if i not in _COCO_UNUSED]


def records_to_host(detections, counts):
    """device / host records -> (numpy [B,K,6], numpy [B])."""
    det = detections.detach().cpu().numpy() if hasattr(detections, 'detach') else np.asarray(detections)
    cnt = counts.detach().cpu().numpy() if hasattr(counts, 'detach') else np.asarray(counts)
    return det, cnt.astype(np.int64)


def voc_result_lines(image_ids, detections, counts, num_classes=21):
    """-> {class index: [lines]}; one line per detection '{id} {score:.3f} {x1+1:.1f} {y1+1:.1f} {x2+1:.1f} {y2+1:.1f}'
    (the VOCdevkit expects 1-based pixels, pascal_eval_files_utils.py:118-122), images in `image_ids` order and, inside
    an image, the class's detections in record order (descending score)."""
    det, cnt = records_to_host(detections, counts)
    lines = {c: [] for c in range(1, num_classes)}
    for b, image_id in enumerate(image_ids):
        d = det[b, :cnt[b]]
        cls = d[:, 5].astype(np.int64)
        for k in range(d.shape[0]):
            lines[int(cls[k])].append('{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n'.format(
                str(image_id), d[k, 4], d[k, 0] + 1, d[k, 1] + 1, d[k, 2] + 1, d[k, 3] + 1))
    return lines


def write_voc_results(result_file_format, image_ids, detections, counts, class_list=PASCAL_CLASSES):
    """`result_file_format.format(class_name)` per foreground class (pascal_eval_files_utils.py:109-122)."""
    lines = voc_result_lines(image_ids, detections, counts, len(class_list))
    paths = []
    for c, name in enumerate(class_list):
        if c == 0:
            continue
        path = result_file_format.format(name)
        with open(path, 'wt') as f:
            f.writelines(lines[c])
        paths.append(path)
    return paths


def coco_results(image_ids, detections, counts, label_to_category_id=None):
    """-> list of {'image_id', 'category_id', 'bbox': [x, y, w+1, h+1], 'score'} (scripts/eval_coco.py:157-165)."""
    det, cnt = records_to_host(detections, counts)
    cat = coco_category_ids() if label_to_category_id is None else label_to_category_id
    out = []
    for b, image_id in enumerate(image_ids):
        for x1, y1, x2, y2, s, c in det[b, :cnt[b]]:
            out.append({'image_id': int(image_id), 'category_id': int(cat[int(c)]),
                        'bbox': [float(x1), float(y1), float(x2 - x1 + 1), float(y2 - y1 + 1)], 'score': float(s)})
    return out


def write_coco_results(path, image_ids, detections, counts, label_to_category_id=None):
    res = coco_results(image_ids, detections, counts, label_to_category_id)
    with open(path, 'w') as f:
        json.dump(res, f)                                          # scripts/eval_coco.py:167-168
    return res


# ------------------------------------------------------------------------------------------------ PASCAL VOC AP
def parse_rec(filename):
    """One VOC annotation xml -> list of objects (detectron_pascal_evaluation_utils.py:34-50)."""
    objects = []
    for obj in ET.parse(filename).findall('object'):
        bb = obj.find('bndbox')
        objects.append({'name': obj.find('name').text, 'pose': obj.find('pose').text,
                        'truncated': int(obj.find('truncated').text), 'difficult': int(obj.find('difficult').text),
                        'bbox': [int(bb.find(k).text) for k in ('xmin', 'ymin', 'xmax', 'ymax')]})
    return objects


def voc_ap(rec, prec, use_07_metric=False):
    """detectron_pascal_evaluation_utils.py:53-83: 11-point VOC07 metric or the area under the precision envelope."""
    rec = np.asarray(rec, np.float64); prec = np.asarray(prec, np.float64)
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            m = rec >= t
            ap = ap + (np.max(prec[m]) if m.any() else 0) / 11.
        return ap
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]                 # precision envelope
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def voc_match(image_ids, confidence, boxes, class_recs, npos, ovthresh=0.5, use_07_metric=True):
    """The matching core of voc_eval (:165-222).  image_ids [D]; confidence [D]; boxes [D,4]; class_recs
    {image id: {'bbox' [G,4], 'difficult' [G] bool}}.  Detections in descending confidence (stable on -confidence, as
    np.argsort there); a detection is a TP when its best-overlap ground truth (+1 pixel areas, first maximum) exceeds
    `ovthresh`, is not difficult and is still unmatched."""
    confidence = np.asarray(confidence, np.float64)
    boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
    order = np.argsort(-confidence)
    boxes = boxes[order]
    ids = [image_ids[i] for i in order]
    taken = {k: np.zeros(len(v['bbox']), bool) for k, v in class_recs.items()}
    nd = len(ids)
    tp = np.zeros(nd); fp = np.zeros(nd)
    for d in range(nd):
        rec = class_recs[ids[d]]
        gt = np.asarray(rec['bbox'], np.float64).reshape(-1, 4)
        bb = boxes[d]
        ovmax, jmax = -np.inf, -1
        if gt.size > 0:
            iw = np.maximum(np.minimum(gt[:, 2], bb[2]) - np.maximum(gt[:, 0], bb[0]) + 1., 0.)
            ih = np.maximum(np.minimum(gt[:, 3], bb[3]) - np.maximum(gt[:, 1], bb[1]) + 1., 0.)
            inters = iw * ih
            uni = ((bb[2] - bb[0] + 1.) * (bb[3] - bb[1] + 1.) + (gt[:, 2] - gt[:, 0] + 1.) * (gt[:, 3] - gt[:, 1] + 1.)
                   - inters)
            overlaps = inters / uni
            jmax = int(np.argmax(overlaps))
            ovmax = overlaps[jmax]
        if ovmax > ovthresh:
            if not rec['difficult'][jmax]:
                if not taken[ids[d]][jmax]:
                    tp[d] = 1.
                    taken[ids[d]][jmax] = True
                else:
                    fp[d] = 1.
        else:
            fp[d] = 1.
    fp = np.cumsum(fp); tp = np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)


def voc_eval(detpath, annopath, imagesetfile, classname, cachedir, ovthresh=0.5, use_07_metric=True):
    """Same signature and files as detectron_pascal_evaluation_utils.py:86-222 (annotation cache pickle included)."""
    if not os.path.isdir(cachedir):
        os.mkdir(cachedir)
    imageset = os.path.splitext(os.path.basename(imagesetfile))[0]
    cachefile = os.path.join(cachedir, imageset + '_annots.pkl')
    with open(imagesetfile, 'r') as f:
        imagenames = [x.strip() for x in f.readlines()]
    if not os.path.isfile(cachefile):
        recs = {name: parse_rec(annopath.format(name)) for name in imagenames}
        with open(cachefile, 'wb') as f:
            pickle.dump(recs, f)
    else:
        with open(cachefile, 'rb') as f:
            recs = pickle.load(f)
    class_recs, npos = {}, 0
    for name in imagenames:
        objs = [o for o in recs[name] if o['name'] == classname]
        difficult = np.array([o['difficult'] for o in objs]).astype(bool)
        npos += int(np.sum(~difficult))
        class_recs[name] = {'bbox': np.array([o['bbox'] for o in objs]), 'difficult': difficult}
    with open(detpath.format(classname), 'r') as f:
        split = [x.strip().split(' ') for x in f.readlines()]
    ids = [x[0] for x in split]
    conf = np.array([float(x[1]) for x in split])
    bb = np.array([[float(z) for z in x[2:]] for x in split])
    return voc_match(ids, conf, bb, class_recs, npos, ovthresh, use_07_metric)


def get_prediction_files(cur_model, eval_dataset, image_sets, result_file_format='/path/to/results/{:s}.txt',
                         score_threshold=0.0, iou_threshold=0.5, max_objects_per_class=50, max_objects_per_image=50,
                         target_means=None, target_stds=None, min_size=10, class_list=PASCAL_CLASSES, group=None):
    """`evaluation/pascal_eval_files_utils.py:19-122` `get_prediction_files` from the model down: for every
    `(img, img_scale, raw_h, raw_w)` of `eval_dataset` (the reference builds it from `dataset_type` / `data_root_path`; data
    pipelines are out of scope, so the iterable and its `image_sets` ids are passed in) run the model's `im_detect`, the
    per-image detection filtering of the loop body (:76-106, on the device, no host synchronisation per image) and write the
    per-class VOC result files (:109-122).  Under torch.distributed every rank passes ITS shard of the dataset
    (`distributed.shard_bounds` order); the records are all-gathered once at the end and rank 0 writes the files for the
    full `image_sets`.  Returns (records [n,rows,6], counts [n]) of all images (host numpy)."""
    import torch
    import torch.distributed as dist
    from . import distributed as bxd
    from .prediction import eval_loop_detections
    num_classes = len(class_list)
    rows = (num_classes - 1) * max_objects_per_class       # room for every tie at the per-image cut (:99-106)
    recs, cnts = [], []
    for img, img_scale, raw_h, raw_w in eval_dataset:
        sm, tx, rois, roi_counts = cur_model.im_detect_batched(img)
        dev = sm.device
        det, cnt = eval_loop_detections(sm, tx, rois, torch.tensor([float(img_scale)], device=dev),
                                        torch.tensor([[float(raw_h), float(raw_w)]], device=dev), target_means, target_stds,
                                        score_threshold=score_threshold, iou_threshold=iou_threshold,
                                        max_objects_per_class=max_objects_per_class,
                                        max_objects_per_image=max_objects_per_image, min_size=min_size, loop='voc',
                                        out_rows=rows, roi_counts=roi_counts)
        recs.append(det)
        cnts.append(cnt)
    if recs:
        rec, cnt = torch.cat(recs), torch.cat(cnts)
    else:
        rec = torch.zeros((0, rows, 6), device='cuda'); cnt = torch.zeros((0,), dtype=torch.int32, device='cuda')
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if distributed:
        rec, cnt = bxd.allgather_detections(rec, cnt, group=group)
    rec_h, cnt_h = records_to_host(rec, cnt)
    if not distributed or dist.get_rank(group) == 0:
        if len(image_sets) != rec_h.shape[0]:
            raise ValueError('get_prediction_files: %d image ids for %d images' % (len(image_sets), rec_h.shape[0]))
        write_voc_results(result_file_format, image_sets, rec_h, cnt_h, class_list)
    return rec_h, cnt_h
