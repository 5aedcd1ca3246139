"""f4: detection file formats + PASCAL VOC AP against goldens produced by the reference's own voc_eval
(oracle/make_golden.py:voc_golden) on the synthetic VOC tree of oracle/voc_fixture.py."""
import hashlib
import json
import os

import numpy as np

from oracle.voc_fixture import synthetic_voc, write_voc_tree
from tf_eager_object_detection_b200 import evaluation as ev, synthetic as syn


def _tree(tmp_path):
    classes = ev.PASCAL_CLASSES[:5]
    gts, det, cnt = synthetic_voc(np.random.default_rng(syn.seed_for(1, 80)))
    names = ['%06d' % (i + 1) for i in range(len(gts))]
    root = str(tmp_path)
    write_voc_tree(root, names, gts, classes)
    ev.write_voc_results(os.path.join(root, 'det_{:s}.txt'), names, det, cnt, classes)
    return root, classes, names, gts, det, cnt


def test_voc_files_and_ap_match_reference(golden, tmp_path):
    root, classes, names, gts, det, cnt = _tree(tmp_path)
    blob = b''.join(open(os.path.join(root, 'det_%s.txt' % c), 'rb').read() for c in classes[1:])
    assert np.array_equal(np.frombuffer(hashlib.sha256(blob).digest(), np.uint8), golden['voc_det_sha'])
    first = open(os.path.join(root, 'det_%s.txt' % classes[1])).readline().split(' ')
    assert len(first) == 6 and first[0] in names and len(first[1].split('.')[1]) == 3      # id score(.3f) 4 coords(.1f)
    for metric07 in (True, False):
        for c in classes[1:]:
            rec, prec, ap = ev.voc_eval(os.path.join(root, 'det_{:s}.txt'), os.path.join(root, 'Annotations', '{:s}.xml'),
                                        os.path.join(root, 'test.txt'), c, os.path.join(root, 'cache'), 0.5, metric07)
            tag = 'voc_%s_%s' % (c, '07' if metric07 else 'area')
            assert np.array_equal(rec, golden[tag + '_rec']) and np.array_equal(prec, golden[tag + '_prec'])
            assert ap == float(golden[tag + '_ap'])
    assert os.path.isfile(os.path.join(root, 'cache', 'test_annots.pkl'))                    # annotation cache, :118-141


def test_voc_ap_edge_cases():
    assert ev.voc_ap([], [], True) == 0 and ev.voc_ap([], [], False) == 0
    assert np.isclose(ev.voc_ap([0.5, 1.0], [1.0, 1.0], False), 1.0)
    assert np.isclose(ev.voc_ap([0.5, 1.0], [1.0, 1.0], True), 1.0)
    rec, prec, ap = ev.voc_match([], [], np.zeros((0, 4)), {}, 3)
    assert rec.size == 0 and ap == 0
    # a detection whose best match is a difficult object is neither TP nor FP; duplicates of a taken gt are FPs
    recs = {'a': {'bbox': np.array([[10, 10, 50, 50], [100, 100, 150, 150]]), 'difficult': np.array([False, True])}}
    rec, prec, ap = ev.voc_match(['a', 'a', 'a'], [0.9, 0.8, 0.7],
                                 [[10, 10, 50, 50], [12, 10, 50, 50], [100, 100, 150, 150]], recs, 1)
    assert rec.tolist() == [1.0, 1.0, 1.0] and prec.tolist() == [1.0, 0.5, 0.5]


def test_coco_results_format(tmp_path):
    ids = ev.coco_category_ids()
    assert len(ids) == 81 and ids[1] == 1 and ids[12] == 13 and ids[80] == 90 and ids[45] == 50
    det = np.zeros((2, 3, 6), np.float32)
    det[0, 0] = [10, 20, 30.5, 60, 0.9, 12]
    det[1, 0] = [0, 0, 5, 5, 0.4, 1]
    det[1, 1] = [1, 2, 3, 4, 0.3, 80]
    res = ev.write_coco_results(str(tmp_path / 'r.json'), [139, 285], det, np.array([1, 2]))
    assert json.load(open(tmp_path / 'r.json')) == res and len(res) == 3
    assert res[0] == {'image_id': 139, 'category_id': 13, 'bbox': [10.0, 20.0, 21.5, 41.0], 'score': float(np.float32(0.9))}
    assert res[2]['category_id'] == 90 and res[2]['bbox'] == [1.0, 2.0, 3.0, 3.0]
