"""The two callers of the box path — `BaseFasterRcnn` and `BaseFPN` mirrors (tf_eager_object_detection_b200/models.py) —
with small torch stand-ins for the parts that stay outside the path (extractor, neck, RPN head, RoI head): the evaluation
forward must equal the composition of the stage oracles on the tensors that flow between the stages, `im_detect` must agree
with it, and the training forward must return four finite losses whose gradients reach every stand-in."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import boxpath_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu

CFG = dict(num_classes=21, weight_decay=0.0001, ratios=[0.5, 1.0, 2.0], rpn_proposal_means=[0, 0, 0, 0],
           rpn_proposal_stds=[1, 1, 1, 1], rpn_proposal_num_pre_nms_train=12000, rpn_proposal_num_post_nms_train=2000,
           rpn_proposal_num_pre_nms_test=6000, rpn_proposal_num_post_nms_test=300, rpn_proposal_nms_iou_threshold=0.7,
           rpn_sigma=3.0, rpn_training_pos_iou_threshold=0.7, rpn_training_neg_iou_threshold=0.3,
           rpn_training_total_num_samples=256, rpn_training_max_pos_samples=128, roi_proposal_means=[0, 0, 0, 0],
           roi_proposal_stds=[0.1, 0.1, 0.2, 0.2], roi_pool_size=7, roi_sigma=1.0, roi_training_pos_iou_threshold=0.5,
           roi_training_neg_iou_threshold=0.0, roi_training_total_num_samples=128, roi_training_max_pos_samples=32,
           prediction_max_objects_per_image=50, prediction_max_objects_per_class=50, prediction_nms_iou_threshold=0.3,
           prediction_score_threshold=0.0)


def _nhwc(conv, x):
    return conv(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).contiguous()


class _RoiHead(torch.nn.Module):
    def __init__(self, c, num_classes):
        super().__init__()
        self.cls = torch.nn.Linear(c, num_classes)
        self.reg = torch.nn.Linear(c, 4 * num_classes)
        torch.nn.init.normal_(self.cls.weight, std=0.5); torch.nn.init.normal_(self.reg.weight, std=0.05)

    def forward(self, f, training=None):
        v = f.mean(dim=(1, 2))
        return self.cls(v), self.reg(v)


class _RpnHead(torch.nn.Module):
    def __init__(self, c, a, pairs):
        super().__init__()
        self.score = torch.nn.Conv2d(c, 2 * a, 1)
        self.box = torch.nn.Conv2d(c, 4 * a, 1)
        torch.nn.init.normal_(self.score.weight, std=0.3); torch.nn.init.normal_(self.box.weight, std=0.03)
        self.a, self.pairs = a, pairs

    def forward(self, f, training=None):
        s = _nhwc(self.score, f)
        b = _nhwc(self.box, f).reshape(-1, 4)
        return (s.reshape(-1, 2) if self.pairs else s.reshape(-1, 2 * self.a)), b


def _faster_rcnn(c=64):
    from tf_eager_object_detection_b200.models import BaseFasterRcnn

    class Net(BaseFasterRcnn):
        def _get_extractor(self):
            conv = torch.nn.Conv2d(3, c, 16, stride=16).cuda()
            self.extractor_conv = conv
            return lambda img, training=None: _nhwc(conv, img)

        def _get_roi_head(self):
            self.roi_head_mod = _RoiHead(c, 21).cuda()
            return self.roi_head_mod

        def _get_rpn_head(self):
            self.rpn_head_mod = _RpnHead(c, 9, pairs=False).cuda()
            return self.rpn_head_mod
    torch.manual_seed(3)
    cfg = dict(CFG, scales=[8, 16, 32], extractor_stride=16, roi_pooling_max_pooling_flag=False)
    return Net(**cfg)


def test_faster_rcnn_eval_and_training_forward():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from tf_eager_object_detection_b200 import ops, _lib
    from tf_eager_object_detection_b200.region_proposal import RegionProposal
    net = _faster_rcnn()
    g = torch.Generator(device='cuda'); g.manual_seed(5)
    image = torch.randn((1, 320, 480, 3), device='cuda', generator=g)
    with torch.no_grad():
        boxes, labels, scores = net(image, training=False)
        sm, tx, rois, counts = net.im_detect_batched(image)
        s1, t1, r1 = net.im_detect(image, 1.6)
    k = int(counts[0])
    assert k == 300 and boxes is not None and boxes.shape[0] == labels.shape[0] == scores.shape[0] <= 50
    assert torch.equal(s1, sm[0, :k]) and torch.equal(t1, tx[0, :k]) and torch.equal(r1, rois[0, :k] / 1.6)
    # proposals: the fused front equals the mirror class fed with the foreground probabilities the device computes
    with torch.no_grad():
        feat = net._extractor(image)
        rpn_score, rpn_bbox = net._rpn_head(feat)
        fg = ops.rpn_scores(rpn_score[None], _lib.RPN_CAFFE, 9)[0]
        anchors = net._anchor_generator(net._anchor_base, 16, 20, 30)
        ref_rois = RegionProposal(num_post_nms_test=300)((rpn_bbox, anchors, fg, [320, 480]), training=False)
    assert torch.equal(ref_rois, rois[0, :k])
    _, idx = orc.region_proposal(rpn_bbox.cpu().numpy(), anchors.cpu().numpy(), fg.cpu().numpy(), (320, 480), 300)
    dec = orc.bboxes_clip_filter(orc.decode_bbox(anchors.cpu().numpy(), rpn_bbox.cpu().numpy()), 0, 320, 480)[0]
    np.testing.assert_allclose(rois[0, :k].cpu().numpy(), dec[idx], rtol=1e-5, atol=1e-3)
    # RoI features the head saw, and the final detections, against the stage oracles on the same intermediate tensors
    feats = ops.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, rois, stride=16.0, roi_counts=counts)
    assert np.array_equal(feats.cpu().numpy(), orc.roi_pool_c4(feat.cpu().numpy(), rois[0].cpu().numpy(), 16, 7, False))
    pb, pc, ps = orc.post_ops_prediction(sm[0].cpu().numpy(), tx[0].cpu().numpy(), rois[0].cpu().numpy(), (320, 480),
                                         stds=(0.1, 0.1, 0.2, 0.2), max_num_per_class=50, max_num_per_image=50,
                                         nms_iou_threshold=0.3, score_threshold=0.0, extractor_stride=16)
    assert np.array_equal(labels.cpu().numpy(), pc) and np.array_equal(scores.cpu().numpy(), ps)
    np.testing.assert_allclose(boxes.cpu().numpy(), pb, rtol=1e-5, atol=1e-3)
    # training forward: four finite losses, gradients into the extractor, the RPN head and the RoI head
    rng = np.random.default_rng(9)
    gt = torch.as_tensor(np.float32([[30, 40, 200, 260], [250, 60, 460, 300], [100, 100, 180, 170]])).cuda()
    gl = torch.tensor([3, 7, 12], dtype=torch.int32, device='cuda')
    losses = net((image, gt, gl), training=True, seed=11)
    assert len(losses) == 4 and all(torch.isfinite(l) for l in losses)
    sum(losses).backward()
    for mod in (net.extractor_conv, net.rpn_head_mod.score, net.rpn_head_mod.box, net.roi_head_mod.cls, net.roi_head_mod.reg):
        assert mod.weight.grad is not None and torch.isfinite(mod.weight.grad).all() and float(mod.weight.grad.abs().sum()) > 0
    sampled = net.predict_roi(image, gt, gl, seed=11)
    assert sampled[0].shape == (128, 4) and sampled[1].shape == (128,)
    del rng


def test_fpn_eval_and_training_forward():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from tf_eager_object_detection_b200 import ops
    from tf_eager_object_detection_b200.models import BaseFPN
    c = 64

    class Net(BaseFPN):
        def _get_extractor(self):
            self.convs = torch.nn.ModuleList([torch.nn.Conv2d(3, c, s, stride=s) for s in (4, 8, 16, 32)]).cuda()
            return lambda img, training=None: [_nhwc(cv, img) for cv in self.convs]

        def _get_neck(self):
            # P2..P5 = C2..C5, P6 = stride-2 subsampling of P5 (resnet_fpn.py:389-404 does a 1x1 max pool with stride 2)
            return lambda c_list, training=None: list(c_list) + [c_list[-1][:, ::2, ::2].contiguous()]

        def _get_roi_head(self):
            self.roi_head_mod = _RoiHead(c, 21).cuda()
            return self.roi_head_mod

        def _get_rpn_head(self):
            self.rpn_head_mod = _RpnHead(c, 3, pairs=True).cuda()
            return self.rpn_head_mod
    torch.manual_seed(4)
    net = Net(**dict(CFG, scales=[1.0], rpn_proposal_num_post_nms_test=200))
    g = torch.Generator(device='cuda'); g.manual_seed(6)
    image = torch.randn((1, 256, 384, 3), device='cuda', generator=g)
    with torch.no_grad():
        boxes, labels, scores = net(image, training=False)
        sm, tx, rois = net.im_detect(image, 2.0)
        image_shape, p_list, anchors, fs, fb, prop, counts = net._front(image, False)
    k = int(counts[0])
    assert anchors.shape[0] == fs.shape[0] == 3 * (64 * 96 + 32 * 48 + 16 * 24 + 8 * 12 + 4 * 6) and k == 200
    # level-major order of what the head saw, and its features, against the oracle
    plain = prop[0, :k].cpu().numpy()
    lv, _, order = orc.assign_levels(plain)
    assert np.array_equal(rois.cpu().numpy() * 2.0, plain[order])
    feats, _, dev_order, _ = ops.fpn_roi_features(p_list[:4], prop[0, :k], image_shape)
    assert np.array_equal(dev_order.cpu().numpy(), order)
    pos = 0
    for l in range(4):
        sel = order[lv[order] == l + 2]
        if sel.size:
            want = orc.roi_pool_fpn(p_list[l].cpu().numpy(), plain[sel], image_shape, 7)
            assert np.array_equal(feats[pos:pos + sel.size].cpu().numpy(), want)
        pos += sel.size
    pb, pc, ps = orc.post_ops_prediction(sm.cpu().numpy(), tx.cpu().numpy(), plain[order], (256, 384), stds=(0.1, 0.1, 0.2, 0.2),
                                         max_num_per_class=50, max_num_per_image=50, nms_iou_threshold=0.3,
                                         score_threshold=0.0, extractor_stride=16)
    assert np.array_equal(labels.cpu().numpy(), pc) and np.array_equal(scores.cpu().numpy(), ps)
    np.testing.assert_allclose(boxes.cpu().numpy(), pb, rtol=1e-5, atol=1e-3)
    gt = torch.as_tensor(np.float32([[20, 30, 120, 200], [200, 40, 370, 240]])).cuda()
    gl = torch.tensor([5, 9], dtype=torch.int32, device='cuda')
    losses = net((image, gt, gl), training=True, seed=2)
    assert len(losses) == 4 and all(torch.isfinite(l) for l in losses)
    sum(losses).backward()
    assert all(cv.weight.grad is not None and torch.isfinite(cv.weight.grad).all() for cv in net.convs)
    assert float(net.roi_head_mod.cls.weight.grad.abs().sum()) > 0 and float(net.rpn_head_mod.box.weight.grad.abs().sum()) > 0
    fg_anchors = net.predict_rpns([256, 384], gt, seed=2, device=image.device)
    assert fg_anchors.shape[1] == 4 and fg_anchors.shape[0] > 0
    assert net.predict_rois(image, gt, gl, seed=2).shape == (128, 4)


def test_get_prediction_files_reproduces_the_reference_files(golden, tmp_path):
    """evaluation.get_prediction_files — model.im_detect -> per-image filtering on the device -> VOC result files — fed
    with the synthetic roi-head outputs the golden was generated from, against the text of the files the reference's own
    get_prediction_files wrote (scores and ids identical, corners printed to 0.1 px)."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from oracle.voc_fixture import eval_loop_inputs
    from tf_eager_object_detection_b200 import evaluation as ev
    imgs = eval_loop_inputs()
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()  # noqa: E731

    class Model:
        def im_detect_batched(self, i):
            im = imgs[i]
            return (cu(im['scores'])[None], cu(im['deltas'].reshape(300, -1))[None], cu(im['rois'])[None],
                    torch.tensor([300], dtype=torch.int32, device='cuda'))
    names = ['%06d' % (i + 1) for i in range(len(imgs))]
    dataset = [(i, im['scale'], im['raw_h'], im['raw_w']) for i, im in enumerate(imgs)]
    for tag, max_img in (('eval_voc', 50), ('eval_voc_nocut', 0)):
        d = tmp_path / tag
        d.mkdir()
        rec, cnt = ev.get_prediction_files(Model(), dataset, names, str(d / '{:s}.txt'), score_threshold=0.05, iou_threshold=0.3,
                                           max_objects_per_class=50, max_objects_per_image=max_img, min_size=10)
        got = ''.join(open(d / ('%s.txt' % c)).read() for c in ev.PASCAL_CLASSES[1:]).splitlines()
        ref_lines = bytes(golden[tag + '_files']).decode().splitlines()
        assert len(got) == len(ref_lines) and (max_img == 0 or cnt.tolist() == [50, 50, 76])
        for x, y in zip(got, ref_lines):
            gx, gy = x.split(), y.split()
            assert gx[:2] == gy[:2] and all(abs(float(p) - float(q)) <= 0.1001 for p, q in zip(gx[2:], gy[2:])), (x, y)
