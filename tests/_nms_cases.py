"""Adversarial inputs for `inter / union > thr` decided exactly at the threshold (shared by the oracle witness test and
the GPU parity test; test infrastructure)."""
import numpy as np

F = np.float32


def near_threshold_cases():
    """Pairs of integer-cornered boxes whose NMS IoU inter/union is EXACTLY a short decimal, with thresholds at, one ulp
    above and one ulp below the fp32 quotient — the decisions `iou > thr` TF makes with a correctly rounded division and
    a strict compare.  Returned as (boxes [2,4], thr, suppressed) triples; shared with the GPU parity test."""
    cases = []
    geoms = []
    for scale in (1.0, 3.0, 7.0, 64.0, 0.125, 1e-3, 1e-12, 1e-17, 1e-19, 1e-20, 3e-21):          # tiny scales: areas down to the denormal range
        # A = 10x10, B = 10x7 inside-overlapping: inter 70, union 100 -> 0.7 ; inter 50 / union 100 -> 0.5 ...
        for inter_w, b_h in ((7, 10), (5, 10), (3, 10)):
            a = np.float64([0, 0, 10, 10]) * scale
            b = np.float64([10 - inter_w, 0, 20 - inter_w, b_h]) * scale   # union = 100 + 100 - 10*inter_w
            geoms.append((a, b))
        a = np.float64([0, 0, 10, 10]) * scale                             # B inside A: inter 70 / union 100
        b = np.float64([0, 0, 10, 7]) * scale
        geoms.append((a, b))
        a = np.float64([0, 0, 10, 3]) * scale                              # inter 21 / union 30 = 0.7 exactly as a ratio
        b = np.float64([0, 0, 7, 3]) * scale
        geoms.append((a, b))
    rng = np.random.default_rng(2026)
    for _ in range(120):                                                   # random integer geometry: many distinct quotients
        w1, h1, w2, h2 = rng.integers(2, 400, 4)
        dx, dy = rng.integers(0, max(1, min(w1, w2))), rng.integers(0, max(1, min(h1, h2)))
        sc = float(rng.choice([1.0, 0.25, 16.0]))
        geoms.append((np.float64([0, 0, w1, h1]) * sc, np.float64([dx, dy, dx + w2, dy + h2]) * sc))
    for a, b in geoms:
        bx = np.stack([a, b]).astype(F)
        lo = np.minimum(bx[:, :2], bx[:, 2:]); hi = np.maximum(bx[:, :2], bx[:, 2:])
        area = (hi[:, 0] - lo[:, 0]) * (hi[:, 1] - lo[:, 1])
        iw = max(F(0), min(hi[0, 0], hi[1, 0]) - max(lo[0, 0], lo[1, 0]))
        ih = max(F(0), min(hi[0, 1], hi[1, 1]) - max(lo[0, 1], lo[1, 1]))
        inter = F(iw * ih)
        uni = F(F(area[0] + area[1]) - inter)
        if not (area[0] > 0 and area[1] > 0 and uni > 0):
            continue                                                       # flushed to zero: covered by the degenerate tests
        iou = F(inter / uni)
        for thr in (iou, np.nextafter(iou, F(2)), np.nextafter(iou, F(-1)), F(0.7), F(0.5)):
            if 0.0 <= thr <= 1.0:
                cases.append((bx, F(thr), bool(iou > thr)))
    return cases


