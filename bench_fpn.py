#!/usr/bin/env python
"""bench_fpn.py — the FPN configurations of BASELINE.json (configs[2] "cfg3" and configs[4] "cfg5") with the same JSON
contract as bench.py (which stays on the headline configuration, cfg2).

    python bench_fpn.py [--workload cfg3|cfg5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = proposals over the concatenated P2..P6 anchors (decode, clip, global NMS -> 1000 rois per image) + level
assignment + the FPN RoI extractor (14x14 crop, 2x2 max pool, 7x7x256) over one batch.  cfg3: 600x1000, 150 111 anchors,
batch 16 per GPU (weak scaling).  cfg5: 800x1333, 267 069 anchors, batch 64 sharded over the GPUs (strong scaling), the
per-image detection records all-gathered over NCCL each step.  Through the public Python API (ops.*)."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bench import ClockSampler, METRIC, UNIT, hbm_peak  # noqa: E402
from tf_eager_object_detection_b200 import synthetic as syn  # noqa: E402

WORKLOADS = {
    'cfg3': dict(name='cfg3: ResNet-101 FPN 600x1000, 150111 anchors P2-P6, global NMS -> 1000 rois/image, RoI extractor '
                      '14x14 crop + 2x2 max -> 7x7x256 over P2-P5, batch 16/GPU',
                 cfg=3, image_hw=(600, 1000), batch=16, scaling='weak'),
    'cfg5': dict(name='cfg5: COCO-shape FPN 800x1333, 267069 anchors P2-P6, global NMS -> 1000 rois/image, RoI extractor '
                      '14x14 crop + 2x2 max -> 7x7x256 over P2-P5, batch 64 sharded over the GPUs',
                 cfg=5, image_hw=(800, 1333), batch=64, scaling='strong'),
}
POST, P, C, NLEV = 1000, 7, 256, 4
NSRC = 4            # distinct synthetic images; the batch cycles through them (features are device-side random)


def alg_bytes(hw, n):
    """B_prop = 36N + 20K; B_roi = 4*C*sum(h*w) + 16R + 4*R*P^2*C  (K = R = 1000)."""
    shapes = syn.fpn_feature_shapes(hw)[:NLEV]
    b_prop = 36 * n + 20 * POST
    b_roi = 4 * C * sum(h * w for h, w in shapes) + 16 * POST + 4 * POST * P * P * C
    return b_prop, b_roi


def load_cpu():
    so = os.path.join(ROOT, 'oracle', 'c', 'libboxpath_ref.so')
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle', 'c')])
    lib = ctypes.CDLL(so)
    lib.orc_fpn_proposal_roi.restype = ctypes.c_int
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    lib.orc_set_threads(ctypes.c_int(ncpu))
    lib.orc_max_threads.restype = ctypes.c_int
    return lib


def cpu_step_fn(w, b):
    """The oracle's C twin on b images of the workload (allowed here only as the cpu_baseline / reference arm)."""
    lib = load_cpu()
    hw = w['image_hw']
    ims = [syn.fpn_image(w['cfg'], i % NSRC, hw, with_features=False) for i in range(b)]
    anchors = ims[0]['anchors']; n = anchors.shape[0]
    deltas = np.ascontiguousarray(np.stack([im['deltas'] for im in ims]))
    scores = np.ascontiguousarray(np.stack([im['scores'] for im in ims]))
    shapes = syn.fpn_feature_shapes(hw)[:NLEV]
    rng = np.random.default_rng(1)
    feats = [rng.standard_normal((b, h, wd, C), dtype=np.float32) for h, wd in shapes]
    fptr = (ctypes.c_void_p * NLEV)(*[f.ctypes.data for f in feats])
    fh = (ctypes.c_int * NLEV)(*[s[0] for s in shapes]); fw = (ctypes.c_int * NLEV)(*[s[1] for s in shapes])
    means, stds = np.zeros(4, np.float32), np.ones(4, np.float32)
    o_rois = np.zeros((b, POST, 4), np.float32); o_idx = np.zeros((b, POST), np.int32); o_cnt = np.zeros(b, np.int32)
    o_feat = np.zeros((b * POST, P, P, C), np.float32); o_ord = np.zeros(b * POST, np.int32)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731

    def step(_keep=(feats, ims)):                      # the raw pointers in fptr must outlive cpu_step_fn
        rc = lib.orc_fpn_proposal_roi(vp(anchors), vp(deltas), vp(scores), fptr, fh, fw, b, n, C, vp(means), vp(stds),
                                      hw[0], hw[1], 0, POST, ctypes.c_float(0.7), P, vp(o_rois), vp(o_idx), vp(o_cnt),
                                      vp(o_feat), vp(o_ord))
        assert rc == 0
    return step, lib.orc_max_threads(), (o_rois, o_idx, o_cnt, o_feat, o_ord)


def run_cpu(w, b, budget_s):
    step, cores, _ = cpu_step_fn(w, b)
    step()
    t0 = time.perf_counter(); reps = 0
    while True:
        step(); reps += 1
        el = time.perf_counter() - t0
        if el > budget_s or reps >= 20:
            break
    return reps * b / el, cores, '%d passes over %d images of the workload (%.1f s), oracle C twin: one NMS thread per ' \
                                 'image, crop_and_resize sharded over boxes with OpenMP' % (reps, b, el)


def run_reference_arm(args, w):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    b = 8
    step, cores, _ = cpu_step_fn(w, b)
    for _ in range(min(args.warmup, 2)):
        step()
    steps = min(args.steps, 10)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    el = time.perf_counter() - t0
    val = steps * b / el
    sample = '%d steps x %d images of the workload per step' % (steps, b)
    print(json.dumps(dict(metric=METRIC, value=round(val, 2), unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=min(args.warmup, 2),
                          ms_per_step=round(1e3 * el / steps, 3), higher_is_better=True, scaling=w['scaling'],
                          vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                          config=dict(workload=w['name'], images_per_step=b),
                          cpu_baseline=dict(value=round(val, 2), unit=UNIT, cores=cores, kind='port', sample=sample),
                          e2e=dict(value=round(val, 2), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    from tf_eager_object_detection_b200 import _lib, ops

    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench_fpn.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    hw = w['image_hw']
    B = w['batch'] if w['scaling'] == 'weak' else max(1, w['batch'] // world)
    lib = _lib.load()
    ims = [syn.fpn_image(w['cfg'], rank * 100 + i, hw, with_features=False) for i in range(NSRC)]
    n = ims[0]['anchors'].shape[0]
    anchors = torch.as_tensor(ims[0]['anchors']).to(dev)
    rep = (B + NSRC - 1) // NSRC
    NBUF = 2
    shapes = syn.fpn_feature_shapes(hw)[:NLEV]
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    d_in = []
    for k in range(NBUF):
        order = [(i + k) % NSRC for i in range(NSRC)]
        d_in.append(dict(
            deltas=torch.as_tensor(np.stack([ims[i]['deltas'] for i in order])).to(dev).repeat(rep, 1, 1)[:B].contiguous(),
            scores=torch.as_tensor(np.stack([ims[i]['scores'] for i in order])).to(dev).repeat(rep, 1)[:B].contiguous(),
            feats=[torch.randn((B, h, wd, C), device=dev, generator=g) for h, wd in shapes]))
    bi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(POST)
    NSTREAM = max(1, args.streams)
    streams = [torch.cuda.Stream(dev) for _ in range(NSTREAM)]
    last = [None] * NSTREAM
    rec_bytes = B * POST * 16 + ((B * 4 + 15) // 16) * 16
    gathered = [torch.empty((world * rec_bytes,), dtype=torch.uint8, device=dev) for _ in range(NSTREAM)] if world > 1 else None

    dbg = lambda *a: print(*a, file=sys.stderr, flush=True) if os.environ.get('BX_BENCH_TRACE') else None  # noqa: E731
    dbg('inputs ready')

    def launch(step):
        s = step % NSTREAM
        din = d_in[step % NBUF]
        with torch.cuda.stream(streams[s]):
            rois, idx, cnt = ops.proposals(anchors, din['deltas'], din['scores'], hw, POST)
            out = ops.fpn_roi_features(din['feats'], rois.view(-1, 4), hw, box_ind=bi)
            if world > 1:
                rec = torch.cat([rois.view(-1).view(torch.uint8), cnt.view(torch.uint8),
                                 torch.zeros(rec_bytes - B * POST * 16 - B * 4, dtype=torch.uint8, device=dev)])
                dist.all_gather_into_tensor(gathered[s], rec)
            last[s] = (rois, cnt, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    main = torch.cuda.current_stream(dev)

    def timed(nsteps, first, single=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for k in range(nsteps):
            launch((first + k) * NSTREAM if single else first + k)
        for s in streams:
            ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    handles = [_lib.handle(local, s.cuda_stream) for s in streams]
    W = max(3, args.warmup)
    dbg('handles')
    timed(W, 0)
    dbg('warm')
    l0 = sum(int(lib.bx_launch_count(h)) for h in handles)
    sampler = ClockSampler(local); sampler.start()
    ms = timed(args.steps, W)
    gpu_launches = sum(int(lib.bx_launch_count(h)) for h in handles) - l0
    dbg('timed')
    _lib.check(lib.bx_profile_roi(handles[0], 1, args.steps + 2))
    ms_single = timed(args.steps, 1, single=True)
    sampler.stop_flag = True; sampler.join()
    buf = (ctypes.c_float * (args.steps + 4))(); cnt_ = ctypes.c_int()
    _lib.check(lib.bx_profile_read(handles[0], buf, args.steps + 4, ctypes.byref(cnt_)))
    roi_ms = list(buf[:cnt_.value])
    dbg('roofline leg')
    _lib.check(lib.bx_profile_roi(handles[0], 0, 0))
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * args.steps * B / (ms_max * 1e-3)
    for o in last:
        if o is not None:
            assert bool((o[1] == POST).all()), 'a step kept fewer than 1000 proposals'

    # ---- e2e: pinned host inputs -> device, results (rois, counts, pooled features) back to pinned host memory
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    Be = min(B, 8)                                      # bounded host staging: 8 images per e2e step
    bie = bi[:Be * POST]
    hin = dict(deltas=d_in[0]['deltas'][:Be].cpu().pin_memory(), scores=d_in[0]['scores'][:Be].cpu().pin_memory(),
               feats=[f[:Be].cpu().pin_memory() for f in d_in[0]['feats']])
    hout = (torch.empty((Be, POST, 4)).pin_memory(), torch.empty((Be,), dtype=torch.int32).pin_memory(),
            torch.empty((Be * POST, P, P, C)).pin_memory())
    h2d = hin['deltas'].numel() * 4 + hin['scores'].numel() * 4 + sum(f.numel() * 4 for f in hin['feats'])
    d2h = sum(t_.numel() * t_.element_size() for t_ in hout)

    def e2e_step():
        d = hin['deltas'].to(dev, non_blocking=True); s_ = hin['scores'].to(dev, non_blocking=True)
        fs = [f.to(dev, non_blocking=True) for f in hin['feats']]
        rois, idx, cnt = ops.proposals(anchors, d, s_, hw, POST)
        out, order, lv, counts = ops.fpn_roi_features(fs, rois.view(-1, 4), hw, box_ind=bie)
        hout[0].copy_(rois, non_blocking=True); hout[1].copy_(cnt, non_blocking=True); hout[2].copy_(out, non_blocking=True)
    e2e_step(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(main)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * Be / (float(t.item()) * 1e-3)
    assert int(hout[1].min()) == POST

    if rank == 0:
        b_prop, b_roi = alg_bytes(hw, n)
        peak, which = hbm_peak()
        roi_avg = sum(roi_ms) / max(1, len(roi_ms))
        achieved = B * b_roi / (roi_avg * 1e-3) / 1e9 if roi_avg > 0 else 0.0
        if world == 1:
            cpu_v, cores, sample = run_cpu(w, 8, 15.0)
        else:
            cpu_v, cores, sample = None, 0, 'not run at N > 1: measured on rank 0 at N = 1 only'
        line = dict(metric=METRIC, value=round(value, 1), unit=UNIT, n_gpus=world, steps=args.steps, warmup=W,
                    ms_per_step=round(ms_max / args.steps, 5), higher_is_better=True, scaling=w['scaling'],
                    vs_baseline=None, dtype='f32', data='synthetic',
                    config=dict(workload=w['name'], images_per_step_per_gpu=B, streams=NSTREAM,
                                l2='working set %.0f MB/step > 126 MB L2' % (B * (b_prop + b_roi) / 1e6),
                                algorithmic_bytes_per_image=b_prop + b_roi,
                                composite_hbm_frac=round(B * (b_prop + b_roi) * args.steps / (ms_max * 1e-3) / 1e9 / peak, 4),
                                single_stream_ms_per_step=round(ms_single / args.steps, 5)),
                    roofline=dict(bound='hbm', kernel='roi_pool2_kernel (FPN RoI extractor of one batch)',
                                  achieved=round(achieved, 1), peak=peak, peak_source=which, unit='GB/s',
                                  frac=round(achieved / peak, 4), traffic=None, launches_timed=len(roi_ms),
                                  avg_launch_ms=round(roi_avg, 5), algorithmic_bytes_per_launch=B * b_roi),
                    cpu_baseline=dict(value=None if cpu_v is None else round(cpu_v, 2), unit=UNIT, cores=cores, kind='port', sample=sample),
                    e2e=dict(value=round(e2e_value, 1), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             steps=e2e_steps, images_per_step=Be,
                             api='ops.proposals + ops.fpn_roi_features, pinned host buffers, one stream'),
                    gpu_launches=gpu_launches, clocks=sampler.summary())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='cfg5', choices=sorted(WORKLOADS))
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--streams', type=int, default=2)
    ap.add_argument('--e2e-steps', type=int, default=3)
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference_arm(args, w)
    else:
        run_ours(args, w)


if __name__ == '__main__':
    main()
