#!/usr/bin/env python
"""bench.py — proposal + NMS + RoI pooling throughput (BASELINE.json metric) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one pass of the hot path over one batch of synthetic images (workload cfg2 of BASELINE.json: ResNet-50 C4,
600x1000, 21 546 anchors, pre-NMS 6000 -> post-NMS 300, crop 7x7x1024, batch 8 per GPU).  Prints ONE JSON line (rank 0).
Device-resident `value`, host-buffer `e2e`, `roofline` of the dominant kernel (RoI pooling), `cpu_baseline` (oracle C
twin on the host cores), clocks.  `--impl reference` times the CPU restatement of the reference path instead.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tf_eager_object_detection_b200 import synthetic as syn  # noqa: E402

WORKLOAD = dict(name='cfg2: ResNet-50 C4 600x1000, 21546 anchors, pre-NMS 6000 -> post-NMS 300, crop 7x7x1024, batch 8/GPU',
                cfg=2, batch=8, image_hw=(600, 1000), stride=16, channels=1024, pre_nms=6000, post_nms=300, pool=7,
                iou_thr=0.7)
METRIC = 'proposal+NMS+RoIAlign images/s'
UNIT = 'images/s'
FALLBACK_HBM_GBS = 6650.0
NCU_TRAFFIC_BYTES_PER_LAUNCH = 516283136   # profiles/r1_ncu_full_raw.csv: 92.8 MB read + 423.5 MB written by roi_band_kernel


def algorithmic_bytes(w, n, fh, fw):
    """B_prop = 36N + 20K ; B_roi = 4*C*h*w + 16R + 4*R*P^2*C (per image, K = R = post_nms)."""
    K = R = w['post_nms']
    b_prop = 36 * n + 20 * K
    b_roi = 4 * w['channels'] * fh * fw + 16 * R + 4 * R * w['pool'] ** 2 * w['channels']
    return b_prop, b_roi


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured'
        except Exception:
            pass
    return FALLBACK_HBM_GBS, 'fallback'


def make_batch(w, first_index, with_features=True):
    imgs = [syn.c4_image(w['cfg'], first_index + i, w['image_hw'], w['stride'], w['channels'], with_features)
            for i in range(w['batch'])]
    out = dict(anchors=imgs[0]['anchors'], deltas=np.stack([im['deltas'] for im in imgs]),
               scores=np.stack([im['scores'] for im in imgs]), feat_hw=imgs[0]['feat_hw'])
    if with_features:
        out['feat'] = np.stack([im['feat'] for im in imgs])
    return out


# ----------------------------------------------------------------------------------------------------- CPU arm
def load_cpu_oracle():
    """The oracle's C twin (oracle/c): allowed here only for the cpu_baseline / --impl reference legs."""
    so = os.path.join(ROOT, 'oracle', 'c', 'libboxpath_ref.so')
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle', 'c')])
    lib = ctypes.CDLL(so)
    lib.orc_c4_proposal_roi.restype = ctypes.c_int
    lib.orc_c4_proposal_roi.argtypes = ([ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 2 +
                                        [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                              ctypes.c_int] + [ctypes.c_void_p] * 4)
    lib.orc_max_threads.restype = ctypes.c_int
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    lib.orc_set_threads(ctypes.c_int(ncpu))      # torchrun sets OMP_NUM_THREADS=1; the CPU arm uses every host core
    return lib


def cpu_step_fn(w, batch_np, n_images):
    lib = load_cpu_oracle()
    n = batch_np['anchors'].shape[0]
    fh, fw = batch_np['feat_hw']
    post, P, c = w['post_nms'], w['pool'], w['channels']
    b = n_images
    means, stds = np.zeros(4, np.float32), np.ones(4, np.float32)
    o_rois = np.zeros((b, post, 4), np.float32); o_idx = np.zeros((b, post), np.int32)
    o_cnt = np.zeros(b, np.int32); o_feat = np.zeros((b * post, P, P, c), np.float32)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    deltas = np.ascontiguousarray(batch_np['deltas'][:b]); scores = np.ascontiguousarray(batch_np['scores'][:b])
    feat = np.ascontiguousarray(batch_np['feat'][:b])

    def step():
        rc = lib.orc_c4_proposal_roi(ptr(batch_np['anchors']), ptr(deltas), ptr(scores), ptr(feat), b, n, fh, fw, c,
                                     ptr(means), ptr(stds), w['image_hw'][0], w['image_hw'][1], w['pre_nms'], post,
                                     w['iou_thr'], float(w['stride']), P, 0, ptr(o_rois), ptr(o_idx), ptr(o_cnt), ptr(o_feat))
        assert rc == 0
    return step, lib.orc_max_threads(), (o_rois, o_idx, o_cnt, o_feat)


def run_cpu_baseline(w, batch_np, budget_s=12.0):
    step, cores, _ = cpu_step_fn(w, batch_np, w['batch'])
    step()  # warm
    t0 = time.perf_counter(); reps = 0
    while True:
        step(); reps += 1
        el = time.perf_counter() - t0
        if el > budget_s or reps >= 50:
            break
    out = dict(value=round(reps * w['batch'] / el, 2), unit=UNIT, cores=cores, kind='port',
               sample='%d passes over one %d-image batch of the same workload (%.1f s), oracle C twin: '
                      'single-thread NMS per image, crop_and_resize sharded over boxes with OpenMP'
                      % (reps, w['batch'], el))
    # second CPU form: the numpy restatement with the reference's eager-style temporaries, one image
    try:
        from oracle import boxpath_oracle as orc
        t0 = time.perf_counter()
        rois, _ = orc.region_proposal(batch_np['deltas'][0], batch_np['anchors'], batch_np['scores'][0], w['image_hw'],
                                      w['post_nms'], w['iou_thr'], pre_nms_top_k=w['pre_nms'])
        orc.roi_pool_c4(batch_np['feat'][:1], rois, w['stride'], w['pool'], False)
        out['numpy_restatement'] = dict(value=round(1.0 / (time.perf_counter() - t0), 2), unit=UNIT, cores=1,
                                        sample='one image of the workload through oracle/boxpath_oracle.py')
    except Exception as e:                                   # the oracle is test infrastructure: never fatal here
        out['numpy_restatement'] = dict(error=str(e)[:80])
    return out


def run_reference_arm(args):
    """`--impl reference`: the CPU restatement of the reference's path (the real TF path is not installable: SURVEY §8c),
    all host threads, same workload/metric.  Rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    w = WORKLOAD
    batch_np = make_batch(w, 0)
    n_img = w['batch']
    step, cores, _ = cpu_step_fn(w, batch_np, n_img)
    t0 = time.perf_counter(); step(); one = time.perf_counter() - t0
    if one * (args.steps + args.warmup) > 240.0:          # keep the whole run within a few minutes
        n_img = max(1, int(n_img * 240.0 / (one * (args.steps + args.warmup))))
        step, cores, _ = cpu_step_fn(w, batch_np, n_img)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    val = args.steps * n_img / el
    sample = '%d steps x %d images of the workload batch per step' % (args.steps, n_img)
    line = dict(metric=METRIC, value=round(val, 2), unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=round(1e3 * el / args.steps, 3), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=w['name'], images_per_step=n_img,
                            note='CPU restatement of the reference TF ops (oracle C twin); TensorFlow itself is not installable here'),
                cpu_baseline=dict(value=round(val, 2), unit=UNIT, cores=cores, kind='port', sample=sample),
                e2e=dict(value=round(val, 2), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.stop_flag = index, period, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake'}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=[], samples=0)
        return dict(sm_mhz=int(statistics.median(self.samples)), sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU, so that the pinned host buffers
    of the e2e leg are allocated on the GPU's own NUMA node (first touch) and the copies do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from tf_eager_object_detection_b200 import _lib, ops

    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    w = WORKLOAD
    B, post, P, C = w['batch'], w['post_nms'], w['pool'], w['channels']
    lib = _lib.load()

    # ---- inputs: NBUF distinct batches resident in HBM, rotated so no step re-reads the previous step's inputs from L2
    NBUF = max(4, args.streams)          # one distinct input batch per stream: concurrent steps never share inputs in L2
    host_batches = [make_batch(w, rank * 10000 + k * B) for k in range(NBUF)]
    n = host_batches[0]['anchors'].shape[0]
    fh, fw = host_batches[0]['feat_hw']
    anchors = torch.as_tensor(host_batches[0]['anchors']).to(dev)
    d_in = [dict(deltas=torch.as_tensor(hb['deltas']).to(dev), scores=torch.as_tensor(hb['scores']).to(dev),
                 feat=torch.as_tensor(hb['feat']).to(dev)) for hb in host_batches]
    NSTREAM = max(1, args.streams)
    # per-image detection records of a step = kept boxes [B,post,4] fp32 + counts [B] int32.  N > 1: the records of G
    # consecutive steps land in one bucket (a ring of NBK buckets) and are all-gathered together — same bytes, 1/G of the
    # NCCL launches (the all-gathers of one communicator serialise, so their launch latency is what N > 1 adds per step)
    G = max(1, args.gather_every) if world > 1 else 1
    NBK = max(2, (2 * NSTREAM + G - 1) // G + 1)         # a bucket is reused only after >= 2 * NSTREAM later steps
    bk_boxes = [torch.empty((G * B, post, 4), device=dev) for _ in range(NBK)]
    bk_counts = [torch.empty((G * B,), dtype=torch.int32, device=dev) for _ in range(NBK)]
    idx_bufs = [torch.empty((B, post), dtype=torch.int32, device=dev) for _ in range(NSTREAM)]
    feat_bufs = [torch.empty((B * post, P, P, C), device=dev) for _ in range(NSTREAM)]

    def outs_of(step):
        b, slot, s = (step // G) % NBK, step % G, step % NSTREAM
        return (bk_boxes[b][slot * B:(slot + 1) * B], idx_bufs[s], bk_counts[b][slot * B:(slot + 1) * B], feat_bufs[s])
    params = ops.proposal_params(w['image_hw'], post, w['iou_thr'], pre_nms_top_k=w['pre_nms'])
    streams = [torch.cuda.Stream(dev) for _ in range(NSTREAM)]
    handles = []
    for _ in range(NSTREAM):
        hh = ctypes.c_void_p()
        _lib.check(lib.bx_create(local, ctypes.byref(hh)))
        handles.append(hh)

    # N > 1: the only collective of the path — all-gather of the per-image detection records (kept boxes + counts)
    # bx_allgather_detections on torch's own ncclComm_t: boxes + counts in one fused NCCL group on the step's stream
    gathered = [(torch.empty((world * G * B, post, 4), device=dev), torch.empty((world * G * B,), dtype=torch.int32, device=dev))
                for _ in range(NBK)] if world > 1 else None
    comm = None
    if world > 1:
        from tf_eager_object_detection_b200.distributed import nccl_comm_ptr
        comm = nccl_comm_ptr()
        comm = ctypes.c_void_p(comm) if comm is not None else None   # None: torch build without _comm_ptr -> torch collectives
    step_done = {}                                       # step -> event recorded behind its kernels (N > 1)
    bucket_free = [None] * NBK                           # event behind the last all-gather that read the bucket

    def gather_bucket(last_step, n_in_bucket):
        """All-gather the bucket whose last filled slot belongs to `last_step`, on that step's stream."""
        b, s = (last_step // G) % NBK, last_step % NSTREAM
        for k in range(last_step - n_in_bucket + 1, last_step):
            ev_k = step_done.pop(k, None)
            if ev_k is not None:
                streams[s].wait_event(ev_k)
        step_done.pop(last_step, None)
        rows = n_in_bucket * B
        if comm is not None:
            _lib.check(lib.bx_allgather_detections(handles[s], comm, bk_boxes[b].data_ptr(), bk_counts[b].data_ptr(), rows,
                                                   post, 4, world, gathered[b][0].data_ptr(), gathered[b][1].data_ptr(),
                                                   ctypes.c_void_p(streams[s].cuda_stream)))
        else:
            with torch.cuda.stream(streams[s]):
                dist.all_gather_into_tensor(gathered[b][0][:world * rows], bk_boxes[b][:rows])
                dist.all_gather_into_tensor(gathered[b][1][:world * rows], bk_counts[b][:rows])
        ev = torch.cuda.Event(); ev.record(streams[s]); bucket_free[b] = ev

    def launch(step, last_of_run=False):
        s = step % NSTREAM
        din, o = d_in[step % NBUF], outs_of(step)
        if world > 1 and step % G == 0 and bucket_free[(step // G) % NBK] is not None:
            for s2 in range(NSTREAM):                    # nobody refills the bucket before its all-gather has read it
                streams[s2].wait_event(bucket_free[(step // G) % NBK])
            bucket_free[(step // G) % NBK] = None
        _lib.check(lib.bx_c4_proposal_roi(handles[s], anchors.data_ptr(), din['deltas'].data_ptr(),
                                          din['scores'].data_ptr(), din['feat'].data_ptr(), B, n, fh, fw, C,
                                          ctypes.byref(params), float(w['stride']), P, _lib.POOL_NONE, o[0].data_ptr(),
                                          o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(),
                                          ctypes.c_void_p(streams[s].cuda_stream)))
        if world > 1:
            if step % G == G - 1 or last_of_run:
                gather_bucket(step, step % G + 1)
            else:
                ev = torch.cuda.Event(); ev.record(streams[s]); step_done[step] = ev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    main = torch.cuda.current_stream(dev)

    def timed(nsteps, first):
        """K steps round-robin over the streams; device time from a start event (main stream, all streams wait on it)
        to an end event recorded after every stream has been joined back into the main stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for k in range(nsteps):
            launch(first + k, last_of_run=(k == nsteps - 1))
        for s in streams:
            ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    def timed_single_stream(nsteps, first):
        """The same steps on ONE stream: per-launch CUDA-event durations of the RoI kernels without co-running launches."""
        saved = NSTREAM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main)
        streams[0].wait_event(e0)
        for k in range(nsteps):
            launch((first + k) * saved * G, last_of_run=True)   # multiple of NSTREAM * G -> stream 0, slot 0, own gather
        ev = torch.cuda.Event(); ev.record(streams[0]); main.wait_event(ev)
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    timed(NSTREAM, 0)                                   # priming, not a warm-up step: every stream's handle allocates its workspace
    timed(max(3, args.warmup), 0)                       # warm-up (>= 3 steps)
    launches_w = [int(lib.bx_launch_count(hh)) for hh in handles]
    sampler = ClockSampler(local); sampler.start()
    ms = timed(args.steps, max(3, args.warmup))         # ---- the timed region: K steps pipelined over the streams
    launches_t = [int(lib.bx_launch_count(hh)) for hh in handles]
    gpu_launches = sum(launches_t) - sum(launches_w)
    # ---- roofline leg: the dominant kernel pair (plan + band) bracketed by CUDA events on its launch stream, K steps
    #      issued on one stream so that every launch is timed alone (compared with the burst HBM peak)
    _lib.check(lib.bx_profile_roi(handles[0], 1, args.steps + 2))
    ms_single = timed_single_stream(args.steps, 1)
    sampler.stop_flag = True; sampler.join()
    buf = (ctypes.c_float * (args.steps + 4))(); cnt = ctypes.c_int()
    _lib.check(lib.bx_profile_read(handles[0], buf, args.steps + 4, ctypes.byref(cnt)))
    roi_ms = list(buf[:cnt.value])
    _lib.check(lib.bx_profile_roi(handles[0], 0, 0))
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * args.steps * B / (ms_max * 1e-3)

    # ---- sanity inside the bench: every image filled its quota (otherwise the work measured is not the workload)
    for c_ in bk_counts if world > 1 else [outs_of(k_)[2] for k_ in range(NSTREAM)]:
        assert bool((c_ == post).all()), 'a step kept fewer than post_nms proposals'
    if world > 1:                                       # a gathered bucket holds this rank's records at its slot
        torch.cuda.synchronize()
        gather_bucket(G - 1, G)                         # bucket 0, all G slots
        torch.cuda.synchronize()
        assert torch.equal(gathered[0][0][rank * G * B:(rank + 1) * G * B], bk_boxes[0]), 'all-gather mismatch (boxes)'
        assert torch.equal(gathered[0][1][rank * G * B:(rank + 1) * G * B], bk_counts[0]), 'all-gather mismatch (counts)'

    # ---- e2e: the public Python API with HOST (pinned) buffers; H2D inputs + D2H outputs inside the timed region
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    hb = host_batches[0]
    # pinned staging buffers are allocated (first touch) while the process is bound to the CPUs local to its GPU
    old_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    NE = 2   # two streams, two sets of pinned buffers: the D2H of step i overlaps the H2D + kernels of step i+1
    h_in = [dict(deltas=pin(b_['deltas']), scores=pin(b_['scores']), feat=pin(b_['feat'])) for b_ in host_batches[:NE]]
    h_outs = [(torch.empty((B, post, 4)).pin_memory(), torch.empty((B, post), dtype=torch.int32).pin_memory(),
               torch.empty((B,), dtype=torch.int32).pin_memory(), torch.empty((B * post, P, P, C)).pin_memory())
              for _ in range(NE)]
    if numa is not None:
        os.sched_setaffinity(0, old_affinity)
    h2d = sum(t_.numel() * t_.element_size() for t_ in h_in[0].values())
    d2h = sum(t_.numel() * t_.element_size() for t_ in h_outs[0])
    e_streams = [torch.cuda.Stream(dev) for _ in range(NE)]

    def e2e_step(k):
        hi = h_in[k % NE]
        with torch.cuda.stream(e_streams[k % NE]):
            ops.c4_proposal_roi_host(anchors, hi['deltas'], hi['scores'], hi['feat'], w['image_hw'], post, h_outs[k % NE],
                                     stride=float(w['stride']), pool_size=P, pre_nms_top_k=w['pre_nms'],
                                     iou_threshold=w['iou_thr'])
    for k in range(NE):
        e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s_ in e_streams:
        s_.wait_event(e0)
    for k in range(e2e_steps):
        e2e_step(k)
    for s_ in e_streams:
        ev = torch.cuda.Event(); ev.record(s_); main.wait_event(ev)
    e1.record(main)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    wall_ms = (time.perf_counter() - t0) * 1e3
    assert all(int(o[2].min()) == post for o in h_outs)
    t = torch.tensor([max(e2e_ms, 0.0)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * B / (float(t.item()) * 1e-3)

    if rank == 0:
        b_prop, b_roi = algorithmic_bytes(w, n, fh, fw)
        peak, which = hbm_peak()
        roi_avg_ms = sum(roi_ms) / max(1, len(roi_ms))
        roi_bytes = B * b_roi                           # one launch pools the whole batch
        achieved = roi_bytes / (roi_avg_ms * 1e-3) / 1e9 if roi_avg_ms > 0 else 0.0
        step_bytes = B * (b_prop + b_roi)
        # the CPU arm is timed on rank 0 at N=1 only (all host cores belong to the one process there)
        cpu = run_cpu_baseline(w, host_batches[0]) if world == 1 else dict(
            value=None, unit=UNIT, cores=0, kind='port', sample='not run at N > 1: measured on rank 0 at N = 1 only')
        line = dict(metric=METRIC, value=round(value, 1), unit=UNIT, n_gpus=world, steps=args.steps,
                    warmup=max(3, args.warmup), ms_per_step=round(ms_max / args.steps, 5), higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                    config=dict(workload=w['name'], images_per_step_per_gpu=B, streams=NSTREAM,
                                parallelism='images sharded over GPUs, no data-path collective; NCCL all-gather of the '
                                            'per-image detection records, %d steps per bucket (bx_allgather_detections on the process group\'s '
                                            'ncclComm_t)' % G if world > 1 else 'single GPU',
                                l2='working set %.0f MB/step (inputs rotate over %d batches, outputs %.0f MB) > 126 MB L2'
                                   % (step_bytes / 1e6, NBUF, B * post * P * P * C * 4 / 1e6),
                                algorithmic_bytes_per_image=b_prop + b_roi,
                                composite_hbm_frac=round(step_bytes * args.steps / (ms_max * 1e-3) / 1e9 / peak, 4)),
                    roofline=dict(bound='hbm', kernel='roi_plan_kernel + roi_band_kernel (RoI pooling of one batch)',
                                  achieved=round(achieved, 1), peak=peak, peak_source=which, unit='GB/s',
                                  frac=round(achieved / peak, 4), traffic=NCU_TRAFFIC_BYTES_PER_LAUNCH,
                                  traffic_source='profiles/ (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum)',
                                  launches_timed=len(roi_ms), avg_launch_ms=round(roi_avg_ms, 5),
                                  algorithmic_bytes_per_launch=roi_bytes,
                                  timed='CUDA events around every launch, same K steps issued on one stream '
                                        '(kernel timed alone); single-stream step %.5f ms' % (ms_single / args.steps)),
                    cpu_baseline=cpu,
                    e2e=dict(value=round(e2e_value, 1), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             steps=e2e_steps, ms_per_step=round(e2e_ms / e2e_steps, 3), wall_ms_per_step=round(wall_ms / e2e_steps, 3),
                             api='ops.c4_proposal_roi_host -> bx_c4_proposal_roi_host (pinned host buffers, 2 streams)',
                             numa_bound_cpus=numa),
                    gpu_launches=gpu_launches, clocks=sampler.summary())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--streams', type=int, default=8, help='steps are issued round-robin over this many CUDA streams')
    ap.add_argument('--e2e-steps', type=int, default=20)
    ap.add_argument('--gather-every', type=int, default=4,
                    help='N > 1: detection records of this many consecutive steps are all-gathered together')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
