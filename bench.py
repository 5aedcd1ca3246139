#!/usr/bin/env python
"""bench.py — proposal + NMS + RoI pooling throughput (BASELINE.json metric) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workloads all|none|cfg3,cfg5,cfg4,f3]

A "step" = one pass of the hot path over one batch of synthetic images.  Headline workload: cfg2 of BASELINE.json
(ResNet-50 C4, 600x1000, 21 546 anchors, pre-NMS 6000 -> post-NMS 300, crop 7x7x1024, batch 8 per GPU).  Prints ONE JSON
line (rank 0): device-resident `value`, host-buffer `e2e`, `roofline` of the dominant kernel (RoI pooling) measured in
the same regime as `value`, `regimes` (pipelined / single stream, structured), `cpu_baseline` (oracle C twin on the host
cores), clocks, and `workloads` — the FPN configurations cfg3 / cfg5, the training-target configuration cfg4 and the
RoI-pooling backward pass (f3) measured in the same run.  `--impl reference` times the CPU restatement of the reference path instead.

How the timed region is issued: the K steps are captured ONCE into a CUDA graph (steps dealt round-robin over S streams,
fork/join inside the graph; at N > 1 the all-gather of the per-image detection records runs on its own branch of the
same graph, on a dedicated communicator, and no compute stream ever waits for it) and the timed region is one replay of
that graph between two CUDA events — no host code between the first and the last kernel.  `--no-graph` issues the same
launches eagerly.
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
NCU_TRAFFIC_BYTES_PER_LAUNCH = 517105152   # profiles/r2_ncu_band_raw.csv: 92.3 MB read + 424.8 MB written by roi_band_kernel


def headline_config(w):
    """The `config` object — identical in the GPU arm and in `--impl reference` (the driver compares them)."""
    return dict(workload=w['name'], images_per_step_per_gpu=w['batch'],
                l2='inputs larger than L2: working set 566 MB/step (82 MB inputs rotating over distinct batches, 482 MB outputs) > 126 MB')


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
                dtype='f32', data='synthetic', impl='reference', config=headline_config(w),
                details=dict(images_per_step=n_img,
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


def gpu_numa_cpus(index):
    """CPUs local to GPU `index`: NVML's affinity mask when it is a proper subset of the visible CPUs, else the cpulist
    of the GPU's PCI device NUMA node from sysfs, else None (single node / not exposed in this container)."""
    allowed = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1} & allowed
        if cpus and cpus != allowed:
            return cpus, 'nvml'
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(':')[0]) == 8:
            bus = bus[4:]
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
        if node >= 0:
            cpus = set()
            for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus |= set(range(int(lo), int(hi or lo) + 1))
            cpus &= allowed
            if cpus and cpus != allowed:
                return cpus, 'sysfs node %d' % node
    except Exception:
        pass
    return None, 'one NUMA domain visible (affinity covers every CPU of the container)'


# ----------------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """Process-wide state of the GPU arm: rank / device / library / process groups."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from tf_eager_object_detection_b200 import _lib
        self.torch, self.dist, self._lib = torch, dist, _lib
        self.rank = int(os.environ.get('RANK', '0')); self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        self.comm = None
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
            # the C-ABI all-gather runs on a communicator nothing else uses (distributed.new_detection_group)
            from tf_eager_object_detection_b200 import distributed as bxd
            self.det_group = bxd.new_detection_group()
            dist.all_reduce(torch.zeros(1, device=self.dev), group=self.det_group)   # forces the communicator to exist
            torch.cuda.synchronize()
            ptr = bxd.nccl_comm_ptr(self.det_group)
            if ptr is None:
                raise SystemExit('bench.py: the process group exposes no ncclComm_t (torch without ProcessGroupNCCL._comm_ptr)')
            self.comm = ctypes.c_void_p(ptr)
        self.lib = _lib.load()
        self.sync_buf = torch.zeros(1, device=self.dev)
        self.main = torch.cuda.current_stream(self.dev)
        self.use_graph = not args.no_graph
        self.use_priority = bool(args.priority)
        self.graph_error = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def new_handle(self):
        hh = ctypes.c_void_p()
        self._lib.check(self.lib.bx_create(self.local, ctypes.byref(hh)))
        return hh

    def max_over_ranks(self, ms):
        """(max over ranks, per-rank list) of a device-timed duration."""
        torch = self.torch
        if self.world == 1:
            return ms, [round(ms, 5)]
        mine = torch.tensor([ms], device=self.dev)
        allv = torch.empty((self.world,), device=self.dev)
        self.dist.all_gather_into_tensor(allv, mine)
        v = [round(float(x), 5) for x in allv.tolist()]
        return max(v), v


class Pipeline:
    """A region of steps dealt round-robin over S CUDA streams (one library handle per stream), with the all-gather of
    the region's detection records on a separate communication stream at N > 1.

    launch(step, pos, stream_index) enqueues the kernels of one step on stream `stream_index`; `pos` is the step's
    position in the region and selects its slot in the record accumulators.  region() forks the streams from `root`,
    issues the steps and joins everything back — the same code eagerly (root = main stream) or under CUDA-graph capture
    (root = capturing stream; events recorded / waited inside the capture become graph edges)."""

    def __init__(self, ctx, n_streams, launch, records=None, gather_every=0, side_streams=False):
        torch = ctx.torch
        self.ctx, self.launch = ctx, launch
        self.streams = [torch.cuda.Stream(ctx.dev) for _ in range(n_streams)]
        self.handles = [ctx.new_handle() for _ in range(n_streams)]
        # optional high-priority side stream per stream (+ its own handle): small latency-critical kernels (the 8-CTA proposal
        # kernel) issued there are scheduled ahead of the queued CTAs of the bandwidth kernels of other steps
        self.hi_streams = [torch.cuda.Stream(ctx.dev, priority=-1) for _ in range(n_streams)] if side_streams else None
        self.hi_handles = [ctx.new_handle() for _ in range(n_streams)] if side_streams else []
        self.records = records          # dict(boxes [cap*B,k,4], counts [cap*B], B, k, out_boxes, out_counts) or None
        self.gather_every = gather_every
        self.comm_stream = torch.cuda.Stream(ctx.dev) if (records is not None and ctx.world > 1) else None
        self.comm_handle = ctx.new_handle() if self.comm_stream is not None else None
        self.graphs = {}

    def _gather(self, first_pos, n_pos, done_events):
        """All-gather of the records of positions [first_pos, first_pos + n_pos) on the communication stream."""
        ctx, r = self.ctx, self.records
        for ev in done_events:
            self.comm_stream.wait_event(ev)
        rows = n_pos * r['B']
        lo = first_pos * r['B']
        # gathered layout per chunk: [world, rows, k, 4] at row offset world * lo
        ctx._lib.check(ctx.lib.bx_allgather_detections(
            self.comm_handle, ctx.comm, r['boxes'][lo:lo + rows].data_ptr(), r['counts'][lo:lo + rows].data_ptr(), rows,
            r['k'], 4, ctx.world, r['out_boxes'][ctx.world * lo:].data_ptr(), r['out_counts'][ctx.world * lo:].data_ptr(),
            ctypes.c_void_p(self.comm_stream.cuda_stream)))

    def region(self, nsteps, first, root, only_stream=None):
        torch = self.ctx.torch
        streams = self.streams if only_stream is None else [self.streams[only_stream]]
        if self.hi_streams is not None:
            streams = streams + (self.hi_streams if only_stream is None else [self.hi_streams[only_stream]])
        ev0 = torch.cuda.Event(); ev0.record(root)
        for s in streams:
            s.wait_event(ev0)
        if self.comm_stream is not None:
            self.comm_stream.wait_event(ev0)
        G = self.gather_every if self.gather_every > 0 else nsteps
        pending = []
        for k in range(nsteps):
            si = (first + k) % len(self.streams) if only_stream is None else only_stream
            self.launch(first + k, k, si)
            if self.comm_stream is not None:
                ev = torch.cuda.Event(); ev.record(self.streams[si]); pending.append(ev)
                if len(pending) == G or k == nsteps - 1:
                    self._gather(k + 1 - len(pending), len(pending), pending)
                    pending = []
        for s in streams + ([self.comm_stream] if self.comm_stream is not None else []):
            ev = torch.cuda.Event(); ev.record(s); root.wait_event(ev)

    def graph(self, nsteps, first, key):
        """The region as one CUDA graph (captured once per key); None when capture is disabled or failed."""
        ctx, torch = self.ctx, self.ctx.torch
        if not ctx.use_graph:
            return None
        if key in self.graphs:
            return self.graphs[key]
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(ctx.dev)
        try:
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=cap, capture_error_mode='thread_local'):
                self.region(nsteps, first, cap)
        except Exception as e:                                   # report and fall back to eager issue for the whole run
            ctx.graph_error = '%s: %s' % (type(e).__name__, str(e)[:200])
            ctx.use_graph = False
            torch.cuda.synchronize()
            return None
        self.graphs[key] = g
        return g

    def timed(self, nsteps, first, key=None, only_stream=None, prime_graph=True, settle=0.0):
        """Device time (ms) of the region: CUDA events on the main stream around one graph replay (or the eager issue)."""
        ctx, torch = self.ctx, self.ctx.torch
        g = self.graph(nsteps, first, key) if (key is not None and only_stream is None) else None
        if g is not None and prime_graph and not getattr(g, '_bx_primed', False):
            g.replay()                                           # untimed: instantiation upload of the executable graph
            g._bx_primed = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        if settle:
            time.sleep(settle)       # every measured region starts from an idle device (same power / clock state: a region
            ctx.barrier()            # measured right behind another one inherits its power-cap state — sw_power_cap at K >= 200)
        if ctx.world > 1:
            # device-side aligned start: a tiny all-reduce queued in front of the start event completes on all ranks
            # within microseconds of each other, so the ranks' timed regions begin together on the DEVICE and the host
            # jitter behind the barrier (tens of microseconds, 1 - 2 % of a 20-step region) is not charged to anybody
            ctx.dist.all_reduce(ctx.sync_buf)
        e0.record(ctx.main)
        if g is not None:
            g.replay()
        else:
            self.region(nsteps, first, ctx.main, only_stream)
        e1.record(ctx.main)
        ctx.barrier()
        return e0.elapsed_time(e1)

    def launches(self):
        return sum(int(self.ctx.lib.bx_launch_count(h)) for h in self.handles + self.hi_handles + ([self.comm_handle] if self.comm_handle else []))

    def close(self):
        self.graphs.clear()
        for h in self.handles + self.hi_handles + ([self.comm_handle] if self.comm_handle else []):
            self.ctx.lib.bx_destroy(h)
        self.handles = []
        self.hi_handles = []


def make_records(ctx, cap_steps, B, k):
    """Accumulators of the per-image detection records of a region (kept boxes [B,k,4] + counts [B] per step) and, at
    N > 1, their all-gathered copies — what the reference accumulates in one process over an evaluation run
    (evaluation/pascal_eval_files_utils.py:73-107)."""
    torch = ctx.torch
    r = dict(B=B, k=k, cap=cap_steps,
             boxes=torch.zeros((cap_steps * B, k, 4), device=ctx.dev),
             counts=torch.zeros((cap_steps * B,), dtype=torch.int32, device=ctx.dev))
    if ctx.world > 1:
        r['out_boxes'] = torch.zeros((ctx.world * cap_steps * B, k, 4), device=ctx.dev)
        r['out_counts'] = torch.zeros((ctx.world * cap_steps * B,), dtype=torch.int32, device=ctx.dev)
    return r


def verify_gather(ctx, r, nsteps, G):
    """Every rank checks the gathered records of EVERY rank: an exact integer checksum (bit patterns summed in int64) of
    each rank's own accumulator is exchanged with torch.distributed and compared with the checksum of that rank's slots
    in the gathered buffer."""
    torch, dist = ctx.torch, ctx.dist
    B, world = r['B'], ctx.world
    rows_total = nsteps * B
    own = torch.stack([r['boxes'][:rows_total].reshape(-1).view(torch.int32).to(torch.int64).sum(),
                       r['counts'][:rows_total].to(torch.int64).sum()])
    allsum = torch.empty((world, 2), dtype=torch.int64, device=ctx.dev)
    dist.all_gather_into_tensor(allsum, own)
    got = torch.zeros((world, 2), dtype=torch.int64, device=ctx.dev)
    G = G if G > 0 else nsteps
    for first in range(0, nsteps, G):
        n_pos = min(G, nsteps - first)
        rows, lo = n_pos * B, first * B
        blk_b = r['out_boxes'][world * lo: world * lo + world * rows].reshape(world, -1).view(torch.int32).to(torch.int64).sum(dim=1)
        blk_c = r['out_counts'][world * lo: world * lo + world * rows].reshape(world, rows).to(torch.int64).sum(dim=1)
        got[:, 0] += blk_b
        got[:, 1] += blk_c
    assert torch.equal(got, allsum), 'all-gather mismatch: gathered records differ from their owners (rank %d sees %s vs %s)' % (
        ctx.rank, got.tolist(), allsum.tolist())
    return True


def pcie_probe(ctx, mb=256, reps=4):
    """Host-link diagnostic for the e2e leg: pinned <-> device copy rates of rank 0 alone and of all ranks at once
    (GB/s, CUDA events; the all-ranks figures are sums over the ranks).  Explains why the host-buffer path scales with
    the host side of the box (PCIe switches / root ports / host memory), not with the GPUs."""
    torch, dist = ctx.torch, ctx.dist
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=ctx.dev)
    s2 = torch.cuda.Stream(ctx.dev)

    def run(h2d, d2h):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(ctx.main)
        s2.wait_event(e0)
        for _ in range(reps):
            if h2d:
                d.copy_(h, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h.copy_(d, non_blocking=True) if not h2d else h2.copy_(d2, non_blocking=True)
        ev = torch.cuda.Event(); ev.record(s2); ctx.main.wait_event(ev)
        e1.record(ctx.main)
        torch.cuda.synchronize()
        return (int(h2d) + int(d2h)) * reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device=ctx.dev)
    out = {}
    for name, (a_, b_) in dict(h2d=(True, False), d2h=(False, True), bidir=(True, True)).items():
        run(a_, b_)
        ctx.barrier()
        if ctx.rank == 0:
            out['rank0_alone_' + name] = round(run(a_, b_), 1)
        ctx.barrier()
        v = run(a_, b_)
        if ctx.world > 1:
            t = torch.tensor([v], device=ctx.dev)
            dist.all_reduce(t)
            v = float(t.item())
        out['all_ranks_' + name] = round(v, 1)
        ctx.barrier()
    return out


def bench_c4(ctx, args):
    """Headline workload (cfg2)."""
    torch, lib, _lib = ctx.torch, ctx.lib, ctx._lib
    from tf_eager_object_detection_b200 import ops
    w = WORKLOAD
    B, post, P, C = w['batch'], w['post_nms'], w['pool'], w['channels']
    K, W = args.steps, max(3, args.warmup)
    NSTREAM = max(1, args.streams)
    NBUF = max(4, NSTREAM)               # one distinct input batch per stream: concurrent steps never share inputs in L2
    host_batches = [make_batch(w, ctx.rank * 10000 + k * B) for k in range(NBUF)]
    n = host_batches[0]['anchors'].shape[0]
    fh, fw = host_batches[0]['feat_hw']
    dev = ctx.dev
    anchors = torch.as_tensor(host_batches[0]['anchors']).to(dev)
    d_in = [dict(deltas=torch.as_tensor(hb['deltas']).to(dev), scores=torch.as_tensor(hb['scores']).to(dev),
                 feat=torch.as_tensor(hb['feat']).to(dev)) for hb in host_batches]
    rec = make_records(ctx, max(K, W, NSTREAM), B, post)
    idx_bufs = [torch.empty((B, post), dtype=torch.int32, device=dev) for _ in range(NSTREAM)]
    feat_bufs = [torch.empty((B * post, P, P, C), device=dev) for _ in range(NSTREAM)]
    params = ops.proposal_params(w['image_hw'], post, w['iou_thr'], pre_nms_top_k=w['pre_nms'])
    G = args.gather_every if args.gather_every > 0 else max(K, W, NSTREAM)   # default: ONE all-gather, at the end of the region

    pipe = None

    def launch(step, pos, si):
        din = d_in[step % NBUF]
        lo = pos * B
        if pipe.hi_streams is not None:
            # proposals on the stream's high-priority side stream, RoI pooling behind an event on the stream itself
            hs, s_ = pipe.hi_streams[si], pipe.streams[si]
            if args.priority == 1:                       # 1: the side stream follows the stream's order; 2: its own chain —
                ev_in = torch.cuda.Event(); ev_in.record(s_); hs.wait_event(ev_in)   # proposals of step i+S overlap the RoI kernels of step i
            _lib.check(lib.bx_proposals(pipe.hi_handles[si], anchors.data_ptr(), din['deltas'].data_ptr(), din['scores'].data_ptr(),
                                        B, n, ctypes.byref(params), rec['boxes'][lo:lo + B].data_ptr(), idx_bufs[si].data_ptr(),
                                        rec['counts'][lo:lo + B].data_ptr(), ctypes.c_void_p(hs.cuda_stream)))
            ev = torch.cuda.Event(); ev.record(hs); s_.wait_event(ev)
            _lib.check(lib.bx_roi_pool(pipe.handles[si], _lib.ROI_STRIDE_NORM, _lib.POOL_NONE, P, din['feat'].data_ptr(), B, fh, fw,
                                       C, rec['boxes'][lo:lo + B].data_ptr(), None, rec['counts'][lo:lo + B].data_ptr(), B * post,
                                       float(w['stride']), w['image_hw'][0], w['image_hw'][1], feat_bufs[si].data_ptr(),
                                       ctypes.c_void_p(s_.cuda_stream)))
            return
        _lib.check(lib.bx_c4_proposal_roi(pipe.handles[si], anchors.data_ptr(), din['deltas'].data_ptr(),
                                          din['scores'].data_ptr(), din['feat'].data_ptr(), B, n, fh, fw, C,
                                          ctypes.byref(params), float(w['stride']), P, _lib.POOL_NONE,
                                          rec['boxes'][lo:lo + B].data_ptr(), idx_bufs[si].data_ptr(),
                                          rec['counts'][lo:lo + B].data_ptr(), feat_bufs[si].data_ptr(),
                                          ctypes.c_void_p(pipe.streams[si].cuda_stream)))

    def launch_roi_only(step, pos, si):
        """The RoI-pooling launch pair of a step (plan + band kernel) on the rois the full steps produced."""
        din = d_in[step % NBUF]
        lo = pos * B
        _lib.check(lib.bx_roi_pool(pipe_roi.handles[si], _lib.ROI_STRIDE_NORM, _lib.POOL_NONE, P, din['feat'].data_ptr(), B, fh,
                                   fw, C, rec['boxes'][lo:lo + B].data_ptr(), None, rec['counts'][lo:lo + B].data_ptr(),
                                   B * post, float(w['stride']), w['image_hw'][0], w['image_hw'][1],
                                   feat_bufs[si].data_ptr(), ctypes.c_void_p(pipe_roi.streams[si].cuda_stream)))

    pipe = Pipeline(ctx, NSTREAM, launch, rec, G, side_streams=ctx.use_priority)
    pipe_roi = Pipeline(ctx, NSTREAM, launch_roi_only)
    pipe.timed(NSTREAM, 0)                                   # priming (eager): every handle allocates its workspace
    pipe_roi.timed(NSTREAM, 0)
    pipe.timed(W, 0, key=('warm', W))                        # warm-up: W steps (>= 3)
    launches0 = pipe.launches()
    g_timed = pipe.graph(K, W, ('timed', K))                 # capture (untimed) + one untimed priming replay
    if g_timed is not None:
        g_timed.replay(); g_timed._bx_primed = True
        torch.cuda.synchronize()
        launches0 = pipe.launches()
    sampler = ClockSampler(ctx.local); sampler.start()
    ms = pipe.timed(K, W, key=('timed', K))                  # ---- the timed region: K steps
    gpu_launches = (pipe.launches() - launches0) if g_timed is None else None
    ms_max, ms_ranks = ctx.max_over_ranks(ms)
    value = ctx.world * K * B / (ms_max * 1e-3)
    assert bool((rec['counts'][:K * B] == post).all()), 'a step kept fewer than post_nms proposals'
    gather_ok = verify_gather(ctx, rec, K, G) if ctx.world > 1 else None
    launches_per_step = 3                                    # proposals_kernel, roi_plan_kernel, roi_band_kernel
    if gpu_launches is None:                                 # graph replay: kernel nodes of the captured region
        gpu_launches = K * launches_per_step + (((K + G - 1) // G) if ctx.world > 1 else 0)

    # ---- roofline leg, same regime as `value`: the dominant kernel pair (plan + band) of the same K steps, pipelined
    #      over the same streams, one graph replay; average launch duration = region time / K
    ms_roi = pipe_roi.timed(K, W, key=('roi', K))
    ms_roi_max, _ = ctx.max_over_ranks(ms_roi)
    # ---- single-stream regime: the same K steps on ONE stream, every RoI launch bracketed by CUDA events (timed alone)
    _lib.check(lib.bx_profile_roi(pipe.handles[0], 1, K + 2))
    ms_single = pipe.timed(K, W, only_stream=0)
    buf = (ctypes.c_float * (K + 4))(); cnt = ctypes.c_int()
    _lib.check(lib.bx_profile_read(pipe.handles[0], buf, K + 4, ctypes.byref(cnt)))
    roi_ms_alone = list(buf[:cnt.value])
    _lib.check(lib.bx_profile_roi(pipe.handles[0], 0, 0))
    ms_single_max, _ = ctx.max_over_ranks(ms_single)
    sampler.stop_flag = True; sampler.join()
    stats = _lib.stats(pipe.handles[0])

    # ---- e2e: the public Python API with HOST (pinned) buffers; H2D inputs + D2H outputs inside the timed region
    e2e_steps = max(3, min(K, args.e2e_steps))
    cpus, numa_how = gpu_numa_cpus(ctx.local) if ctx.world > 1 else (None, 'single process')
    old_affinity = os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)                         # first touch of the pinned buffers on the GPU's own node
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    NE = 2   # two streams, two sets of pinned buffers: the D2H of step i overlaps the H2D + kernels of step i+1
    h_in = [dict(deltas=pin(b_['deltas']), scores=pin(b_['scores']), feat=pin(b_['feat'])) for b_ in host_batches[:NE]]
    h_outs = [(torch.empty((B, post, 4)).pin_memory(), torch.empty((B, post), dtype=torch.int32).pin_memory(),
               torch.empty((B,), dtype=torch.int32).pin_memory(), torch.empty((B * post, P, P, C)).pin_memory())
              for _ in range(NE)]
    if cpus:
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
    ctx.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.main)
    for s_ in e_streams:
        s_.wait_event(e0)
    for k in range(e2e_steps):
        e2e_step(k)
    for s_ in e_streams:
        ev = torch.cuda.Event(); ev.record(s_); ctx.main.wait_event(ev)
    e1.record(ctx.main)
    ctx.barrier()
    e2e_ms = e0.elapsed_time(e1)
    wall_ms = (time.perf_counter() - t0) * 1e3
    assert all(int(o[2].min()) == post for o in h_outs)
    e2e_ms_max, e2e_ranks = ctx.max_over_ranks(max(e2e_ms, 0.0))
    e2e_value = ctx.world * e2e_steps * B / (e2e_ms_max * 1e-3)

    probe = pcie_probe(ctx) if (ctx.world > 1 or args.pcie_probe) else None
    out = None
    if ctx.rank == 0:
        b_prop, b_roi = algorithmic_bytes(w, n, fh, fw)
        peak, which = hbm_peak()
        roi_bytes = B * b_roi                                 # one launch pair pools the whole batch
        step_bytes = B * (b_prop + b_roi)
        roi_avg_ms = ms_roi_max / K
        achieved = roi_bytes / (roi_avg_ms * 1e-3) / 1e9
        alone_ms = sum(roi_ms_alone) / max(1, len(roi_ms_alone))
        frac = lambda bytes_, ms_: round(bytes_ / (ms_ * 1e-3) / 1e9 / peak, 4)  # noqa: E731
        cpu = run_cpu_baseline(w, host_batches[0]) if ctx.world == 1 else dict(
            value=None, unit=UNIT, cores=0, kind='port', sample='not run at N > 1: measured on rank 0 at N = 1 only')
        out = dict(
            metric=METRIC, value=round(value, 1), unit=UNIT, n_gpus=ctx.world, steps=K, warmup=W,
            ms_per_step=round(ms_max / K, 5), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
            data='synthetic', config=headline_config(w),
            details=dict(streams=NSTREAM, proposal_side_streams=args.priority,
                         issue='one CUDA-graph replay of the K steps' if g_timed is not None else 'eager launches',
                         graph_error=ctx.graph_error,
                         priming='one eager step per stream (workspace allocation) and one untimed replay of the step graph '
                                 '(executable-graph upload) precede the W warm-up steps / the timed replay',
                         parallelism=('images sharded over GPUs, no data-path collective; the detection records (kept boxes + counts) the '
                                      'region accumulates are all-gathered every %d steps (K = %d: once, at the end of the region, inside '
                                      'the timed graph) by bx_allgather_detections on a dedicated ncclComm_t, on a communication branch no '
                                      'compute stream waits for; every rank verifies every rank\'s gathered records' % (G, K)) if ctx.world > 1 else 'single GPU',
                         gather_verified=gather_ok, ms_per_rank=ms_ranks,
                         rank_skew=round((max(ms_ranks) - min(ms_ranks)) / max(ms_ranks), 4),
                         algorithmic_bytes_per_image=b_prop + b_roi, band_launches=stats['band_launches'],
                         band_fallbacks=stats['band_fallbacks']),
            regimes=dict(
                pipelined=dict(ms_per_step=round(ms_max / K, 5), images_per_s=round(value, 1),
                               composite_hbm_frac=frac(step_bytes * K, ms_max), streams=NSTREAM,
                               roi_launch_ms=round(roi_avg_ms, 5), roi_hbm_frac=round(achieved / peak, 4)),
                single_stream=dict(ms_per_step=round(ms_single_max / K, 5),
                                   images_per_s=round(ctx.world * K * B / (ms_single_max * 1e-3), 1),
                                   composite_hbm_frac=frac(step_bytes * K, ms_single_max), streams=1,
                                   note='one lane: one stream (plus its proposal side stream when --priority > 0), steps back to back',
                                   roi_launch_ms=round(alone_ms, 5), roi_hbm_frac=frac(roi_bytes, alone_ms),
                                   launches_timed=len(roi_ms_alone))),
            roofline=dict(bound='hbm', kernel='roi_plan_kernel + roi_band_kernel (RoI pooling of one batch)',
                          achieved=round(achieved, 1), peak=peak, peak_source=which, unit='GB/s',
                          frac=round(achieved / peak, 4), traffic=NCU_TRAFFIC_BYTES_PER_LAUNCH,
                          traffic_source='profiles/ (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum)',
                          launches_timed=K, avg_launch_ms=round(roi_avg_ms, 5), algorithmic_bytes_per_launch=roi_bytes,
                          regime='pipelined (the regime of `value`): the K RoI launch pairs of the timed steps, same streams, '
                                 'one graph replay between CUDA events; avg_launch_ms = region time / K <= ms_per_step',
                          alone_launch_ms=round(alone_ms, 5), alone_frac=frac(roi_bytes, alone_ms)),
            cpu_baseline=cpu,
            e2e=dict(value=round(e2e_value, 1), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     steps=e2e_steps, ms_per_step=round(e2e_ms_max / e2e_steps, 3), wall_ms_per_step=round(wall_ms / e2e_steps, 3),
                     h2d_gbs_per_gpu=round(h2d * e2e_steps / (e2e_ms_max * 1e-3) / 1e9, 2),
                     d2h_gbs_per_gpu=round(d2h * e2e_steps / (e2e_ms_max * 1e-3) / 1e9, 2),
                     host_link_gbs_all_gpus=round(ctx.world * (h2d + d2h) * e2e_steps / (e2e_ms_max * 1e-3) / 1e9, 1),
                     bound='PCIe: the step moves 82 MB in and 482 MB out per 8 images; the kernels take 0.19 ms of it',
                     ms_per_rank=e2e_ranks,
                     api='ops.c4_proposal_roi_host -> bx_c4_proposal_roi_host (pinned host buffers, 2 streams)',
                     numa=(('bound to %d CPUs (%s)' % (len(cpus), numa_how)) if cpus else numa_how),
                     host_link_probe_gbs=probe),
            gpu_launches=gpu_launches, clocks=sampler.summary())
    pipe.close(); pipe_roi.close()
    del d_in, feat_bufs, h_outs, h_in, rec
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------- other workloads
FPN = dict(
    cfg3=dict(name='cfg3: ResNet-101 FPN 600x1000, 150111 anchors P2-P6, one global NMS -> 1000 rois/image (reference behaviour), '
                   'level assignment + RoI extractor 14x14 crop + 2x2 max -> 7x7x256 over P2-P5, batch 16/GPU',
              cfg=3, image_hw=(600, 1000), batch=16, scaling='weak'),
    cfg5=dict(name='cfg5: COCO-shape FPN 800x1333, 267069 anchors P2-P6, one global NMS -> 1000 rois/image, level assignment + RoI '
                   'extractor 14x14 crop + 2x2 max -> 7x7x256 over P2-P5, batch 64 sharded over the GPUs, detection all-gather',
              cfg=5, image_hw=(800, 1333), batch=64, scaling='strong'))
FPN_POST, FPN_P, FPN_C, FPN_NLEV, FPN_NSRC = 1000, 7, 256, 4, 4


def bench_fpn(ctx, args, name):
    """cfg3 / cfg5: proposals over the P2..P6 concatenation (base_fpn_model.py:219-225) + level assignment + FPN RoI
    extractor (:152-161, :303-324), straight through the C ABI with preallocated outputs."""
    torch, lib, _lib = ctx.torch, ctx.lib, ctx._lib
    from tf_eager_object_detection_b200 import ops
    w = FPN[name]
    hw = w['image_hw']
    B = w['batch'] if w['scaling'] == 'weak' else max(1, w['batch'] // ctx.world)
    if args.batch_override > 0:
        B = args.batch_override                  # experiments only: e.g. the per-GPU share of cfg5 at N = 8 on one GPU
    K, W = max(3, min(args.steps, args.workload_steps)), 3
    dev = ctx.dev
    POST, P, C, NLEV = FPN_POST, FPN_P, FPN_C, FPN_NLEV
    ims = [syn.fpn_image(w['cfg'], ctx.rank * 100 + i, hw, with_features=False) for i in range(FPN_NSRC)]
    n = ims[0]['anchors'].shape[0]
    anchors = torch.as_tensor(ims[0]['anchors']).to(dev)
    rep = (B + FPN_NSRC - 1) // FPN_NSRC
    shapes = syn.fpn_feature_shapes(hw)[:NLEV]
    NSTREAM = max(1, args.fpn_streams)
    NBUF = 2
    g = torch.Generator(device=dev); g.manual_seed(1234 + ctx.rank)
    d_in = []
    for k in range(NBUF):
        order = [(i + k) % FPN_NSRC for i in range(FPN_NSRC)]
        d_in.append(dict(
            deltas=torch.as_tensor(np.stack([ims[i]['deltas'] for i in order])).to(dev).repeat(rep, 1, 1)[:B].contiguous(),
            scores=torch.as_tensor(np.stack([ims[i]['scores'] for i in order])).to(dev).repeat(rep, 1)[:B].contiguous(),
            feats=[torch.randn((B, h, wd, C), device=dev, generator=g) for h, wd in shapes]))
    for d in d_in:
        d['fptr'] = (ctypes.c_void_p * NLEV)(*[f.data_ptr() for f in d['feats']])
    fh = (ctypes.c_int * NLEV)(*[s[0] for s in shapes]); fw = (ctypes.c_int * NLEV)(*[s[1] for s in shapes])
    bi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(POST)
    rec = make_records(ctx, max(K, W, NSTREAM), B, POST)
    idx_bufs = [torch.empty((B, POST), dtype=torch.int32, device=dev) for _ in range(NSTREAM)]
    out_bufs = [torch.empty((B * POST, P, P, C), device=dev) for _ in range(NSTREAM)]
    lv_bufs = [(torch.empty((B * POST,), dtype=torch.int32, device=dev), torch.empty((B * POST,), dtype=torch.int32, device=dev),
                torch.zeros((NLEV,), dtype=torch.int32, device=dev)) for _ in range(NSTREAM)]
    params = ops.proposal_params(hw, POST, 0.7)
    G = max(K, W, NSTREAM)                                  # one all-gather of the region's records, at its end
    pipe = None

    def launch(step, pos, si):
        din = d_in[step % NBUF]
        lo = pos * B
        st = ctypes.c_void_p(pipe.streams[si].cuda_stream)
        rois = rec['boxes'][lo:lo + B]
        # (the proposal stage as its own chain on a side stream, which pays at cfg2, measured SLOWER here: cfg5 23.2 k vs
        #  26.1 k images/s, cfg3 25.8 k vs 26.9 k — its six prefilter launches then compete with the extractor of the same lane)
        _lib.check(lib.bx_proposals(pipe.handles[si], anchors.data_ptr(), din['deltas'].data_ptr(), din['scores'].data_ptr(), B,
                                    n, ctypes.byref(params), rois.data_ptr(), idx_bufs[si].data_ptr(),
                                    rec['counts'][lo:lo + B].data_ptr(), st))
        lv, order, counts = lv_bufs[si]
        _lib.check(lib.bx_fpn_roi_features(pipe.handles[si], din['fptr'], fh, fw, NLEV, 2, B, C, rois.data_ptr(), bi.data_ptr(),
                                           B * POST, hw[0], hw[1], P, out_bufs[si].data_ptr(), lv.data_ptr(), order.data_ptr(),
                                           counts.data_ptr(), st))

    pipe = Pipeline(ctx, NSTREAM, launch, rec, G)
    pipe.timed(NSTREAM, 0)                                   # priming (eager)
    pipe.timed(W, 0, key=('warm', W))
    ms = pipe.timed(K, W, key=('timed', K))
    ms_max, ms_ranks = ctx.max_over_ranks(ms)
    assert bool((rec['counts'][:K * B] == POST).all()), '%s: a step kept fewer than 1000 proposals' % name
    gather_ok = verify_gather(ctx, rec, K, G) if ctx.world > 1 else None
    _lib.check(lib.bx_profile_roi(pipe.handles[0], 1, K + 2))
    ms_single = pipe.timed(K, W, only_stream=0)
    buf = (ctypes.c_float * (K + 4))(); cnt = ctypes.c_int()
    _lib.check(lib.bx_profile_read(pipe.handles[0], buf, K + 4, ctypes.byref(cnt)))
    roi_ms = list(buf[:cnt.value])
    _lib.check(lib.bx_profile_roi(pipe.handles[0], 0, 0))
    ms_single_max, _ = ctx.max_over_ranks(ms_single)
    out = None
    if ctx.rank == 0:
        peak, _ = hbm_peak()
        b_prop = 36 * n + 20 * POST
        b_roi = 4 * C * sum(h * wd for h, wd in shapes) + 16 * POST + 4 * POST * P * P * C
        roi_avg = sum(roi_ms) / max(1, len(roi_ms))
        value = ctx.world * K * B / (ms_max * 1e-3)
        out = dict(workload=w['name'], scaling=w['scaling'], images_per_step_per_gpu=B, steps=K, warmup=W, streams=NSTREAM,
                   ms_per_step=round(ms_max / K, 5), images_per_s=round(value, 1),
                   composite_hbm_frac=round(B * (b_prop + b_roi) * K / (ms_max * 1e-3) / 1e9 / peak, 4),
                   algorithmic_bytes_per_image=b_prop + b_roi,
                   single_stream=dict(ms_per_step=round(ms_single_max / K, 5),
                                      images_per_s=round(ctx.world * K * B / (ms_single_max * 1e-3), 1)),
                   roofline=dict(bound='hbm', kernel='FPN RoI extractor launch of one batch (level assignment + pooled crop kernel)',
                                 avg_launch_ms=round(roi_avg, 5), algorithmic_bytes_per_launch=B * b_roi,
                                 achieved=round(B * b_roi / (roi_avg * 1e-3) / 1e9, 1) if roi_avg > 0 else None,
                                 frac=round(B * b_roi / (roi_avg * 1e-3) / 1e9 / peak, 4) if roi_avg > 0 else None,
                                 launches_timed=len(roi_ms), regime='single stream, CUDA events around every launch'),
                   gather_verified=gather_ok, ms_per_rank=ms_ranks)
    pipe.close()
    del d_in, out_bufs, rec
    torch.cuda.empty_cache()
    return out


def bench_targets(ctx, args):
    """cfg4: anchor_target (21 546 anchors x 100 gt) + proposal_target (2000 training proposals x 100 gt), batch 16,
    fixed-permutation sampling (anchor_target.py:29-107, proposal_target.py:32-124)."""
    torch, lib, _lib = ctx.torch, ctx.lib, ctx._lib
    from tf_eager_object_detection_b200 import ops
    name = ('cfg4: training targets, anchor_target (21546 anchors x 100 gt, 256 samples) + proposal_target (2000 rois x 100 gt, '
            '128 samples, 21 classes), fixed-permutation sampling, batch 16/GPU')
    B, M, Kr, S, NC = 16, 100, 2000, 128, 21
    K, W = max(3, min(args.steps, args.workload_steps)), 3
    dev = ctx.dev
    imgs = [syn.c4_image(4, ctx.rank * 100 + i, with_features=False) for i in range(B)]
    anchors = torch.as_tensor(imgs[0]['anchors']).to(dev)
    n = anchors.shape[0]
    rng = np.random.default_rng(syn.seed_for(4, 900 + ctx.rank))
    gts, gls = zip(*[syn.gt_boxes(rng, M, (600, 1000)) for _ in range(B)])
    gt = torch.as_tensor(np.stack(gts)).to(dev); gl = torch.as_tensor(np.stack(gls)).to(dev)
    perm_a = torch.as_tensor(np.stack([rng.permutation(n) for _ in range(B)]).astype(np.int32)).to(dev)
    perm_r = torch.as_tensor(np.stack([rng.permutation(Kr) for _ in range(B)]).astype(np.int32)).to(dev)
    deltas = torch.as_tensor(np.stack([im['deltas'] for im in imgs])).to(dev)
    scores = torch.as_tensor(np.stack([im['scores'] for im in imgs])).to(dev)
    rois, _, rc = ops.proposals(anchors, deltas, scores, (600, 1000), Kr)            # the 2000 training proposals (untimed)
    torch.cuda.synchronize()
    NSTREAM = max(1, args.target_streams)
    o = [dict(lab=torch.empty((B, n), device=dev), tg=torch.empty((B, n, 4), device=dev), iw=torch.empty((B, n, 4), device=dev),
              ow=torch.empty((B, n, 4), device=dev), cnt=torch.empty((B, 2), dtype=torch.int32, device=dev),
              pr=torch.empty((B, S, 4), device=dev), pl=torch.empty((B, S), dtype=torch.int32, device=dev),
              pt=torch.empty((B, S, 4 * NC), device=dev), pi=torch.empty((B, S, 4 * NC), device=dev),
              po=torch.empty((B, S, 4 * NC), device=dev), pk=torch.empty((B, S), dtype=torch.int32, device=dev),
              pc=torch.empty((B, 2), dtype=torch.int32, device=dev)) for _ in range(NSTREAM)]
    f4 = _lib.f4
    ap = _lib.AnchorTargetParams(0.7, 0.3, 256, 128, f4((0, 0, 0, 0)), f4((1, 1, 1, 1)), 600, 1000)
    pp = _lib.ProposalTargetParams(NC, 0.5, 0.0, S, 32, f4((0, 0, 0, 0)), f4((0.1, 0.1, 0.2, 0.2)))
    pipe = None

    def launch(step, pos, si):
        st = ctypes.c_void_p(pipe.streams[si].cuda_stream)
        b = o[si]
        _lib.check(lib.bx_anchor_target(pipe.handles[si], anchors.data_ptr(), n, gt.data_ptr(), None, B, M, perm_a.data_ptr(),
                                        ctypes.byref(ap), b['lab'].data_ptr(), b['tg'].data_ptr(), b['iw'].data_ptr(),
                                        b['ow'].data_ptr(), b['cnt'].data_ptr(), st))
        _lib.check(lib.bx_proposal_target(pipe.handles[si], rois.data_ptr(), rc.data_ptr(), Kr, gt.data_ptr(), gl.data_ptr(), None,
                                          B, M, perm_r.data_ptr(), ctypes.byref(pp), b['pr'].data_ptr(), b['pl'].data_ptr(),
                                          b['pt'].data_ptr(), b['pi'].data_ptr(), b['po'].data_ptr(), b['pk'].data_ptr(),
                                          b['pc'].data_ptr(), st))

    pipe = Pipeline(ctx, NSTREAM, launch)
    pipe.timed(NSTREAM, 0)
    pipe.timed(W, 0, key=('warm', W))
    ms = pipe.timed(K, W, key=('timed', K))
    ms_max, ms_ranks = ctx.max_over_ranks(ms)
    ms_single = pipe.timed(K, W, only_stream=0)
    ms_single_max, _ = ctx.max_over_ranks(ms_single)
    assert bool((o[0]['pc'][:, 1] == 0).all()) and bool((o[0]['cnt'].sum(dim=1) == 256).all())
    out = None
    if ctx.rank == 0:
        peak, _ = hbm_peak()
        b_atgt = 16 * (n + M) + 52 * n
        b_ptgt = 16 * (Kr + M) + 4 * M + S * (16 + 4 + 3 * 16 * NC)
        out = dict(workload=name, scaling='weak', images_per_step_per_gpu=B, steps=K, warmup=W, streams=NSTREAM,
                   ms_per_step=round(ms_max / K, 5), images_per_s=round(ctx.world * K * B / (ms_max * 1e-3), 1),
                   composite_hbm_frac=round(B * (b_atgt + b_ptgt) * K / (ms_max * 1e-3) / 1e9 / peak, 4),
                   algorithmic_bytes_per_image=b_atgt + b_ptgt,
                   single_stream=dict(ms_per_step=round(ms_single_max / K, 5),
                                      images_per_s=round(ctx.world * K * B / (ms_single_max * 1e-3), 1)),
                   roofline=dict(bound='latency (7 dependent launches over 1.9 MB per image; not bandwidth-bound)',
                                 kernel='bx_anchor_target (5 launches) + bx_proposal_target (2 launches), one batch',
                                 avg_launch_ms=round(ms_single_max / K, 5), algorithmic_bytes_per_launch=B * (b_atgt + b_ptgt),
                                 frac=round(B * (b_atgt + b_ptgt) * K / (ms_single_max * 1e-3) / 1e9 / peak, 4)),
                   ms_per_rank=ms_ranks)
    pipe.close()
    torch.cuda.empty_cache()
    return out


def bench_backward(ctx, args):
    """f3: gradient of the cfg2 RoI extractor w.r.t. the feature map (what scripts/train.py:99-103 back-propagates through
    tf.image.crop_and_resize), 8 images x 300 rois per GPU and step, grad_out [2400,7,7,1024] -> grad_feat [8,38,63,1024].
    One stream (a call fills the device); the default kernel at this shape is the row-owned, atomic-free one, the scatter
    kernel (BX_ROI_GRAD_ATOMIC=1, read per call) is timed beside it."""
    torch, lib, _lib = ctx.torch, ctx.lib, ctx._lib
    name = ('f3: RoI-pooling backward at the cfg2 shape, 8 images x 300 rois per GPU and step, grad_out [2400,7,7,1024] -> '
            'grad_feat [8,38,63,1024] (ResNet-101 C4 map of 600x1000), one stream')
    B, R, P, C, fh, fw = 8, 300, 7, 1024, 38, 63
    K, W = max(3, min(args.steps, args.workload_steps)), 3
    dev = ctx.dev
    rng = np.random.default_rng(syn.seed_for(2, 700 + ctx.rank))
    rois = torch.as_tensor(np.stack([syn.random_rois(rng, R, (600, 1000)) for _ in range(B)]).reshape(-1, 4)).to(dev)
    counts = torch.full((B,), R, dtype=torch.int32, device=dev)
    go = [torch.randn((B * R, P, P, C), device=dev) for _ in range(2)]          # 2 x 481 MB, alternated: inputs larger than L2
    gf = torch.empty((B, fh, fw, C), device=dev)
    pipe = None

    def launch(step, pos, si):
        _lib.check(lib.bx_roi_pool_grad(pipe.handles[si], _lib.ROI_STRIDE_NORM, _lib.POOL_NONE, P, None, B, fh, fw, C,
                                        rois.data_ptr(), None, counts.data_ptr(), B * R, 16.0, 0, 0, go[step & 1].data_ptr(),
                                        gf.data_ptr(), ctypes.c_void_p(pipe.streams[si].cuda_stream)))

    pipe = Pipeline(ctx, 1, launch)
    res = {}
    for tag, env in (('row_owned', '0'), ('scatter', '1')):
        os.environ['BX_ROI_GRAD_ATOMIC'] = env
        pipe.timed(1, 0)
        pipe.timed(W, 0, key=('warm', tag))
        ms = pipe.timed(K, W, key=('timed', tag))
        res[tag], ranks = ctx.max_over_ranks(ms)
        if tag == 'row_owned':
            ms_ranks = ranks
            first = gf.clone()
            pipe.timed(1, W + K - 1)
            assert torch.equal(first, gf), 'row-owned backward kernel: two runs differ'
    os.environ.pop('BX_ROI_GRAD_ATOMIC', None)
    out = None
    if ctx.rank == 0:
        peak, _ = hbm_peak()
        alg = 4 * (B * R * P * P * C + B * fh * fw * C)
        ms = res['row_owned'] / K
        out = dict(workload=name, scaling='weak', images_per_step_per_gpu=B, steps=K, warmup=W, streams=1,
                   ms_per_step=round(ms, 5), images_per_s=round(ctx.world * B / (ms * 1e-3), 1),
                   scatter_kernel_ms_per_step=round(res['scatter'] / K, 5), bit_reproducible=True,
                   roofline=dict(bound='hbm', kernel='roi_grad_rows_kernel<NONE, 2, 4, 7, false> (1 launch per step)',
                                 avg_launch_ms=round(ms, 5), algorithmic_bytes_per_launch=alg,
                                 achieved=round(alg / (ms * 1e-3) / 1e9, 1), peak=peak, unit='GB/s',
                                 frac=round(alg / (ms * 1e-3) / 1e9 / peak, 4)),
                   ms_per_rank=ms_ranks)
    pipe.close()
    del go, gf
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    ctx = Ctx(args)
    line = bench_c4(ctx, args)
    names = ['cfg3', 'cfg5', 'cfg4', 'f3'] if args.workloads == 'all' else [s for s in args.workloads.split(',') if s and s != 'none']
    extra = {}

    def give_up():
        """Watchdog (every rank): the extra workloads may never take the headline line down — if they have not finished in
        time (a rank failed inside a collective, say), rank 0 prints the headline with what is there and everybody exits."""
        if ctx.rank == 0:
            extra['error'] = 'extra workloads did not finish within %d s; headline line printed without them' % args.workload_timeout
            line['workloads'] = extra
            print(json.dumps(line), flush=True)
        os._exit(0)
    watchdog = threading.Timer(args.workload_timeout, give_up)
    watchdog.daemon = True
    if names:
        watchdog.start()
    for nm in names:
        try:
            extra[nm] = bench_targets(ctx, args) if nm == 'cfg4' else bench_backward(ctx, args) if nm == 'f3' else bench_fpn(ctx, args, nm)
        except Exception as e:                               # a secondary workload never takes the headline line down
            extra[nm] = dict(error='%s: %s' % (type(e).__name__, str(e)[:300]))
            ctx.torch.cuda.synchronize()
    watchdog.cancel()
    if ctx.rank == 0:
        line['workloads'] = extra
        line['details']['graph_error'] = ctx.graph_error
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--streams', type=int, default=16, help='steps are issued round-robin over this many CUDA streams')
    ap.add_argument('--fpn-streams', type=int, default=4)
    ap.add_argument('--target-streams', type=int, default=8, help='streams of the cfg4 (training targets) workload')
    ap.add_argument('--e2e-steps', type=int, default=20)
    ap.add_argument('--gather-every', type=int, default=0,
                    help='N > 1: detection records of this many consecutive steps are all-gathered together (0: all K, one gather at the end of the region)')
    ap.add_argument('--workloads', default='all', help='all | none | comma list of cfg3,cfg5,cfg4,f3 (measured after the headline)')
    ap.add_argument('--workload-steps', type=int, default=100, help='cap on the steps of the extra workloads')
    ap.add_argument('--workload-timeout', type=int, default=240, help='seconds after which the extra workloads are abandoned')
    ap.add_argument('--batch-override', type=int, default=0, help='experiments: images per GPU and step of the FPN workloads')
    ap.add_argument('--pcie-probe', action='store_true', help='N = 1: also run the host-link copy probe of the e2e block')
    ap.add_argument('--priority', type=int, default=2, help='1: proposal kernels on high-priority side streams; 2: as independent chains (proposals of a later step overlap the RoI kernels of the lane)')
    ap.add_argument('--no-graph', action='store_true', help='issue the timed region eagerly instead of as one CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
