#!/usr/bin/env python
"""bench.py -- the proposal-and-RoI hot path on synthetic inputs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
                    [--no-extra-workloads] [--no-cpu-baseline] [--no-graphs]

A step = one pass of the hot path over one batch: RPN proposals (decode, clip, min-size, top-k, NMS,
pad/gather) for B images, then the RoI head's coordinate map + RoI gather of every proposal.
Headline workload = BASELINE.json configs[1] (cfg2): B=16 600x600 images, ResNet-50 stride-16 features
[16,1024,38,38], 3000 pre-NMS / 300 post-NMS RoIs per image, RoIPool 14x14.  The backbone is not part of
the path: features / RPN conv outputs are synthetic tensors of the right shape (data: synthetic).

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same
through the module API (ProposalCreator.batched + HarNetRoIHead.forward) from pinned host buffers with
the host<->device copies inside the timed region; `roofline` = the RoI gather kernel (name queried from
the library) against the measured HBM copy peak; `workloads` = the other BASELINE configs (cfg3 training
step, cfg4 RoIAlign, cfg5 proposal stress) timed the same way in the same run; `cpu_baseline` = the oracle
port on the host cores on a bounded sample (+ torchvision's own CPU nms / roi_pool beside it).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: B per GPU, image S, feature C/H/W, n_pre, n_post, roi op, pooled P, extra
    "cfg2": dict(batch=16, img=600, C=1024, H=38, W=38, n_pre=3000, n_post=300, op="pool", P=14,
                 desc="ResNet-50 Faster R-CNN batched inference, batch 16 synthetic 600x600, 300 post-NMS "
                      "RoIs/image, RoIPool 14x14"),
    "cfg3": dict(batch=8, img=600, C=512, H=38, W=38, n_pre=12000, n_post=600, op="pool", P=7, train=True, n_gt=8,
                 desc="training-step hot path: proposals 12000->600, AnchorTargetCreator(256), "
                      "ProposalTargetCreator(128), RoIPool 7x7 (+argmax) on the 128 sampled RoIs, 8 images per GPU"),
    "cfg4": dict(batch=32, img=800, C=512, H=50, W=50, n_pre=3000, n_post=300, op="align", P=7,
                 desc="HarDNet Faster R-CNN inference, batch 32 synthetic 800x800, RoIAlign 7x7 (sr=2)"),
    "cfg5": dict(batch=8, img=1024, C=512, H=64, W=64, n_pre=30000, n_post=2000, op="pool", P=7,
                 desc="proposal stress: 1024x1024, 36864 anchors, 30k pre / 2k post NMS, 8 images per GPU"),
}
METRIC = "images/sec (RPN proposals + RoI gather hot path)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def config_dict(name, cfg, world):
    """The `config` object of the JSON line: identical for both arms (the driver compares them)."""
    B, H, W, C, P = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"]
    in_mb = (B * H * W * 9 * 6 + B * C * H * W) * 4 / 1e6
    return {"workload": f"{name}: {cfg['desc']}", "batch_per_gpu": B, "global_batch": world * B,
            "anchors_per_image": H * W * 9, "n_pre_nms": cfg["n_pre"], "n_post_nms": cfg["n_post"],
            "feature": [B, C, H, W], "roi_op": f"{cfg['op']} {P}x{P}",
            "l2": f"3 input sets rotated (3 x {in_mb:.0f} MB) and the gather output rewritten every step: a step "
                  "never finds its inputs in L2 (126 MB)",
            "parallelism": f"dp{world} (images sharded per GPU; one NCCL all_gather of the rois per step when N>1, "
                           "issued after the proposal layer, waited for two steps later)"}


# ------------------------------------------------------------------------------------------------
# host placement: the rank's CPU threads and (first-touch) its pinned buffers on the GPU's NUMA node
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    """Best effort: restrict this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned
    buffer is allocated (first touch then places the pages on that node).  Returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local
        if vis:
            tok = vis.split(",")[local].strip()
            h = pynvml.nvmlDeviceGetHandleByIndex(int(tok)) if tok.isdigit() else pynvml.nvmlDeviceGetHandleByUUID(tok)
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]  # 0000:xx:yy.z
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"node": None, "note": "platform reports no NUMA affinity for the GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed), "pci": bdf}
    except Exception as exc:  # containers without sysfs NUMA files, no NVML, ...
        return {"node": None, "note": f"not bound: {type(exc).__name__}"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every ~2 ms; nvidia-smi is too
    slow for a 20 ms region and is only the fallback)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index = index
        self.sm, self.bits, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._ready = threading.Event()  # first sample taken: the timed region may start
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                tok = vis.split(",")[self.index].strip()
                if tok.isdigit():
                    idx = int(tok)
                else:
                    uuid = tok
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                self._ready.set()
                self._stop.wait(0.002)
        except Exception:
            self._smi_loop()

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while True:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    r = [c.strip() for c in out.stdout.strip().splitlines()[0].split(",")]
                    self.sm.append(float(r[0]))
                    self.max_mhz = float(r[1])
                    self._ready.set()
                    for i, n in enumerate(names):
                        if r[2 + i].lower().startswith("active"):
                            self.bits |= self.REASONS[n]
            except Exception:
                pass
            if self._stop.wait(0.2):
                break

    def start(self):
        """Opens NVML and takes the first sample; call BEFORE the barrier that aligns the ranks."""
        self._t.start()
        self._ready.wait(timeout=5.0)
        return self

    def mark(self):
        """The timed region starts now: samples taken while waiting do not belong to it."""
        self.sm.clear()
        self.bits = 0

    def stop(self):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in self.REASONS.items() if self.bits & bit], "samples": len(sm)}


def make_inputs(cfg, seed, device=None, pin=False):
    """Synthetic RPN conv outputs + backbone features of the workload's shape."""
    g = torch.Generator().manual_seed(seed)
    B, N = cfg["batch"], cfg["H"] * cfg["W"] * 9
    loc = (torch.randn(B, N, 4, generator=g) * 0.2).float()
    logits = torch.randn(B, N, 2, generator=g).float()
    feat = torch.relu(torch.randn(B, cfg["C"], cfg["H"], cfg["W"], generator=g)).float()
    ts = [loc, logits, feat]
    if pin:
        ts = [t.pin_memory() for t in ts]
    if device is not None:
        ts = [t.to(device) for t in ts]
    return ts


def input_seed(rank, s):
    return 1000 * rank + s


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Workload:
    """One BASELINE config on one GPU: rotating input sets, preallocated gather output, step(i)."""

    def __init__(self, name, dev, rank, world):
        from two_stage_object_detection_b200 import functional as F
        self.F, self.name, self.cfg, self.dev, self.rank, self.world = F, name, WORKLOADS[name], dev, rank, world
        cfg = self.cfg
        B, H, W, C, P = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"]
        self.S, self.n_post, self.train = cfg["img"], cfg["n_post"], bool(cfg.get("train"))
        self.base = F.base_anchors(device=dev)
        self.idx = torch.arange(B, dtype=torch.int32, device=dev)
        # three input sets rotated between steps (3 x ~100 MB > L2) and a GB-sized output per step: no step
        # finds its inputs in L2
        self.sets = [make_inputs(cfg, input_seed(rank, s), device=dev) for s in range(3)]
        self.pkw = dict(clip_x_max=self.S, clip_y_max=self.S, n_pre_nms=cfg["n_pre"], n_post_nms=self.n_post,
                        nms_iou=0.7, min_size=16.0, base=self.base, feat_stride=16, feat_hw=(H, W),
                        score_is_logits=True)
        self.rois_per_image = 128 if self.train else self.n_post
        self.K = B * self.rois_per_image
        self.pooled = torch.empty((self.K, C, P, P), dtype=torch.float32, device=dev)
        if self.train:  # G boxes per image, centre U(0,S)^2, w,h U(50,250), clipped; labels in [0,20)
            gg = torch.Generator().manual_seed(77 + rank)
            G = cfg["n_gt"]
            ctr = torch.rand(B, G, 2, generator=gg) * self.S
            wh = 50 + torch.rand(B, G, 2, generator=gg) * 200
            self.gt_box = torch.cat([ctr - wh / 2, ctr + wh / 2], -1).clamp(0, self.S).to(dev)
            self.gt_lab = torch.randint(0, 20, (B, G), generator=gg).to(dev)
            self.n_gt = torch.full((B,), G, dtype=torch.int32, device=dev)
        # the path's only collective: all-gather of the rois.  Two buffers: step i's gather is waited for at step
        # i + 2, so a late rank delays nobody inside a step and the collective has a whole step to finish
        self.gathered = [torch.empty((world * B, self.n_post, 4), dtype=torch.float32, device=dev)
                         for _ in range(2)] if world > 1 else None
        self.works = [None, None]
        self.wait_ev = []
        self.graphs = None           # capture(): the step replayed from CUDA graphs
        self.replayed_launches = 0   # kernels inside the graphs replayed so far (counted at capture time)
        n_alg = self.K * C * P * P * 4 * (2 if self.train else 1) + self.K * 20 + B * C * H * W * 4
        self.alg_bytes = n_alg  # gather kernel: output (+ int32 argmax when training) + rois + features once

    def proposals_only(self, i, sets=None):
        loc, logits, _ = (sets or self.sets)[i % 3]
        return self.F.proposals(loc, logits, **self.pkw)

    def gather(self, feat, rois):
        F, cfg = self.F, self.cfg
        H, W, P = cfg["H"], cfg["W"], cfg["P"]
        rois5 = F.roi_head_coords(rois, self.idx, (self.S, self.S), (H, W))
        if self.train:
            F.roi_pool_forward(feat, rois5, P, 1.0, with_argmax=True, out=self.pooled, rois_per_image=self.rois_per_image)
        elif cfg["op"] == "pool":
            F.roi_pool_forward(feat, rois5, P, 1.0, out=self.pooled, rois_per_image=self.rois_per_image)
        else:
            F.roi_align_forward(feat, rois5, P, 1.0, 2, False, out=self.pooled, rois_per_image=self.rois_per_image)

    def front(self, s):
        """Proposal layer (+ target assignment when training) on input set s -> (rois, status, RoIs for the gather)."""
        F, cfg = self.F, self.cfg
        loc, logits, _ = self.sets[s]
        rois, src, n_keep, status = F.proposals(loc, logits, **self.pkw)
        sel = rois
        if self.train:
            F.anchor_targets(self.gt_box, self.n_gt, base=self.base, feat_stride=16, feat_hw=(cfg["H"], cfg["W"]))
            sel, _, _, _, _ = F.proposal_targets(rois, self.gt_box, self.gt_lab, self.n_gt)
        return rois, status, sel

    def capture(self):
        """The step as CUDA graphs, one pair per input set: (proposal layer [+ targets]) and (RoI gather).  The only
        thing between the two is the NCCL all-gather of the rois (N > 1), issued from the stream as before.  A step
        is ~14-25 launches of 2-90 us kernels: replaying them removes the launch gaps (cfg2: ~30 us of 840)."""
        from two_stage_object_detection_b200 import _lib
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up on a capture-like stream: per-stream workspaces, smem attributes
            for s in range(3):
                self.gather(self.sets[s][2], self.front(s)[2])
        cur.wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for s in range(3):
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(ga):
                rois, status, sel = self.front(s)
            n1 = _lib.launch_count()
            with torch.cuda.graph(gb):
                self.gather(self.sets[s][2], sel)
            n2 = _lib.launch_count()
            graphs.append(dict(front=ga, back=gb, rois=rois, status=status, launches=n2 - n0))
        self.graphs = graphs

    def step(self, i, ev=None, collective=True):
        import torch.distributed as dist
        F, cfg = self.F, self.cfg
        loc, logits, feat = self.sets[i % 3]
        g = self.graphs[i % 3] if self.graphs else None
        if ev:
            ev["p0"][i].record()
        if g:
            g["front"].replay()
            rois, status = g["rois"], g["status"]
            self.replayed_launches += g["launches"]
        else:
            rois, status, sel = self.front(i % 3)
        if ev:
            ev["p1"][i].record()
        if self.world > 1 and collective:
            slot = i % 2
            if self.works[slot] is not None:  # the gather issued two steps ago wrote this buffer
                w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                w0.record()
                self.works[slot].wait()  # stream-side: the host does not block
                w1.record()
                self.wait_ev.append((w0, w1))
            # the detections are final once the proposal layer is done: their all-gather runs on NCCL's stream
            # under this step's RoI gather and the next step's proposal layer
            self.works[slot] = dist.all_gather_into_tensor(self.gathered[slot], rois, async_op=True)
        if ev:
            ev["r0"][i].record()
        if g:
            g["back"].replay()
        else:
            self.gather(feat, sel)
        if ev:
            ev["r1"][i].record()
        return rois, status

    def drain(self):
        for k in range(2):
            if self.works[k] is not None:
                self.works[k].wait()
                self.works[k] = None


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def measure(wl, steps, warmup, clk=None):
    """W warm-up steps, then exactly `steps` timed ones between two aligned (barrier + synchronize) points.
    Returns the per-rank numbers; the caller max-reduces ms over ranks."""
    from two_stage_object_detection_b200 import _lib
    total = steps + warmup
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(total)] for k in ("p0", "p1", "r0", "r1")}
    for i in range(warmup):
        wl.step(i, ev)
    wl.drain()
    wl.wait_ev.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(wl.world)  # ranks aligned, streams idle: nothing (NVML start-up, warm-up tails) leaks into the region
    if clk:
        clk.mark()
    n0 = _lib.launch_count() + wl.replayed_launches
    t0.record()
    for i in range(warmup, total):
        rois, status = wl.step(i, ev)
    wl.drain()
    t1.record()
    n1 = _lib.launch_count() + wl.replayed_launches
    barrier(wl.world)
    ms = t0.elapsed_time(t1)
    sl = slice(warmup, total)
    out = {"ms_total": ms, "ms_step": ms / steps, "launches": n1 - n0,
           "roi_ms": float(np.mean([a.elapsed_time(b) for a, b in zip(ev["r0"][sl], ev["r1"][sl])])),
           "prop_ms": float(np.mean([a.elapsed_time(b) for a, b in zip(ev["p0"][sl], ev["p1"][sl])])),
           "kernel": _lib.last_roi_kernel(), "last_i": total - 1,
           "allgather_wait_ms": (float(np.mean([a.elapsed_time(b) for a, b in wl.wait_ev])) if wl.wait_ev else None)}
    assert not status.cpu().numpy().any(), "proposal layer flagged an index error"
    return out, rois


def traffic_for(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of exactly this kernel instance, from a committed
    `ncu --set full` capture (profiles/traffic.json, keyed by kernel name so a changed kernel reads null)."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None
    with open(tpath) as f:
        for rec in json.load(f).values():
            if rec.get("kernel", "").replace(" ", "") == kernel_name.replace(" ", ""):
                return rec.get("dram_bytes")
    return None


def roofline_obj(wl, m):
    peak, peak_src = peaks()
    achieved = wl.alg_bytes / (m["roi_ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": m["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic_for(m["kernel"]), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": wl.alg_bytes, "kernel_ms": m["roi_ms"]}


def max_over_ranks(vals, dev, world):
    if world == 1:
        return [float(v) for v in vals]
    import torch.distributed as dist
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def run_ours(args):
    import torch.distributed as dist
    from two_stage_object_detection_b200 import functional as F
    from two_stage_object_detection_b200.nets import HarNetRoIHead, ProposalCreator
    from two_stage_object_detection_b200.nets.frcnn import GlobalAvgClassifier

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    numa = bind_to_gpu_numa(local)  # before the first pinned allocation
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload
    cfg = WORKLOADS[name]
    B, H, W, C, P = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"]
    S, n_post = cfg["img"], cfg["n_post"]
    wl = Workload(name, dev, rank, world)
    launch_mode = "stream launches (--no-graphs)"
    if not args.no_graphs:
        try:
            wl.capture()
            launch_mode = ("CUDA graph replay: one graph for the proposal layer, one for the RoI gather, per input set"
                           + ("; the NCCL all-gather is issued from the stream between them" if world > 1 else ""))
        except Exception as exc:  # the kernel-by-kernel path is always there
            wl.graphs = None
            launch_mode = f"stream launches (graph capture failed: {type(exc).__name__}: {exc})"
    clk = ClockSampler(local).start()  # NVML is open and sampling BEFORE the ranks are aligned
    m, rois = measure(wl, args.steps, args.warmup, clk)
    clk.stop()
    ms_step, = max_over_ranks([m["ms_step"]], dev, world)

    # ---- multi-GPU: the gathered detections are the other ranks' own results, bit for bit ---------------------
    sharded_parity = None
    if world > 1:
        # every rank regenerates the inputs of the LAST rank's last timed step from their seed, recomputes that shard
        # on its own GPU and compares it with the slice the all-gather delivered
        r = world - 1
        i = m["last_i"]
        other = make_inputs(cfg, input_seed(r, i % 3), device=dev)
        exp, _, _, _ = F.proposals(other[0], other[1], **wl.pkw)
        got = wl.gathered[i % 2][r * B:(r + 1) * B]
        mine = wl.gathered[i % 2][rank * B:(rank + 1) * B]
        ok = torch.equal(got, exp) and torch.equal(mine, rois)
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        sharded_parity = bool(flag.item() == 1.0)
        del other, exp

    # ---- the same step launched kernel by kernel from the stream (informational: what the graphs save) ----------
    stream_ms = None
    if wl.graphs and world == 1:
        saved, wl.graphs = wl.graphs, None
        ms2, _ = measure(wl, min(args.steps, 20), 3)
        stream_ms = ms2["ms_step"]
        wl.graphs = saved

    # ---- the head's gather when its classifier is the global average (HarDNet): fused kernel, [K,C] out ------
    fused_ms = None
    if not wl.train:
        loc, logits, feat = wl.sets[0]
        rois_f, _, _, _ = F.proposals(loc, logits, **wl.pkw)
        rois5_f = F.roi_head_coords(rois_f, wl.idx, (S, S), (H, W))
        fused = (lambda: F.roi_pool_mean(feat, rois5_f, P, 1.0, rois_per_image=n_post)) if cfg["op"] == "pool" else \
                (lambda: F.roi_align_mean(feat, rois5_f, P, 1.0, 2, False, rois_per_image=n_post))
        for _ in range(3):
            fused()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        f0.record()
        for _ in range(20):
            fused()
        f1.record()
        torch.cuda.synchronize()
        fused_ms = f0.elapsed_time(f1) / 20

    # ---- end to end through the module API, from pinned host memory --------------------------------
    e2e = e2e_leg(args, wl, dev, rank, world, ProposalCreator, HarNetRoIHead, GlobalAvgClassifier)
    roof = roofline_obj(wl, m)
    launches = m["launches"]
    out = {
        "metric": METRIC, "value": world * B / (ms_step * 1e-3),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, cfg, world),
        "proposals_per_sec": world * B * n_post / (ms_step * 1e-3),
        "breakdown_ms": {"proposals": m["prop_ms"], "roi_gather": m["roi_ms"]},
        "fused_head_gather_ms": fused_ms,  # RoI gather + global-average classifier in one kernel (informational)
        "launch_mode": launch_mode,
        "stream_launch_ms_per_step": stream_ms,  # the same step without graphs (informational)
        "e2e": e2e,
        "gpu_launches": launches,  # counted by the library (frcnn_launch_count) over the timed region of this rank
        "roofline": roof,
        "clocks": clk.summary(),
        "numa": numa,
    }
    if world > 1:
        out["sharded_parity"] = sharded_parity
        out["allgather_wait_ms"] = m["allgather_wait_ms"]
    del wl
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, same code path, same timing rules, in the same record ----------------------
    if not args.no_extra_workloads:
        extra = {}
        for other in ("cfg3", "cfg4", "cfg5"):
            if other == name:
                continue
            w2 = Workload(other, dev, rank, world)
            if not args.no_graphs:
                try:
                    w2.capture()
                except Exception:
                    w2.graphs = None
            m2, _ = measure(w2, min(args.steps, 20), max(3, min(args.warmup, 5)))
            ms2, = max_over_ranks([m2["ms_step"]], dev, world)
            c2 = WORKLOADS[other]
            extra[other] = {"desc": c2["desc"], "ms_per_step": ms2, "images_per_s": world * c2["batch"] / (ms2 * 1e-3),
                            "proposals_per_s": world * c2["batch"] * c2["n_post"] / (ms2 * 1e-3),
                            "breakdown_ms": {"proposals": m2["prop_ms"], "roi_gather": m2["roi_ms"]},
                            "gpu_launches": m2["launches"], "graphs": w2.graphs is not None,
                            "roofline": roofline_obj(w2, m2)}
            del w2
            torch.cuda.empty_cache()
        out["workloads"] = extra
        try:
            out["ddp_train_step"] = ddp_train_leg(dev, rank, world, local)
        except Exception as exc:  # informational leg: never takes the line down
            out["ddp_train_step"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(cfg, images=cfg["batch"])
            tv = torchvision_baseline(cfg, images=cfg["batch"])
            if tv:
                out["cpu_baseline_torchvision"] = tv
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ddp_train_leg(dev, rank, world, local, steps=6):
    """BASELINE config 3 as a whole training step: FasterRCNNTrainer (batched trainer glue: proposals, anchor / proposal
    targets, RoIPool forward AND backward in the sm_100a kernels) on 8 synthetic 600x600 images per GPU, stock DDP
    gradient all-reduce over NCCL when N > 1, AdamW step.  The backbone is a two-conv stand-in (backbones are out of
    scope), so the figure is about the path + the collective, not about a ResNet.  Checks: finite loss, replicas
    bit-identical after the steps."""
    import torch.distributed as dist
    from torch import nn
    from two_stage_object_detection_b200.nets import FasterRCNNTrainer

    class TinyExtractor(nn.Module):
        def __init__(self):
            super().__init__()
            self.net = nn.Sequential(nn.Conv2d(3, 32, 3, 4, 1), nn.ReLU(), nn.Conv2d(32, 512, 3, 4, 1), nn.ReLU())

        def forward(self, x):
            return self.net(x)

    torch.manual_seed(0)
    model = FasterRCNNTrainer("train", num_classes=20, extractor=TinyExtractor()).to(dev)
    ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, weight_decay=1e-4)
    g = torch.Generator().manual_seed(300 + rank)
    B, S, G = 8, 600, 8
    imgs = [torch.rand(3, S, S, generator=g).to(dev) for _ in range(B)]
    boxes, labels = [], []
    for _ in range(B):
        c = torch.rand(G, 2, generator=g) * S
        wh = 50 + torch.rand(G, 2, generator=g) * 200
        boxes.append(torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, S).to(dev))
        labels.append(torch.randint(0, 20, (G,), generator=g).to(dev))
    finite = True
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for step in range(steps + 2):
        if step == 2:
            barrier(world)
            t0.record()
        losses, *_ = ddp(imgs, boxes, labels)
        opt.zero_grad(set_to_none=True)
        losses[-1].backward()
        opt.step()
        finite = finite and bool(torch.isfinite(losses[-1].detach()))
    t1.record()
    barrier(world)
    ms, = max_over_ranks([t0.elapsed_time(t1) / steps], dev, world)
    same = True
    if world > 1:
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([1.0 if torch.equal(flat, ref) and finite else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item() == 1.0)
    return {"ms_per_step": ms, "images_per_s": world * B / (ms * 1e-3), "batch_per_gpu": B, "steps": steps,
            "loss_finite": finite, "replicas_identical": same,
            "what": "FasterRCNNTrainer forward + backward + AdamW on 8 synthetic 600x600 images per GPU with 8 GT boxes each "
                    "(stand-in two-conv extractor), DDP gradient all-reduce when N > 1"}


def e2e_leg(args, wl, dev, rank, world, ProposalCreator, HarNetRoIHead, GlobalAvgClassifier):
    cfg = wl.cfg
    B, H, W, C, P, S, n_post = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"], cfg["img"], cfg["n_post"]
    creator = ProposalCreator("test", n_test_pre_nms=cfg["n_pre"], n_test_post_nms=n_post)
    head = HarNetRoIHead(n_class=21, roi_size=P, spatial_scale=1, classifier=GlobalAvgClassifier(),
                         in_features=C, roi_op=cfg["op"], sampling_ratio=2).to(dev).eval()
    # ONE pinned slab per rank (allocated after the NUMA binding: first touch puts it next to the GPU), two input
    # sets carved out of it
    shapes = [(B, H * W * 9, 4), (B, H * W * 9, 2), (B, C, H, W)]
    sizes = [int(np.prod(s)) for s in shapes]
    slab = torch.empty(2 * sum(sizes), dtype=torch.float32).pin_memory()
    host, off = [], 0
    for s in range(2):
        src = make_inputs(cfg, input_seed(rank, 10 + s))
        views = []
        for t, n, shp in zip(src, sizes, shapes):
            v = slab[off:off + n].view(shp)
            v.copy_(t)
            views.append(v)
            off += n
        host.append(views)
    h2d = sum(sizes) * 4

    def e2e_compute(loc, logits, feat):
        rois, _, _, st = creator.batched(loc, logits, (3, S, S), 1.0, base=wl.base, feat_stride=16, feat_hw=(H, W),
                                         score_is_logits=True)
        with torch.no_grad():
            cls_locs, scores = head(feat, rois, None, (S, S))
        return [rois, cls_locs, scores, st]

    def e2e_step(i):  # serial form: copy in, compute, copy out, host waits
        return [o.cpu() for o in e2e_compute(*(t.to(dev, non_blocking=True) for t in host[i % 2]))]

    e2e_steps = max(3, min(args.steps, 20))
    for i in range(3):
        outs = e2e_step(i)
    d2h = sum(o.numel() * o.element_size() for o in outs)
    barrier(world)
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier(world)
    e2e_serial_ms = (time.perf_counter() - w0) * 1e3 / e2e_steps

    # pipelined form (what a serving loop does): step i+1's host->device copy runs on a copy stream while
    # step i computes; every step still copies its own inputs in and its results out, and the host reads
    # each step's results (one step later).  Double-buffered device inputs and pinned host outputs.
    copy_s, comp_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dev_in = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host[0]] for _ in range(2)]
    probe = e2e_compute(*dev_in[0])
    host_out = [[torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in probe] for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_done = [torch.cuda.Event() for _ in range(2)]
    copy_t = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    torch.cuda.synchronize()

    def e2e_pipe(n, timed=False):
        seen = 0.0
        for i in range(n):
            s = i % 2
            with torch.cuda.stream(copy_s):
                if i >= 2:
                    copy_s.wait_event(in_free[s])
                if timed:
                    copy_t[i][0].record(copy_s)
                for d, h in zip(dev_in[s], host[i % 2]):
                    d.copy_(h, non_blocking=True)
                if timed:
                    copy_t[i][1].record(copy_s)
                in_ready[s].record(copy_s)
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(in_ready[s])
                outs = e2e_compute(*dev_in[s])
                in_free[s].record(comp_s)
                for ho, o in zip(host_out[s], outs):
                    ho.copy_(o, non_blocking=True)
                out_done[s].record(comp_s)
            if i >= 1:  # the host consumes step i-1's results while step i is in flight
                out_done[1 - s].synchronize()
                seen += float(host_out[1 - s][0][0, 0, 0])
        out_done[(n - 1) % 2].synchronize()
        return seen

    e2e_pipe(4)
    barrier(world)
    w0 = time.perf_counter()
    e2e_pipe(e2e_steps, timed=True)
    barrier(world)
    e2e_ms = (time.perf_counter() - w0) * 1e3 / e2e_steps
    copy_ms = float(np.median([a.elapsed_time(b) for a, b in copy_t]))
    e2e_ms, e2e_serial_ms, copy_ms = max_over_ranks([e2e_ms, e2e_serial_ms, copy_ms], dev, world)
    return {"value": world * B / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "serial_ms_per_step": e2e_serial_ms,
            "h2d_copy_ms": copy_ms, "h2d_gbs_per_gpu": h2d / (copy_ms * 1e-3) / 1e9,
            "api": "ProposalCreator.batched + HarNetRoIHead.forward (RoI gather + global-average classifier fused "
                   "into one kernel, Linear heads) from ONE pinned host slab per rank (allocated on the GPU's NUMA "
                   "node); copies of step i+1 overlap the compute of step i (copy stream + compute stream, "
                   "double-buffered); host-link-bound: h2d_gbs_per_gpu is the slowest rank's measured copy rate"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (kind "port"; /root/reference does not exist on the GPU box)
# ------------------------------------------------------------------------------------------------
def cpu_step(cfg, O, loc, logits, feat, images):
    """Reference algorithm for `images` images: softmax, anchors, decode, proposal layer, RoI gather."""
    B, H, W, P, S = images, cfg["H"], cfg["W"], cfg["P"], cfg["img"]
    anchor = O.shifted_anchors(O.base_anchors(), 16, H, W)
    fg = O.fg_scores(logits[:B])
    dec = np.stack([O.decode(anchor, loc[b]) for b in range(B)])
    rois, _, _, rc = O.proposal_layer_batch_from_boxes(dec, fg, (3, S, S), 1.0, 0.7, cfg["n_pre"], cfg["n_post"], 16)
    assert not rc.any()
    op = "pool" if cfg["op"] == "pool" else "align"
    kw = {} if op == "pool" else dict(sampling_ratio=2, aligned=False)
    return O.roi_head_gather(feat[:B], rois, np.arange(B), (S, S), roi_size=P, spatial_scale=1.0, op=op, **kw)


def cpu_baseline(cfg, images=2, budget_s=12.0, max_reps=40):
    from oracle import ref_port as O
    O.set_threads(len(os.sched_getaffinity(0)))
    loc, logits, feat = (t.numpy() for t in make_inputs(cfg, 7))
    images = min(images, cfg["batch"])
    cpu_step(cfg, O, loc, logits, feat, 2)  # warm-up (builds / loads the C library)
    times, start = [], time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - start < budget_s and len(times) < max_reps):  # ~10 s of CPU work
        t0 = time.perf_counter()
        cpu_step(cfg, O, loc, logits, feat, images)
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    return {"value": images / dt, "unit": "images/s", "cores": O.max_threads(), "kind": "port",
            "sample": f"{images} of {cfg['batch']} images of the same workload per pass (oracle/ref_port.py + "
                      f"frcnn_oracle.c, OpenMP over images and RoIs), median of {len(times)} passes, {dt:.2f} s each, "
                      f"{sum(times):.1f} s of CPU work"}


def torchvision_baseline(cfg, images=2, budget_s=20.0):
    """The reference's own native callables on the host cores, where they exist on this box: torchvision's CPU
    `nms` (nets/rpn.py:63) and `roi_pool` / `roi_align` (nets/classify.py:43) inside the same per-image flow
    (decode / sort by the oracle's numpy, as the reference does them with ATen)."""
    try:
        import torchvision
        from torchvision.ops import nms as tv_nms, roi_align as tv_roi_align, roi_pool as tv_roi_pool
    except Exception:
        return None
    from oracle import ref_port as O
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    loc, logits, feat = make_inputs(cfg, 7)
    H, W, P, S = cfg["H"], cfg["W"], cfg["P"], cfg["img"]
    anchor = O.shifted_anchors(O.base_anchors(), 16, H, W)

    def one(b):
        fg = torch.softmax(logits[b], -1)[:, 1]
        dec, valid = O.clip_filter(O.decode(anchor, loc[b].numpy()), (3, S, S), 16.0)
        valid = torch.from_numpy(valid)
        roi = torch.from_numpy(dec)[valid]
        sc = fg[valid]
        order = torch.argsort(sc, descending=True, stable=True)[:cfg["n_pre"]]
        roi, sc = roi[order], sc[order]
        keep = tv_nms(roi, sc, 0.7)
        if len(keep) < cfg["n_post"]:
            keep = torch.cat([keep, torch.arange(cfg["n_post"] - len(keep))])
        roi = roi[keep[:cfg["n_post"]]]
        r5 = torch.cat([torch.zeros(len(roi), 1), roi / S * W], 1)
        fb = feat[b:b + 1]
        return tv_roi_pool(fb, r5, (P, P), 1.0) if cfg["op"] == "pool" else tv_roi_align(fb, r5, (P, P), 1.0, 2, False)

    one(0)
    n, t0 = 0, time.perf_counter()
    while n < images and time.perf_counter() - t0 < budget_s:
        one(n % cfg["batch"])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "reference-native-ops",
            "sample": f"{n} images one after the other (the reference is batch-1): numpy decode, torch argsort, "
                      f"torchvision {torchvision.__version__} CPU nms + roi_{cfg['op']}, {dt:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import ref_port as O
    O.set_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: take the host's cores back
    cfg = WORKLOADS[args.workload]
    images = cfg["batch"]  # one full batch per step: OpenMP spreads images / RoIs over every host core
    loc, logits, feat = (t.numpy() for t in make_inputs(cfg, 7))
    for _ in range(args.warmup):
        cpu_step(cfg, O, loc, logits, feat, images)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_step(cfg, O, loc, logits, feat, images)
        times.append(time.perf_counter() - t0)
    dt = float(np.sum(times)) / args.steps
    v = images / dt
    sample = f"{images} of {cfg['batch']} images per step, {args.steps} steps after {args.warmup} warm-up steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v,
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args.workload, cfg, world),
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": O.max_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from the stream instead of replaying CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if args.steps > 40:  # the CPU arm's default: ~0.4 s per step on cfg2
            args.steps = 20
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
