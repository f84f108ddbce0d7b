#!/usr/bin/env python
"""bench.py -- the proposal-and-RoI hot path on synthetic inputs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A step = one pass of the hot path over one batch: RPN proposals (decode, clip, min-size, top-k, NMS,
pad/gather) for B images, then the RoI head's coordinate map + RoIPool gather of every proposal.
Default workload = BASELINE.json configs[1]: B=16 600x600 images, ResNet-50 stride-16 features
[16,1024,38,38], 3000 pre-NMS / 300 post-NMS RoIs per image, RoIPool 14x14.  The backbone is not part of
the path: features / RPN conv outputs are synthetic tensors of the right shape (data: synthetic).

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same
through the module API (ProposalCreator.batched + HarNetRoIHead.forward) from pinned host buffers with
the host<->device copies inside the timed region; `roofline` = the RoIPool gather kernel against the
measured HBM copy peak; `cpu_baseline` = the oracle port on the host cores on a bounded sample.
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


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every ~2 ms; nvidia-smi is too
    slow for a 50 ms region and is only the fallback)."""
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

    def __enter__(self):
        self._t.start()
        self._ready.wait(timeout=5.0)  # NVML opened and sampling before the timed region starts
        self.sm.clear()  # idle-clock samples taken while waiting do not belong to the region
        return self

    def __exit__(self, *a):
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


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from two_stage_object_detection_b200 import functional as F
    from two_stage_object_detection_b200.nets import HarNetRoIHead, ProposalCreator
    from two_stage_object_detection_b200.nets.frcnn import GlobalAvgClassifier

    cfg = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, H, W, C, P = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"]
    S, n_post = cfg["img"], cfg["n_post"]
    N = H * W * 9
    K = B * n_post
    base = F.base_anchors(device=dev)
    idx = torch.arange(B, dtype=torch.int32, device=dev)
    # three input sets rotated between steps (3 x ~100 MB > L2) and a 3.85 GB output per step: no step
    # finds its inputs in L2
    sets = [make_inputs(cfg, 1000 * rank + s, device=dev) for s in range(3)]
    pooled = torch.empty((K, C, P, P), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * B, n_post, 4), dtype=torch.float32, device=dev) if world > 1 else None
    pkw = dict(clip_x_max=S, clip_y_max=S, n_pre_nms=cfg["n_pre"], n_post_nms=n_post, nms_iou=0.7, min_size=16.0,
               base=base, feat_stride=16, feat_hw=(H, W), score_is_logits=True)
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + args.warmup)] for k in
          ("p0", "p1", "r0", "r1")}
    train = bool(cfg.get("train"))
    if train:  # G boxes per image, centre U(0,S)^2, w,h U(50,250), clipped; labels in [0,20)
        gg = torch.Generator().manual_seed(77 + rank)
        G = cfg["n_gt"]
        ctr = torch.rand(B, G, 2, generator=gg) * S
        wh = 50 + torch.rand(B, G, 2, generator=gg) * 200
        gt_box = torch.cat([ctr - wh / 2, ctr + wh / 2], -1).clamp(0, S).to(dev)
        gt_lab = torch.randint(0, 20, (B, G), generator=gg).to(dev)
        n_gt = torch.full((B,), G, dtype=torch.int32, device=dev)
        n_roi = 128
        K = B * n_roi
        pooled = torch.empty((K, C, P, P), dtype=torch.float32, device=dev)

    def step(i):
        loc, logits, feat = sets[i % 3]
        ev["p0"][i].record()
        rois, src, n_keep, status = F.proposals(loc, logits, **pkw)
        if train:
            F.anchor_targets(gt_box, n_gt, base=base, feat_stride=16, feat_hw=(H, W))
            sample, _, _, _, _ = F.proposal_targets(rois, gt_box, gt_lab, n_gt)
            ev["p1"][i].record()
            work = dist.all_gather_into_tensor(gathered, rois, async_op=True) if world > 1 else None
            rois5 = F.roi_head_coords(sample, idx, (S, S), (H, W))
            ev["r0"][i].record()
            F.roi_pool_forward(feat, rois5, P, 1.0, with_argmax=True, out=pooled, rois_per_image=n_roi)
            ev["r1"][i].record()
            if work is not None:
                work.wait()
            return rois, status
        ev["p1"][i].record()
        # the detections are final once the proposal layer is done: their all-gather (the path's only
        # collective) runs on NCCL's stream under the RoI gather; the step ends when both have finished
        work = dist.all_gather_into_tensor(gathered, rois, async_op=True) if world > 1 else None
        rois5 = F.roi_head_coords(rois, idx, (S, S), (H, W))
        ev["r0"][i].record()
        if cfg["op"] == "pool":
            F.roi_pool_forward(feat, rois5, P, 1.0, out=pooled, rois_per_image=n_post)
        else:
            F.roi_align_forward(feat, rois5, P, 1.0, 2, False, out=pooled, rois_per_image=n_post)
        ev["r1"][i].record()
        if work is not None:
            work.wait()  # stream-side wait: the compute stream depends on the collective, the host does not block
        return rois, status

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        t0.record()
        for i in range(args.warmup, args.warmup + args.steps):
            rois, status = step(i)
        t1.record()
        barrier()
    ms = t0.elapsed_time(t1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    assert not status.cpu().numpy().any(), "proposal layer flagged an index error"
    ms_step = ms / args.steps
    sl = slice(args.warmup, args.warmup + args.steps)
    roi_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev["r0"][sl], ev["r1"][sl])]))
    prop_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev["p0"][sl], ev["p1"][sl])]))

    # ---- the same step replayed from CUDA graphs (one per input set; informational) -------------------
    graph_ms = None
    if world == 1 and not train:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    step(0)
            torch.cuda.current_stream().wait_stream(side)
            graphs = []
            for i in range(3):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    loc, logits, feat = sets[i]
                    rois_g, _, _, _ = F.proposals(loc, logits, **pkw)
                    rois5_g = F.roi_head_coords(rois_g, idx, (S, S), (H, W))
                    if cfg["op"] == "pool":
                        F.roi_pool_forward(feat, rois5_g, P, 1.0, out=pooled, rois_per_image=n_post)
                    else:
                        F.roi_align_forward(feat, rois5_g, P, 1.0, 2, False, out=pooled, rois_per_image=n_post)
                graphs.append(gr)
            for i in range(3):
                graphs[i].replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for i in range(args.steps):
                graphs[i % 3].replay()
            g1.record()
            torch.cuda.synchronize()
            graph_ms = g0.elapsed_time(g1) / args.steps
        except Exception as exc:  # graphs are an extra, never the reported value
            graph_ms = f"unavailable: {exc}"

    # ---- the head's gather when its classifier is the global average (HarDNet): fused kernel, [K,C] out ------
    fused_ms = None
    if not train:
        loc, logits, feat = sets[0]
        rois_f, _, _, _ = F.proposals(loc, logits, **pkw)
        rois5_f = F.roi_head_coords(rois_f, idx, (S, S), (H, W))
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
    creator = ProposalCreator("test", n_test_pre_nms=cfg["n_pre"], n_test_post_nms=n_post)
    head = HarNetRoIHead(n_class=21, roi_size=P, spatial_scale=1, classifier=GlobalAvgClassifier(),
                         in_features=C, roi_op=cfg["op"], sampling_ratio=2).to(dev).eval()
    host = [make_inputs(cfg, 1000 * rank + 10 + s, pin=True) for s in range(2)]
    h2d = sum(t.numel() * 4 for t in host[0])
    d2h = 0

    def e2e_compute(loc, logits, feat):
        rois, _, _, st = creator.batched(loc, logits, (3, S, S), 1.0, base=base, feat_stride=16, feat_hw=(H, W),
                                         score_is_logits=True)
        with torch.no_grad():
            cls_locs, scores = head(feat, rois, None, (S, S))
        return [rois, cls_locs, scores, st]

    def e2e_step(i):  # serial form: copy in, compute, copy out, host waits
        nonlocal d2h
        outs = [o.cpu() for o in e2e_compute(*(t.to(dev, non_blocking=True) for t in host[i % 2]))]
        d2h = sum(o.numel() * o.element_size() for o in outs)
        return outs

    e2e_steps = max(3, min(args.steps, 20))
    for i in range(3):
        e2e_step(i)
    barrier()
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
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
    torch.cuda.synchronize()

    def e2e_pipe(n):
        seen = 0.0
        for i in range(n):
            s = i % 2
            with torch.cuda.stream(copy_s):
                if i >= 2:
                    copy_s.wait_event(in_free[s])
                for d, h in zip(dev_in[s], host[i % 2]):
                    d.copy_(h, non_blocking=True)
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
    barrier()
    w0 = time.perf_counter()
    e2e_pipe(e2e_steps)
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3 / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_ms, e2e_serial_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_serial_ms = (float(v) for v in tt.tolist())

    # ---- roofline of the dominant kernel (RoI gather) -----------------------------------------------
    peak, peak_src = peaks()
    alg_bytes = K * C * P * P * 4 * (2 if train else 1) + K * 20 + B * C * H * W * 4  # + int32 argmax when training
    achieved = alg_bytes / (roi_ms * 1e-3) / 1e9
    rows = min(cfg["n_pre"], N)
    # NMS super-block schedule of csrc/proposals.cu (run_nms_sorted): first block ~2*n_post, then doubling
    need = -(-rows // 256) * 256
    s0 = min(max(-(-2 * n_post // 256) * 256, 256), 2048, need)
    smax = max(s0, min(2048, need))
    n_sb, c0, ln = 0, 0, s0
    while c0 < rows:
        n_sb, c0, ln = n_sb + 1, c0 + ln, min(2 * ln, smax)
    launches = 1 + 1 + n_sb + 1 + 1 + 1  # decode, topk, nms block kernel per super-block, finalize, coords, gather
    if train:
        launches += 3  # anchor_iou, anchor_label, proposal_target
    # <P, threads, channels per CTA, CTAs per SM, argmax, table levels[, bins per thread, mbarrier hand-off]> as
    # csrc/roi_ops.cu picks them
    kernel_name = "roi_pool_tab_kernel<14,392,4,2,false,2,2,true>" if (cfg["op"], P) == ("pool", 14) else (
        ("roi_pool_tab_kernel<7,392,4,2,true,1>" if train else
         ("roi_pool_tab_kernel<7,784,4,1,false,2,1,true,true>" if H * W > 3000 else "roi_pool_tab_kernel<7,392,4,2,false,2,1,true>"))
        if cfg["op"] == "pool" else "roi_align_tab_kernel<7,2,392,2>")
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from one ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get("dram_bytes")
    in_mb = sum(t.numel() * 4 for t in sets[0]) / 1e6
    out_mb = pooled.numel() * 4 * (2 if train else 1) / 1e6
    out = {
        "metric": "images/sec (RPN proposals + RoI gather hot path)", "value": world * B / (ms_step * 1e-3),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfg['desc']}", "batch_per_gpu": B, "global_batch": world * B,
                   "anchors_per_image": N, "n_pre_nms": cfg["n_pre"], "n_post_nms": n_post,
                   "feature": [B, C, H, W], "roi_op": f"{cfg['op']} {P}x{P}",
                   "l2": f"3 input sets rotated (3 x {in_mb:.0f} MB) + {out_mb:.0f} MB of gather output written per "
                         "step: a step never finds its inputs in L2 (126 MB)",
                   "parallelism": f"dp{world} (images sharded per GPU; NCCL all_gather of the rois inside the step, overlapped with the RoI gather, when N>1)"},
        "proposals_per_sec": world * K / (ms_step * 1e-3),
        "breakdown_ms": {"proposals": prop_ms, "roi_gather": roi_ms},
        "fused_head_gather_ms": fused_ms,  # RoI gather + global-average classifier in one kernel (informational)
        "cuda_graph_ms_per_step": graph_ms,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "serial_ms_per_step": e2e_serial_ms,
                "api": "ProposalCreator.batched + HarNetRoIHead.forward (RoIPool + global-average classifier fused "
                       "into one kernel, Linear heads) from pinned host buffers; copies of step i+1 overlap the "
                       "compute of step i (copy stream + compute stream, double-buffered); PCIe-bound: "
                       "h2d_bytes_per_step / ms_per_step is the host link's measured ~55 GB/s"},
        "gpu_launches": launches * args.steps,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": roi_ms},
        "clocks": clk.summary(),
    }
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(cfg, images=cfg["batch"])
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


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


def cpu_baseline(cfg, images=2, reps=1):
    from oracle import ref_port as O
    O.set_threads(len(os.sched_getaffinity(0)))
    loc, logits, feat = (t.numpy() for t in make_inputs(cfg, 7))
    images = min(images, cfg["batch"])
    cpu_step(cfg, O, loc, logits, feat, 2)  # warm-up (builds / loads the C library)
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_step(cfg, O, loc, logits, feat, images)
    dt = (time.perf_counter() - t0) / reps
    return {"value": images / dt, "unit": "images/s", "cores": O.max_threads(), "kind": "port",
            "sample": f"{images} of {cfg['batch']} images of the same workload (oracle/ref_port.py + frcnn_oracle.c, "
                      f"OpenMP over images and RoIs), {dt:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_port as O
    O.set_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: take the host's cores back
    cfg = WORKLOADS[args.workload]
    images = cfg["batch"]  # one full batch per step: OpenMP spreads images / RoIs over every host core
    loc, logits, feat = (t.numpy() for t in make_inputs(cfg, 7))
    for _ in range(min(args.warmup, 1) or 1):
        cpu_step(cfg, O, loc, logits, feat, images)
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(cfg, O, loc, logits, feat, images)
    dt = (time.perf_counter() - t0) / steps
    v = images / dt
    sample = f"{images} of {cfg['batch']} images per step, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "images/sec (RPN proposals + RoI gather hot path)", "value": v,
        "unit": "images/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": 1,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"{args.workload}: {cfg['desc']}", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": O.max_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
