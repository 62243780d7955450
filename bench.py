#!/usr/bin/env python
"""bench.py - CorrIFNet train throughput on B200 (contract: the task statement / DESIGN.md section 5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--only full|block]

MAIN LINE = BASELINE.json's metric, "CorrIFNet train imgs/s @256^2", on BASELINE configs[2]: one optimizer step over a
global batch of 64 synthetic DSTL-shaped 256x256 tiles = 8 micro-batches of 8 (the semantic micro-batch is never
split), through the drop-in ``mmvit4.MMVit4`` + ``corrif_b200.train.TrainStep`` (forward, fused
BCE/Jaccard tail, backward, bucketed gradient all-reduce overlapped with the backward, one-kernel Adam).  With N GPUs
rank r runs micro-batches r, r+N, ... of the SAME 8 (strong scaling: the step's work is fixed) and the 341 MB of
gradients are all-reduced over NCCL.  ``value`` = 64 / step time with the inputs resident in HBM; ``e2e`` = the same
step fed from pinned host buffers through corrif_b200.staging.PinnedPipeline with the loss read back every step.

Beside it, in the same JSON line:
  fusion_block   BASELINE configs[1]: the fusion hot path alone (mmvit4.py:456-529 fwd+bwd, batch 16, dropout 0.1):
                 imgs/s, kernel breakdown and the roofline of its dominant kernel family (``roofline``)
  eager_b200     stock-PyTorch eager on the same GPU (baseline/eager_mmvit4.py): the fusion block at batch 16 in fp32
                 and with allow_tf32, and the whole train step at micro-batch 8 - "the existing Blackwell path"
  metric_kernels BASELINE configs[3]: Jaccard / confusion-matrix kernels on 64 tiles of 256^2, bit-exact flag
  cpu_baseline   the reference's CPU path (its oracle port: /root/reference is not on the GPU box) on a bounded sample

``--impl reference`` times the CPU port of the same train step on the host cores (rank 0 only).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DROPIN = os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin")

METRIC = "CorrIFNet train imgs/s @256x256 (full train step: fwd + loss/metric + bwd + grad all-reduce + Adam)"
UNIT = "imgs/s"
MICRO_BATCH, MICRO_BATCHES = 8, 8          # BASELINE configs[2]: global batch 64 = 8 micro-batches of 8
TILE = 256
FWD_GFLOP_PER_SAMPLE = 24.495              # fusion block, SURVEY.md section 8a (GEMM FLOPs only)
STEP_GFLOP_PER_SAMPLE = 73.484             # fwd + dgrad + wgrad


def shared_config(args):
    """Identical in both arms (the driver compares them)."""
    return {"workload": "CorrIFNet (mmvit4) full train step, BASELINE configs[2]: global batch %d = %d micro-batches "
                        "of %d synthetic DSTL-shaped %dx%d tiles, Adam lr 1e-4, dropout 0.1"
                        % (MICRO_BATCH * MICRO_BATCHES, MICRO_BATCHES, MICRO_BATCH, TILE, TILE),
            "global_batch": MICRO_BATCH * MICRO_BATCHES, "micro_batch": MICRO_BATCH, "micro_batches_per_step": MICRO_BATCHES,
            "tile": "%dx%d" % (TILE, TILE), "dropout": 0.1, "optimizer": "Adam(lr=1e-4)", "parallelism": "dp%d" % args.gpus}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"],
                "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_tiles(batch, seed, tile=TILE):
    """SURVEY.md section 8d synthetic inputs: zero-centred tiles [B,3,3,T,T]; one binary 224^2 mask replicated per
    modality (the model output is fixed at 224^2, mmvit4.py:263)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, 3, tile, tile, generator=g)
    masks = (torch.rand(batch, 1, 1, 224, 224, generator=g) < 0.3).float().repeat(1, 3, 1, 1, 1)
    return images, masks


def make_block_inputs(batch, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    x6 = [torch.randn(batch, 64, 8, 8, 8, generator=g) for _ in range(3)]
    fused = torch.randn(batch, 192, 8, 8, 8, generator=g)
    gout = torch.randn(batch, 192, 8, 8, 8, generator=g)
    return x6, fused, gout


# ---------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference's train step (baseline, not target)
# ---------------------------------------------------------------------------------------------
def cpu_train_steps(sample_batch, steps, warmup, budget_s=None):
    """Times ``steps`` CPU train steps (after ``warmup``) of the full model on ``sample_batch`` tiles of 256^2.
    Returns (mean s/step, cores, steps actually timed)."""
    import torch
    from oracle import corrif_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    with open(os.path.join(ROOT, "tests", "golden", "mmvit4_state_dict_inventory.json")) as f:
        inv = json.load(f)
    state = {k: (v.requires_grad_(True) if v.is_floating_point() and not k.endswith(("running_mean", "running_var")) else v)
             for k, v in O.make_full_model_state(7, inv).items()}
    images, masks = make_tiles(sample_batch, 11)
    times, t_begin = [], time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step_cpu(state, images, masks, lr=1e-4, dropout_p=0.1)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and times and time.perf_counter() - t_begin + dt > budget_s:
            break
    return sum(times) / len(times), torch.get_num_threads(), len(times)


def run_reference(args):
    """--impl reference: the reference's CPU train step (oracle port) on this box's host cores, rank 0 only.
    Each step is a bounded sample of the workload: one micro-batch of 2 tiles (BASELINE configs[0]'s batch)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sample_b = 2
    t, cores, timed = cpu_train_steps(sample_b, args.steps, max(1, args.warmup), budget_s=float(os.environ.get("CORRIF_CPU_BUDGET_S", "270")))
    val = sample_b / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": timed, "warmup": max(1, args.warmup), "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "oracle port of MMVit4.forward + BCE + backward + Adam (mmvit4.py:441-532, "
                                   "F4_TRAIN.py:52-71), torch CPU fp32, one micro-batch of %d tiles of 256^2 per step, "
                                   "%d timed steps (the run is bounded to a few minutes)" % (sample_b, timed)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def _event_ms(fn, steps, sync):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    sync()
    return e0.elapsed_time(e1) / steps


def measure_tf32_peak(dev):
    """cuBLAS TF32 GEMM 8192^3 on this box (torch.matmul with allow_tf32): the measured counterpart of the
    "TF32 = bf16 / 2" denominator.  Best of 10 (burst) and the mean of a 1-second loop (sustained)."""
    import torch
    saved = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a, b = torch.randn(8192, 8192, device=dev), torch.randn(8192, 8192, device=dev)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            best = min(best, _event_ms(lambda i: a @ b, 1, torch.cuda.synchronize))
        n, t0 = 0, time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.perf_counter() - t0 < 1.0:
            for _ in range(10):
                a @ b
            n += 10
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        flop = 2.0 * 8192 ** 3
        return {"burst_tflops": flop / (best * 1e-3) / 1e12, "sustained_tflops": flop * n / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS), best of 10 / 1 s loop"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = saved


def bench_fusion_block(dev, steps, warmup, peaks):
    """BASELINE configs[1]: fusion block fwd+bwd, batch 16, dropout 0.1, inputs resident in HBM, replayed as two CUDA
    graphs; plus the per-kernel-family device times of one extra step (CUDA events around every launch)."""
    import torch
    from corrif_b200 import fusion, module, ops
    B = 16
    torch.manual_seed(1234)
    blk = module.CorrIFusionBlock(dropout_rate=0.1, precision="tf32").to(dev)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith("_pos"):
                p.normal_(0, 0.02)
    blk.train()
    names = fusion.param_names()
    named = dict(blk.named_parameters())
    params = {n: named[n].detach() for n in names}
    eng = fusion.FusionBlockEngine(params, dropout_p=0.1, precision="tf32", use_graphs=True)
    flat, grads = eng.new_grad_buffers()
    hx6, hfused, hgout = make_block_inputs(B, 100)
    dx6, dfused, dgout = [t.to(dev) for t in hx6], hfused.to(dev), hgout.to(dev)

    def step(i):
        eng.set_seed(1000 + i)
        flat.zero_()
        eng.forward(dx6, dfused)
        eng.backward(dgout, grads)

    for i in range(max(3, warmup)):
        step(i)
    l0 = ops.launch_count()
    ms = _event_ms(lambda i: step(100 + i), steps, torch.cuda.synchronize)
    launches = (ops.launch_count() - l0) // steps
    with ops.profile() as rec:
        step(999)
    summ = rec.summary()
    peak_tf32 = peaks["bf16_sustained"] / 2.0
    total_prof = sum(v[1] for v in summ.values())

    def family(prefix, label):
        fam = {k: v for k, v in summ.items() if k.startswith(prefix)}
        n_, ms_, fl_ = (sum(v[i] for v in fam.values()) for i in range(3))
        ach = fl_ / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0
        return {"bound": "tensor", "achieved": ach, "peak": peak_tf32, "unit": "TFLOP/s", "frac": ach / peak_tf32,
                "frac_of_burst_half": ach / (peaks["bf16_burst"] / 2.0),
                "traffic": None, "kernel": label % n_, "flops_per_step": fl_, "kernel_ms_per_step": ms_,
                "kernel_share_of_step": ms_ / total_prof if total_prof else None,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 = half the 16-bit rate), %s"
                               % peaks["source"]}

    roof_gemm = family("gemm_", "tcgen05 TF32 GEMM family (CTA-pair cta_group::2 kernel + small-N variants), all %d "
                                "launches of one fusion-block step at batch 16")
    roof_attn = family("attn_", "fused tcgen05 attention family (fwd + bwd), all %d launches of one fusion-block step "
                                "at batch 16; algorithmic FLOPs 4 (fwd) + 8 (bwd) x B*8*N^2*64, recomputation not counted")
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        roof_gemm["traffic"] = tj.get("gemm", {}).get("dram_bytes_per_launch")
        roof_attn["traffic"] = tj.get("attention", {}).get("dram_bytes_per_launch")
    tc = ("gemm", "attn")
    breakdown = {k: {"launches": v[0], "ms": round(v[1], 4),
                     ("tflops" if k.startswith(tc) else "gbs"):
                     round(v[2] / (v[1] * 1e-3) / (1e12 if k.startswith(tc) else 1e9), 2) if v[1] > 0 else 0.0}
                 for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1])}
    del eng
    return {"workload": "BASELINE configs[1]: fusion block (mmvit4.py:456-529) fwd+bwd, batch 16, dropout 0.1, tf32",
            "value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "launches_per_step": launches,
            "algorithmic_tflops": B * STEP_GFLOP_PER_SAMPLE / 1e3 / (ms * 1e-3),
            "kernel_breakdown": breakdown}, roof_gemm, roof_attn


def bench_eager(dev, steps):
    """Stock-PyTorch eager on the same B200 (baseline/eager_mmvit4.py): fusion block at batch 16 (fwd+bwd, dropout
    0.1) in strict fp32 and with allow_tf32; whole train step (fwd + BCE + bwd + torch Adam) at micro-batch 8."""
    import torch
    from baseline.eager_mmvit4 import EagerFusionBlock, EagerMMVit4
    out = {}
    hx6, hfused, hgout = make_block_inputs(16, 100)
    xs = [t.to(dev).requires_grad_(True) for t in hx6]
    fx, go = hfused.to(dev).requires_grad_(True), hgout.to(dev)
    torch.manual_seed(1234)
    blk = EagerFusionBlock(0.1).to(dev).train()

    def block_step(_):
        for p in blk.parameters():
            p.grad = None
        blk(xs, fx).backward(go)

    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for name, flag in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = flag
        torch.backends.cudnn.allow_tf32 = flag
        for i in range(3):
            block_step(i)
        ms = _event_ms(block_step, steps, torch.cuda.synchronize)
        out["fusion_block_b16_" + name] = {"ms_per_step": ms, "imgs_per_s": 16 / (ms * 1e-3)}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    del blk, xs, fx, go
    torch.cuda.empty_cache()
    # whole model, PyTorch defaults (cuDNN convs allow TF32, matmuls fp32): what `python F2_MAIN.py` of the reference runs
    torch.manual_seed(0)
    model = EagerMMVit4().to(dev).train()
    optim = torch.optim.Adam(model.parameters(), 1e-4)
    images, masks = (t.to(dev) for t in make_tiles(MICRO_BATCH, 5))

    def full_step(_):
        optim.zero_grad()
        loss = torch.nn.functional.binary_cross_entropy_with_logits(model(images), masks)
        loss.backward()
        optim.step()

    for i in range(3):
        full_step(i)
    ms = _event_ms(full_step, max(3, steps // 2), torch.cuda.synchronize)
    out["full_step_micro_batch_8"] = {"ms_per_step": ms, "imgs_per_s": MICRO_BATCH / (ms * 1e-3),
                                      "note": "one micro-batch of 8 per optimizer step (no accumulation), torch.optim.Adam, "
                                              "PyTorch default TF32 flags"}
    out["impl"] = "baseline/eager_mmvit4.py: nn.Linear / nn.Conv3d / nn.LayerNorm / F.interpolate eager, torch %s" % torch.__version__
    del model, optim
    torch.cuda.empty_cache()
    return out


def bench_metric_kernels(dev, peaks):
    """BASELINE configs[3]: 10-class confusion matrix / per-class Jaccard over 64 tiles of 256^2, bit-exact against a
    host bincount."""
    import numpy as np
    import torch
    from corrif_b200 import metrics
    g = torch.Generator().manual_seed(3)
    n = 64 * 256 * 256
    label = torch.randint(0, 10, (n,), generator=g, dtype=torch.uint8)
    label[label == 7] = 3                                           # one empty class: the inversion branch
    pred = torch.where(torch.rand(n, generator=g) < 0.7, label, torch.randint(0, 10, (n,), generator=g, dtype=torch.uint8))
    dl, dp = label.to(dev), pred.to(dev)
    y, yp = (dl == 3).float(), (dp == 3).float()
    res = {}
    for name, fn, nbytes in (("confusion_counts_u8", lambda: metrics.confusion_matrix(dl, dp, 10), 2 * n),
                             ("jaccard2_f32", lambda: metrics.Jaccard2(y, yp), 8 * n)):
        for _ in range(3):
            fn()
        ms = _event_ms(lambda i: fn(), 20, torch.cuda.synchronize)
        res[name] = {"us": ms * 1e3, "bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9,
                     "frac_of_hbm": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    cm = metrics.confusion_matrix(dl, dp, 10).cpu().numpy()
    host = np.bincount(label.numpy().astype(np.int64) * 10 + pred.numpy().astype(np.int64), minlength=100).reshape(10, 10)
    res["bit_exact_vs_host_bincount"] = bool((cm == host).all())
    res["pixels"] = n
    return res


def summarise_volume(details, peaks):
    """details: [(class, shape string, ms, algorithmic work)] of one micro-batch step.  Roofline of the convolution
    family: the tcgen05 line convolution (forward + data gradient of the 128^3 / 64^3 layers: HBM-bound, judged on its
    algorithmic bytes) beside the whole family (incl. the warp-level weight-gradient kernel), and a per-class table."""
    agg = {}
    for cls, det, ms, work in details:
        a = agg.setdefault(cls, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += work
    conv = {k: v for k, v in agg.items() if k in ("conv3d_fwd", "conv3d_dgrad", "conv3d_wgrad")}
    n_, ms_, fl_ = (sum(v[i] for v in conv.values()) for i in range(3))
    # algorithmic bytes of the convolution launches: the forward's shape string carries them; dgrad / wgrad move the
    # same two tensors (+ the tiny weights)
    byts = 0.0
    tc_n, tc_ms, tc_bytes, tc_flops = 0, 0.0, 0.0, 0.0
    for cls, det, ms, work in details:
        if cls == "conv3d_fwd" and "bytes=" in det:
            byts += 3.0 * float(det.split("bytes=")[1])
        if cls in ("conv3d_fwd", "conv3d_dgrad") and " tc " in det and "bytes=" in det:
            tc_n += 1; tc_ms += ms; tc_bytes += float(det.split("bytes=")[1]); tc_flops += work
    tc = ("gemm", "attn", "conv3d_fwd", "conv3d_dgrad", "conv3d_wgrad")
    table = {k: {"launches": v[0], "ms": round(v[1], 3),
                 ("tflops" if k.startswith(tc) and k != "conv3d_dgrad_border" else "gbs"):
                 round(v[2] / (v[1] * 1e-3) / (1e12 if k.startswith(tc) and k != "conv3d_dgrad_border" else 1e9), 1) if v[1] > 0 else 0.0}
             for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]}
    ach = fl_ / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0
    gbs = byts / (ms_ * 1e-3) / 1e9 if ms_ > 0 else 0.0
    tc_gbs = tc_bytes / (tc_ms * 1e-3) / 1e9 if tc_ms > 0 else 0.0
    peak_tf32 = peaks["bf16_sustained"] / 2.0
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    traffic = None
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("conv3d_tc", {}).get("dram_bytes_per_launch")
    roof = {"bound": "hbm", "achieved": tc_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": tc_gbs / peaks["hbm_gbs"], "traffic": traffic,
            "kernel": "conv3d_tc_kernel: tcgen05 line convolution, forward + data gradient of the 3x3x3 layers at 128^3 / "
                      "64^3 voxels (%d launches of one micro-batch of 8); achieved = algorithmic bytes (sources read once "
                      "+ result written once) / CUDA-event time" % tc_n,
            "algorithmic_bytes_per_step": tc_bytes, "flops_per_step": tc_flops, "kernel_ms_per_step": tc_ms,
            "tensor_tflops": tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0,
            "family": {"what": "every conv3d launch of the step: line convolution, warp-level forward / data gradient of the "
                               "small layers, warp-level weight gradient (%d launches)" % n_,
                       "ms": ms_, "tflops": ach, "frac_of_tf32_tcgen05_peak": ach / peak_tf32,
                       "gbs_algorithmic": gbs, "frac_of_hbm": gbs / peaks["hbm_gbs"],
                       "mma_sync_peak_tflops_measured": 270.0},
            "note": "these layers have 8-64 output channels: ~90 FLOP per HBM byte puts them on the HBM side of the "
                    "tcgen05 ridge; DESIGN.md section 4.4"}
    return roof, table


def run_ours(args):
    saved_stdout = os.dup(1)                 # keep stdout clean for the single JSON line (NCCL prints a banner)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sys.path.insert(0, DROPIN)
    import mmvit4                                    # the drop-in
    from corrif_b200 import ops, train
    from corrif_b200.staging import PinnedPipeline
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    line = {}
    if args.only in ("all", "full"):
        if MICRO_BATCHES % world:
            raise SystemExit("--gpus must divide %d micro-batches" % MICRO_BATCHES)
        torch.manual_seed(0)                         # identical initial weights on every rank
        model = mmvit4.MMVit4(num_cls=1).to(dev).train()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("_pos"):
                    p.normal_(0, 0.02)
        train.broadcast_module(model)
        optim = torch.optim.Adam(model.parameters(), 1e-4)
        stepper = train.TrainStep(model, optim, lim=224, graphs=not args.no_graphs)
        mine, total = train.shard_micro_batches(MICRO_BATCHES, MICRO_BATCHES, rank, world)[0]
        host = [tuple(t.pin_memory() for t in make_tiles(MICRO_BATCH, 100 + j)) for j in mine]
        resident = [(im.to(dev), ma.to(dev)) for im, ma in host]

        fus_spans = []
        from corrif_b200 import fusion
        for name in ("forward", "backward"):         # fusion-block share: CUDA events around the engine calls
            orig = getattr(fusion.FusionBlockEngine, name)

            def timed(self, *a, _orig=orig, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = _orig(self, *a, **k)
                e1.record()
                fus_spans.append((e0, e1))
                return r
            setattr(fusion.FusionBlockEngine, name, timed)

        def step(_):
            return stepper(resident, total_micro_batches=total)

        # W warm-up steps, and at least 10 micro-batches per rank before the clock starts: the fusion engine captures its
        # CUDA graphs on its third call after FlatAdam has re-pointed the parameters (first optimizer step), which with
        # one micro-batch per rank (N = 8) would otherwise land inside the timed region (measured: 108 vs 85 ms per step)
        warm = max(args.warmup, -(-10 // max(1, len(mine))))
        for i in range(warm):
            step(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        fus_spans.clear()
        l0 = ops.launch_count()
        ms = max_over_ranks(_event_ms(step, args.steps, barrier))
        launches = ops.launch_count() - l0
        fus_ms = sum(a.elapsed_time(b) for a, b in fus_spans) / args.steps

        # exposed (non-overlapped) gradient exchange: the step with the all-reduce hooks muted
        exposed = None
        if world > 1:
            saved_launch = stepper.buckets._launch
            stepper.buckets._launch = lambda b: b.__setitem__("launched", True)
            for i in range(2):
                step(i)
            ms_nocomm = max_over_ranks(_event_ms(step, max(3, args.steps // 4), barrier))
            stepper.buckets._launch = saved_launch
            train.broadcast_module(model)            # the muted steps let the ranks drift: re-synchronise
            exposed = ms - ms_nocomm

        # ---- the library's own kernels inside one micro-batch step (CUDA events per launch, stream launches)
        vol_prof = None
        if world == 1:
            saved_graphs = stepper.graphs
            stepper.graphs = False                       # per-launch events need real launches
            fwd_graphed = model.forward
            if stepper._eager_forward is not None:
                model.forward = stepper._eager_forward
            saved_streams, saved_ws = mmvit4._ENC_STREAMS, os.environ.get("CORRIF_WGRAD_STREAM")
            mmvit4._ENC_STREAMS = False                  # one stream: a launch's events must not span its neighbours' work
            os.environ["CORRIF_WGRAD_STREAM"] = "0"
            with ops.profile() as rec:
                stepper([resident[0]], total_micro_batches=1)
            vol_prof = rec.details()
            mmvit4._ENC_STREAMS = saved_streams
            if saved_ws is None:
                os.environ.pop("CORRIF_WGRAD_STREAM", None)
            else:
                os.environ["CORRIF_WGRAD_STREAM"] = saved_ws
            model.forward, stepper.graphs = fwd_graphed, saved_graphs

        # ---- end to end: every step's inputs from pinned host memory, loss read back every step
        pipe = PinnedPipeline(dev)
        hloss = torch.zeros(1).pin_memory()
        bytes_in = sum(im.numel() * 4 + ma.numel() * 4 for im, ma in host)

        def e2e_step(_):
            mbs = []
            pipe.prefetch(list(host[0]))
            for j in range(len(host)):
                im, ma = pipe.get()
                mbs.append((im.clone(), ma.clone()) if len(host) > 1 else (im, ma))
                if len(host) > 1:
                    pipe.release()
                if j + 1 < len(host):
                    pipe.prefetch(list(host[j + 1]))
            out = stepper(mbs, total_micro_batches=total)
            if len(host) == 1:
                pipe.release()
            pipe.put(out["loss"].reshape(1), hloss)
            torch.cuda.current_stream().wait_stream(pipe.copy_stream)        # the D2H read is part of the step

        for i in range(2):
            e2e_step(i)
        e2e_ms = max_over_ranks(_event_ms(e2e_step, args.steps, barrier))
        clocks = sampler.stop() if rank == 0 else None
        loss_val = float(hloss[0])
        grad_bytes = stepper.buckets.grad_bytes()
        peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
        imgs = MICRO_BATCH * MICRO_BATCHES
        line.update({
            "metric": METRIC, "value": imgs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic", "config": shared_config(args),
            "details": {
                "local_micro_batches": len(mine), "warmup_steps_run": warm,
                "cuda_graphs": "model forward + backward replayed as two CUDA graphs per micro-batch (TrainStep(graphs=True), "
                               "captured after the first step)" if stepper._graph_shape is not None else "stream launches", "fusion_block_ms_per_step": fus_ms, "fusion_block_share": fus_ms / ms,
                "grad_allreduce_bytes": grad_bytes if world > 1 else 0,
                "exposed_allreduce_ms": exposed, "peak_mem_GiB": peak_mem, "loss_last_step": loss_val,
                "l2": "activations of one micro-batch ~25 GB >> 126 MB L2 (no explicit flush needed)",
                "model": "dropin/mmvit4.MMVit4: fusion block, loss/Jaccard tail and Adam on libcorrif_b200 kernels; see "
                         "DESIGN.md section 4 for which of the encoder / early-fusion / decoder blocks run on them"},
            "clocks": clocks,
            "e2e": {"value": imgs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 4,
                    "api": "dropin mmvit4.MMVit4 + corrif_b200.train.TrainStep fed by corrif_b200.staging.PinnedPipeline "
                           "(what dropin/F4_TRAIN.train_model runs per step); loss read back to pinned host memory"},
            "gpu_launches": launches,
        })
        if vol_prof is not None:
            line["roofline_volume"], line["kernel_breakdown_full_step"] = summarise_volume(vol_prof, peaks)
            if fus_ms == 0.0:
                # under whole-model graphs the engine's Python entry points are not called in the timed region: take the
                # fusion block's kernels from the per-launch table of the profiled micro-batch instead
                fus_cls = ("gemm_tf32/linear", "gemm_tf32/dgrad", "gemm_tf32/wgrad", "attn_", "layernorm", "inter_corr",
                           "colsum", "dropout_colsum", "batchsum", "add_rows")
                per_mb = sum(ms_ for cls, _, ms_, _ in vol_prof if cls.startswith(fus_cls))
                line["details"]["fusion_block_ms_per_step"] = per_mb * len(mine)
                line["details"]["fusion_block_share"] = per_mb * len(mine) / ms
                line["details"]["fusion_block_share_how"] = ("sum of the fusion block's kernel times (CUDA events, one profiled "
                                                             "micro-batch on stream launches) x micro-batches / step time")
        del stepper, optim, model, resident
        torch.cuda.empty_cache()
    barrier()
    if rank == 0 and args.only in ("all", "block"):
        fb, roof_gemm, roof_attn = bench_fusion_block(dev, args.steps, args.warmup, peaks)
        tf32 = measure_tf32_peak(dev)
        for r in (roof_gemm, roof_attn):
            r["tf32_cublas_measured"] = tf32
            r["frac_of_measured_tf32_cublas_sustained"] = r["achieved"] / tf32["sustained_tflops"]
            r["frac_of_measured_tf32_cublas_burst"] = r["achieved"] / tf32["burst_tflops"]
        dominant, other, other_key = ((roof_attn, roof_gemm, "roofline_gemm")
                                      if roof_attn["kernel_ms_per_step"] >= roof_gemm["kernel_ms_per_step"]
                                      else (roof_gemm, roof_attn, "roofline_attention"))
        line.update({"fusion_block": fb, "roofline": dominant, other_key: other})
        if args.only == "block":
            line.update({"metric": "CorrIFNet fusion-block train imgs/s (configs[1])", "value": fb["value"], "unit": UNIT,
                         "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": fb["ms_per_step"],
                         "higher_is_better": True, "dtype": "tf32", "data": "synthetic"})
    if rank == 0 and args.only == "all":
        try:
            line["eager_b200"] = bench_eager(dev, min(args.steps, 10))
        except Exception as e:                       # a baseline must never take the product's line down
            line["eager_b200"] = {"error": repr(e)[:300]}
        line["metric_kernels"] = bench_metric_kernels(dev, peaks)
        if world == 1:
            t, cores, timed = cpu_train_steps(2, 2, 1)
            line["cpu_baseline"] = {"value": 2 / t, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "oracle port of the reference's train step (MMVit4.forward + BCE + backward "
                                              "+ Adam, torch CPU fp32), one micro-batch of 2 tiles of 256^2, %d timed steps "
                                              "after 1 warm-up, %.2f s/step" % (timed, t)}
    barrier()
    if rank == 0:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graphs", action="store_true", help="stream launches instead of whole-model CUDA graphs")
    ap.add_argument("--only", default="all", choices=["all", "full", "block"],
                    help="full: the train step only; block: the fusion-block microbench only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
