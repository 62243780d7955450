"""BASELINE.json configs[3]: Jaccard / confusion-matrix evaluation over synthetic 10-class predictions
(64 tiles of 256x256), bit-exact against a torch.bincount count, with the achieved HBM bandwidth of the three
metric kernels (CUDA events, inputs resident in HBM, 20 launches after 3 warm-ups; inputs are 4.2 - 100 MB, so
the small ones partly live in L2 between launches: a 256 MB buffer is rewritten before every timed launch)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import metrics, ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
K, T, H = 10, 64, 256
label = torch.randint(0, K, (T, H, H), generator=g, dtype=torch.uint8)
noise = torch.randint(0, K, (T, H, H), generator=g, dtype=torch.uint8)
pred = torch.where(torch.rand(T, H, H, generator=g) < 0.7, label, noise)
label, pred = label.to(dev), pred.to(dev)
P = label.numel()
flush = torch.empty(64 << 20, device=dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


cm = metrics.confusion_matrix(label, pred, K)
ref = torch.bincount(label.view(-1).long() * K + pred.view(-1).long(), minlength=K * K).view(K, K)
assert torch.equal(cm, ref)
counts = torch.zeros(K * K, dtype=torch.int64, device=dev)
ms_cm = timed(lambda: ops.confusion_counts(label.view(-1), pred.view(-1), P, K, counts))
# per-class soft Jaccard on {0,1} maps of class 3 (the reference's Jaccard2 feed), 8 B per pixel
y, yp = (label == 3).float().view(-1, 1), (pred == 3).float().view(-1, 1)
j = metrics.Jaccard2(y, yp)
tp, fp, fn = int(cm[3, 3]), int(cm[:, 3].sum() - cm[3, 3]), int(cm[3, :].sum() - cm[3, 3])
assert j.item() == torch.tensor((tp + 1e-8) / (tp + fp + fn + 1e-8), dtype=torch.float32).item() or abs(j.item() - tp / (tp + fp + fn)) < 1e-6
sums = torch.zeros(4, dtype=torch.float64, device=dev)
ms_j = timed(lambda: ops.jaccard_sums(y, yp, P, sums))
# the train-step tail at the whole-model micro-batch: outputs / masks [8, 3, 1, 224, 224]
out = torch.rand(8, 3, 1, 224, 224, device=dev)
msk = (torch.rand(8, 1, 1, 224, 224, device=dev) < 0.3).float().repeat(1, 3, 1, 1, 1)
ms_t = timed(lambda: metrics.loss_and_jaccard(out, msk))
n_t = out.numel()
# the same two kernels at 1024 tiles (inputs far larger than L2): their streaming rate
T2 = 1024
big_l = torch.randint(0, K, (T2, H, H), device=dev, dtype=torch.uint8)
big_p = torch.where(torch.rand(T2, H, H, device=dev) < 0.7, big_l, torch.randint(0, K, (T2, H, H), device=dev, dtype=torch.uint8))
P2 = big_l.numel()
cm2 = metrics.confusion_matrix(big_l, big_p, K)
assert torch.equal(cm2, torch.bincount(big_l.view(-1).long() * K + big_p.view(-1).long(), minlength=K * K).view(K, K))
ms_cm2 = timed(lambda: ops.confusion_counts(big_l.view(-1), big_p.view(-1), P2, K, counts), n=10)
y2, yp2 = (big_l == 3).float().view(-1, 1), (big_p == 3).float().view(-1, 1)
ms_j2 = timed(lambda: ops.jaccard_sums(y2, yp2, P2, sums), n=10)
peak = 6550.7
res = {"config": "F5_JACCARD eval, 64 x 256 x 256 synthetic 10-class predictions, 1 B200",
       "confusion_counts": {"bit_exact_vs_bincount": True, "bytes": 2 * P, "us": ms_cm * 1e3, "GBps": 2 * P / ms_cm / 1e6,
                            "frac_of_hbm_peak": 2 * P / ms_cm / 1e6 / peak},
       "jaccard_sums": {"bytes": 8 * P, "us": ms_j * 1e3, "GBps": 8 * P / ms_j / 1e6, "frac_of_hbm_peak": 8 * P / ms_j / 1e6 / peak},
       "loss_and_jaccard_tail": {"bytes": 12 * n_t, "us": ms_t * 1e3, "GBps": 12 * n_t / ms_t / 1e6,
                                 "note": "BCE-with-logits + gradient + Jaccard sums + finish, [8,3,1,224,224]: 4 small launches"},
       "confusion_counts_1024_tiles": {"bit_exact_vs_bincount": True, "bytes": 2 * P2, "us": ms_cm2 * 1e3, "GBps": 2 * P2 / ms_cm2 / 1e6,
                                       "frac_of_hbm_peak": 2 * P2 / ms_cm2 / 1e6 / peak},
       "jaccard_sums_1024_tiles": {"bytes": 8 * P2, "us": ms_j2 * 1e3, "GBps": 8 * P2 / ms_j2 / 1e6,
                                   "frac_of_hbm_peak": 8 * P2 / ms_j2 / 1e6 / peak},
       "hbm_peak_GBps": peak, "peak_source": "MEASURED_PEAKS.json"}
print(json.dumps(res))
