"""Profiling driver: a few warm-up steps of the bench workload, then exactly one step between
cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).  `--gemm M N K [a_mn b_mn]` runs one
stand-alone GEMM shape instead (for the `ncu --set full` capture of the dominant kernel)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import fusion, module, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--gemm", type=int, nargs="*", default=None)
ap.add_argument("--attn", type=int, nargs=2, default=None, help="B N: fused attention fwd+bwd only")
ap.add_argument("--pdrop", type=float, default=0.1)
ap.add_argument("--epi", type=int, default=0, help="--gemm: epilogue mode (include/corrif.h)")
ap.add_argument("--gdrop", type=float, default=0.0, help="--gemm: fused epilogue dropout p")
args = ap.parse_args()
dev = torch.device("cuda:0")

if args.gemm:
    M, N, K = args.gemm[:3]
    a_mn, b_mn = (bool(args.gemm[3]), bool(args.gemm[4])) if len(args.gemm) >= 5 else (False, False)
    A = torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)
    B = torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)
    D = torch.empty(M, N, device=dev)
    kw = dict(M=M, N=N, K=K, lda=M if a_mn else K, ldb=N if b_mn else K, ldd=N, a_mn=a_mn, b_mn=b_mn)
    if args.epi:
        kw.update(epilogue=args.epi, bias=torch.randn(N, device=dev), residual=torch.randn(M, N, device=dev), ldr=N,
                  aux=torch.randn(M, N, device=dev), ldaux=N)
        if args.gdrop > 0:
            kw.update(drop_p=args.gdrop, drop_sites=(1, 2 if args.epi == 3 else ops.NO_SITE), drop_seed=5)
    for _ in range(3):
        ops.gemm(A, B, D, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(3):
        ops.gemm(A, B, D, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

if args.attn:
    B, N = args.attn
    p = args.pdrop
    qkv = torch.randn(B * N, 1536, device=dev)
    dO = torch.randn(B * N, 512, device=dev)
    O = torch.empty(B * N, 512, device=dev)
    lse = torch.empty(B * 8, N, device=dev)
    delta = torch.empty(B * 8, N, device=dev)
    bits = torch.zeros(B * 8, N, N // 32, dtype=torch.int32, device=dev) if p > 0 else None
    dqkv = torch.empty(B * N, 1536, device=dev)
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        ops.attention_fwd(qkv, O, lse, bits, B, N, 8, 64, 0.125, p, seed=1, site=0)
        ops.attention_bwd(qkv, O, dO, lse, bits, delta, dqkv, B, N, 8, 64, 0.125, p)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

torch.manual_seed(0)
blk = module.CorrIFusionBlock(dropout_rate=args.dropout).to(dev)
named = dict(blk.named_parameters())
params = {n: named[n].detach() for n in fusion.param_names()}
eng = fusion.FusionBlockEngine(params, dropout_p=args.dropout, precision="tf32")
B = args.batch
x6 = [torch.randn(B, 64, 8, 8, 8, device=dev) for _ in range(3)]
fused, gout = torch.randn(B, 192, 8, 8, 8, device=dev), torch.randn(B, 192, 8, 8, 8, device=dev)
for i in range(2):
    eng.seed = i
    eng.forward(x6, fused)
    eng.backward(gout)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.seed = 7
eng.forward(x6, fused)
eng.backward(gout)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, launches so far:", ops.launch_count())
