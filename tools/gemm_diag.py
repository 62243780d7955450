"""GPU diagnostic: run the tcgen05 GEMM in every layout on small shapes, print error structure.
Used during bring-up (python tools/gemm_diag.py > gpurun_out/gemm_diag.log)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def run(M, N, K, a_mn, b_mn, prec=ops.GEMM_TF32, ints=False):
    g = torch.Generator().manual_seed(1)
    if ints:   # small integers: exactly representable in TF32 -> any layout bug shows as O(1) error
        A = torch.randint(-3, 4, (M, K), generator=g).float().to(dev)
        B = torch.randint(-3, 4, (N, K), generator=g).float().to(dev)
    else:
        A = torch.randn(M, K, generator=g).to(dev)
        B = torch.randn(N, K, generator=g).to(dev)
    Am = A.t().contiguous() if a_mn else A
    Bm = B.t().contiguous() if b_mn else B
    D = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(Am, Bm, D, M=M, N=N, K=K, lda=M if a_mn else K, ldb=N if b_mn else K, ldd=N, a_mn=a_mn,
             b_mn=b_mn, precision=prec)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (D.double() - ref)
    rel = float(err.norm() / ref.norm())
    signed = float((err * ref.sign()).sum() / ref.abs().sum())
    nan = int(torch.isnan(D).sum())
    print(f"M{M:5d} N{N:5d} K{K:5d} a_mn={int(a_mn)} b_mn={int(b_mn)} ints={int(ints)}  rel {rel:.3e}  "
          f"signed-bias {signed:+.3e}  nan {nan}")
    if rel > 5e-3 and not nan:
        bad = (err.abs() > 1e-2 * ref.abs().mean()).float()
        rows = bad.mean(1)
        cols = bad.mean(0)
        print("   bad rows frac by 32-row group:", [round(float(rows[i:i + 32].mean()), 2) for i in range(0, min(M, 256), 32)])
        print("   bad cols frac by 32-col group:", [round(float(cols[i:i + 32].mean()), 2) for i in range(0, min(N, 256), 32)])
        print("   D[0,:8]  ", D[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].float().tolist())
    return rel


if __name__ == "__main__":
    ops.check_device()
    for ints in (True, False):
        for (a_mn, b_mn) in ((False, False), (False, True), (True, False), (True, True)):
            for (M, N, K) in ((128, 128, 32), (128, 128, 64), (128, 64, 128), (256, 256, 256), (512, 192, 512)):
                try:
                    run(M, N, K, a_mn, b_mn, ints=ints)
                except Exception as e:  # noqa: BLE001
                    print("EXC", M, N, K, a_mn, b_mn, repr(e))
    # fp32 checking mode sanity
    run(256, 256, 256, False, False, prec=ops.GEMM_FP32)
    run(256, 256, 256, True, True, prec=ops.GEMM_FP32)
