#!/usr/bin/env python
"""Is the dense GEMM bound by its epilogue?  Same MMA work, different epilogues (STORE / SWIGLU write the tile row by row
from registers, LMHEAD writes two floats per row), next to torch.matmul (cuBLAS) on the same shape.  One JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from spacer_b200 import ops  # noqa: E402
from spacer_b200.ops import EPI_LMHEAD, EPI_STORE, EPI_SWIGLU  # noqa: E402


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    shapes = [(4096, 37888, 3584, "gate|up fwd"), (4096, 3584, 18944, "down fwd"), (4096, 4608, 3584, "qkv fwd"),
              (8192, 3840, 1280, "ViT qkv"), (8192, 5120, 1280, "ViT fc1"), (4096, 152064, 3584, "lm_head")]
    for M, N, K, what in shapes:
        a = (torch.randn(M, K, device=dev, generator=g) * 0.1).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) * 0.02).bfloat16()
        fl = 2.0 * M * N * K
        res = {"shape": [M, N, K], "what": what}
        if N * M * 2 < (3 << 30):
            out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
            res["store_tflops"] = round(fl / timed(lambda: ops.gemm(a, w, out=out, epilogue=EPI_STORE)) / 1e9, 1)
            res["cublas_tflops"] = round(fl / timed(lambda: torch.matmul(a, w.t(), out=out)) / 1e9, 1)
        if N % 256 == 0 and N * M * 2 < (3 << 30):
            o2 = torch.empty((M, N // 2), device=dev, dtype=torch.bfloat16)
            aux = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
            res["swiglu_aux_tflops"] = round(fl / timed(lambda: ops.gemm(a, w, out=o2, aux=aux, epilogue=EPI_SWIGLU)) / 1e9, 1)
            res["swiglu_noaux_tflops"] = round(fl / timed(lambda: ops.gemm(a, w, out=o2, epilogue=EPI_SWIGLU)) / 1e9, 1)
        nt = (N + 255) // 256
        part = torch.empty((M, nt, 2), device=dev, dtype=torch.float32)
        tl = torch.zeros(M, device=dev, dtype=torch.float32)
        tgt = torch.randint(0, N, (M,), device=dev, dtype=torch.int32)
        res["lmhead_epilogue_tflops"] = round(fl / timed(lambda: ops.gemm(a, w, epilogue=EPI_LMHEAD, targets=tgt, lse_part=part, tgt_logit=tl)) / 1e9, 1)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
