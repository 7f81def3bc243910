#!/usr/bin/env python
"""Run one GEMM shape a few times (ncu target).  usage: gemm_one.py M N K [bf16]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykaldi2_b200.models import lstm as L
M, N, K = (int(x) for x in sys.argv[1:4])
bf16 = len(sys.argv) > 4
dev = torch.device("cuda", 0)
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
b = torch.randn(N, K, device=dev).to(torch.bfloat16)
c = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if bf16 else torch.float32)
for _ in range(3):
    L._gemm(a, b, c, None, M, N, K, K, K, N, bf16_out=bf16)
torch.cuda.synchronize()
