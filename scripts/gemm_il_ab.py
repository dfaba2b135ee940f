"""A/B of the fp16-split projection GEMM's operand layouts: planes (64-byte TMA rows) vs interleaved (hi, lo) blocks (128-byte rows).
MP_GEMM_IL is read per call by mp_gemm_bias mode 3."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobileposer_b200 import _cabi

lib = _cabi.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 76800
for N, K in [(2048, 512), (2048, 256), (1024, 256), (512, 64)]:
    torch.manual_seed(1)
    A = torch.randn(M, K, device='cuda')
    W = torch.randn(N, K, device='cuda') / K ** 0.5
    b = torch.randn(N, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    ref = (A[:2048].double() @ W.double().T + b.double())
    C = torch.zeros(M, N, device='cuda')
    _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, 1, s))
    torch.cuda.synchronize()
    print(f'M={M} N={N} K={K} fp32 FFMA kernel: max err {(C[:2048].double() - ref).abs().max().item():.2e} rms {(C[:2048].double() - ref).pow(2).mean().sqrt().item():.2e}')
    for il in (0, 1, 2, 3):
        os.environ.pop('MP_GEMM_IL', None)
        os.environ['MP_GEMM_PAIR_TEST'] = '1' if il >= 2 else '0'
        if il in (1, 3):
            os.environ['MP_GEMM_IL'] = '1'
        C = torch.zeros(M, N, device='cuda')
        for _ in range(3):
            _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, 3, s))
        torch.cuda.synchronize()
        err = (C[:2048].double() - ref).abs().max().item()
        rms = (C[:2048].double() - ref).pow(2).mean().sqrt().item()
        tail = (C[-128:].double() - (A[-128:].double() @ W.double().T + b.double())).abs().max().item()
        _cabi.check(lib.mp_profile_enable(1))
        for _ in range(10):
            _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, 3, s))
        prof = _cabi.profile_collect()
        _cabi.check(lib.mp_profile_enable(0))
        for name, v in prof.items():
            if name.startswith('gemm'):
                ms = v['total_ms'] / v['launches']
                print(f'M={M} N={N} K={K} il={il & 1} pair={il >> 1} {name}: {ms:.4f} ms  {6 * M * N * K / ms / 1e9:.0f} TFLOP/s of fp16 products  max err {err:.2e} / tail {tail:.2e} rms {rms:.2e}')
