"""Isolated timing of mp_gemm_bias (mode 1 = FFMA, 2 = tcgen05 3xTF32) for the shapes of the cfg3 step.
A/B switches of the FFMA kernel (read once per process): MP_GEMM_FFMA2=0 (plain FFMA), MP_GEMM_WIDE=1 (128-wide tile for every N)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobileposer_b200 import _cabi

lib = _cabi.lib()
M = 76800
shapes = [(72, 512), (96, 512), (72, 256), (64, 132), (2, 128), (256, 60), (256, 132)]
if '--all' in sys.argv:
    shapes += [(2048, 512), (2048, 256), (1024, 256)]
for N, K in shapes:
    A = torch.randn(M, K, device='cuda')
    W = torch.randn(N, K, device='cuda') / K ** 0.5
    b = torch.randn(N, device='cuda')
    C = torch.empty(M, N, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for mode in ((1, 2) if (N % 256 == 0 and K % 16 == 0) or os.environ.get('MP_GEMM') == 'tc' and N % 4 == 0 and N >= 16 and K % 16 == 0 else (1,)):
        for _ in range(3):
            _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, mode, s))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, mode, s))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'M={M} N={N} K={K} mode={mode}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s (fp32-equivalent)')
