"""Per-kernel device times (mp_profile_*: CUDA events around every launch) of the projection GEMMs at the cfg3 shapes:
tcgen05 3xTF32 (mode 2) against tcgen05 3xFP16 split (mode 3: split_f16 of A and W + gemm_f16x3)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobileposer_b200 import _cabi

lib = _cabi.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 76800
for N, K in [(2048, 512), (2048, 256), (1024, 256), (1024, 512), (512, 64), (512, 128)]:
    A = torch.randn(M, K, device='cuda')
    W = torch.randn(N, K, device='cuda') / K ** 0.5
    b = torch.randn(N, device='cuda')
    C = torch.empty(M, N, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for mode in (2, 3):
        if mode == 2 and K % 16:
            continue
        for _ in range(3):
            _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, mode, s))
        torch.cuda.synchronize()
        _cabi.check(lib.mp_profile_enable(1))
        for _ in range(10):
            _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, mode, s))
        prof = _cabi.profile_collect()
        _cabi.check(lib.mp_profile_enable(0))
        for name, v in prof.items():
            ms = v['total_ms'] / v['launches']
            extra = f'  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s fp32-equivalent' if name.startswith('gemm') else f'  {v["algorithmic_bytes"] / v["launches"] / ms / 1e6:.0f} GB/s'
            print(f'M={M} N={N} K={K} mode={mode} {name}: {ms:.4f} ms ({v["launches"]} launches){extra}')
