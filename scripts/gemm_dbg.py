"""Clock stamps of the projection GEMM's issuing thread (MP_GEMM_DBG): operand starvation vs epilogue gap, effective SM clock."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobileposer_b200 import _cabi

lib = _cabi.lib()
M = 76800
for N, K in [(2048, 512), (2048, 256)]:
    A = torch.randn(M, K, device='cuda')
    W = torch.randn(N, K, device='cuda') / K ** 0.5
    b = torch.randn(N, device='cuda')
    C = torch.zeros(M, N, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for pair in ('0', '1'):
        for il in (0, 1):
            os.environ.pop('MP_GEMM_IL', None)
            if il:
                os.environ['MP_GEMM_IL'] = '1'
            os.environ['MP_GEMM_PAIR_TEST'] = pair
            os.environ.pop('MP_GEMM_DBG', None)
            for _ in range(5):
                _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, 3, s))
            os.environ['MP_GEMM_DBG'] = '1'
            for _ in range(2):
                _cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, 3, s))
            torch.cuda.synchronize()
