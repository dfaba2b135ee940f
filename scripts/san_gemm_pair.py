"""Small CTA-pair projection GEMM for compute-sanitizer (mp_gemm_bias mode 3, MP_GEMM_PAIR_TEST=1): 2 x 4 pair tiles, checked against fp64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobileposer_b200 import _cabi

os.environ['MP_GEMM_PAIR_TEST'] = '1'
lib = _cabi.lib()
M, N, K = 1000, 512, 256
A = torch.randn(M, K, device='cuda')
W = torch.randn(N, K, device='cuda') / K ** 0.5
b = torch.randn(N, device='cuda')
C = torch.zeros(M, N, device='cuda')
_cabi.check(lib.mp_gemm_bias(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, 3, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print('pair gemm max err', (C.double() - (A.double() @ W.double().T + b.double())).abs().max().item())
