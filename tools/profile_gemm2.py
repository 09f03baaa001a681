import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from samcarriestheburden_b200 import _lib
lib = _lib.load()
M, N, K = 32768, 1280, 5120
A = torch.randn((M, K), device="cuda").bfloat16()
W = (torch.randn((N, K), device="cuda") / K ** 0.5).bfloat16()
b = torch.randn((N,), device="cuda")
out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
for mc in (0, -1, 0, -1):
    _lib.check(lib.b200sam_gemm_bf16(A.data_ptr(), W.data_ptr(), out.data_ptr(), b.data_ptr(), None, M, N, K, K, K, N, N, 0, 0, 1, mc, _lib.current_stream()))
torch.cuda.synchronize()
print("done")
