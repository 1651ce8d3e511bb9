"""ncu target: Compress_d + ByteEncode_d and the inverse on 1 Mi polynomials (d = 10)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools_b200.compression import byte_code_dev
dev = torch.device("cuda:0")
npoly, d = 1 << 20, 10
x = torch.randint(0, 3329, (npoly * 256,), dtype=torch.int32, device=dev).to(torch.int16)
p = torch.empty(npoly * 32 * d, dtype=torch.uint8, device=dev)
z = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    byte_code_dev(x.data_ptr(), p.data_ptr(), npoly, 3329, d, st)
    byte_code_dev(p.data_ptr(), z.data_ptr(), npoly, 3329, d, st, decode=True)
torch.cuda.synchronize()
