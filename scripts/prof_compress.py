import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools_b200.compression import compress_dev
dev = torch.device("cuda:0")
count = 1 << 28
x = torch.randint(0, 3329, (count,), dtype=torch.int32, device=dev).to(torch.int16)
y = torch.empty_like(x); z = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    compress_dev(x.data_ptr(), y.data_ptr(), count, 3329, 11, st)
    compress_dev(y.data_ptr(), z.data_ptr(), count, 3329, 11, st, decompress=True)
torch.cuda.synchronize()
