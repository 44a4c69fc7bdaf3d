"""One forward+backward of the per-cell LSTM at the bench shape (for ncu captures of the LSTM kernels)."""
import sys, torch
sys.path.insert(0, ".")
from mpgcn_b200 import ops
B, T, N = 4, 12, 1000
dev = torch.device("cuda:0")
torch.manual_seed(0)
lstm = torch.nn.LSTM(1, 32, 1, batch_first=True).to(dev)
ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
x = torch.rand(B, T, N, N, 1, device=dev) * 4
d = torch.randn(B * N * N, 32, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    h = ops.lstm_last(x, *ws, precision="fp16")
    h.backward(d)
torch.cuda.synchronize()
print("ok", float(h.abs().mean()))
