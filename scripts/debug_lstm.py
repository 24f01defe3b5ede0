"""Stage-by-stage run of the persistent LSTM recurrence (debug aid; run under CUDA_LAUNCH_BLOCKING=1 / compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_error_detection_b200.lstm_stack import lstm_last_hidden

B, W = int(sys.argv[1]) if len(sys.argv) > 1 else 700, int(sys.argv[2]) if len(sys.argv) > 2 else 16
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
torch.manual_seed(0)
F, H = 58, 128
lstm = torch.nn.LSTM(F, H, num_layers=3, batch_first=True, dropout=p).cuda()
x = torch.randn(B, F, W, device="cuda")
gh = torch.randn(B, H, device="cuda")
seed = torch.tensor([5], dtype=torch.int32, device="cuda")
with torch.no_grad():
    h = lstm_last_hidden(x, lstm, training=p > 0, seed_dev=seed)
torch.cuda.synchronize(); print("inference fwd ok", float(h.abs().mean()))
xo = x.clone().requires_grad_(True)
h = lstm_last_hidden(xo, lstm, training=True, seed_dev=seed)
torch.cuda.synchronize(); print("train fwd ok", float(h.abs().mean()))
h.backward(gh)
torch.cuda.synchronize(); print("bwd ok", float(xo.grad.abs().mean()))
if p == 0:
    xr = x.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(enabled=False):
        out, _ = lstm(xr.transpose(1, 2).contiguous())
    g0 = {k: v.grad.clone() for k, v in lstm.named_parameters()}
    lstm.zero_grad()
    out[:, -1, :].backward(gh)
    nrel = lambda a, b: float((a.float() - b).norm() / b.norm())
    print("h", nrel(h.detach(), out[:, -1, :].detach()), "dx", nrel(xo.grad, xr.grad))
    for k, v in lstm.named_parameters():
        print(k, nrel(g0[k], v.grad))
