import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec
n = next((int(a) for a in sys.argv[1:] if a.isdigit()), 256) << 20
ALG = next((a for a in sys.argv[1:] if a in ("cheetah", "lion")), "cheetah")
C = density_b200.CODECS[ALG]
d_in = synth.synth_mixed(n, device="cuda") if "--mixed" in sys.argv else synth.synth_text(n, device="cuda")
d_out = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
for _ in range(2): codec.encode_device(ALG, d_in, d_out, d_sz, path=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): codec.encode_device(ALG, d_in, d_out, d_sz, path=1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
m = int(d_sz.item())
print(f"{ALG} encode {n>>20} MiB: {ms:.3f} ms  {n/ms/1e6:.2f} GB/s  ratio {n/max(m,1):.3f} (size 0 = copy map did not settle)")
if "--check" in sys.argv:
    import oracle
    want = oracle.encode(ALG, d_in.cpu().numpy())
    print("  bit-exact vs oracle:", m == want.size and bool((d_out[:m].cpu().numpy() == want).all()))
