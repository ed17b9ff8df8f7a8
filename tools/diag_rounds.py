"""Copy-map iteration diagnostics: per fixed-point round, the first block whose copy status changed and how many changed
(density_b200_prot_debug), for Chameleon encode of mixed text / binary data and of noise."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec
lib = density_b200.load()
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 256) << 20
for kind in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["mixed", "random"]):
    d_in = synth.synth_mixed(n, device="cuda") if kind == "mixed" else synth.random_bytes(n, 5, device="cuda")
    d_out = torch.empty(density_b200.Chameleon.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
    d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
    codec.encode_device("chameleon", d_in, d_out, d_sz); torch.cuda.synchronize()
    dbg = (ctypes.c_uint64 * 32)(); lib.density_b200_prot_debug(dbg)
    st = (ctypes.c_uint64 * 6)(); lib.density_b200_encode_status(st)
    nblocks = n // 256
    rounds = [(k, dbg[2 * k], dbg[2 * k + 1]) for k in range(16) if dbg[2 * k + 1]]
    print(f"{kind} {n >> 20} MiB ({nblocks} blocks): converged={st[4]}  rounds with changes: " +
          ", ".join(f"it{k}: {c} blocks from {f} ({100.0 * f / nblocks:.1f} %)" for k, f, c in rounds), flush=True)
