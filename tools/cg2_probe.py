"""cta_group::2 bring-up: correctness of the pair MMA against torch and its dispatch cost (cycles per MMA, one cluster alone)"""
import ctypes, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
lib = ctypes.CDLL(R + "/tetra-nerf_b200/csrc/libtetranerf_b200.so")
lib.tn_last_error.restype = ctypes.c_char_p
dev = torch.device("cuda:0")
torch.manual_seed(0)
P = torch.randn(256, 128, device=dev); Q = torch.randn(128, 128, device=dev)
ref = (P.double() @ Q.double().T)
cyc = (ctypes.c_longlong * 2)()
vp = ctypes.c_void_p
for ts in (0, 1):
    for bswap in (0, 1):
        out = torch.zeros(256, 128, device=dev)
        rc = lib.tn_debug_cg2(0, 1, bswap, ts, vp(P.data_ptr()), vp(Q.data_ptr()), vp(out.data_ptr()), cyc)
        if rc: print("error", lib.tn_last_error()); continue
        err = (out.double() - ref).abs().max().item()
        # if the halves are swapped the columns 0..63 / 64..127 come out exchanged
        err_sw = (torch.cat([out[:, 64:], out[:, :64]], 1).double() - ref).abs().max().item()
        print(f"ts={ts} bswap={bswap}: max|err| {err:.3e}  (columns swapped: {err_sw:.3e})  rows0-127 err {(out[:128].double()-ref[:128]).abs().max().item():.2e} rows128-255 err {(out[128:].double()-ref[128:]).abs().max().item():.2e}")
    for nrep in (1, 8, 32):
        out = torch.zeros(256, 128, device=dev)
        lib.tn_debug_cg2(0, nrep, 0, ts, vp(P.data_ptr()), vp(Q.data_ptr()), vp(out.data_ptr()), cyc)
        print(f"ts={ts} nrep={nrep}: issue {cyc[0] / (24 * nrep):.1f} cyc/MMA, issue+complete {cyc[1] / (24 * nrep):.1f} cyc/MMA")
