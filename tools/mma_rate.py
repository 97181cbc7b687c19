import ctypes, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(R + "/tetra-nerf_b200/csrc/libtetranerf_b200.so")
out = (ctypes.c_longlong * 2)()
for mode, name in ((0, "TS"), (4, "TS + concurrent tcgen05.ld/st by 4 warps"), (2, "SS")):
    for boff in (0, 65536, 98304, 131072, 147456):
        nrep = 24
        lib.tn_debug_mma_rate(0, nrep, mode, ctypes.c_uint32(boff), out)
        n = nrep * 8
        print(f"{name:42s} B@{boff:6d}: issue+complete {out[1]/n:7.1f} cyc/MMA")
