import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from distributed_sac_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
lib.b200sac_tc_gemm_timeline.argtypes = [C.c_int32] * 4 + [C.c_void_p]
torch.cuda.init(); torch.zeros(1, device="cuda")
shapes = [(0, 256, 256, 256), (2, 256, 256, 256)]
if len(sys.argv) > 1:      # e.g. 0,4096,400,400 1,2048,400,400 (mode,M,N,K); B200SAC_TC_BN picks the tile width
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for mode, M, N, K in shapes:
    out = (C.c_longlong * 96)()
    rc = lib.b200sac_tc_gemm_timeline(mode, M, N, K, out)
    t = list(out); t0 = t[0]
    rel = lambda i: (t[i] - t0) if t[i] else None
    nk = min(16, (K + 31) // 32)
    print(f"mode {mode} {M}x{N}x{K} rc={rc}: setup {rel(1)}  accum_ready {rel(82)}  epi_done {rel(83)}  end {rel(84)} (cycles)")
    print("  epilogue: first ld", rel(85), " all ld", rel(86), " scratch", rel(87), " rows stored", rel(88))
    print("  tma   ", [rel(2 + k) for k in range(nk)])
    print("  full  ", [rel(18 + 2 * k) for k in range(nk)])
    print("  split ", [rel(19 + 2 * k) for k in range(nk)])
    print("  ready ", [rel(50 + 2 * k) for k in range(nk)])
    print("  commit", [rel(51 + 2 * k) for k in range(nk)])
