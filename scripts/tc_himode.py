"""Which fp32->tf32 conversion does tcgen05 kind::tf32 apply to smem operands?  Accuracy of the 3xTF32
GEMM when the raw fp32 tile is left in place as the "hi" operand and lo is computed under two hypotheses."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from distributed_sac_b200 import _lib
lib = _lib.load()
from test_gpu_tc_gemm import run_tc, rel
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(512, 256, device="cuda", generator=g); W = torch.randn(256, 256, device="cuda", generator=g)
ref = A.double() @ W.double().T
out, _ = run_tc(lib, 0, A, W, 512, 256, 256)
print("hi_mode", os.environ.get("B200SAC_TC_HIMODE", "0"), "rel err vs fp64:", rel(out, ref))
