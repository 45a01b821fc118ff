import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch
from vacnic_b200 import kernels as K
from gemm_bench import bench, lin_fwd
M = 16384
for act, nm in ((K.ACT_NONE, "none"), (K.ACT_GELU, "gelu")):
    for tn in (256, 1256, 128, 1128):
        bench(f"fc1 act={nm} tile_n={tn}", lin_fwd(M, 4096, 1024, act, tn), 2 * M * 4096 * 1024)
