"""torch.profiler view of one training step of tools/train_step_bench.py (top CUDA ops by self time)."""
import os, sys
R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0]]
sys.path[:0] = [R_ + "/tools"]
import importlib.util
spec = importlib.util.spec_from_file_location("tsb", R_ + "/tools/train_step_bench.py")
tsb = importlib.util.module_from_spec(spec); spec.loader.exec_module(tsb)
import torch
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tsb.step(0); tsb.step(1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=60))
