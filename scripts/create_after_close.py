"""dtl_create right after a big engine was closed (what bench.py's end-to-end run does): DTL_CREATE_TRACE=1 prints the sections"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dlsim_mrm_b200 as D
from bench import ensure_workload
p = D.read_gmns(ensure_workload("grid100k"), 8)
for rep in range(3):
    e = D.Engine(p, max_columns_per_od=40)
    e.iterations(0, 25)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e.close(); t1 = time.perf_counter()
    print(f"--- rep {rep}: close {1e3 * (t1 - t0):.1f} ms; create follows", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    e2 = D.Engine(p, max_columns_per_od=40)
    torch.cuda.synchronize()
    print(f"--- rep {rep}: create {1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr, flush=True)
    e2.close()
