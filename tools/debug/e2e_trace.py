"""host-side wall-clock trace of the e2e (JPEG) run_pages call: RETTO_B200_HOST_TRACE=1 python tools/debug/e2e_trace.py"""
import io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench

def main():
    sys.argv = ["bench.py", "--no-cpu-baseline", "--no-variants", "--steps", "3", "--warmup", "3", "--no-forward"] + sys.argv[1:]
    bench.main()

main()
