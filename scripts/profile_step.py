"""Warm-up + one measured step of the default workload through bench.py's own path, for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-reward", "--no-torch-baseline"] + sys.argv[1:]
import bench
print(bench.run_b200(bench.parse()))
