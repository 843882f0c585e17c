"""One warm-up step + one measured step of the default workload, for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-reward"] + sys.argv[1:]
import bench
bench.run_b200(bench.parse())
