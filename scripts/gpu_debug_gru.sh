#!/bin/bash
T="tests/test_gpu_sizes.py::test_ithor_sound_branch_backward_at_four_row_tiles"
for i in 1 2; do
  VAR_DEBUG=1 timeout 300 python -m pytest $T -q -x -s 2>&1 | grep -E "var\]|passed|failed|grad rel" | sort | uniq -c | head -12
done
echo "--- VAR_GRU_H16=0"; VAR_GRU_H16=0 timeout 300 python -m pytest $T -q -x -s 2>&1 | grep -E "passed|failed|grad rel" | head -3
echo "--- whole sizes file"; VAR_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_sizes.py -q -x -s 2>&1 | grep -E "ksplit|passed|failed|grad rel" | sort | uniq -c | head -20
