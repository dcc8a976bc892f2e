#!/bin/bash
O=gpurun_out/r2_17; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
timeout 600 python tests/sample_types_check.py $L 2>&1 | tail -4 | tee $O/sample_types.txt
timeout 300 python tools/kbench.py $L 512 8 10 2>&1 | tail -1 | cut -c1-260 | tee -a $O/kbench.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest.txt
