#!/bin/bash
O=gpurun_out/r2_10; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q --durations=10 2>&1 | tail -25 | tee $O/pytest.txt
