LAMEGPU_TIMING=1 timeout 300 python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 512 8 10 2>&1 | tail -4
timeout 300 python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 4096 8 5 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
