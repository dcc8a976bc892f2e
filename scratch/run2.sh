for v in $VARS; do timeout 300 python tools/kbench.py scratch/variants/lib_$v.so 512 8 10 2>&1 | tail -1; done
for v in $VARS4K; do timeout 300 python tools/kbench.py scratch/variants/lib_$v.so 4096 8 5 2>&1 | tail -1; done
