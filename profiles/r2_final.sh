#!/bin/bash
# what the driver runs at round end, in its order; then the handles' throughput record (defaults: depth 5, crowded launches capped at 2)
O=gpurun_out/r2_final2; mkdir -p $O; rm -f $O/*
REF=oracle/_ref/libmp3lame_ref.so
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3 | tee $O/smoke.txt
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; tail -1 $O/bench_ref.json | cut -c1-200
timeout 900 python bench.py > $O/bench_ours.json 2> $O/bench_ours.err; tail -1 $O/bench_ours.json | cut -c1-200
for t in "512 256 1152" "512 128 2304"; do LAMEGPU_LANES=512 timeout 400 tests/c/bin/handles_mt $t 128 $REF 2>&1 | tail -1 | tee -a $O/handles.txt; done
for t in "512 32 1152" "64 128 1152" "1 256 1152"; do HANDLES_MT_RATE_ONLY=1 LAMEGPU_LANES=512 timeout 400 tests/c/bin/handles_mt $t 128 $REF 2>&1 | tail -1 | tee -a $O/handles.txt; done
echo "the synchronous call (LAMEGPU_HANDLE_DEPTH=1, no cap) for comparison" | tee -a $O/handles.txt
for t in "512 256 1152" "64 128 1152" "1 256 1152"; do HANDLES_MT_RATE_ONLY=1 LAMEGPU_HANDLE_DEPTH=1 LAMEGPU_HANDLE_CROWD_CAP=0 LAMEGPU_LANES=512 timeout 400 tests/c/bin/handles_mt $t 128 $REF 2>&1 | tail -1 | tee -a $O/handles.txt; done
echo "VBR -V2 / ABR 128 / VBR-old -V4, checked" | tee -a $O/handles.txt
for m in "4 2" "3 4" "2 4"; do LAMEGPU_LANES=512 timeout 200 tests/c/bin/handles_mt 512 40 1152 128 $REF $m 2>&1 | tail -1 | tee -a $O/handles.txt; done
