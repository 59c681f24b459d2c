set -u
mkdir -p gpurun_out/s1
O=gpurun_out/s1
nvidia-smi -L > $O/gpus.txt 2>&1
( GLOC_KNN_PAIR=1 timeout 900 python -m pytest tests/test_knn_gpu.py -x -q -m gpu > $O/pair_tests.log 2>&1; echo "pair_tests rc=$?" >> $O/status.txt )
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_base.log 2>&1; echo "bench_base rc=$?" >> $O/status.txt )
( GLOC_KNN_PAIR=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_pair.log 2>&1; echo "bench_pair rc=$?" >> $O/status.txt )
( timeout 600 python bench.py --workload verify --no-cpu-baseline > $O/bench_verify.log 2>&1; echo "bench_verify rc=$?" >> $O/status.txt )
( GLOC_CSM_TIMING=1 timeout 600 python bench.py --workload verify --steps 2 --warmup 3 --no-cpu-baseline > $O/verify_timing.log 2>&1; echo "verify_timing rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --workload describe --steps 5 --warmup 3 > $O/bench_describe.log 2>&1; echo "bench_describe rc=$?" >> $O/status.txt )
( timeout 600 python bench.py --workload stream --no-cpu-baseline > $O/bench_stream.log 2>&1; echo "bench_stream rc=$?" >> $O/status.txt )
cat $O/status.txt
