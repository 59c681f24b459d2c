set -u
O=gpurun_out/s11
mkdir -p $O
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_shortlist_gemm -s 4 -c 1 -f -o $O/full_gemm python bench.py --workload retrieval --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_gemm.log 2>&1; echo "ncu_gemm rc=$?" >> $O/status.txt )
( timeout 900 python -m pytest tests -x -q -m gpu > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -3 $O/tests_all.log
