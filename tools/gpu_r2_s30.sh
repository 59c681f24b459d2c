set -u
O=gpurun_out/s30
mkdir -p $O
( timeout 300 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py -x -q -m gpu > $O/tests_csm.log 2>&1; echo "tests_csm rc=$?" >> $O/status.txt )
( GLOC_CSM_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/loc_timing.log 2>&1; echo "loc_timing rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -2 $O/tests_csm.log; grep "csm\] pairs" $O/loc_timing.log | tail -3 | cut -c1-200; grep '^{' $O/loc_timing.log | tail -1 | cut -c1-160
