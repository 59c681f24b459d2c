set -u
O=gpurun_out/s18
mkdir -p $O
( timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py tests/test_host_cpp.py tests/test_driver.py -x -q -m gpu > $O/tests_csm.log 2>&1; echo "tests_csm rc=$?" >> $O/status.txt )
( GLOC_CSM_NO_PAIRED=1 timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py -x -q -m gpu > $O/tests_csm_unpaired.log 2>&1; echo "tests_csm_unpaired rc=$?" >> $O/status.txt )
( GLOC_CSM_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/loc_timing.log 2>&1; echo "loc_timing rc=$?" >> $O/status.txt )
( timeout 900 python bench.py > $O/loc.log 2>&1; echo "loc rc=$?" >> $O/status.txt )
( GLOC_CSM_NO_FUSED_BUILD=1 timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py -x -q -m gpu > $O/tests_csm_unfused.log 2>&1; echo "tests_csm_unfused rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -3 $O/tests_csm.log; grep "csm\] pairs" $O/loc_timing.log | tail -3
