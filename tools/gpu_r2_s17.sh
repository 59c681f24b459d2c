set -u
O=gpurun_out/s17
mkdir -p $O
( timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py tests/test_host_cpp.py tests/test_driver.py -x -q -m gpu > $O/tests_csm.log 2>&1; echo "tests_csm rc=$?" >> $O/status.txt )
( GLOC_CSM_NO_PAIRED=1 timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py -x -q -m gpu > $O/tests_csm_unpaired.log 2>&1; echo "tests_csm_unpaired rc=$?" >> $O/status.txt )
( GLOC_CSM_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/loc_timing.log 2>&1; echo "loc_timing rc=$?" >> $O/status.txt )
( timeout 900 python bench.py > $O/loc.log 2>&1; echo "loc rc=$?" >> $O/status.txt )
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 8250 -c 260 --csv --log-file $O/launches_localize.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu_launches rc=$?" >> $O/status.txt )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:csm_coarse_bits -s 2 -c 1 -f -o $O/full_coarse_paired python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_coarse.log 2>&1; echo "ncu_coarse rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -3 $O/tests_csm.log; grep "csm\] pairs" $O/loc_timing.log | tail -3
