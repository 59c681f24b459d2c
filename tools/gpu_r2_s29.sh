set -u
O=gpurun_out/s29
mkdir -p $O
( timeout 300 python -m pytest tests/test_bev_gpu.py tests/test_driver.py -x -q -m gpu > $O/tests_bev.log 2>&1; echo "tests_bev rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -5 $O/tests_bev.log
