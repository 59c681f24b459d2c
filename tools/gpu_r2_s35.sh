set -u
O=gpurun_out/s35
mkdir -p $O
( timeout 100 python -m pytest tests/test_localize_gpu.py -x -q -m gpu > $O/tests_loc.log 2>&1; echo "tests_loc rc=$?" >> $O/status.txt )
( timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/loc.log 2>&1; echo "loc rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -2 $O/tests_loc.log; grep '^{' $O/loc.log | tail -1 | cut -c1-200
