set -u
O=gpurun_out/s2
mkdir -p $O
( timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py tests/test_grid_store.py tests/test_bev_gpu.py -x -q -m gpu > $O/tests_csm.log 2>&1; echo "tests_csm rc=$?" >> $O/status.txt )
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt )
( GLOC_CSM_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_loc_timing.log 2>&1; echo "bench_loc_timing rc=$?" >> $O/status.txt )
( timeout 1200 python bench.py > $O/bench_loc.log 2>&1; echo "bench_loc rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --loc-policy first --no-cpu-baseline > $O/bench_loc_first.log 2>&1; echo "bench_loc_first rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_loc_ref.log 2>&1; echo "bench_loc_ref rc=$?" >> $O/status.txt )
cat $O/status.txt
tail -5 $O/tests_csm.log
