set -u
O=gpurun_out/s32
mkdir -p $O
( timeout 600 python -m pytest tests -x -q -m gpu > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/status.txt )
( timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt )
( timeout 400 python bench.py > $O/loc.log 2>&1; echo "loc rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -3 $O/tests_all.log; tail -1 $O/smoke.log; grep '^{' $O/loc.log | tail -1 | cut -c1-200
