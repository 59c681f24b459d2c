set -u
O=gpurun_out/s28
mkdir -p $O
( timeout 600 python -m pytest tests -x -q -m gpu > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/status.txt )
( timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt )
( timeout 400 python bench.py > $O/loc.log 2>&1; echo "loc rc=$?" >> $O/status.txt )
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/loc_ref.log 2>&1; echo "loc_ref rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -3 $O/tests_all.log; tail -2 $O/smoke.log; for f in loc loc_ref; do grep '^{' $O/$f.log | tail -1 | cut -c1-200; done
