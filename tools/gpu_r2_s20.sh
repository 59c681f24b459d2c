set -u
O=gpurun_out/s20
mkdir -p $O
( GLOC_CSM_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/loc_timing.log 2>&1; echo "loc_timing rc=$?" >> $O/status.txt )
( timeout 900 python bench.py > $O/loc.log 2>&1; echo "loc rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/loc_ref.log 2>&1; echo "loc_ref rc=$?" >> $O/status.txt )
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 8250 -c 260 --csv --log-file $O/launches_localize.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu_launches rc=$?" >> $O/status.txt )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:csm_coarse_bits -s 2 -c 1 -f -o $O/full_coarse_paired python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_coarse.log 2>&1; echo "ncu_coarse rc=$?" >> $O/status.txt )
cat $O/status.txt; grep "csm\] pairs" $O/loc_timing.log | tail -3; grep '^{' $O/loc.log | tail -1 | cut -c1-200
