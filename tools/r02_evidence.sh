#!/bin/bash
# Round-2 evidence run on ONE B200 (developer tool): bench lines, launch lists, ncu captures, latency probes.
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests.log 2>&1
python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2>> $O/r02_bench.err
python bench.py --workload pairing --pairs 2 --steps 3 --no-cpu > $O/r02_pairing_k2.json 2>> $O/r02_bench.err
python bench.py --workload pairing --pairs 16 --steps 3 --no-cpu > $O/r02_pairing_k16.json 2>> $O/r02_bench.err
python bench.py --logn 16 --steps 5 --no-cpu > $O/r02_g1_2p16.json 2>> $O/r02_bench.err
python bench.py --logn 18 --steps 5 --no-cpu > $O/r02_g1_2p18.json 2>> $O/r02_bench.err
python bench.py --logn 12 --steps 5 --no-cpu > $O/r02_g1_2p12.json 2>> $O/r02_bench.err
for c in 1 128 1024 4096; do python bench.py --workload pairing --calls $c --steps 5 --no-cpu > $O/r02_pairing_c$c.json 2>> $O/r02_bench.err; done
python tools/quick_bench.py single checked check batch > $O/r02_quick.log 2>&1
python tools/quick_bench.py micro latency > $O/r02_micro.log 2>&1
# launch lists (cold-cache, serialised: shares of the step only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --secondary 0 --no-cpu > $O/r02_ncu_launch1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_g2_launches.csv python bench.py --workload g2msm --steps 1 --warmup 1 --no-cpu > $O/r02_ncu_launch3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_pairing_launches.csv python bench.py --workload pairing --steps 1 --warmup 1 --no-cpu > $O/r02_ncu_launch2.log 2>&1
# full captures of the dominant kernels
# (the .ncu-rep files are summarised ON THE BOX and deleted: gpurun only brings back 64 MiB)
T=/tmp/r02_ncu
mkdir -p $T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -c 1 -o $T/k_accumulate python tools/profile_msm.py 1 20 2 > $O/r02_ncu_full1.log 2>&1
python tools/ncu_summary.py $T/k_accumulate.ncu-rep --traffic k_accumulate_g1_2p20 $O/r02_traffic.json > $O/r02_k_accumulate_ncu_full.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pairing_accumulate_dot6|k_pairing_final_dot6|k_pairing_lines_slots|k_pairing_decode' -c 4 -o $T/pairing python bench.py --workload pairing --steps 1 --warmup 1 --no-cpu > $O/r02_ncu_full2.log 2>&1
python tools/ncu_summary.py $T/pairing.ncu-rep > $O/r02_pairing_kernels_ncu_full.txt 2>&1
python tools/ncu_summary.py $T/pairing.ncu-rep k_pairing_accumulate_dot6 --traffic k_pairing_accumulate_16384 $O/r02_traffic.json > /dev/null 2>&1
ncu -i $T/pairing.ncu-rep --page source --csv > $O/r02_pairing_kernels_source.csv 2>/dev/null
python tools/criterion_replay.py > $O/r02_criterion.txt 2>&1
ls -la $T $O | tail -30
