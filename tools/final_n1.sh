# Final single-GPU evidence of a round, in one gpurun call:  bash tools/final_n1.sh   (outputs under gpurun_out/final/)
set -x
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/gpu_tests.txt; cat $O/gpu_tests.txt
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/reference_arm.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:implicit_kernel --launch-skip 6 --launch-count 2 -f -o $O/implicit python tools/profile_step.py 6 > $O/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python tools/profile_step.py 4 --all > $O/ncu_list.log 2>&1
python tools/timeline.py > $O/timeline.txt 2>&1
python tools/timeline.py SQ_PHASES 2>&1 | tail -10 > $O/phases.txt
SQ_DENSE=1 python tools/timeline.py 2>&1 | head -8 > $O/timeline_dense.txt
python tests/tools/parity_sweep.py --explicit --seeds 12 --out $O/parity_sweep.json > $O/parity_sweep.log 2>&1
for s in 0 1 2 3 8 10 11 14; do timeout 300 python tests/tools/parity_fuzz.py --cases 120 --seed $s --out $O/fuzz_$s.json 2>&1 | tail -1; done > $O/fuzz.log
timeout 300 python tests/tools/parity_fuzz_other.py --cases 80 --seed 4 --out $O/fuzz_other.json 2>&1 | tail -2 > $O/fuzz_other.log
timeout 300 python tests/tools/parity_dense.py --out $O/parity_dense.json 2>&1 | tail -3 > $O/parity_dense.log
tail -3 $O/ncu_full.log; cat $O/fuzz.log $O/fuzz_other.log $O/parity_dense.log; cat $O/bench_n1.json | cut -c1-600
