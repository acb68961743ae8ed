set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:implicit_kernel --launch-skip 6 --launch-count 2 -f -o gpurun_out/r2_final2_implicit python tools/profile_step.py 6 > gpurun_out/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py 4 --all > gpurun_out/ncu_list.log 2>&1
python tools/timeline.py > gpurun_out/timeline_final.txt 2>&1
tail -3 gpurun_out/ncu_full.log
