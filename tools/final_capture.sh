#!/bin/bash
# tools/final_capture.sh TAG -- the evidence run for a frozen build, most important first (GPU budget may cut the tail):
#  1. ncu --set full of k_fine<2> on the headline frame  -> gpurun_out/prof_fine_TAG.ncu-rep, profiles/fine_ncu.json (this k_fine.cu hash)
#  2. pytest -m gpu                                        -> gpurun_out/pytest_TAG.log
#  3. the default bench line                               -> gpurun_out/bench_TAG.json
#  4. ncu launch list of the same command                  -> gpurun_out/launches_TAG.csv
tag=$1
timeout -s KILL 100 ncu --set full --clock-control none --import-source on -k regex:k_fine -s 3 -c 1 -o gpurun_out/prof_fine_$tag -f \
    python bench.py --profile-only --steps 2 --warmup 2 > gpurun_out/ncu_$tag.log 2>&1
tail -1 gpurun_out/ncu_$tag.log
python tools/fine_ncu.py gpurun_out/prof_fine_$tag.ncu-rep "paris-like-30k 4096x4096 MSAA16" > gpurun_out/fine_ncu_$tag.json 2>&1
timeout -s KILL 170 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$tag.log 2>&1
tail -2 gpurun_out/pytest_$tag.log
timeout -s KILL 120 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cut -c1-300 gpurun_out/bench_$tag.json
timeout -s KILL 60 ncu --metrics gpu__time_duration.sum --clock-control none -s 54 -c 36 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --profile-only --steps 3 --warmup 3 > gpurun_out/ncu2_$tag.log 2>&1
tail -2 gpurun_out/launches_$tag.csv | cut -c1-200
