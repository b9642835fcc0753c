#!/bin/bash
# One GPU-box pass that produces the evidence kept under profiles/: every program first runs WITHOUT ncu (a number
# printed under a profiler is never a bench value), then the launch list and one --set full capture per main kernel.
# usage (from the repo root, on the GPU box): bash tools/profile_round.sh <tag>
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err || exit 1
python tools/kbench.py > $out/${tag}_kbench.txt 2>&1 || exit 1
small="--tracks 64 --seconds 20 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $small > $out/${tag}_bench_small.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches_bench.csv python bench.py $small > $out/${tag}_ncu_launches.log 2>&1
# dominant kernel at full size: 256 tracks x 180 s through the pitch node
N=256 SECS=180 python tools/prof_st.py || exit 1
N=256 SECS=180 ncu --set full --clock-control none --import-source on -k regex:tds_offsets -c 1 -o $out/${tag}_tds_offsets_fullsize -f python tools/prof_st.py > $out/${tag}_ncu_tds.log 2>&1
N=256 SECS=60 ncu --set full --clock-control none --import-source on -k regex:st_post -c 1 -o $out/${tag}_st_post -f python tools/prof_st.py > $out/${tag}_ncu_post.log 2>&1
python tools/prof_resample.py > /dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:"resample_tile2" -c 2 -o $out/${tag}_resample -f python tools/prof_resample.py > $out/${tag}_ncu_rs.log 2>&1
python tools/prof_stft.py > $out/${tag}_stft_timing.txt || exit 1
ncu --set full --clock-control none --import-source on -k regex:"stft4096" -c 2 -o $out/${tag}_stft -f python tools/prof_stft.py > $out/${tag}_ncu_stft.log 2>&1
tail -q -n 2 $out/${tag}_ncu_tds.log $out/${tag}_ncu_post.log $out/${tag}_ncu_rs.log $out/${tag}_ncu_stft.log
