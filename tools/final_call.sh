#!/bin/bash
# last GPU pass of a round: the suite, the bench line (both arms), smoke, and a fresh ncu capture of the fused tail
set -u
tag=${1:-r2g}
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 $out/${tag}_pytest_gpu.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?"; tail -n 2 $out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference rc=$?"; head -c 300 $out/${tag}_bench_reference.json; echo
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
N=256 SECS=60 timeout 300 ncu --set full --clock-control none --import-source on -k regex:st_post -c 1 -o $out/${tag}_st_post -f python tools/prof_st.py > $out/${tag}_ncu_post.log 2>&1; tail -n 1 $out/${tag}_ncu_post.log
N=256 SECS=180 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tds_offsets -c 1 -o $out/${tag}_tds_offsets_fullsize -f python tools/prof_st.py > $out/${tag}_ncu_tds.log 2>&1; tail -n 1 $out/${tag}_ncu_tds.log
python tools/kbench.py > $out/${tag}_kbench.txt 2>&1; tail -n 3 $out/${tag}_kbench.txt
