#!/bin/bash
set -u
for st in 0 3000 6000 12000; do
  NODEY_ST_STAGGER=$st N=256 SECS=60 timeout 200 python tools/prof_st_time.py 2>&1 | tail -n 1
done
