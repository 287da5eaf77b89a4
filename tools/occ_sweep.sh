#!/bin/bash
# occupancy sweep of the event loop: launch-bound variants built into build/lib_m{4,5,6}.so
for m in 4 5 6; do
  R=$((148*4*m))
  DMDB_LIB=build/lib_m$m.so python tools/prof_run.py $R 5000 20000
done
DMDB_LIB=build/lib_m4.so python tools/prof_run.py 4736 5000 20000
DMDB_LIB=build/lib_m4.so python tools/prof_run.py 1184 5000 20000
DMDB_LIB=build/lib_m4.so python tools/prof_run.py 148 5000 20000
DMDB_LIB=build/lib_m4.so python tools/prof_run.py 1 5000 20000
