#!/bin/bash
# One process per setting (the knobs are read once); output to gpurun_out/r2_pair_sweep.log
out=gpurun_out/r2_pair_sweep.log
: > $out
timeout 20 python tools/pair_probe.py >> $out 2>&1
MUNK_PANEL_BIG_UNTIL=6144 timeout 20 python tools/pair_probe.py >> $out 2>&1
MUNK_PANEL_BIG_UNTIL=2048 MUNK_PANEL_SWITCH=2048 timeout 20 python tools/pair_probe.py >> $out 2>&1
MUNK_PANEL_BIG_UNTIL=16384 timeout 20 python tools/pair_probe.py >> $out 2>&1
cat $out
