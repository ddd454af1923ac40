#!/bin/bash
# per-kernel durations of the label + pairs stage under ncu (cold-cache, serialised): tools/pair_kernel_times.sh [lib.so]
[ -n "$1" ] && export MDVC_LIB=$1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_row_pairs|k_task_pairs|k_strip_link|k_tile_link_x|k_link_round" --csv --log-file /tmp/pt.csv python tools/label_stage.py 1024 2 > /dev/null 2>&1
grep -v "^==" /tmp/pt.csv | tail -6 | python -c "
import csv,sys
for r in csv.reader(sys.stdin): print(f'{r[4][:28]:30s} {float(r[-1])/1e3:8.1f} us')"
