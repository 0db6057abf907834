#!/bin/bash
# per-kernel durations of the labelling chain (ncu launch list of the last remove_unconnected call), for tuning:
#   profiles/tools/ccl_kernels.sh [lib.so] [shape] [foam]
LIB=$1; SHAPE=${2:-256,2048,2048}; FOAM=$3
[ -n "$LIB" ] && export CICLOPE_B200_LIB=$LIB
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_ccl|k_scan_words|k_pack' --csv \
    python profiles/tools/bench_ccl.py $SHAPE $FOAM 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
acc={}
for r in rows[1:]:
    acc.setdefault(r[ik].split('(')[0],[]).append(float(r[iv].replace(',','')))
print('${LIB:-default} $SHAPE $FOAM', {k: round(sum(v[-3:])/len(v[-3:])/1000,1) for k,v in acc.items()})
"
