#!/bin/bash
# ncu --set full of K1 for one workload: args workload skip count name [ENV=.. ...]
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
wl=$1; skip=$2; cnt=$3; name=$4; shift 4
CMD="python bench.py --workload $wl --warm-cache --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config"
env "$@" $CMD > $OUT/ncu_plain_$name.log 2>&1 && \
env "$@" ncu --set full --clock-control none --import-source on -k regex:qcut_tally_sliced -s $skip -c $cnt -f -o /tmp/prof_$name $CMD > $OUT/ncu_$name.log 2>&1
echo "ncu $name exit $?"
python scripts/ncu_summary.py /tmp/prof_$name.ncu-rep > $OUT/r02_${name}_ncu_summary.txt 2>&1
python scripts/ncu_opmix.py /tmp/prof_$name.ncu-rep > $OUT/r02_${name}_opcode_mix.txt 2>&1
