#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
export MBF_BENCH_READS=${MBF_BENCH_READS:-50000000}
python tools/profile_target.py > /dev/null 2>&1
for T in 8 16; do
echo "== read threads $T"
MBF_BAM_READ_THREADS=$T MBF_PROFILE_RESIDENT=0 MBF_BAM_TRACE=1 python tools/profile_target.py 2>&1 | tail -42 | grep -v "finish:" 
done
