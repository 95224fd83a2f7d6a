#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
export MBF_BENCH_READS=${MBF_BENCH_READS:-50000000}
python tools/profile_target.py 2>&1 | tail -1
MBF_PROFILE_RESIDENT=0 MBF_BAM_TRACE=1 python tools/profile_target.py 2>&1 | grep -E "win [5-7] |records"
MBF_PROFILE_RESIDENT=0 python tools/profile_target.py 2>&1 | tail -1
python tools/profile_target.py 2>&1 | tail -1
