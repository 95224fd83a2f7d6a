#!/bin/bash
# Round evidence in one GPU call: GPU test suite, smoke, bench (both arms), ncu launch list of the bench command,
# ncu --set full of one steady-state window -> gpurun_out/ (tools/write_profiles_r02.sh turns them into profiles/)
TAG=${1:-r02f}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_${TAG}_ref.json 2>> gpurun_out/bench_$TAG.err
python - <<P
import json
j = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1]); r = json.loads(open("gpurun_out/bench_${TAG}_ref.json").read().strip().splitlines()[-1])
print("ours %.1f M reads/s (%.1f ms), resident %.1f ms; reference arm %.1f M reads/s; parity %s" % (j["value"] / 1e6, j["ms_per_step"], j["resident"]["ms_per_step"], r["value"] / 1e6, j["parity"].get("result")))
P
# launch list of the bench command (cold-cache, serialised: compare shares), then the same command plain
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-parity > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --no-parity > gpurun_out/plain_$TAG.json 2> /dev/null
# one steady-state window, --set full
export MBF_BENCH_READS=20000000
timeout 300 python tools/profile_target.py > gpurun_out/plain_target_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tokens_par|k_par_headers|k_lz_resolve|k_frame_walk|k_count_reads|k_scan|k_frame_fix' --launch-skip ${SKIP:-20} -c ${COUNT:-10} -f -o gpurun_out/prof_$TAG python tools/profile_target.py > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_full_$TAG.log
