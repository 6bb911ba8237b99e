#!/bin/bash
# Round 2, GPU call A: validate the round's CPU-side changes and take a fresh baseline.   bash scripts/gpu_r2a.sh
set -u
TAG=r02a
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/${TAG}_smoke.log 2>&1; tail -6 $O/${TAG}_smoke.log
timeout 600 python scripts/parity_fullsize.py --out $O/${TAG}_parity_fullsize.json > $O/${TAG}_parity.log 2>&1; echo "parity rc=$?"; tail -3 $O/${TAG}_parity.log | cut -c1-400
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $O/${TAG}_bench.json
IGGT_FUSED_TAIL=0 timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_unfused_tail.json 2> $O/${TAG}_bench_unfused_tail.err; echo "bench(unfused tail) rc=$?"; cut -c1-200 $O/${TAG}_bench_unfused_tail.json
timeout 300 python bench.py --part --size 532 --no-cpu-baseline > $O/${TAG}_bench_part532.json 2> $O/${TAG}_bench_part532.err; echo "bench(part) rc=$?"; cut -c1-200 $O/${TAG}_bench_part532.json
timeout 200 python scripts/microbench.py > $O/${TAG}_mb.json 2> $O/${TAG}_mb.err; echo "mb rc=$?"
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; echo "ref rc=$?"; cut -c1-300 $O/${TAG}_ref.json
# every kernel once under ncu (full set), extracted on the box (the report itself is too big to bring back with sources)
timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/prof_${TAG}_all \
    python scripts/ncu_targets.py > $O/${TAG}_ncu_all.log 2>&1; echo "ncu rc=$?"; tail -2 $O/${TAG}_ncu_all.log
ncu -i /tmp/prof_${TAG}_all.ncu-rep --page raw --csv > $O/${TAG}_ncu_all_raw.csv 2>/dev/null
ls -la /tmp/prof_${TAG}_all.ncu-rep; SZ=$(stat -c %s /tmp/prof_${TAG}_all.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -lt 30000000 ] && [ "$SZ" -gt 0 ]; then cp /tmp/prof_${TAG}_all.ncu-rep $O/; fi
# the two kernels being worked on, with sources
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/prof_${TAG}_tail \
    python scripts/ncu_targets.py --only dpt_tail_fused,conv_nhwc_tail_generic > $O/${TAG}_ncu_tail.log 2>&1
# launch list of one forward (per-launch device times; shares of the step)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --quick --warmup 1 --steps 1 > $O/${TAG}_ncu_list.log 2>&1; echo "launch list rc=$?"
ls -la $O | grep ${TAG}
