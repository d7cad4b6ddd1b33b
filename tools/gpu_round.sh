#!/bin/bash
# One GPU-box session of a round: GPU tests, the benchmark line, the ncu launch list and the ncu --set full captures of the
# dominant kernels (each only after the same command has exited 0 without ncu).  Usage: tools/gpu_round.sh <tag> [outdir]
tag=${1:-r02}; out=${2:-gpurun_out}
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "bench ref rc=$?"
python bench.py --steps 2 --warmup 1 --no-per-config --no-parity > $out/bench_short_$tag.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-per-config --no-parity > $out/ncu_launches_$tag.log 2>&1; echo "launch list rc=$?"
for q in P33450 P07327; do
  python tools/profile_target.py $q sw 570000 > $out/target_$q.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:sw32_kernel -s 2 -c 1 \
      -o $out/prof_${tag}_$q -f python tools/profile_target.py $q sw 570000 > $out/ncu_$q.log 2>&1; echo "ncu $q rc=$?"; cat $out/target_$q.log
done
