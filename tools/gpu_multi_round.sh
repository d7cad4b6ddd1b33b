#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus N`): sharded-parity tests, bench.py at 2 / 4 / N GPUs, and the unchanged C
# benchmark driver with LIBSSA_B200_GPUS=1 and =N.  Usage: tools/gpu_multi_round.sh <outdir>
out=${1:-gpurun_out/multi}; mkdir -p $out
n=$(python -c "import torch; print(torch.cuda.device_count())"); echo "gpus $n" | tee $out/gpus.txt
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $out/pytest_multi.log
for g in $n 4 2; do
  [ $g -le $n ] || continue
  [ -f $out/bench_n$g.json ] && continue
  extra=""; [ $g -ne $n ] && extra="--no-parity"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) bench.py --gpus $g --steps 5 --warmup 3 $extra \
      > $out/bench_n$g.json 2> $out/bench_n$g.err; echo "bench n=$g rc=$?"
done
python bench.py --steps 5 --warmup 3 --no-per-config --no-parity > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench n=1 rc=$?"
make -s -C examples ssa_benchmark
db=$(ls /tmp/libssa_b200_bench_S570000_*.fas | head -1)
for g in 1 $n; do
  LIBSSA_B200_GPUS=$g examples/ssa_benchmark -n 10 -b 8,16 -t SW -d $db tests/golden/data/P07327 tests/golden/data/P33450 > $out/cbench_g$g.csv 2> $out/cbench_g$g.err; echo "cbench g=$g rc=$?"
done
for g in 1 $n; do echo "== LIBSSA_B200_GPUS=$g"; cat $out/cbench_g$g.csv; done
python - <<P
import json,glob
for f in sorted(glob.glob("$out/bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["e2e"]["value"], d["roofline"]["per_search_kernel_gcups"], d.get("parity_check"))
    except Exception as e: print(f, "ERR", e)
P
