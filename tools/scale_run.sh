#!/bin/bash
# One 8-GPU box: the multi-GPU numbers of a round, written to gpurun_out/scale_*.json(l).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/scale_run.sh'
# (a) the N-way concurrent pinned H2D floor, (b) weak scaling of config 2 at N = 8 (e2e against that
# floor), (c) STRONG scaling of one book (config 3) and of an 18-book job at N = 1, 2, 4, 8.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
run() { # N, out, args...
  local n=$1 out=$2; shift 2
  port=$((port + 1))
  if [ "$n" = 1 ]; then python "$@" >> gpurun_out/$out 2>> gpurun_out/scale.err
  else $TR --nproc-per-node $n --master-port $port "$@" >> gpurun_out/$out 2>> gpurun_out/scale.err; fi
}
: > gpurun_out/scale.err
for n in 1 2 4 8; do run $n scale_h2d_floor.jsonl tools/h2d_floor.py; done
run 8 scale_h2d_floor_nobind.jsonl tools/h2d_floor.py --no-bind
for n in 4 8; do run $n scale_weak_segments.jsonl bench.py --gpus $n --steps 10 --no-cpu-baseline --no-extras; done
for n in 1 2 4 8; do run $n scale_strong_book.jsonl bench.py --gpus $n --steps 5 --workload chapters --scaling strong --no-cpu-baseline --no-extras; done
for n in 8 4 2; do run $n scale_strong_books18.jsonl bench.py --gpus $n --steps 3 --workload books --scaling strong --no-cpu-baseline --no-extras; done
tail -3 gpurun_out/scale.err
for f in gpurun_out/scale_*.jsonl; do echo "== $f"; cut -c1-400 $f; done
