#!/bin/bash
O=gpurun_out/r19; mkdir -p $O
nvidia-smi -L | tee $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench2.json 2> $O/bench2.err; echo "bench2 exit $?" | tee $O/summary.txt; cut -c1-400 $O/bench2.json; tail -5 $O/bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 > $O/bench2_ref.json 2> $O/bench2_ref.err; echo "bench2 ref exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench2_ref.json
timeout 600 python -m pytest tests -q -m gpu -x -k "shard or multi or dist" > $O/tests.log 2>&1; tail -3 $O/tests.log
