#!/bin/bash
O=gpurun_out/r29; mkdir -p $O
nvidia-smi -L | wc -l | tee $O/ngpu.txt
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 > $O/bench$n.json 2> $O/bench$n.err; echo "bench$n exit $?" | tee -a $O/summary.txt; cut -c1-330 $O/bench$n.json; grep -o '"e2e": {[^}]*}' $O/bench$n.json
done
