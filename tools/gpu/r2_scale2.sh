#!/bin/bash
# multi-GPU sanity after the attention changes: bench.py at N GPUs under torchrun
N=${1:-2}
O=gpurun_out/r2_scale2; mkdir -p $O
nvidia-smi -L | wc -l | tee $O/ngpu_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 30 --warmup 5 > $O/bench$N.json 2> $O/bench$N.err; echo "bench$N exit $?" | tee -a $O/summary_$N.txt
cut -c1-300 $O/bench$N.json; grep -o '"e2e": {[^}]*}' $O/bench$N.json; tail -3 $O/bench$N.err
