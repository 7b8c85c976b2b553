#!/bin/bash
# multi-GPU pass: bench.py at N GPUs (+ the 1868-clip dataset-scale run, BASELINE config 4)
N=${1:-8}
O=gpurun_out/r2_scale; mkdir -p $O
nvidia-smi -L | wc -l | tee $O/ngpu_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --warmup 5 > $O/bench$N.json 2> $O/bench$N.err; echo "bench$N exit $?" | tee -a $O/summary_$N.txt
cut -c1-300 $O/bench$N.json; grep -o '"e2e": {[^}]*}' $O/bench$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/sample_dataset.py --clips 1868 --batch 64 > $O/dataset$N.json 2> $O/dataset$N.err; echo "dataset$N exit $?" | tee -a $O/summary_$N.txt
cat $O/dataset$N.json; tail -3 $O/dataset$N.err
