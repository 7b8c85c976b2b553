#!/bin/bash
O=gpurun_out/r2_44; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "torch_custom" > $O/t.log 2>&1; echo "exit $?"; tail -40 $O/t.log
