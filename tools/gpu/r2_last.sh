#!/bin/bash
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention_core or attention_cond" 2>&1 | tail -1
