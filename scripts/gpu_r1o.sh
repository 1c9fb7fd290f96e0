#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -x -k "train_step" 2>&1 | tail -60 > gpurun_out/r1o_pytest_train.log
tail -60 gpurun_out/r1o_pytest_train.log
