#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_trsv.py tests/test_gpu_spmv.py -x -q -m gpu 2>&1 | tail -3
