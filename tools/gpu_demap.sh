#!/bin/bash
python tools/quick_perf.py 16 2>&1 | grep "^demap"
python -m pytest tests/test_gpu_modem.py -m gpu -x -q 2>&1 | tail -3
