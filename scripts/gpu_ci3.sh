#!/bin/bash
python -m pytest tests/test_gpu_lookup.py tests/test_gpu_xline.py -x -q 2>&1 | tail -15
