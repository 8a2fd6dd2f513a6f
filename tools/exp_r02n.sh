#!/bin/bash
python -m pytest tests/test_gpu_sr.py tests/test_gpu_one_collective.py tests/test_reference_text.py tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -15
