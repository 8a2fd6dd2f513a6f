#!/bin/bash
python -m pytest tests/test_gpu_sr.py tests/test_reference_text.py -x -q -m gpu 2>&1 | tail -15
