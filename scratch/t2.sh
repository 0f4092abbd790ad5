#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_parse.py -m gpu -q 2>&1 | tail -12 | cut -c1-250
