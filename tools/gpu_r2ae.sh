#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "two_stream" 2>&1 | tail -3
