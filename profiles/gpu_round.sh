#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "padded or chunked" 2>&1 | tail -6
