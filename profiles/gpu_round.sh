#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "iteration_cap" 2>&1 | tail -3
