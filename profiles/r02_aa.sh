#!/bin/bash
# round 2, call AA: GPU suite with the microbial / immobile additions (RMicrobial, RImmobileDecay, immobile dofs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_aa_pytest_gpu.log
cat gpurun_out/r02_aa_pytest_gpu.log
