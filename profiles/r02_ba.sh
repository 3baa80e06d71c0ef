#!/bin/bash
# round 2, call BA: the three single-cell differences of the second soak, classified
mkdir -p gpurun_out
timeout 900 python profiles/parity_soak.py 88003 1.0 hanford300a_stoich hanford300a_kinsrf mineral_prefactor > gpurun_out/r02_ba_parity_soak_seed2_three.jsonl 2>/dev/null
cat gpurun_out/r02_ba_parity_soak_seed2_three.jsonl
