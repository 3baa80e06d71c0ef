#!/bin/bash
# round 2, call AN: parity soak of the two ill-conditioned chemistries with the oracle's own last-bit sensitivity beside it
mkdir -p gpurun_out
timeout 1200 python profiles/parity_soak.py 77001 1.0 mineral_prefactor ascem > gpurun_out/r02_an_parity_soak_illcond.jsonl 2> gpurun_out/r02_an_parity_soak.err
cat gpurun_out/r02_an_parity_soak_illcond.jsonl; tail -2 gpurun_out/r02_an_parity_soak.err
