#!/bin/bash
# round 2, call H: the tests that failed in call G (verbose), synccheck per kernel family
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "report_failed_cells or unsupported or group_widths" 2>&1 | tail -60 > gpurun_out/r02_h_pytest.log
cat gpurun_out/r02_h_pytest.log
CS=/usr/local/cuda/bin/compute-sanitizer
RXN_TM=0 timeout 900 $CS --tool synccheck --print-limit 5 python profiles/sanitize_run.py all 40000 > gpurun_out/r02_h_synccheck_notm.log 2>&1
echo "== synccheck without the tensor-memory kernel rc=$?"; tail -4 gpurun_out/r02_h_synccheck_notm.log
timeout 900 $CS --tool synccheck --print-limit 3 python profiles/sanitize_run.py react 20000 > gpurun_out/r02_h_synccheck_tm.log 2>&1
echo "== synccheck tensor-memory kernel rc=$?"; grep -c "Barrier error" gpurun_out/r02_h_synccheck_tm.log; tail -4 gpurun_out/r02_h_synccheck_tm.log
