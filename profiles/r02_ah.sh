#!/bin/bash
# round 2, call AH: RTUpdateKineticState with the multirate sorbed totals on the streaming kernel k_kinmr_update: parity, kernel time before / after
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "global_implicit_entry_points or time_stepped" 2>&1 | tail -3 > gpurun_out/r02_ah_pytest.log; cat gpurun_out/r02_ah_pytest.log
RXN_KINMR_PER_CELL=1 timeout 300 python profiles/bench_kinstate.py hanford300a_mr 500000 > gpurun_out/r02_ah_kinstate_mr_per_cell.json 2>/dev/null; cat gpurun_out/r02_ah_kinstate_mr_per_cell.json
timeout 300 python profiles/bench_kinstate.py hanford300a_mr 500000 > gpurun_out/r02_ah_kinstate_mr_stream.json 2>/dev/null; cat gpurun_out/r02_ah_kinstate_mr_stream.json
timeout 300 python profiles/bench_kinstate.py hanford300a_mr 2000000 > gpurun_out/r02_ah_kinstate_mr_stream_2m.json 2>/dev/null; cat gpurun_out/r02_ah_kinstate_mr_stream_2m.json
