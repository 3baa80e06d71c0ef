#!/bin/bash
# round 2, call AV: work order under imperfect prediction (two noise realisations of the same cells alternate)
mkdir -p gpurun_out
timeout 900 python profiles/bench_order_prediction.py hanford300a_eq 2000000 > gpurun_out/r02_av_order_prediction_300a.json 2> gpurun_out/r02_av_order_prediction.err; cat gpurun_out/r02_av_order_prediction_300a.json
timeout 900 python profiles/bench_order_prediction.py calcite 4000000 > gpurun_out/r02_av_order_prediction_calcite.json 2>> gpurun_out/r02_av_order_prediction.err; cat gpurun_out/r02_av_order_prediction_calcite.json
timeout 900 python profiles/bench_order_prediction.py hanford300a_mr 1000000 > gpurun_out/r02_av_order_prediction_mr.json 2>> gpurun_out/r02_av_order_prediction.err; cat gpurun_out/r02_av_order_prediction_mr.json
tail -3 gpurun_out/r02_av_order_prediction.err
