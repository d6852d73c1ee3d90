set -x
NCU=/usr/local/cuda/bin/ncu
# training step with the fused forward chain: launch list of one eager vanilla step, full capture of the two TRAIN launches
timeout 400 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step_fused.csv python tools/train_timing.py --steps 1 --no-graph --only vanilla:tc > gpurun_out/ncu_train2.log 2>&1
timeout 400 $NCU --set full --clock-control none --import-source on -k regex:render_tc_kernel -c 2 -o gpurun_out/r2_train_fwd -f env AON_PROF_ONLY=train python tools/prof_fwd_train.py 2048 > gpurun_out/ncu_train_fwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
