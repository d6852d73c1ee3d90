set -x
NCU=/usr/local/cuda/bin/ncu
# launch list of the bench command (2 steps, 1 warm-up)
timeout 300 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_f16x3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-train --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full capture of the fused kernel, parity mode and fast mode
timeout 400 $NCU --set full --clock-control none --import-source on -k regex:render_tc_kernel -c 1 -o gpurun_out/r2_fused_f16x3 -f python bench.py --steps 1 --warmup 3 --no-extras --no-train --no-cpu-baseline > gpurun_out/ncu_x3.log 2>&1
timeout 400 $NCU --set full --clock-control none --import-source on -k regex:render_tc_kernel -c 1 -o gpurun_out/r2_fused_f16 -f python bench.py --precision f16 --steps 1 --warmup 3 --no-extras --no-train --no-cpu-baseline > gpurun_out/ncu_f16.log 2>&1
# training: launch list of one step, full captures of the three GEMM shapes
timeout 300 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/train_timing.py --steps 1 > gpurun_out/ncu_train.log 2>&1
timeout 300 $NCU --set full --clock-control none --import-source on -k regex:gemm_tc -c 6 -o gpurun_out/r2_gemm_tc -f python tools/prof_gemm.py --reps 1 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
