set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_moments_fused.py tests/test_gpu_work_order.py tests/test_gpu_parity_algs.py tests/test_gpu_sde_adaptive.py -x -q -m gpu > gpurun_out/r2b_gputest2.log 2>&1
tail -15 gpurun_out/r2b_gputest2.log
(python tools/prof_one.py f32 random 1000000 | tail -1; python tools/prof_one.py f64 random 1000000 | tail -1; python tools/prof_net16.py 200000 101 | tail -2) > gpurun_out/r2b_timing2.log 2>&1
cat gpurun_out/r2b_timing2.log
for d in f32 f64; do ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:b2_ensemble --csv python tools/prof_one.py $d random 1000000 2>&1 | grep -v "^==" | tail -3; done > gpurun_out/r2b_window_dram_default.log 2>&1
cat gpurun_out/r2b_window_dram_default.log
python tools/bench_configs.py 4 > gpurun_out/r2b_configs4b.log 2>&1; cat gpurun_out/r2b_configs4b.log
python tools/summary_fused_probe.py cfg5_101 200000 > gpurun_out/r2b_fused_moments.log 2>&1
python tools/summary_fused_probe.py cfg5_1001 100000 >> gpurun_out/r2b_fused_moments.log 2>&1
python tools/summary_fused_probe.py lorenz401 1000000 >> gpurun_out/r2b_fused_moments.log 2>&1
cat gpurun_out/r2b_fused_moments.log
