set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_tstops.py tests/test_save_idxs.py tests/test_gpu_moments_fused.py tests/test_gpu_reference_core.py tests/test_gpu_split.py tests/test_gpu_edge_cases.py tests/test_gpu_parity_tsit5.py -q -m gpu > gpurun_out/r2b_gputest4.log 2>&1
tail -25 gpurun_out/r2b_gputest4.log
(python tools/prof_one.py f32 random 1000000 | tail -1; python tools/prof_one.py f64 random 1000000 | tail -1; python tools/prof_net16.py 200000 101 | tail -2) > gpurun_out/r2b_timing4.log 2>&1
cat gpurun_out/r2b_timing4.log
