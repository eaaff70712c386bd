set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_work_order.py tests/test_gpu_parity_algs.py tests/test_gpu_split.py tests/test_gpu_parity_tsit5.py -x -q -m gpu > gpurun_out/r2b_gputest1.log 2>&1
tail -3 gpurun_out/r2b_gputest1.log
for w in 0 65536 131072 262144 524288; do echo "window $w"; B200ENS_WORK_WINDOW=$w python tools/prof_one.py f32 random 1000000 | tail -1; B200ENS_WORK_WINDOW=$w python tools/prof_one.py f64 random 1000000 | tail -1; done > gpurun_out/r2b_window_sweep.log 2>&1
cat gpurun_out/r2b_window_sweep.log
for w in 0 131072 262144; do B200ENS_WORK_WINDOW=$w ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:b2_ensemble --csv python tools/prof_one.py f32 random 1000000 2>&1 | grep -v "^==" | tail -9; done > gpurun_out/r2b_window_dram.log 2>&1
cat gpurun_out/r2b_window_dram.log
python tools/prof_net16.py 200000 101 > gpurun_out/r2b_net16_batch5.log 2>&1; B200ENS_DEFINES=B2_EVENT_BATCH=1 python tools/prof_net16.py 200000 101 > gpurun_out/r2b_net16_batch1.log 2>&1; B200ENS_DEFINES=B2_EVENT_BATCH=10 python tools/prof_net16.py 200000 101 > gpurun_out/r2b_net16_batch10.log 2>&1
tail -2 gpurun_out/r2b_net16_batch*.log
python tools/bench_configs.py 4 > gpurun_out/r2b_configs4.log 2>&1; cat gpurun_out/r2b_configs4.log
ncu --set full --clock-control none --import-source on -k regex:b2_ensemble -s 2 -c 1 -o gpurun_out/r2b_gbm_em_f32 python tools/prof_sde.py f32 2000000 > gpurun_out/r2b_ncu_gbm.log 2>&1
tail -3 gpurun_out/r2b_ncu_gbm.log
