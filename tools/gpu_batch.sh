set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_tstops.py tests/test_gpu_work_order.py -q -m gpu > gpurun_out/r2b_gputest5.log 2>&1
tail -8 gpurun_out/r2b_gputest5.log
(for i in 1 2; do python tools/prof_one.py f32 random 1000000 | tail -1; python tools/prof_one.py f64 random 1000000 | tail -1; done
for mb in 5 6; do echo "f64 MINBLOCKS=$mb"; B200ENS_MINBLOCKS=$mb python tools/prof_one.py f64 random 1000000 | tail -1; done
for mb in 6 8; do echo "f32 MINBLOCKS=$mb"; B200ENS_MINBLOCKS=$mb python tools/prof_one.py f32 random 1000000 | tail -1; done
python tools/prof_net16.py 200000 101 | tail -2) > gpurun_out/r2b_timing5.log 2>&1
cat gpurun_out/r2b_timing5.log
