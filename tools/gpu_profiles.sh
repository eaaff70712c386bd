# round-2 profile set: launch list of the bench command + ncu --set full of the headline kernels (f32 / f64)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_f32.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r2_launches_bench.log 2>&1
for dt in f32 f64; do
  ncu --set full --clock-control none --import-source on -k regex:b2_ensemble_kernel_adaptive --launch-skip 2 -c 1 -f -o gpurun_out/r2_tsit5_${dt}_final python tools/prof_one.py $dt random 1000000 > gpurun_out/r2_prof_one_${dt}.log 2>&1
done
ls -la gpurun_out
