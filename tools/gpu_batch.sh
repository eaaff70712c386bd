set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r2s3_sanitizer_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/r2s3_sanitizer_memcheck.log
tail -30 gpurun_out/r2s3_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r2s3_sanitizer_racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/r2s3_sanitizer_racecheck.log
tail -12 gpurun_out/r2s3_sanitizer_racecheck.log
