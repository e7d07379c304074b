set -x
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_pipeline.py -k "rearranged" -x -q > gpurun_out/sanitizer_seams.log 2>&1; echo "memcheck seams rc=$?"; tail -4 gpurun_out/sanitizer_seams.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_lanes.py -x -q > gpurun_out/sanitizer_lanes.log 2>&1; echo "memcheck lanes rc=$?"; tail -4 gpurun_out/sanitizer_lanes.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_pipeline.py -k "dpx_classes" -x -q > gpurun_out/sanitizer_race.log 2>&1; echo "racecheck dpx rc=$?"; tail -8 gpurun_out/sanitizer_race.log
