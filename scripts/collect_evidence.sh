#!/bin/bash
# One gpurun call that produces the raw files behind profiles/ (then: python scripts/make_profiles.py 148 r02).
#   /usr/local/graft/bin/gpurun --timeout 3600 -- 'bash scripts/collect_evidence.sh'
# SANITIZE=1 adds the compute-sanitizer passes (memcheck ~1 min, racecheck ~10 min).
timeout 600 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 70 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo launches rc=$?
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_ -s 10 -c 10 -o gpurun_out/prof_r02 -f python scripts/profile_step.py 148 2 > /dev/null 2>&1; echo ncu rc=$?
timeout 300 python scripts/realistic_report.py gpurun_out/realistic_r02.json > /dev/null 2>&1; echo realistic rc=$?
timeout 600 python scripts/bench_configs.py gpurun_out/configs_r02.json > /dev/null 2>&1; echo configs rc=$?
if [ -n "$SANITIZE" ]; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 9 python -m pytest tests/test_gpu_stereo.py tests/test_gpu_masked.py tests/test_gpu_stored_cells.py tests/test_gpu_realistic.py tests/test_gpu_decode.py tests/test_gpu_match.py -m gpu -x -q -k "not 4096 and not 8192" > gpurun_out/memcheck_r02.log 2>&1; echo memcheck rc=$?; tail -3 gpurun_out/memcheck_r02.log
  timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_stored_cells.py "tests/test_gpu_stereo.py::test_stereo_batch_host_and_device" "tests/test_gpu_masked.py::test_stereo_pipeline_row_band_flag" tests/test_gpu_decode.py -m gpu -x -q > gpurun_out/racecheck_r02.log 2>&1; echo racecheck rc=$?; tail -3 gpurun_out/racecheck_r02.log
fi
