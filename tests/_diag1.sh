set -x
timeout 200 python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | grep -v "^$" | tail -40
SDV_TRACK_IMPL=v1 timeout 200 python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -3
timeout 60 python tests/_gpu_perf_track.py 592 128 1
timeout 60 python tests/_gpu_perf_track.py 1184 128 1; echo "rc=$?"
SDV_TRACK_IMPL=v1 timeout 60 python tests/_gpu_perf_track.py 1184 128 1; echo "rc=$?"
timeout 250 compute-sanitizer --tool racecheck --racecheck-report all python tests/_gpu_perf_track.py 8 128 1 2>&1 | tail -40
timeout 200 compute-sanitizer --tool memcheck python tests/_gpu_perf_track.py 8 128 1 2>&1 | tail -30
