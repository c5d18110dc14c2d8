set -x
timeout 300 python -m pytest tests/test_sequence_parity.py tests/test_gpu_tracker.py -m gpu -x -q -s -k "gpu_replays or set_calib" 2>&1 | tail -12
for ns in 0 8000 16000 32000 64000; do
  SDV_TRACK_STAGGER_NS=$ns timeout 60 python tests/_gpu_perf_track.py 592 128 1
  SDV_TRACK_STAGGER_NS=$ns timeout 60 python tests/_gpu_perf_track.py 1184 128 1
done
timeout 60 python tests/_gpu_perf_track.py 1184 64 1
timeout 60 python tests/_gpu_perf_track.py 1776 128 1
timeout 60 python tests/_gpu_perf_track.py 2368 128 1
