set -x
timeout 60 python tests/_gpu_perf_track.py 592 128 1; echo "rc=$?"
timeout 60 python tests/_gpu_perf_track.py 1184 128 1; echo "rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_sequence_parity.py 2>&1 | tail -30
timeout 600 python -m pytest tests/test_sequence_parity.py -m gpu -x -q -s 2>&1 | tail -15
