set -x
timeout 300 python -m pytest tests/test_sequence_parity.py tests/test_gpu_tracker.py -m gpu -x -q -s -k "gpu_replays or set_calib" 2>&1 | tail -15
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench_default.json; tail -5 gpurun_out/r2_bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"track_cluster|pyr_|h2d_words" -s 8336 -c 160 --csv --log-file gpurun_out/r2_launches_bench_b1184.csv python bench.py --steps 2 --warmup 3 --batches 2 --no-cpu-baseline --ba-windows 0 --no-refine --no-extra-legs > gpurun_out/r2_launches_bench.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_cluster_kernel -s 3 -c 1 -o gpurun_out/r2_track_full python bench.py --steps 1 --warmup 3 --batches 1 --no-cpu-baseline --ba-windows 0 --no-refine --no-extra-legs > gpurun_out/r2_track_full.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/
