"""profiling probe (not a test): the keyframe-rate stages of bench.keyframe_leg once, for an ncu launch list
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/kf_launches.csv python tests/_gpu_prof_keyframe.py [B]"""
import sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import sdv_loam_b200
from sdv_loam_b200 import synth, api
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
w, h = synth.KITTI_WH
ctx = api.Context(synth.KITTI_K, w, h, max_frames=8)
print(json.dumps(bench.keyframe_leg(ctx, api, synth, B, reps=1, warm=1, cpu=False)))
