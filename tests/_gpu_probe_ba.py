import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import sdv_loam_b200
from sdv_loam_b200 import synth, api
import orc
from conftest import cached_sequence
seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
w,h = synth.KITTI_WH
win = synth.make_ba_window(seq, list(range(7)), n_per_frame=250, seed=3, pose_noise=(0.005,0.0003), match_noise=0.1, prior_scale=1e-3)
frames=[orc.Frame(seq.images[k],4) for k in win['kf_idx']]
ob=orc.BAWindow(win, frames)
ctx=api.Context(synth.KITTI_K,w,h,max_frames=8)
for i,k in enumerate(win['kf_idx']): ctx.makeImages(100+i, seq.images[k])
gb=api.EnergyFunctional(ctx, win, [100+i for i in range(7)])
def cmp(name,a,b,rel=True):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64)
    d=np.abs(a-b).max() if a.size else 0; s=np.abs(a).max() if a.size else 1
    print(f"  {name}: maxabs diff {d:.3e} (scale {s:.3e})", "EXACT" if d==0 else "")
for (hh,tt) in ((0,1),(3,6),(6,2)):
    po=ob.precalc(hh,tt); pg=gb.precalc(hh,tt)
    for k in po: cmp(f"precalc[{hh},{tt}].{k}", po[k], pg[k])
ob.reset_oob(); gb.reset_oob()
eo=ob.linearizeAll(False); eg=gb.linearizeAll(False); print("linearize energy", eo, eg, abs(eo-eg)/eo)
ro=ob.residuals(); rg=gb.residuals()
print("  new_state equal:", np.array_equal(ro['new_state'],rg['new_state']), np.bincount(ro['new_state'],minlength=3))
cmp("J", ro['J'], rg['J']); cmp("energies", ro['energies'], rg['energies']); cmp("center", ro['center'], rg['center'])
print("  frameEnergyTH", ob.frames()['frameEnergyTH'], gb.frames()['frameEnergyTH'])
print("  L/M energy", ob.calcLEnergy(), ob.calcMEnergy(), gb.energies())
ob.applyRes(); gb.applyRes()
ro=ob.residuals(); rg=gb.residuals(); print("  active equal", np.array_equal(ro['active'],rg['active'])); cmp("efJ", ro['efJ'], rg['efJ']); cmp("JpJdF", ro['JpJdF'], rg['JpJdF'])
ob.backupState(); gb.backupState()
for it,lam in ((0,0.1),(2,0.025)):
    HAo,bAo,Hsco,bsco = ob.accumulate(); xo,HSo,bSo = ob.solveSystem(it,lam)
    xg,HSg,bSg,(HAg,bAg,Hscg,bscg) = gb.solveSystem(it,lam)
    print("solve it",it); cmp("HA",HAo,HAg); cmp("bA",bAo,bAg); cmp("Hsc",Hsco,Hscg); cmp("bsc",bsco,bscg); cmp("HS",HSo,HSg); cmp("bS",bSo,bSg); cmp("x",xo,xg)
    po=ob.points(); pg=gb.points(); cmp("HdiF",po['HdiF'],pg['HdiF']); cmp("bdSumF",po['bdSumF'],pg['bdSumF']); cmp("pt step",po['step'],pg['step'])
    fo=ob.frames(); fg=gb.frames(); cmp("frame step", fo['step'], fg['step'])
cbo=ob.doStepFromBackup(1.0); cbg=gb.doStepFromBackup(1.0); print("canbreak", cbo, cbg)
fo=ob.frames(); fg=gb.frames(); cmp("state after step", fo['state'], fg['state']); cmp("PRE_w2c", fo['PRE_worldToCam'], fg['PRE_worldToCam'])
eo=ob.linearizeAll(False); eg=gb.linearizeAll(False); print("linearize energy after step", eo, eg)
# full optimize on fresh windows
ob2=orc.BAWindow(win, frames); gb2=api.EnergyFunctional(ctx, win, [100+i for i in range(7)])
t0=time.time(); r_o=ob2.optimize(6); t1=time.time(); r_g=gb2.optimize(6); t2=time.time()
print("optimize oracle", r_o, "cpu s", t1-t0); print("optimize gpu   ", r_g, "wall s", t2-t1)
fo=ob2.frames(); fg=gb2.frames(); cmp("final T_eval", fo['T_eval'], fg['T_eval']); cmp("final state", fo['state'], fg['state']); cmp("final frameEnergyTH", fo['frameEnergyTH'], fg['frameEnergyTH'])
po=ob2.points(); pg=gb2.points(); cmp("final idepth", po['idepth'], pg['idepth']); cmp("maxRelBaseline", po['maxRelBaseline'], pg['maxRelBaseline']); print("  numGood equal", np.array_equal(po['numGood'],pg['numGood']))
ro=ob2.residuals(); rg=gb2.residuals(); print("  final state equal", np.array_equal(ro['state'],rg['state']), "toRemove equal", np.array_equal(ro['toRemove'],rg['toRemove']), np.bincount(ro['state'],minlength=3))
