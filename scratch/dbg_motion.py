import sys; sys.path[:0]=['.','oracle','tests']
import numpy as np, torch
from test_gpu_motion import load, _artifact, CASES, _inject
from fetalsyngen_b200.generator.artifacts import simulate_reco as SR
from fetalsyngen_b200 import _lib
name='motion_default'
g=load(name)
for k in range(int(g['n_attempts'])):
    s=g[f'fwd_mask_sums_{k}']; print('golden', k, np.round(s,1), 'kept', g[f'noise_mask_{k}'].size*8//(96*96) if f'noise_mask_{k}' in g else None)
o=_lib.call
def call(name,*a):
    r=o(name,*a)
    if name=='fsg_slice_sums':
        torch.cuda.synchronize()
    return r
art=_artifact(*CASES[name])
img, seg = torch.from_numpy(g["image"]).cuda(), torch.from_numpy(g["seg"].astype(np.float32)).cuda()
o_sums = None
import fetalsyngen_b200.generator.artifacts.simulate_reco as S
orig_acq = S.slice_acquisition
def acq(mat, vol, psf, shape, res, out=None):
    r = orig_acq(mat, vol, psf, shape, res, out)
    print('ours acq sums', np.round(r.sum((1,2,3)).cpu().numpy(),1)[:30], 'ntaps', (psf!=0).sum())
    return r
S.slice_acquisition = acq
np.random.seed(int(g['seed'])); np.random.rand()
try:
    out, meta = SR.simulate_motion(art, img, seg, [0.5]*3, inject=_inject(g))
except Exception as e:
    print('ERR', e)
