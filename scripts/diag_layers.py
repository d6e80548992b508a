import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from helpers import load_golden, state_dict_from_golden, rel_err
import ref_models
import sc2bench_b200 as s2
g = load_golden('fp_bottleneck_small.npz')
sd = state_dict_from_golden(g)
ref = ref_models.build_fp_bottleneck(3, 8, 32); ref.load_state_dict(sd); ref.eval(); ref = ref.double()
mine = s2.get_layer('FPBasedResNetBottleneck', num_input_channels=3, num_bottleneck_channels=8, num_target_channels=32)
mine.load_state_dict(sd); mine.eval().cuda()
x = torch.from_numpy(g['x'])
with torch.no_grad():
    a = x.cuda(); b = x.double()
    for lm, lr in zip(mine.encoder, ref.encoder):
        b_in = b
        b = lr(b)
        # chained
        a = s2.models.run_transform(torch.nn.Sequential(lm), a)
        # isolated: feed the fp64 reference input
        iso = s2.models.run_transform(torch.nn.Sequential(lm), b_in.float().cuda())
        print(type(lm).__name__, 'chained', rel_err(a.cpu(), b), 'isolated', rel_err(iso.cpu(), b), 'max', float(b.abs().max()))
        if not isinstance(lm, torch.nn.Conv2d):
            gm, bt = lm.effective_params(); gr = lr.gamma_reparam(lr.gamma); br = lr.beta_reparam(lr.beta)
            print('   gamma err', rel_err(gm.cpu(), gr), 'beta err', rel_err(bt.cpu(), br))
            tg = torch.nn.functional.conv2d(b_in.float().cuda().abs(), gm.reshape(gm.shape[0], gm.shape[0], 1, 1), bt)
            print('   torch-cuda gdn err', rel_err((b_in.float().cuda() / tg).cpu(), b), torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
