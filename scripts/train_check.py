import sys, numpy as np
sys.path.insert(0, '.')
from tests.test_gpu_train import run_case, CSP, PS1
W22B = ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1', 'Y_WFCAM', 'J_WFCAM', 'H_WFCAM']
for args in [('M20_model', CSP, np.arange(-8, 40, 4.0), 12, 2, 10), ('T21_model', PS1, np.arange(-8, 40, 4.0), 9, 1, 15),
             ('W22_model', W22B, np.array([-8.0, 2.0, 12.0, 22.0, 32.0]), 7, 4, 11)]:
    for r in run_case(*args):
        print(args[0], 'lp %.2e gg %.2e gl %.2e' % r[:3], {k: '%.1e' % v for k, v in r[3].items()})
