import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from orcvio_b200 import api, synth
import helpers as H
np.set_printoptions(precision=4, linewidth=200, suppress=False)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N, F = 20, 60
snap = synth.stress_snapshot(N, F, 6, seed=11)
sigma2 = 0.002**2*4
tri = dict(cost_threshold=1e-3, init_final_dist_threshold=100.0)
out = api.snapshot_update(snap, flags=flags, noise_var=sigma2, translation_threshold=-1.0, cost_threshold=1e-3, init_final_dist_threshold=100.0)
ref = H.oracle_snapshot_update(snap, flags, sigma2, tri=dict(translation_threshold=-1.0, **tri))
print('status eq', np.array_equal(out['status'], ref['status']), out['status'][:10], ref['status'][:10])
g = np.abs(out['gamma']-ref['gamma'])/np.abs(ref['gamma']); print('gamma rel', g.max())
R = out['R_thin']; print('R nnz', np.count_nonzero(R), 'diag', np.diag(R)[:12], 'rthin', out['r_thin'][:6])
Hs = ref['H'][:,22:]; G_ref = Hs.T@Hs; G = R.T@R
print('G err', np.abs(G-G_ref).max()/np.abs(G_ref).max())
print('b err', np.abs(R.T@out['r_thin'] - Hs.T@ref['r']).max()/np.abs(Hs.T@ref['r']).max())
print('dx err', np.abs(out['delta_x']-ref['delta_x']).max()/np.abs(ref['delta_x']).max(), out['delta_x'][:6], ref['delta_x'][:6])
print('P err', np.abs(out['P']-ref['P']).max()/np.abs(ref['P']).max())
print('timings us', out['timings_us'])
