import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from orcvio_b200 import api, synth
import helpers as H
config = sys.argv[1] if len(sys.argv) > 1 else "euroc"
pred_only = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ov = dict(if_ZUPT_valid=0, max_features_in_one_grid=0)
if pred_only: ov["prediction_only_flag"] = 1
seq = synth.make_sequence(synth.SynthSpec(config=config, seed=0, n_frames=45, feats_per_frame=120, overrides=ov))
path = H.write_cfg(seq["cfg"])
vio = api.OrcVIO(path); assert vio.initialize()
it = H.run_oracle_sequence(seq)
imu = seq["imu"]; k = 0
for fi, (t_img, feats) in enumerate(seq["frames"]):
    k1 = k
    while k1 < len(imu) and imu[k1][0] <= t_img + 0.02: k1 += 1
    vio.push_imu(imu[k:k1]); k = k1
    vio.processFeatures(t_img, feats)
    ref = next(it)
    st = vio.state(); rs = ref.imu_state
    P = vio.cov()
    upd = [l for l in ref.log if l['kind']=='update']
    print(fi, 'dp %.2e dv %.2e dR %.2e dbg %.2e dP %.2e' % (np.abs(np.array(st.p)-rs.position).max(), np.abs(np.array(st.v)-rs.velocity).max(),
          np.abs(np.array(st.R).reshape(3,3)-rs.orientation).max(), np.abs(np.array(st.bg)-rs.gyro_bias).max(),
          np.abs(P-ref.state_cov).max()/np.abs(ref.state_cov).max()), 'N', st.n_clones, 'max|dx|', (max(np.abs(u['delta_x']).max() for u in upd[-2:]) if upd else 0))
