// Stage 3 -- object keypoint / bounding-box reprojection residuals and their Jacobians.
//
// Reference (evaluated at the object-LM optimum to produce (r, H_f, H_c) for the filter):
//   O1  CameraLM::ErrorFeatureQuadric::{operator(),df}   src/obj/ObjectResJacCam.cpp:153-282
//       project_object_points / _df_camera / project_image_df   include/orcvio/utils/se3_ops.hpp:325-453
//   O2  CameraLM::ErrorBBoxQuadric::{operator(),df}      src/obj/ObjectResJacCam.cpp:308-519
//       bbox2poly / poly2lineh / ellipse_from_shape      src/obj/ObjectLM.cpp:380-414
//       circledCirc / odotOperator                       se3_ops.hpp:229-240, 510-519
//   O3  CameraLM::{operator(),df} stacking, get_valid_camera_pose_mat   ObjectResJacCam.cpp:521-604
//   O4  ObjectLM::ErrorFeatureQuadric::df / ErrorBBoxQuadric::df        src/obj/ObjectLM.cpp:318-371, 503-632
//   O5  OrcVIO::constructObjectResidualJacobians         src/orcvio.cpp:2017-2151
// Row order of the outputs = the reference's: all keypoint rows frame by frame (2 per valid
// keypoint), then 4 bounding-box rows per frame; residual weights 1 and Huber epsilon = inf as
// deployed (ObjectFeatureInitializer.h:40, ObjectInitNode.cpp:293-306).  The new-bbox-residual
// Jacobian reproduces the reference as written (its plane uses K*cTw without wTo, :446).
//
// One CTA per frame: one thread per keypoint (residual + both Jacobians), four threads for the
// bounding-box lines, one thread for log(wTc).  The work is tiny (K <= 12 keypoints): the
// kernel exists so that stage 3 lives on the device next to the filter state, not for speed.
#include "kernels.h"
#include "obj_math.cuh"

namespace ob {

namespace {

using namespace objm;

}  // namespace

// frames_wTc: T x 16; zs: T x K x 2 (NaN = not observed); zb: T x 4; kp_row_off[f] = first keypoint
// row of frame f; rows_kp = total keypoint rows.  Outputs column-major with `rows` rows.
__global__ void __launch_bounds__(64) k_object_rows(const double* frames_wTc, int T, const double* wTo_in,
                                                    const double* shape, const double* kps, int K,
                                                    const double* zs, const double* zb, int flags,
                                                    const int* kp_row_off, int rows_kp, int rows,
                                                    double* fvec, double* fjac_cam, double* fjac_obj,
                                                    double* cam_pose_se3) {
  const int f = blockIdx.x, tid = threadIdx.x;
  const bool left = (flags & 1) != 0, new_res = (flags & 2) != 0;
  const int odim = 9 + 3 * K;
  double wTc[16], cTw[16], wTo[16], Pm[16];
  for (int i = 0; i < 16; ++i) { wTc[i] = frames_wTc[16 * f + i]; wTo[i] = wTo_in[i]; }
  inv_rigid(wTc, cTw);
  mm<4, 4, 4>(cTw, wTo, Pm);                       // cTw * wTo (rows 0..2 = P * wTo with K = I)
  if (tid < K) {
    const double zu = zs[((size_t)f * K + tid) * 2], zv = zs[((size_t)f * K + tid) * 2 + 1];
    if (isfinite(zu) && isfinite(zv)) {
      // row index of this keypoint: valid keypoints before it in this frame
      int before = 0;
      for (int q = 0; q < tid; ++q) {
        const double a = zs[((size_t)f * K + q) * 2], b = zs[((size_t)f * K + q) * 2 + 1];
        if (isfinite(a) && isfinite(b)) ++before;
      }
      const int r0 = kp_row_off[f] + 2 * before;
      const double X[4] = {kps[3 * tid], kps[3 * tid + 1], kps[3 * tid + 2], 1.0};
      double Y[4], Z[4];
      mm<4, 4, 1>(wTo, X, Y);
      mm<4, 4, 1>(cTw, Y, Z);
      double d[6];
      dpi_of(Z, d);
      fvec[r0] = Z[0] / Z[2] - zu;
      fvec[r0 + 1] = Z[1] / Z[2] - zv;
      // camera-pose Jacobian (se3_ops.hpp:430-442): left  -dpi [I 0] cTw odot(wTo X)
      //                                             right -dpi [I 0] odot(cTw wTo X)
      double O[24], M36[18], J26[12];
      if (left) {
        odot(Y, O);
        double cO[24];
        mm<4, 4, 6>(cTw, O, cO);
        for (int i = 0; i < 18; ++i) M36[i] = cO[i];
      } else {
        odot(Z, O);
        for (int i = 0; i < 18; ++i) M36[i] = O[i];
      }
      mm<2, 3, 6>(d, M36, J26);
      for (int c = 0; c < 6; ++c) {
        fjac_cam[(size_t)c * rows + r0] = -J26[c];
        fjac_cam[(size_t)c * rows + r0 + 1] = -J26[6 + c];
      }
      // object-state Jacobian (ObjectLM.cpp:318-349): pose 6 | shape 3 (zero) | keypoints 3K
      for (int c = 0; c < odim; ++c) {
        fjac_obj[(size_t)c * rows + r0] = 0.0;
        fjac_obj[(size_t)c * rows + r0 + 1] = 0.0;
      }
      double PO[18];
      if (left) {                                   // dpi P odot(wTo X)
        odot(Y, O);
        double cO[24];
        mm<4, 4, 6>(cTw, O, cO);
        for (int i = 0; i < 18; ++i) PO[i] = cO[i];
      } else {                                      // dpi P wTo odot(X)
        odot(X, O);
        double cO[24];
        mm<4, 4, 6>(Pm, O, cO);
        for (int i = 0; i < 18; ++i) PO[i] = cO[i];
      }
      mm<2, 3, 6>(d, PO, J26);
      for (int c = 0; c < 6; ++c) {
        fjac_obj[(size_t)c * rows + r0] = J26[c];
        fjac_obj[(size_t)c * rows + r0 + 1] = J26[6 + c];
      }
      double PW[9], J23[6];                         // (P wTo)[:, :3]
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) PW[3 * i + j] = Pm[4 * i + j];
      mm<2, 3, 3>(d, PW, J23);
      for (int c = 0; c < 3; ++c) {
        fjac_obj[(size_t)(9 + 3 * tid + c) * rows + r0] = J23[c];
        fjac_obj[(size_t)(9 + 3 * tid + c) * rows + r0 + 1] = J23[3 + c];
      }
    }
  } else if (tid >= 32 && tid < 36) {
    // ---- bounding-box line i (ObjectResJacCam.cpp:308-499, ObjectLM.cpp:503-612)
    const int i = tid - 32;
    const int row = rows_kp + 4 * f + i;
    const double xmin = zb[4 * f], ymin = zb[4 * f + 1], xmax = zb[4 * f + 2], ymax = zb[4 * f + 3];
    const double px[4] = {xmin, xmax, xmax, xmin}, py[4] = {ymin, ymin, ymax, ymax};
    const double ax = px[i], ay = py[i], bx = px[(i + 1) & 3], by = py[(i + 1) & 3];
    const double line[3] = {ay * 1.0 - 1.0 * by, 1.0 * bx - ax * 1.0, ax * by - ay * bx};   // cross([a,1],[b,1])
    const double v2[3] = {shape[0] * shape[0], shape[1] * shape[1], shape[2] * shape[2]};
    // residual: plane through the projected line in the object frame, (P wTo)^T l
    double ub[4];
    for (int c = 0; c < 4; ++c) ub[c] = (Pm[c] * line[0] + Pm[4 + c] * line[1]) + Pm[8 + c] * line[2];
    if (!new_res) {
      fvec[row] = ((v2[0] * ub[0] * ub[0] + v2[1] * ub[1] * ub[1]) + v2[2] * ub[2] * ub[2]) - ub[3] * ub[3];
    } else {
      const double bn = sqrt((ub[0] * ub[0] + ub[1] * ub[1]) + ub[2] * ub[2]);
      const double sq = sqrt((v2[0] * ub[0] * ub[0] + v2[1] * ub[1] * ub[1]) + v2[2] * ub[2] * ub[2]);
      const double sign = ub[3] > 0 ? 1.0 : -1.0;
      fvec[row] = (ub[3] - sign * sq) / bn;
    }
    // Jacobians: yyw = l P (P = K cTw), yyo = yyw wTo
    double yyw[4], yyo[4];
    for (int c = 0; c < 4; ++c) yyw[c] = (line[0] * cTw[c] + line[1] * cTw[4 + c]) + line[2] * cTw[8 + c];
    for (int c = 0; c < 4; ++c)
      yyo[c] = ((yyw[0] * wTo[c] + yyw[1] * wTo[4 + c]) + yyw[2] * wTo[8 + c]) + yyw[3] * wTo[12 + c];
    double wToT[16], cTwT[16];
    tr<4, 4>(wTo, wToT);
    tr<4, 4>(cTw, cTwT);
    const double Q[4] = {v2[0], v2[1], v2[2], -1.0};
    double Jc[6], Jo[6], Js[3];
    double Cc[24], CcT[24];
    if (!new_res) {
      double yq[4];
      for (int c = 0; c < 4; ++c) yq[c] = 2 * yyo[c] * Q[c];      // 2 yyo Qi
      double t4[4];
      mm<1, 4, 4>(yq, wToT, t4);                                  // 2 yyo Qi wTo^T
      // object pose
      if (left) {
        circ(yyw, Cc);
        tr<6, 4>(Cc, CcT);
        mm<1, 4, 6>(t4, CcT, Jo);
      } else {
        double wy[4];
        mm<4, 4, 1>(wToT, yyw, wy);
        circ(wy, Cc);
        tr<6, 4>(Cc, CcT);
        mm<1, 4, 6>(yq, CcT, Jo);
      }
      for (int c = 0; c < 3; ++c) Js[c] = 2 * shape[c] * (yyo[c] * yyo[c]);
      // camera pose
      if (left) {
        circ(yyw, Cc);
        tr<6, 4>(Cc, CcT);
        mm<1, 4, 6>(t4, CcT, Jc);
        for (int c = 0; c < 6; ++c) Jc[c] = -Jc[c];
      } else {
        const double yp[4] = {line[0], line[1], line[2], 0.0};   // l P' with P' = [I 0]
        double t4b[4];
        mm<1, 4, 4>(t4, cTwT, t4b);
        circ(yp, Cc);
        tr<6, 4>(Cc, CcT);
        mm<1, 4, 6>(t4b, CcT, Jc);
        for (int c = 0; c < 6; ++c) Jc[c] = -Jc[c];
      }
    } else {
      // the reference evaluates the plane of the new residual's Jacobian from P = K cTw (:446)
      double ulb[4] = {yyw[0], yyw[1], yyw[2], yyw[3]};
      const double bn = sqrt((ulb[0] * ulb[0] + ulb[1] * ulb[1]) + ulb[2] * ulb[2]);
      const double sq = sqrt((v2[0] * ulb[0] * ulb[0] + v2[1] * ulb[1] * ulb[1]) + v2[2] * ulb[2] * ulb[2]);
      const double sign = ulb[3] > 0 ? 1.0 : -1.0;
      // p_be_p_ulinea = [0 0 0 1] - sign (ulb^T diag(v2,0)) / sq ; p_ulinea_ulineb = I/bn - ulb ulb^T diag(1,1,1,0)/bn^3
      double pa[4];
      for (int c = 0; c < 4; ++c) pa[c] = (c == 3 ? 1.0 : 0.0) - sign * (c < 3 ? ulb[c] * v2[c] : 0.0) / sq;
      double M44[16];
      const double bn3 = bn * bn * bn;
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) M44[4 * r + c] = (r == c ? 1.0 / bn : 0.0) - (c < 3 ? ulb[r] * ulb[c] : 0.0) / bn3;
      double chain[4];
      mm<1, 4, 4>(pa, M44, chain);
      double D46[24];
      // object pose
      if (left) {
        circ(yyw, Cc);
        tr<6, 4>(Cc, CcT);
        mm<4, 4, 6>(wToT, CcT, D46);
      } else {
        double wy[4];
        mm<4, 4, 1>(wToT, yyw, wy);
        circ(wy, Cc);
        tr<6, 4>(Cc, D46);
      }
      mm<1, 4, 6>(chain, D46, Jo);
      for (int c = 0; c < 3; ++c) Js[c] = (shape[c] * (ulb[c] * ulb[c])) / (bn * sq);
      // camera pose
      if (left) {
        circ(yyw, Cc);
        tr<6, 4>(Cc, CcT);
        mm<4, 4, 6>(wToT, CcT, D46);
      } else {
        const double yp[4] = {line[0], line[1], line[2], 0.0};
        double t44[16];
        mm<4, 4, 4>(wToT, cTwT, t44);
        circ(yp, Cc);
        tr<6, 4>(Cc, CcT);
        mm<4, 4, 6>(t44, CcT, D46);
      }
      mm<1, 4, 6>(chain, D46, Jc);
      for (int c = 0; c < 6; ++c) Jc[c] = -Jc[c];
    }
    for (int c = 0; c < 6; ++c) fjac_cam[(size_t)c * rows + row] = Jc[c];
    for (int c = 0; c < odim; ++c) fjac_obj[(size_t)c * rows + row] = 0.0;
    for (int c = 0; c < 6; ++c) fjac_obj[(size_t)c * rows + row] = Jo[c];
    for (int c = 0; c < 3; ++c) fjac_obj[(size_t)(6 + c) * rows + row] = Js[c];
  } else if (tid == 40) {
    double xi[6];
    se3_log(wTc, xi);
    for (int c = 0; c < 6; ++c) cam_pose_se3[(size_t)f * 6 + c] = xi[c];     // 6 x T column-major
  }
}

// O5: constructObjectResidualJacobians -- re-order rows to per-frame [2 k_f keypoint rows, 4 bbox
// rows] for the frames whose timestamp is in the window and place H_c * dcam/dimu at the clone's
// columns.  One CTA per kept frame; map[f] = (first source kp row, k rows, source bbox row,
// destination row, pose index).
__global__ void __launch_bounds__(128) k_object_construct(const double* jac_sensor, const double* Hf,
                                                          const double* res, int rows_in, int odim,
                                                          const int* map5, const double* dcam_dimu /*per kept frame 36*/,
                                                          int leg, int D, int rows_out, double* Hx_out,
                                                          double* Hf_out, double* res_out) {
  const int f = blockIdx.x, tid = threadIdx.x;
  const int src_kp = map5[5 * f], nkp = map5[5 * f + 1], src_bb = map5[5 * f + 2], dst = map5[5 * f + 3],
            pose = map5[5 * f + 4];
  const double* J = dcam_dimu + 36 * f;      // row-major 6x6
  const int nrow = nkp + 4;
  for (int e = tid; e < nrow * 6; e += blockDim.x) {
    const int i = e / 6, c = e - i * 6;
    const int src = (i < nkp) ? (src_kp + i) : (src_bb + (i - nkp));
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += jac_sensor[(size_t)k * rows_in + src] * J[6 * k + c];
    Hx_out[(size_t)(leg + 6 * pose + c) * rows_out + dst + i] = s;
  }
  for (int e = tid; e < nrow * odim; e += blockDim.x) {
    const int i = e / odim, c = e - i * odim;
    const int src = (i < nkp) ? (src_kp + i) : (src_bb + (i - nkp));
    Hf_out[(size_t)c * rows_out + dst + i] = Hf[(size_t)c * rows_in + src];
  }
  for (int i = tid; i < nrow; i += blockDim.x) {
    const int src = (i < nkp) ? (src_kp + i) : (src_bb + (i - nkp));
    res_out[dst + i] = res[src];
  }
  (void)D;
}

void launch_object_rows(const double* frames_wTc, int T, const double* wTo, const double* shape, const double* kps,
                        int K, const double* zs, const double* zb, int flags, const int* kp_row_off, int rows_kp,
                        int rows, double* fvec, double* fjac_cam, double* fjac_obj, double* cam_pose_se3,
                        cudaStream_t s) {
  k_object_rows<<<T, 64, 0, s>>>(frames_wTc, T, wTo, shape, kps, K, zs, zb, flags, kp_row_off, rows_kp, rows, fvec,
                                 fjac_cam, fjac_obj, cam_pose_se3);
  check_launch("k_object_rows");
}

void launch_object_construct(const double* jac_sensor, const double* Hf, const double* res, int rows_in, int odim,
                             const int* map5, const double* dcam_dimu, int n_kept, int leg, int D, int rows_out,
                             double* Hx_out, double* Hf_out, double* res_out, cudaStream_t s) {
  if (n_kept <= 0) return;
  k_object_construct<<<n_kept, 128, 0, s>>>(jac_sensor, Hf, res, rows_in, odim, map5, dcam_dimu, leg, D, rows_out,
                                            Hx_out, Hf_out, res_out);
  check_launch("k_object_construct");
}

}  // namespace ob
