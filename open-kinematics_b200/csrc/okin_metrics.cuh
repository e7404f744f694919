// On-device state metrics, topology metrics and derivative metrics ("motion ratios") of a solved
// state, evaluated from the positions and the tangents dq/dt_j left in shared memory by the solve.
//
// Reference (paths relative to src/kinematics/core/):
//   corner state metrics   metrics/catalog.py:86-159 with context.py:25-165, angles.py:22-132,
//                          steering_geometry.py:22-76, swing_arms.py:46-88, travel.py:19-62,
//                          anti_geometry.py:33-206; instant axis double_wishbone.py:378-431,
//                          macpherson.py:325-379; vector_utils/geometric.py:216-348
//   axle state metrics     metrics/axle_metrics.py:18-95
//   mechanism metrics      corner/mechanisms.py:379-407, :611-623; axle/mechanisms.py:402-430,
//                          :768-808, :931-938
//   derivative metrics     metrics/derivatives.py:247-352, metrics/kernels.py:66-202,
//                          catalog.py:169-308; response_rate / driver_rate along the tangent whose
//                          target point equals the driver's selector point (strongest |driver rate|,
//                          None below 1e-6)
// The reference evaluates responses on dual numbers (primitives/dual.py); OkinDual below is the
// same forward-mode pair and every response is one template over {double, OkinDual}.
// "None" in the reference is NaN in the output buffer.
#pragma once

#include "okin_defs.h"

#define OKIN_RAD2DEG 57.29577951308232
#define OKIN_GEOM_EPS 1e-6

struct OkinDual {
  double v, d;
};
OKIN_HD OkinDual operator+(OkinDual a, OkinDual b) { return {a.v + b.v, a.d + b.d}; }
OKIN_HD OkinDual operator-(OkinDual a, OkinDual b) { return {a.v - b.v, a.d - b.d}; }
OKIN_HD OkinDual operator-(OkinDual a) { return {-a.v, -a.d}; }
OKIN_HD OkinDual operator*(OkinDual a, OkinDual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
OKIN_HD OkinDual operator*(double s, OkinDual a) { return {s * a.v, s * a.d}; }
OKIN_HD OkinDual operator*(OkinDual a, double s) { return {s * a.v, s * a.d}; }
OKIN_HD OkinDual operator/(OkinDual a, OkinDual b) {
  const double q = a.v / b.v;
  return {q, (a.d - q * b.d) / b.v};
}
OKIN_HD OkinDual okin_sqrt(OkinDual a) {
  const double s = sqrt(a.v);
  return {s, a.d / (2.0 * s)};
}
OKIN_HD double okin_sqrt(double a) { return sqrt(a); }
OKIN_HD OkinDual okin_atan2(OkinDual y, OkinDual x) {
  const double den = x.v * x.v + y.v * y.v;
  return {atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / den};
}
OKIN_HD double okin_atan2(double y, double x) { return atan2(y, x); }
OKIN_HD double okin_val(double a) { return a; }
OKIN_HD double okin_val(OkinDual a) { return a.v; }
OKIN_HD double okin_der(double) { return 0.0; }
OKIN_HD double okin_der(OkinDual a) { return a.d; }

template <typename T>
struct OkinV3 {
  T x, y, z;
};
template <typename T>
OKIN_HD OkinV3<T> operator-(OkinV3<T> a, OkinV3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T>
OKIN_HD OkinV3<T> operator+(OkinV3<T> a, OkinV3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T>
OKIN_HD T okin_dot(OkinV3<T> a, OkinV3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T>
OKIN_HD OkinV3<T> okin_cross(OkinV3<T> a, OkinV3<T> b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T>
OKIN_HD OkinV3<T> okin_scale(OkinV3<T> a, T s) { return {a.x * s, a.y * s, a.z * s}; }

OKIN_HD OkinV3<double> okin_lift(const double* p, const double*, double) { return {p[0], p[1], p[2]}; }
OKIN_HD OkinV3<OkinDual> okin_lift(const double* p, const double* v, OkinDual) {
  return {OkinDual{p[0], v[0]}, OkinDual{p[1], v[1]}, OkinDual{p[2], v[2]}};
}
OKIN_HD double okin_const(double c, double) { return c; }
OKIN_HD OkinDual okin_const(double c, OkinDual) { return {c, 0.0}; }

// ---- dual-safe response kernels (metrics/kernels.py) ------------------------------------------
// camber (kernels.py:115-133): wheel_up = cross(axle, X) * (-side); atan2(up_y, up_z), sign by side.
template <typename T>
OKIN_HD T okin_camber_deg(OkinV3<T> ai, OkinV3<T> ao, double side) {
  const OkinV3<T> a = ao - ai;
  // cross(a, (1,0,0)) = (0, a.z, -a.y)
  const T up_y = a.z * (-side), up_z = (-a.y) * (-side);
  const T ang = okin_atan2(up_y, up_z);
  return (side > 0 ? ang : -ang) * OKIN_RAD2DEG;
}
// toe / roadwheel angle (kernels.py:136-156)
template <typename T>
OKIN_HD T okin_toe_deg(OkinV3<T> ai, OkinV3<T> ao, double side) {
  const OkinV3<T> a = ao - ai;
  return (side > 0 ? okin_atan2(a.x, a.y) : okin_atan2(a.x, -a.y)) * OKIN_RAD2DEG;
}
// caster (kernels.py:159-174), kpi (kernels.py:177-202)
template <typename T>
OKIN_HD T okin_caster_deg(OkinV3<T> lo, OkinV3<T> up) {
  const OkinV3<T> s = up - lo;
  return okin_atan2(-s.x, s.z) * OKIN_RAD2DEG;
}
template <typename T>
OKIN_HD T okin_kpi_deg(OkinV3<T> lo, OkinV3<T> up, double side) {
  const OkinV3<T> s = up - lo;
  return okin_atan2(s.y * (-side), s.z) * OKIN_RAD2DEG;
}
// signed rotation of a point about a fixed axis, degrees (kernels.py:66-85; geometric.py:31-52)
template <typename T>
OKIN_HD T okin_rotation_deg(OkinV3<T> p, const double* design, const double* axis_a, const double* axis_b) {
  double ax[3] = {axis_b[0] - axis_a[0], axis_b[1] - axis_a[1], axis_b[2] - axis_a[2]};
  const double il = 1.0 / sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  ax[0] *= il; ax[1] *= il; ax[2] *= il;
  const double dr[3] = {design[0] - axis_a[0], design[1] - axis_a[1], design[2] - axis_a[2]};
  const double dd = dr[0] * ax[0] + dr[1] * ax[1] + dr[2] * ax[2];
  const double dp[3] = {dr[0] - dd * ax[0], dr[1] - dd * ax[1], dr[2] - dd * ax[2]};
  const T zero = okin_const(0.0, T());
  const OkinV3<T> A = {okin_const(ax[0], T()), okin_const(ax[1], T()), okin_const(ax[2], T())};
  const OkinV3<T> DR = {okin_const(dr[0], T()), okin_const(dr[1], T()), okin_const(dr[2], T())};
  const OkinV3<T> DP = {okin_const(dp[0], T()), okin_const(dp[1], T()), okin_const(dp[2], T())};
  const OkinV3<T> cr = p - OkinV3<T>{okin_const(axis_a[0], T()), okin_const(axis_a[1], T()), okin_const(axis_a[2], T())};
  const OkinV3<T> cp = cr - okin_scale(A, okin_dot(cr, A));
  const T sine = okin_dot(A, okin_cross(DR, cr));
  const T cosine = okin_dot(DP, cp);
  (void)zero;
  return okin_atan2(sine, cosine) * OKIN_RAD2DEG;
}
// T-bar shaft twist in radians (axle/mechanisms.py:731-739, :796-808)
template <typename T>
OKIN_HD T okin_tbar_twist_rad(OkinV3<T> left, OkinV3<T> right, OkinV3<T> pivot) {
  const T half = okin_const(0.5, T());
  const OkinV3<T> center = left + okin_scale(right - left, half);
  OkinV3<T> stem = center - pivot;
  const T inv = okin_const(1.0, T()) / okin_sqrt(okin_dot(stem, stem));
  stem = okin_scale(stem, inv);
  OkinV3<T> cb = left - right;
  cb = cb - okin_scale(stem, okin_dot(stem, cb));
  // lateral reference (0,1,0): cross(ref, cb) = (cb.z, 0, -cb.x)
  const T sine = stem.x * cb.z - stem.z * cb.x;
  const T cosine = cb.y;
  return okin_atan2(sine, cosine);
}
template <typename T>
OKIN_HD T okin_distance(OkinV3<T> a, OkinV3<T> b) {
  const OkinV3<T> d = a - b;
  return okin_sqrt(okin_dot(d, d));
}

// ---- metric program records --------------------------------------------------------------------
// Generic op: int32[OKIN_MOP_STRIDE] =
//   {kind, rtype, p0, p1, p2, p3, d0, d1, drv_point, drv_axis, cand_mask, out_col, caux, 0, 0, 0}
#define OKIN_MOP_STRIDE 16
#define OKIN_MOP_VALUE 0
#define OKIN_MOP_DERIV 1
// response types
#define OKIN_R_COORD 0       // dot(p0, axis) with axis = fconst[caux..caux+3)
#define OKIN_R_DIST 1        // |p0 - p1|
#define OKIN_R_CAMBER 2      // p0 = axle in, p1 = axle out, side = fconst[caux]
#define OKIN_R_TOE 3
#define OKIN_R_CASTER 4      // p0 = lower pivot, p1 = upper pivot
#define OKIN_R_KPI 5         // + side
#define OKIN_R_ROTATION 6    // sign * rotation of p0 about axis (p1 -> p2) from design slot d0; sign = fconst[caux]
#define OKIN_R_ROTATION_DIFF 7  // rotation(p0; d0) - rotation(p3; d1) about axis (p1 -> p2)
#define OKIN_R_MID_X 8       // x of midpoint(p0, p1)
#define OKIN_R_TBAR_TWIST_DEG 9     // degrees(twist(p0 = left, p1 = right, p2 = pivot))                    [derivative response]
#define OKIN_R_TBAR_TWIST_DELTA 10  // degrees(twist(current) - twist(design slots d0 = left, d1 = right))  [state metric]
#define OKIN_R_TBAR_HEAVE 11        // rotation of crossbar centre about (pivot, +Y) from the design centre

// Corner record: int32[OKIN_MCORNER_STRIDE] =
//   {ai, ao, wc, cp, lower, upper, ic_type, ic0..ic5, damper_top, damper_bottom, d_wc, d_cp,
//    out_base, caux, flags, ...}; fconst[caux..] = {side, cg_z, wheelbase, front_brake_bias}
#define OKIN_MCORNER_STRIDE 24
#define OKIN_IC_DW 0   // planes (ic0,ic1,ic2) upper and (ic3,ic4,ic5) lower
#define OKIN_IC_MAC 1  // plane (ic0,ic1,ic2) lower arm; strut axis ic2 -> ic3 through ic3
#define OKIN_IC_NONE 2 // the architecture declares no instant centres: those columns are NaN (None)
#define OKIN_MF_FRONT 1
#define OKIN_MF_REAR 2
#define OKIN_MF_HAS_BIAS 4
#define OKIN_MF_DRIVEN_HERE 8
// Axle record: int32[16] = {wcL, wcR, cpL, cpR, d_wcL, d_wcR, d_cpL, d_cpR, rackL, d_rackL, out_base, 0...}
#define OKIN_MAXLE_STRIDE 16

// ---- camber-shim assembly pre-solve (suspensions/config/shims.py:126-501) -------------------------
OKIN_HD OkinDual okin_sin(OkinDual a) { return {sin(a.v), cos(a.v) * a.d}; }
OKIN_HD OkinDual okin_cos(OkinDual a) { return {cos(a.v), -sin(a.v) * a.d}; }
OKIN_HD double okin_sin(double a) { return sin(a); }
OKIN_HD double okin_cos(double a) { return cos(a); }

// Rodrigues rotation of v by the rotation vector w (geometric.py:351-374).  For |w| -> 0 the
// series forms of sin(a)/a and (1-cos a)/a^2 keep the value and its derivative exact.
template <typename T>
OKIN_HD OkinV3<T> okin_rodrigues(OkinV3<T> v, OkinV3<T> w) {
  const T a2 = okin_dot(w, w);
  T s, c;  // s = sin(a)/a, c = (1 - cos a)/a^2
  if (okin_val(a2) < 1e-12) {
    s = okin_const(1.0, T()) - a2 * (1.0 / 6.0);
    c = okin_const(0.5, T()) - a2 * (1.0 / 24.0);
  } else {
    const T a = okin_sqrt(a2);
    s = okin_sin(a) / a;
    c = (okin_const(1.0, T()) - okin_cos(a)) / a2;
  }
  const OkinV3<T> wxv = okin_cross(w, v);
  const OkinV3<T> wxwxv = okin_cross(w, wxv);
  return v + okin_scale(wxv, s) + okin_scale(wxwxv, c);
}

struct OkinShimCtx {
  double t_setup, n[3], wb_axis[3], hl_in[3], hl_len, lbj[3], uwf[3], uwf_to_ubj[3];
  double ubj_to_a[3], ubj_to_b[3], lbj_to_a[3], lbj_to_b[3], lbj_to_hl_out[3];
  int has_rocker;
  double rk_axis_pt[3], rk_axis[3], rk_to_pr_in[3], lbj_to_pr_out[3], pr_len;
};

template <typename T>
OKIN_HD OkinV3<T> okin_c3(const double* p) { return {okin_const(p[0], T()), okin_const(p[1], T()), okin_const(p[2], T())}; }

// Residuals (shims.py:126-268): datum A closure (3), datum B closure (3), normal alignment (3),
// heading-link length (1), optional pushrod length (1).
template <typename T>
OKIN_HD void okin_shim_residuals(const OkinShimCtx& c, const T* x, T* r) {
  const OkinV3<T> wb = okin_scale(okin_c3<T>(c.wb_axis), x[0]);
  const OkinV3<T> cb = {x[1], x[2], x[3]}, ub = {x[4], x[5], x[6]};
  const OkinV3<T> ubj = okin_c3<T>(c.uwf) + okin_rodrigues(okin_c3<T>(c.uwf_to_ubj), wb);
  const OkinV3<T> lbj = okin_c3<T>(c.lbj);
  const OkinV3<T> ncb = okin_rodrigues(okin_c3<T>(c.n), cb);
  const OkinV3<T> nub = okin_rodrigues(okin_c3<T>(c.n), ub);
  const OkinV3<T> cbA = ubj + okin_rodrigues(okin_c3<T>(c.ubj_to_a), cb);
  const OkinV3<T> cbB = ubj + okin_rodrigues(okin_c3<T>(c.ubj_to_b), cb);
  const OkinV3<T> ubA = lbj + okin_rodrigues(okin_c3<T>(c.lbj_to_a), ub);
  const OkinV3<T> ubB = lbj + okin_rodrigues(okin_c3<T>(c.lbj_to_b), ub);
  const T t = okin_const(c.t_setup, T());
  const OkinV3<T> ra = ubA - cbA - okin_scale(ncb, t);
  const OkinV3<T> rb = ubB - cbB - okin_scale(ncb, t);
  const OkinV3<T> rn = nub - ncb;
  r[0] = ra.x; r[1] = ra.y; r[2] = ra.z; r[3] = rb.x; r[4] = rb.y; r[5] = rb.z;
  r[6] = rn.x; r[7] = rn.y; r[8] = rn.z;
  const OkinV3<T> hl_out = lbj + okin_rodrigues(okin_c3<T>(c.lbj_to_hl_out), ub);
  r[9] = okin_distance(hl_out, okin_c3<T>(c.hl_in)) - okin_const(c.hl_len, T());
  if (c.has_rocker) {
    const OkinV3<T> pr_in = okin_c3<T>(c.rk_axis_pt) + okin_rodrigues(okin_c3<T>(c.rk_to_pr_in), okin_scale(okin_c3<T>(c.rk_axis), x[7]));
    const OkinV3<T> pr_out = lbj + okin_rodrigues(okin_c3<T>(c.lbj_to_pr_out), ub);
    r[10] = okin_distance(pr_out, pr_in) - okin_const(c.pr_len, T());
  }
}

OKIN_HD void okin_rotate_about_axis(double* p, const double* pivot, const double* axis, double angle) {
  // Rodrigues about a unit axis through pivot (geometric.py:377-400)
  const double v[3] = {p[0] - pivot[0], p[1] - pivot[1], p[2] - pivot[2]};
  const double ca = cos(angle), sa = sin(angle);
  const double kv = axis[0] * v[0] + axis[1] * v[1] + axis[2] * v[2];
  const double cr[3] = {axis[1] * v[2] - axis[2] * v[1], axis[2] * v[0] - axis[0] * v[2], axis[0] * v[1] - axis[1] * v[0]};
  for (int k = 0; k < 3; ++k) p[k] = pivot[k] + v[k] * ca + cr[k] * sa + axis[k] * (kv * (1.0 - ca));
}
