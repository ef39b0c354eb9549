// SEEPS of one grid point -- the arithmetic of
// /root/reference/weatherbenchX/metrics/categorical.py:217-290 for ONE element,
// written once for the device (seeps.cu) and for the host-compiled copy that
// tests/test_seeps_host.py checks bit for bit against the oracle.
//
// Categories (categorical.py:217-241), all comparisons in float32:
//   dry   = x <= dry_threshold          (a Python scalar: NumPy compares it as
//                                         float32(dry_threshold_mm / 1000))
//   light = dry_threshold < x < wet     (wet: the climatological threshold of
//                                         the point for the valid time)
//   heavy = x >= wet
// The three tests are evaluated independently, exactly as the reference does:
// when wet <= dry_threshold a value can be in two categories or in none, and
// the 3x3 indicator product `forecast_cat * truth_cat` (:263-266) then has
// several or no non-zero entries.  Score (:268-290) = 0.5 * sum over (f, t) of
// indicator[f][t] * S[f][t](p1) with
//   S = [[0,               1/(1-p1),   4/(1-p1)],
//        [1/p1,            0,          3/(1-p1)],
//        [1/p1 + 3/(2+p1), 3/(2+p1),   0       ]]      (float32 arithmetic)
// accumulated in the order of np.einsum over (forecast_cat, truth_cat) in
// float64 -- the indicators are float64 in the reference (`.where(notnull)` on
// a boolean array promotes) -- and returned as float32 (exact when a single
// term is non-zero, which is every case with wet > dry_threshold).
// NaN where the forecast, the observation or p1 is NaN (p1 is pre-masked with
// NaN outside [min_p1, max_p1], :293-294).  A NaN wet threshold compares false
// everywhere: such a point is never light or heavy.
#ifndef WBX_SEEPS_POINT_H_
#define WBX_SEEPS_POINT_H_

#ifndef WBX_HD
#ifdef __CUDACC__
#define WBX_HD __host__ __device__ __forceinline__
#else
#define WBX_HD inline
#endif
#endif

WBX_HD float wbx_seeps_point(float p, float t, float wet, float p1,
                             float dry_threshold) {
  const bool bad = !(p == p) || !(t == t) || !(p1 == p1);
  const float f_cat[3] = {p <= dry_threshold ? 1.f : 0.f,
                          (p > dry_threshold && p < wet) ? 1.f : 0.f,
                          p >= wet ? 1.f : 0.f};
  const float t_cat[3] = {t <= dry_threshold ? 1.f : 0.f,
                          (t > dry_threshold && t < wet) ? 1.f : 0.f,
                          t >= wet ? 1.f : 0.f};
  // float32 scoring matrix, operation by operation as NumPy evaluates it
  const float one_minus = 1.f - p1;
  const float two_plus = 2.f + p1;
  const float inv_p1 = 1.f / p1;
  const float three_over = 3.f / two_plus;
  const float s[3][3] = {
      {0.f, 0.5f * (1.f / one_minus), 0.5f * (4.f / one_minus)},
      {0.5f * inv_p1, 0.f, 0.5f * (3.f / one_minus)},
      {0.5f * (inv_p1 + three_over), 0.5f * three_over, 0.f}};
  double acc = 0.0;
  for (int f = 0; f < 3; ++f)
    for (int k = 0; k < 3; ++k)
      acc += static_cast<double>(f_cat[f] * t_cat[k]) *
             static_cast<double>(s[f][k]);
  if (bad) {
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7fc00000);
#else
    return __builtin_nanf("");
#endif
  }
  return static_cast<float>(acc);
}

#endif  // WBX_SEEPS_POINT_H_
