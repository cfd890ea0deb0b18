/*
 * teb_resize.h — TimedElasticBand::autoResize on packed pose records, one routine compiled for the host
 * (drop-in TimedElasticBand class, C-ABI tebgpu_auto_resize_host) and for the device (k_auto_resize).
 *
 * Reference behaviour: src/timed_elastic_band.cpp:227-286 (autoResize), pose_se2.h:266 (average),
 * g2o average_angle = atan2(sin a + sin b, cos a + cos b), 0 when both sums vanish.
 *
 * Record layout: rec[4*i+0..3] = x_i, y_i, theta_i, dt_i; dt_i connects pose i and i+1; record n-1 has dt = 0.
 * Inserting pose i+1 together with timediff i+1, and deleting pose i+1 together with timediff i, both move whole
 * records, which is why the packed layout keeps the two containers of the reference consistent for free.
 */
#ifndef TEB_RESIZE_H
#define TEB_RESIZE_H

#include <math.h>

#if defined(__CUDACC__)
#define TEB_HD __host__ __device__
#else
#define TEB_HD
#endif

TEB_HD inline double teb_average_angle(double a, double b) {
  double x = cos(a) + cos(b), y = sin(a) + sin(b);
  if (x == 0 && y == 0) return 0;
  return atan2(y, x);
}

/* returns the new n, or -1 when an insertion would exceed n_cap */
TEB_HD inline int teb_auto_resize_records(double* rec, int n, int n_cap, double dt_ref, double dt_hysteresis,
                                          int min_samples, int max_samples, int fast_mode) {
  bool modified = true;
  for (int rep = 0; rep < 100 && modified; ++rep) {
    modified = false;
    for (int i = 0; i < n - 1; ++i) {  /* n-1 == sizeTimeDiffs(), re-evaluated every pass like the reference */
      const double dti = rec[4 * i + 3];
      if (dti > dt_ref + dt_hysteresis && (n - 1) < max_samples) {
        if (dti > 2 * dt_ref) {
          if (n + 1 > n_cap) return -1;
          const double newtime = 0.5 * dti;
          const double ax = (rec[4 * i] + rec[4 * (i + 1)]) / 2;
          const double ay = (rec[4 * i + 1] + rec[4 * (i + 1) + 1]) / 2;
          const double at = teb_average_angle(rec[4 * i + 2], rec[4 * (i + 1) + 2]);
          for (int k = n - 1; k > i; --k) {  /* shift records i+1.. up by one */
            rec[4 * (k + 1)] = rec[4 * k];
            rec[4 * (k + 1) + 1] = rec[4 * k + 1];
            rec[4 * (k + 1) + 2] = rec[4 * k + 2];
            rec[4 * (k + 1) + 3] = rec[4 * k + 3];
          }
          ++n;
          rec[4 * i + 3] = newtime;
          rec[4 * (i + 1)] = ax;
          rec[4 * (i + 1) + 1] = ay;
          rec[4 * (i + 1) + 2] = at;
          rec[4 * (i + 1) + 3] = newtime;
          --i;  /* check the updated pose diff again */
          modified = true;
        } else {
          if (i < n - 2) rec[4 * (i + 1) + 3] += rec[4 * i + 3] - dt_ref;
          rec[4 * i + 3] = dt_ref;
        }
      } else if (dti < dt_ref - dt_hysteresis && (n - 1) > min_samples) {
        if (i < n - 2) {
          const double merged = rec[4 * (i + 1) + 3] + rec[4 * i + 3];
          for (int k = i + 1; k < n - 1; ++k) {  /* drop record i+1 */
            rec[4 * k] = rec[4 * (k + 1)];
            rec[4 * k + 1] = rec[4 * (k + 1) + 1];
            rec[4 * k + 2] = rec[4 * (k + 1) + 2];
            rec[4 * k + 3] = rec[4 * (k + 1) + 3];
          }
          --n;
          rec[4 * i + 3] = merged;
          --i;
        } else {
          if (i > 0) rec[4 * (i - 1) + 3] += rec[4 * i + 3];
          for (int k = i; k < n - 1; ++k) {  /* drop record i (pose i), goal record moves down */
            rec[4 * k] = rec[4 * (k + 1)];
            rec[4 * k + 1] = rec[4 * (k + 1) + 1];
            rec[4 * k + 2] = rec[4 * (k + 1) + 2];
            rec[4 * k + 3] = rec[4 * (k + 1) + 3];
          }
          --n;
        }
        modified = true;
      }
    }
    if (fast_mode) break;
  }
  if (n >= 1) rec[4 * (n - 1) + 3] = 0.0; /* an empty band stays empty (the reference has no time difference to touch) */
  return n;
}

#endif
